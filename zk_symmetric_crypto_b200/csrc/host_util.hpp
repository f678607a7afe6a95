// Host-side pieces of the prover that stay on the CPU: circle-group points / twiddle tables, the Blake2s Fiat-Shamir
// channel, query sampling and small serialisation helpers.
//
// Upstream modules restated (stwo rev f117d487, un-vendored; see DESIGN.md "oracle pinning"): core/circle.rs,
// core/poly/circle/{canonic,domain}.rs, core/channel/blake2s.rs, core/queries.rs, core/constraints.rs (coset_vanishing).
// Reference call sites: /root/reference/stwo/src/chacha/bitwise/air_stream.rs:66-123 (statement mixing), :185-231.
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include <set>
#include <algorithm>
#include "m31.cuh"
#include "blake2s.cuh"

namespace host {
using namespace m31;

struct Pt {
    uint32_t x, y;
};
inline Pt pt_add(Pt p, Pt q) { return {sub(mul(p.x, q.x), mul(p.y, q.y)), add(mul(p.x, q.y), mul(p.y, q.x))}; }
inline Pt pt_double(Pt p) { return pt_add(p, p); }
constexpr Pt GEN{2, 1268011823};

inline Pt index_to_point(uint32_t idx) {
    idx &= 0x7fffffffu;
    Pt r{1, 0}, g = GEN;
    while (idx) {
        if (idx & 1) r = pt_add(r, g);
        g = pt_double(g);
        idx >>= 1;
    }
    return r;
}
inline uint32_t subgroup_gen(int log_size) { return 1u << (31 - log_size); }

struct Coset {
    uint32_t initial, step;
    int log_size;
    static Coset odds(int k) { return {subgroup_gen(k + 1), subgroup_gen(k), k}; }
    static Coset half_odds(int k) { return {subgroup_gen(k + 2), subgroup_gen(k), k}; }
    uint32_t index_at(uint32_t i) const { return (initial + step * i) & 0x7fffffffu; }
    Pt at(uint32_t i) const { return index_to_point(index_at(i)); }
    // all points in coset order
    std::vector<Pt> points() const {
        size_t n = (size_t)1 << log_size;
        std::vector<Pt> p(n);
        p[0] = index_to_point(initial);
        Pt s = index_to_point(step);
        for (size_t m = 1; m < n; m <<= 1) {
            for (size_t i = 0; i < m; i++) p[m + i] = pt_add(p[i], s);
            s = pt_double(s);
        }
        return p;
    }
};

inline uint32_t bit_reverse(uint32_t i, int bits) {
    uint32_t r = 0;
    for (int b = 0; b < bits; b++) r |= ((i >> b) & 1u) << (bits - 1 - b);
    return r;
}

// canonic circle domain of log size m: at(i) = half_odds(m-1).at(i) for i < half, conjugate otherwise
inline uint32_t canonic_index_at(int m, uint32_t i) {
    Coset h = Coset::half_odds(m - 1);
    uint32_t half = 1u << (m - 1);
    if (i < half) return h.index_at(i);
    return (0x80000000u - h.index_at(i - half)) & 0x7fffffffu;
}

// Flattened twiddle tables for canonic domains up to log size max_log (see kernels_fft.cu header).
struct TwiddleTables {
    std::vector<uint32_t> X, Y, IX, IY;
    int max_log = 0;
};

inline void batch_inverse(std::vector<uint32_t>& v) {
    // Montgomery trick; zeros are left as zeros
    std::vector<uint32_t> pre(v.size());
    uint32_t acc = 1;
    for (size_t i = 0; i < v.size(); i++) {
        pre[i] = acc;
        if (v[i]) acc = mul(acc, v[i]);
    }
    uint32_t inv_acc = inv(acc);
    for (size_t i = v.size(); i-- > 0;) {
        if (!v[i]) continue;
        uint32_t t = mul(inv_acc, pre[i]);
        inv_acc = mul(inv_acc, v[i]);
        v[i] = t;
    }
}

// shifted = false: the tower of canonic half cosets half_odds(k) (initial g_(k+2), step g_k).
// shifted = true:  the tower with initial g_(k+3): its log-k circle domain is the first half (storage rows [0, 2^k)) of the
// canonic domain of log size k+1, i.e. the sub-domain on which the trace coset's vanishing polynomial is constant.
inline TwiddleTables make_twiddles(int max_log, bool shifted = false) {
    TwiddleTables t;
    t.max_log = max_log;
    size_t ny = (size_t)1 << max_log;  // Y[k], k=0..max_log-1 at offset 2^k
    t.Y.assign(ny, 0);
    t.X.assign(ny / 2 + 1, 0);         // X[k], k=1..max_log-1 at offset 2^(k-1)
    for (int k = 0; k < max_log; k++) {
        std::vector<Pt> p = (shifted ? Coset{subgroup_gen(k + 3), subgroup_gen(k), k} : Coset::half_odds(k)).points();
        size_t n = (size_t)1 << k;
        for (size_t j = 0; j < n; j++) t.Y[n + j] = p[bit_reverse((uint32_t)j, k)].y;
        if (k >= 1)
            for (size_t j = 0; j < n / 2; j++) t.X[n / 2 + j] = p[bit_reverse((uint32_t)j, k - 1)].x;
    }
    t.IX = t.X;
    t.IY = t.Y;
    batch_inverse(t.IX);
    batch_inverse(t.IY);
    return t;
}

// core/constraints.rs coset_vanishing(Coset::odds(trace_log), p) for a base-field point p
inline uint32_t coset_vanishing_m31(int trace_log, Pt p) {
    Coset c = Coset::odds(trace_log);
    Pt t = index_to_point((0x80000000u - c.initial + (c.step >> 1)) & 0x7fffffffu);
    Pt q = pt_add(p, t);
    uint32_t x = q.x;
    for (int i = 1; i < trace_log; i++) x = sub(mul(2, mul(x, x)), 1);
    return x;
}

// ---------------------------------------------------------------------------------------------- Blake2s channel
struct Hash32 {
    uint8_t b[32];
};

inline Hash32 blake2s_bytes(const uint8_t* d, size_t n) {
    Hash32 h;
    blake2s::hash(d, n, h.b);
    return h;
}

inline uint32_t load_le32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
// little-endian host (the same assumption as blake2s.cuh): integers are appended as their memory image
inline void put_u32(std::vector<uint8_t>& v, uint32_t x) {
    const uint8_t* b = (const uint8_t*)&x;
    v.insert(v.end(), b, b + 4);
}
inline void put_u64(std::vector<uint8_t>& v, uint64_t x) {
    const uint8_t* b = (const uint8_t*)&x;
    v.insert(v.end(), b, b + 8);
}
inline void put_bytes(std::vector<uint8_t>& v, const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    v.insert(v.end(), b, b + n);
}
inline void put_qm31(std::vector<uint8_t>& v, const QM31& q) { put_bytes(v, q.v, 16); }

// Blake2sChannel as the reference binary behaves (oracle/trace_blake.py):
//   mix: digest = H(digest || payload);  draw: H(digest || n_sent as 4 LE bytes || 0x00), n_sent++
struct Channel {
    Hash32 digest;
    uint32_t n_sent = 0;
    Channel() { memset(digest.b, 0, 32); }
    void update(const Hash32& d) {
        digest = d;
        n_sent = 0;
    }
    void mix_bytes(const uint8_t* p, size_t n) {
        if (n <= 32) {
            uint8_t buf[64];
            memcpy(buf, digest.b, 32);
            memcpy(buf + 32, p, n);
            update(blake2s_bytes(buf, 32 + n));
            return;
        }
        // H(digest || payload) without copying the payload: the first block is digest || payload[0..32), the rest streams
        uint8_t first[64];
        memcpy(first, digest.b, 32);
        memcpy(first + 32, p, 32);
        blake2s::Incremental inc;
        inc.update(first, 64, true);
        Hash32 out;
        inc.update(p + 32, n - 32, false, out.b);
        update(out);
    }
    void mix_root(const Hash32& r) { mix_bytes(r.b, 32); }
    void mix_u64(uint64_t v) {
        uint8_t b[8];
        for (int i = 0; i < 8; i++) b[i] = (uint8_t)(v >> (8 * i));
        mix_bytes(b, 8);
    }
    void mix_felts(const QM31* f, size_t n) {
        // canonical QM31 values are 4 little-endian u32 words each: the array is its own serialisation
        static_assert(sizeof(QM31) == 16, "QM31 layout");
        mix_bytes((const uint8_t*)f, 16 * n);
    }
    void draw_u32s(uint32_t out[8]) {
        uint8_t buf[37];
        memcpy(buf, digest.b, 32);
        for (int i = 0; i < 4; i++) buf[32 + i] = (uint8_t)(n_sent >> (8 * i));
        buf[36] = 0;
        n_sent++;
        Hash32 h = blake2s_bytes(buf, 37);
        for (int i = 0; i < 8; i++)
            out[i] = (uint32_t)h.b[4 * i] | ((uint32_t)h.b[4 * i + 1] << 8) | ((uint32_t)h.b[4 * i + 2] << 16) | ((uint32_t)h.b[4 * i + 3] << 24);
    }
    void draw_base_felts(uint32_t out[8]) {
        for (;;) {
            draw_u32s(out);
            bool ok = true;
            for (int i = 0; i < 8; i++) ok = ok && out[i] < 2 * P;
            if (ok) break;
        }
        for (int i = 0; i < 8; i++) out[i] = out[i] >= P ? out[i] - P : out[i];
    }
    QM31 draw_secure_felt() {
        uint32_t f[8];
        draw_base_felts(f);
        return {{f[0], f[1], f[2], f[3]}};
    }
    // H(POW_PREFIX || 0^12 || digest || n_bits)
    Hash32 pow_prefixed_digest(uint32_t n_bits) const {
        std::vector<uint8_t> buf;
        put_u32(buf, 0x12345678u);
        for (int i = 0; i < 12; i++) buf.push_back(0);
        put_bytes(buf, digest.b, 32);
        put_u32(buf, n_bits);
        return blake2s_bytes(buf.data(), buf.size());
    }
    bool verify_pow_nonce(uint32_t n_bits, uint64_t nonce) const {
        Hash32 pd = pow_prefixed_digest(n_bits);
        uint8_t buf[40];
        memcpy(buf, pd.b, 32);
        for (int i = 0; i < 8; i++) buf[32 + i] = (uint8_t)(nonce >> (8 * i));
        Hash32 r = blake2s_bytes(buf, 40);
        uint32_t tz = 0;
        for (int i = 0; i < 16; i++) {
            if (r.b[i] == 0) { tz += 8; continue; }
            tz += __builtin_ctz(r.b[i]);
            break;
        }
        return tz >= n_bits;
    }
};

struct CirclePointQ {
    QM31 x, y;
};
inline CirclePointQ get_random_point(Channel& ch) {
    QM31 t = ch.draw_secure_felt();
    QM31 t2 = qmul(t, t);
    QM31 iv = qinv(qadd(t2, qone()));
    return {qmul(qsub(qone(), t2), iv), qmul(qadd(t, t), iv)};
}

inline std::vector<uint32_t> queries_generate(Channel& ch, int log_domain, int n_queries) {
    std::set<uint32_t> qs;
    int cnt = 0;
    uint32_t mask = (1u << log_domain) - 1;
    for (;;) {
        uint32_t w[8];
        ch.draw_u32s(w);
        for (int i = 0; i < 8; i++) {
            qs.insert(w[i] & mask);
            if (++cnt == n_queries) return std::vector<uint32_t>(qs.begin(), qs.end());
        }
    }
}

inline std::vector<uint32_t> fold_positions(const std::vector<uint32_t>& p, int n) {
    std::vector<uint32_t> r;
    for (uint32_t v : p)
        if (r.empty() || r.back() != (v >> n)) r.push_back(v >> n);
    return r;
}

inline std::string base64_encode(const uint8_t* d, size_t n) {
    static const char* T = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    std::string out;
    out.reserve((n + 2) / 3 * 4);
    size_t i = 0;
    for (; i + 2 < n; i += 3) {
        uint32_t v = (d[i] << 16) | (d[i + 1] << 8) | d[i + 2];
        out.push_back(T[v >> 18]); out.push_back(T[(v >> 12) & 63]); out.push_back(T[(v >> 6) & 63]); out.push_back(T[v & 63]);
    }
    if (i + 1 == n) {
        uint32_t v = d[i] << 16;
        out.push_back(T[v >> 18]); out.push_back(T[(v >> 12) & 63]); out.push_back('='); out.push_back('=');
    } else if (i + 2 == n) {
        uint32_t v = (d[i] << 16) | (d[i + 1] << 8);
        out.push_back(T[v >> 18]); out.push_back(T[(v >> 12) & 63]); out.push_back(T[(v >> 6) & 63]); out.push_back('=');
    }
    return out;
}

// base64 (RFC 4648 alphabet, canonical padding required) with the error renderings of the reference's decoder
// (base64 0.22 DecodeError Display, wasm_api.rs:628).  Returns "" on success.
inline std::string base64_decode(const char* s, size_t n, std::vector<uint8_t>& out) {
    static int8_t T[256];
    static bool init = false;
    if (!init) {
        memset(T, -1, sizeof T);
        const char* A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; i++) T[(uint8_t)A[i]] = (int8_t)i;
        init = true;
    }
    out.clear();
    out.reserve(n / 4 * 3 + 3);
    size_t body = n;
    while (body > 0 && s[body - 1] == '=' && n - body < 2) body--;
    const size_t pad = n - body;
    for (size_t i = 0; i < body; i++)
        if (T[(uint8_t)s[i]] < 0) return "Invalid symbol " + std::to_string((unsigned)(uint8_t)s[i]) + ", offset " + std::to_string(i) + ".";
    if (body % 4 == 1) return "Invalid input length: " + std::to_string(n);
    if ((body + pad) % 4 != 0) return "Invalid padding";
    size_t i = 0;
    for (; i + 4 <= body; i += 4) {
        uint32_t v = (T[(uint8_t)s[i]] << 18) | (T[(uint8_t)s[i + 1]] << 12) | (T[(uint8_t)s[i + 2]] << 6) | T[(uint8_t)s[i + 3]];
        out.push_back((uint8_t)(v >> 16)); out.push_back((uint8_t)(v >> 8)); out.push_back((uint8_t)v);
    }
    const size_t rem = body - i;
    if (rem == 2) {
        uint32_t v = (T[(uint8_t)s[i]] << 18) | (T[(uint8_t)s[i + 1]] << 12);
        if (v & 0xffff) return "Invalid last symbol " + std::to_string((unsigned)(uint8_t)s[i + 1]) + ", offset " + std::to_string(i + 1) + ".";
        out.push_back((uint8_t)(v >> 16));
    } else if (rem == 3) {
        uint32_t v = (T[(uint8_t)s[i]] << 18) | (T[(uint8_t)s[i + 1]] << 12) | (T[(uint8_t)s[i + 2]] << 6);
        if (v & 0xff) return "Invalid last symbol " + std::to_string((unsigned)(uint8_t)s[i + 2]) + ", offset " + std::to_string(i + 2) + ".";
        out.push_back((uint8_t)(v >> 16)); out.push_back((uint8_t)(v >> 8));
    }
    return "";
}

}  // namespace host
