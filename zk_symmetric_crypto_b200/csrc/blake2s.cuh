// Blake2s-256 (RFC 7693, unkeyed, no personalisation) compression for device and host.
// The reference's Merkle hasher is upstream stwo `Blake2sMerkleHasher` (core/vcs_lifted/blake2_merkle.rs, imported at
// /root/reference/stwo/src/wasm_api.rs:24): leaf = Blake2s(LE u32 column values of the row), node = Blake2s(l || r);
// byte layouts were confirmed by tracing the reference binary (oracle/trace_blake.py).
#pragma once
#include <string.h>
#include <stdint.h>
#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace blake2s {

#define B2S_IV0 0x6A09E667u
#define B2S_IV1 0xBB67AE85u
#define B2S_IV2 0x3C6EF372u
#define B2S_IV3 0xA54FF53Au
#define B2S_IV4 0x510E527Fu
#define B2S_IV5 0x9B05688Cu
#define B2S_IV6 0x1F83D9ABu
#define B2S_IV7 0x5BE0CD19u

B2_HD uint32_t rotr(uint32_t x, int n) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

B2_HD void init(uint32_t h[8]) {
    h[0] = B2S_IV0 ^ 0x01010020u;
    h[1] = B2S_IV1; h[2] = B2S_IV2; h[3] = B2S_IV3; h[4] = B2S_IV4; h[5] = B2S_IV5; h[6] = B2S_IV6; h[7] = B2S_IV7;
}

#define B2S_G(a, b, c, d, x, y) \
    a = a + b + x; d = rotr(d ^ a, 16); c = c + d; b = rotr(b ^ c, 12); \
    a = a + b + y; d = rotr(d ^ a, 8);  c = c + d; b = rotr(b ^ c, 7);

#define B2S_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    B2S_G(v0, v4, v8, v12, m[s0], m[s1]) B2S_G(v1, v5, v9, v13, m[s2], m[s3])           \
    B2S_G(v2, v6, v10, v14, m[s4], m[s5]) B2S_G(v3, v7, v11, v15, m[s6], m[s7])         \
    B2S_G(v0, v5, v10, v15, m[s8], m[s9]) B2S_G(v1, v6, v11, v12, m[s10], m[s11])       \
    B2S_G(v2, v7, v8, v13, m[s12], m[s13]) B2S_G(v3, v4, v9, v14, m[s14], m[s15])

// t = total bytes absorbed including this block; last = final block flag
B2_HD void compress(uint32_t h[8], const uint32_t m[16], uint64_t t, bool last) {
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = B2S_IV0, v9 = B2S_IV1, v10 = B2S_IV2, v11 = B2S_IV3;
    uint32_t v12 = B2S_IV4 ^ (uint32_t)t, v13 = B2S_IV5 ^ (uint32_t)(t >> 32);
    uint32_t v14 = last ? ~B2S_IV6 : B2S_IV6, v15 = B2S_IV7;
    B2S_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2S_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    B2S_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    B2S_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    B2S_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    B2S_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    B2S_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    B2S_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    B2S_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    B2S_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    h[0] ^= v0 ^ v8; h[1] ^= v1 ^ v9; h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

#ifdef __CUDACC__
// Device variant that issues the additions on the FMA pipe: `one` is a run-time 1 (kernel parameter), so a*one+b compiles
// to IMAD instead of IADD3.  Blake2s is otherwise all ALU-pipe work (LOP3/SHF at one warp-instruction per 2 cycles per
// SM sub-partition); moving the 3 additions of every half-G to the FMA pipe leaves 4 ALU instructions per half-G.
#define B2S_GD(a, b, c, d, x, y) \
    a = (a * one + b) * one + x; d = rotr(d ^ a, 16); c = c * one + d; b = rotr(b ^ c, 12); \
    a = (a * one + b) * one + y; d = rotr(d ^ a, 8);  c = c * one + d; b = rotr(b ^ c, 7);

#define B2S_ROUNDD(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    B2S_GD(v0, v4, v8, v12, m[s0], m[s1]) B2S_GD(v1, v5, v9, v13, m[s2], m[s3])           \
    B2S_GD(v2, v6, v10, v14, m[s4], m[s5]) B2S_GD(v3, v7, v11, v15, m[s6], m[s7])         \
    B2S_GD(v0, v5, v10, v15, m[s8], m[s9]) B2S_GD(v1, v6, v11, v12, m[s10], m[s11])       \
    B2S_GD(v2, v7, v8, v13, m[s12], m[s13]) B2S_GD(v3, v4, v9, v14, m[s14], m[s15])

// Lone-warp variant (product-size proofs: one warp hashes all leaves, bound by that warp's issue rate on the 16-lane ALU pipe):
// the 3-input additions stay one IADD3 each (two chained IMADs would lengthen the dependent chain), only c += d moves to the
// FMA pipe: 10 instead of 12 ALU-pipe instructions per G at an unchanged chain length.
#define B2S_GM(a, b, c, d, x, y) \
    a = a + b + x; d = rotr(d ^ a, 16); c = c * one + d; b = rotr(b ^ c, 12); \
    a = a + b + y; d = rotr(d ^ a, 8);  c = c * one + d; b = rotr(b ^ c, 7);

#define B2S_ROUNDM(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    B2S_GM(v0, v4, v8, v12, m[s0], m[s1]) B2S_GM(v1, v5, v9, v13, m[s2], m[s3])           \
    B2S_GM(v2, v6, v10, v14, m[s4], m[s5]) B2S_GM(v3, v7, v11, v15, m[s6], m[s7])         \
    B2S_GM(v0, v5, v10, v15, m[s8], m[s9]) B2S_GM(v1, v6, v11, v12, m[s10], m[s11])       \
    B2S_GM(v2, v7, v8, v13, m[s12], m[s13]) B2S_GM(v3, v4, v9, v14, m[s14], m[s15])

__device__ __forceinline__ void compress_mix(uint32_t h[8], const uint32_t m[16], uint64_t t, bool last, uint32_t one) {
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = B2S_IV0, v9 = B2S_IV1, v10 = B2S_IV2, v11 = B2S_IV3;
    uint32_t v12 = B2S_IV4 ^ (uint32_t)t, v13 = B2S_IV5 ^ (uint32_t)(t >> 32);
    uint32_t v14 = last ? ~B2S_IV6 : B2S_IV6, v15 = B2S_IV7;
    B2S_ROUNDM(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2S_ROUNDM(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    B2S_ROUNDM(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    B2S_ROUNDM(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    B2S_ROUNDM(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    B2S_ROUNDM(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    B2S_ROUNDM(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    B2S_ROUNDM(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    B2S_ROUNDM(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    B2S_ROUNDM(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    h[0] ^= v0 ^ v8; h[1] ^= v1 ^ v9; h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

__device__ __forceinline__ void compress_fma(uint32_t h[8], const uint32_t m[16], uint64_t t, bool last, uint32_t one) {
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = B2S_IV0, v9 = B2S_IV1, v10 = B2S_IV2, v11 = B2S_IV3;
    uint32_t v12 = B2S_IV4 ^ (uint32_t)t, v13 = B2S_IV5 ^ (uint32_t)(t >> 32);
    uint32_t v14 = last ? ~B2S_IV6 : B2S_IV6, v15 = B2S_IV7;
    B2S_ROUNDD(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B2S_ROUNDD(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    B2S_ROUNDD(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    B2S_ROUNDD(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    B2S_ROUNDD(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    B2S_ROUNDD(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    B2S_ROUNDD(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    B2S_ROUNDD(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    B2S_ROUNDD(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    B2S_ROUNDD(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
    h[0] ^= v0 ^ v8; h[1] ^= v1 ^ v9; h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}
#endif

// SIMD compression of full non-final blocks (host_blake2s_simd.cpp; returns 0 when the CPU has no usable vector path)
extern "C" int s2c_host_blake2s_blocks(uint32_t h[8], const uint8_t* data, size_t nblocks, uint64_t t);

// host: incremental form for data that arrives in chunks.  Every update() but the last must carry a multiple of 64 bytes.
struct Incremental {
    uint32_t h[8];
    uint64_t total = 0;
    Incremental() { init(h); }
    // absorbs full blocks; when `more` is false the final (possibly partial, possibly empty only for an empty message) block is
    // compressed with the last-block flag and the digest is written to out
    void update(const uint8_t* data, size_t len, bool more, uint8_t* out = nullptr) {
        uint32_t m[16];
        size_t off = 0;
        {
            const size_t nb = more ? len / 64 : (len ? (len - 1) / 64 : 0);  // full blocks that are not the final one
            if (nb >= 4 && s2c_host_blake2s_blocks(h, data, nb, total)) {
                off = nb * 64;
                total += off;
            }
        }
        while (len - off > (more ? 63 : 64)) {
            memcpy(m, data + off, 64);
            off += 64;
            total += 64;
            compress(h, m, total, false);
        }
        if (more) return;
        uint8_t buf[64] = {0};
        memcpy(buf, data + off, len - off);
        memcpy(m, buf, 64);
        total += len - off;
        compress(h, m, total, true);
        memcpy(out, h, 32);
    }
};

// host convenience: one-shot hash of a byte buffer (little-endian host assumed)
inline void hash(const uint8_t* data, size_t len, uint8_t out[32]) {
    uint32_t h[8];
    init(h);
    uint32_t m[16];
    size_t off = 0;
    {
        const size_t nb = len ? (len - 1) / 64 : 0;  // full blocks that are not the final one
        if (nb >= 4 && s2c_host_blake2s_blocks(h, data, nb, 0)) off = nb * 64;
    }
    while (len - off > 64) {
        memcpy(m, data + off, 64);  // little-endian host (see above)
        off += 64;
        compress(h, m, off, false);
    }
    uint8_t buf[64] = {0};
    for (size_t i = 0; i < len - off; i++) buf[i] = data[off + i];
    for (int i = 0; i < 16; i++)
        m[i] = (uint32_t)buf[4 * i] | ((uint32_t)buf[4 * i + 1] << 8) | ((uint32_t)buf[4 * i + 2] << 16) | ((uint32_t)buf[4 * i + 3] << 24);
    compress(h, m, len, true);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)h[i]; out[4 * i + 1] = (uint8_t)(h[i] >> 8);
        out[4 * i + 2] = (uint8_t)(h[i] >> 16); out[4 * i + 3] = (uint8_t)(h[i] >> 24);
    }
}

}  // namespace blake2s
