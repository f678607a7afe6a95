// AES-128/256-CTR proof driver: replays the reference's prove flow on the GPU backend.
//
// Mirrors /root/reference/stwo/src/wasm_api.rs:652-896 (generate_aes{128,256}_ctr_proof: validation, log_size, lane packing,
// padding-lane keystreams) and /root/reference/stwo/src/aes/lookup/air_ctr.rs:297-422 (prove_aes_ctr_with_inputs_internal:
// preprocessed S-box tree, statement 0, main trace + multiplicities, lookup elements, interaction traces, statement 1,
// interaction tree, LogUp balance check, prove) followed by upstream stwo::prover::prove with two components of different
// sizes.  Everything the reference leaves to the un-vendored stwo rev (size-sorted lifted Merkle leaves,
// lift_and_accumulate of the table component, doubled sample points and periodicity samples of lifted columns, per-sample
// powers of the quotient random coefficient) follows the CPU restatement of the test tree (aes_api.py, prover.py in the oracle directory), which reproduce the reference
// binary byte for byte.  Output = bincode(AESCtrProof{stmt0, stmt1, stark_proof}) (air_ctr.rs:44-184).
//
// This driver stores the LDE of the trace (24,480 / 34,784 columns): it serves log_size <= 19 (AES-128) / 18 (AES-256) on one B200; the
// streaming (tile-by-tile) machinery of the ChaCha prover is not applied to the AES AIR yet.
#include <array>
#include <map>
#include "prover.hpp"

using namespace m31;
using host::Channel;
using host::Hash32;

namespace {

// ---- native cipher pieces the host needs (aes/mod.rs:10-30 S-box, :213-270 key expansion)
struct AesTables {
    uint8_t sbox[256];
    AesTables() {
        uint8_t p = 1, q = 1;
        do {
            p = p ^ (uint8_t)(p << 1) ^ ((p & 0x80) ? 0x1B : 0);
            q ^= q << 1; q ^= q << 2; q ^= q << 4;
            if (q & 0x80) q ^= 0x09;
            uint8_t x = q ^ (uint8_t)((q << 1) | (q >> 7)) ^ (uint8_t)((q << 2) | (q >> 6)) ^ (uint8_t)((q << 3) | (q >> 5)) ^
                        (uint8_t)((q << 4) | (q >> 4));
            sbox[p] = x ^ 0x63;
        } while (p != 1);
        sbox[0] = 0x63;
    }
};
const AesTables& tables() {
    static AesTables t;
    return t;
}
uint8_t xt(uint8_t a) { return (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1B : 0)); }

std::vector<uint8_t> expand_key(const uint8_t* key, int key_len) {
    const int nk = key_len / 4, nr = nk + 6;
    std::vector<std::array<uint8_t, 4>> w(4 * (nr + 1));
    for (int i = 0; i < nk; i++) w[i] = {key[4 * i], key[4 * i + 1], key[4 * i + 2], key[4 * i + 3]};
    uint8_t rc = 1;
    const uint8_t* S = tables().sbox;
    for (int i = nk; i < 4 * (nr + 1); i++) {
        std::array<uint8_t, 4> t = w[i - 1];
        if (i % nk == 0) {
            t = {S[t[1]], S[t[2]], S[t[3]], S[t[0]]};
            t[0] ^= rc;
            rc = xt(rc);
        } else if (nk > 6 && i % nk == 4) {
            t = {S[t[0]], S[t[1]], S[t[2]], S[t[3]]};
        }
        for (int b = 0; b < 4; b++) w[i][b] = w[i - nk][b] ^ t[b];
    }
    std::vector<uint8_t> out(16 * (nr + 1));
    for (int i = 0; i < 4 * (nr + 1); i++)
        for (int b = 0; b < 4; b++) out[4 * i + b] = w[i][b];
    return out;
}

// Column bookkeeping of the AIR (same traversal as the witness / constraint kernels): S-box (input, output) columns.
}  // namespace
AesLayout aes_make_layout(int nr, bool block) {
    static const int SR[16] = {0, 5, 10, 15, 4, 9, 14, 3, 8, 13, 2, 7, 12, 1, 6, 11};
    AesLayout L;
    int col = 16 + 16 * (nr + 1) + (block ? 0 : 32), k = 0;
    auto xor_byte = [&]() { col += 25; k += 35; return col - 1; };
    auto xtime = [&]() { col += 17; k += 26; return col - 1; };
    int s[16], t[16];
    for (int i = 0; i < 16; i++) s[i] = xor_byte();
    for (int rnd = 1; rnd <= nr; rnd++) {
        for (int i = 0; i < 16; i++) {
            L.lk_in.push_back(s[i]);
            s[i] = col++;
            L.lk_out.push_back(s[i]);
        }
        for (int i = 0; i < 16; i++) t[i] = s[SR[i]];
        if (rnd < nr) {
            for (int c = 0; c < 4; c++) {
                // per output byte: the operations of ctr.rs:244-281 in order; only column/constraint counts matter here
                xtime(); xtime(); xor_byte(); xor_byte(); xor_byte(); s[4 * c] = xor_byte();
                xtime(); xtime(); xor_byte(); xor_byte(); xor_byte(); s[4 * c + 1] = xor_byte();
                xtime(); xtime(); xor_byte(); xor_byte(); xor_byte(); s[4 * c + 2] = xor_byte();
                xtime(); xor_byte(); xtime(); xor_byte(); xor_byte(); s[4 * c + 3] = xor_byte();
            }
        } else {
            for (int i = 0; i < 16; i++) s[i] = t[i];
        }
        for (int i = 0; i < 16; i++) s[i] = xor_byte();
    }
    if (!block) {
        for (int i = 0; i < 16; i++) s[i] = xor_byte();
        k += 16;
    }
    L.n_cols = col;
    L.n_constraints = k + (int)L.lk_in.size() / 2;
    return L;
}
// shared with the stage-level C ABI (cb_api_ext.cu)
std::vector<uint8_t> aes_expand_key(const uint8_t* key, int key_len) { return expand_key(key, key_len); }
const uint8_t* aes_sbox() { return tables().sbox; }

namespace {
using Layout = AesLayout;
Layout make_layout(int nr, bool block = false) { return aes_make_layout(nr, block); }

struct PtQ {
    QM31 x, y;
};
PtQ pt_add_m(const PtQ& p, host::Pt s) { return {qsub(qmul_m(p.x, s.x), qmul_m(p.y, s.y)), qadd(qmul_m(p.x, s.y), qmul_m(p.y, s.x))}; }
PtQ pt_double(const PtQ& p) { return {qsub(qmul_m(qmul(p.x, p.x), 2), qone()), qmul_m(qmul(p.x, p.y), 2)}; }
std::array<uint32_t, 8> pt_key(const PtQ& p) {
    return {p.x.v[0], p.x.v[1], p.x.v[2], p.x.v[3], p.y.v[0], p.y.v[1], p.y.v[2], p.y.v[3]};
}

// storage index (bit-reversed circle-domain order) of the i-th point of CanonicCoset(log).coset in natural order
std::vector<uint32_t> coset_order_to_storage(int log) {
    const uint32_t n = 1u << log;
    std::vector<uint32_t> o(n);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t cd = (i % 2 == 0) ? i / 2 : n - 1 - i / 2;
        o[i] = host::bit_reverse(cd, log);
    }
    return o;
}

// LogupTraceGenerator::finalize_last on one QM31 column given as 4 coordinate vectors: returns claimed sum, rewrites the
// column as the inclusive prefix sum (coset order) of (value - claimed_sum / N)
QM31 finalize_last(std::vector<uint32_t>& col4, int log) {
    const size_t n = (size_t)1 << log;
    QM31 claimed = qzero();
    for (int c = 0; c < 4; c++) {
        uint64_t s = 0;
        for (size_t i = 0; i < n; i++) s += col4[c * n + i];
        claimed.v[c] = (uint32_t)(s % P);
    }
    const uint32_t ninv = inv((uint32_t)((uint64_t)n % P));
    const std::vector<uint32_t> order = coset_order_to_storage(log);
    for (int c = 0; c < 4; c++) {
        const uint32_t shift = mul(claimed.v[c], ninv);
        uint32_t run = 0;
        for (size_t i = 0; i < n; i++) {
            uint32_t& v = col4[c * n + order[i]];
            run = add(run, sub(v, shift));
            v = run;
        }
    }
    return claimed;
}

// one committed tree: a list of column groups (equal size, equal stride) in trace order
struct Group {
    uint32_t *coeffs = nullptr, *lde = nullptr;  // [ncols][2^log], [ncols][2^(log+1)]
    int ncols = 0, log = 0;                       // trace log size of the group's columns
};
struct Tree {
    std::vector<Group> groups;
    DevMerkle merkle;
};

}  // namespace

std::string prove_aes_ctr(cb_ctx* ctx, int key_len, const uint8_t* key, const uint8_t nonce[12], uint32_t counter,
                          const uint8_t* plaintext, const uint8_t* ciphertext, size_t len, std::vector<uint8_t>& proof, bool block_air) {
    const PcsConfig cfg;
    const uint32_t num_blocks = (uint32_t)(len / 16);
    int log_size = 8;
    while (((size_t)1 << log_size) < num_blocks) log_size++;
    if (log_size > 24) return "log_size (" + std::to_string(log_size) + ") must be <= MAX_LOG_SIZE (24)";
    const int n = log_size, m = n + 1, nr = key_len == 16 ? 10 : 14;
    const size_t N = (size_t)1 << n, M = (size_t)1 << m;
    static const Layout L128 = make_layout(10), L256 = make_layout(14), L128b = make_layout(10, true);
    if (block_air && (key_len != 16 || len != ((size_t)16 << log_size))) return "block AIR: AES-128, one input block per row";
    const Layout& lay = block_air ? L128b : (nr == 10 ? L128 : L256);
    const int C = lay.n_cols, K = lay.n_constraints, NL = (int)lay.lk_in.size(), NI = 4 * (NL / 2);
    cudaStream_t st = ctx->stream;
    {
        size_t free_b = 0, total_b = 0;
        CB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = ((size_t)C * 3 + NI * 3 + 2 * NL + 64) * N * 4 + ((size_t)1 << 30);
        if (need > free_b) {  // give back what earlier proofs left cached in the stream-ordered pool (and the ChaCha tile arena)
            ctx->sync();
            ctx->close_peers();
            ctx->release_arena();
            cudaMemPool_t pool;
            CB_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
            CB_CUDA(cudaMemPoolTrimTo(pool, 0));
            CB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        }
        if (need > free_b + ctx->arena_bytes)
            throw CbError("AES-CTR proof at log_size " + std::to_string(n) + " needs " + std::to_string(need >> 30) +
                          " GiB for the stored LDE; the streaming path is not built for the AES AIR yet");
        if (need > free_b) ctx->release_arena();
    }
    ctx->ensure_twiddles(m);
    ctx->pending_events.clear();
    ctx->host_marks.clear();
    ctx->host_mark("setup");
    CB_CUDA(aes_upload_sbox(tables().sbox));
    const std::vector<uint8_t> rk = expand_key(key, key_len);

    // ---- witness
    ctx->stage_begin("witness");
    DBuf<uint8_t> d_pt(ctx, len), d_ct(ctx, block_air ? 16 : len);
    DBuf<uint32_t> T(ctx, (size_t)C * N);
    DBuf<unsigned int> d_mults(ctx, 256);
    DBuf<int> d_invalid(ctx, 1);
    CB_CUDA(cudaMemcpyAsync(d_pt.p, plaintext, len, cudaMemcpyHostToDevice, st));
    if (!block_air) CB_CUDA(cudaMemcpyAsync(d_ct.p, ciphertext, len, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemsetAsync(d_mults.p, 0, 256 * 4, st));
    CB_CUDA(cudaMemsetAsync(d_invalid.p, 0, 4, st));
    const uint32_t rows_needed = (num_blocks + 15) / 16;
    CB_CUDA(launch_aes_witness(st, rk.data(), nr, nonce, counter, num_blocks, rows_needed * 16, d_pt.p, d_ct.p, n, T.p, N, d_mults.p,
                               d_invalid.p, block_air ? 1 : 0));
    ctx->launches++;
    ctx->stage_end();
    int invalid = 0;
    std::vector<uint32_t> mults(256);
    CB_CUDA(cudaMemcpyAsync(&invalid, d_invalid.p, 4, cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaMemcpyAsync(mults.data(), d_mults.p, 256 * 4, cudaMemcpyDeviceToHost, st));
    ctx->sync();
    if (invalid && !block_air) return "Ciphertext does not match encryption - invalid witness";
    d_pt.release();
    d_ct.release();

    Channel ch;
    std::vector<Hash32> roots;
    std::vector<Tree> trees(4);

    // commits the groups of a tree: Merkle leaves hash the columns sorted by size, smallest first (stable)
    auto commit_tree = [&](Tree& t, const char* stage) {
        ctx->stage_begin(stage);
        std::vector<int> order(t.groups.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return t.groups[a].log < t.groups[b].log; });
        int lifting = 0;
        for (auto& g : t.groups) lifting = std::max(lifting, g.log + 1);
        LeafGroups lg{};
        for (int gi : order) {
            const Group& g = t.groups[gi];
            lg.g[lg.n++] = {g.lde, (size_t)2 << g.log, g.ncols, g.log + 1, nullptr, nullptr, nullptr};
        }
        t.merkle = build_merkle(ctx, lg, lifting);
        ctx->stage_end();
        roots.push_back(t.merkle.root);
        ch.mix_root(t.merkle.root);
    };
    // interpolate + extend a group of M31 columns (values in `vals`, overwritten by the coefficients)
    auto transform = [&](uint32_t* vals, int ncols, int lg, uint32_t* lde) {
        ColSrc src{SRC_M31, vals, (size_t)1 << lg, 0};
        CB_CUDA(launch_fft(st, src, ncols, lg, 1, 1 | 2 | 4, vals, (size_t)1 << lg, lde, (size_t)2 << lg, ctx->tw, vals,
                           (size_t)1 << lg));
        ctx->launches += (lg + 1 <= 13) ? 1 : 3;
    };

    // large groups: values stay in place (their out-of-domain samples are taken from the values, not from coefficients) and the
    // LDE comes from the register-radix kernels of kernels_fft2.cu, 4 columns per job
    const StageHook hk = ctx->hook();
    auto transform_fast = [&](const uint32_t* vals, int ncols, uint32_t* lde) {
        if (m <= 13) {  // whole columns fit shared memory: one launch of the generic kernel for all columns
            ColSrc src{SRC_M31, vals, N, 0};
            CB_CUDA(launch_fft(st, src, ncols, n, 1, 1 | 4, nullptr, 0, lde, M, ctx->tw, nullptr, 0));
            ctx->launches++;
            return;
        }
        const int njobs = ncols / 4;
        DBuf<uint32_t> scratch(ctx, fft_packed_scratch_words(SRC_M31, njobs < MAX_FFT_JOBS ? njobs : MAX_FFT_JOBS, n));
        std::vector<const uint32_t*> src(njobs);
        std::vector<uint32_t*> out(njobs);
        for (int j = 0; j < njobs; j++) { src[j] = vals + (size_t)4 * j * N; out[j] = lde + (size_t)4 * j * M; }
        int nl = 0;
        CB_CUDA(launch_fft_packed(st, SRC_M31, src.data(), out.data(), njobs, n, ctx->tw, scratch.p, ctx->profile ? &hk : nullptr, &nl));
        ctx->launches += nl;
    };

    // ---- tree 0: preprocessed S-box table (aes/sbox_table.rs:35-48)
    DBuf<uint32_t> pre(ctx, 2 * 256), pre_lde(ctx, 2 * 512);
    {
        std::vector<uint32_t> h(512);
        for (int i = 0; i < 256; i++) { h[i] = i; h[256 + i] = tables().sbox[i]; }
        CB_CUDA(cudaMemcpyAsync(pre.p, h.data(), 512 * 4, cudaMemcpyHostToDevice, st));
        ctx->sync();
        transform(pre.p, 2, 8, pre_lde.p);
        trees[0].groups = {{pre.p, pre_lde.p, 2, 8}};
        commit_tree(trees[0], "preprocessed_commit");
    }

    // ---- statement 0 (air_ctr.rs:156-160, 66-99); the public-input hashes cover the caller's plaintext / ciphertext
    std::vector<uint8_t> stmt;
    host::put_u32(stmt, (uint32_t)log_size);
    ch.mix_u64((uint64_t)log_size);  // AESLookupStatement0::mix_into (aes/lookup/air.rs:65-67) stops here
    if (!block_air) {
        host::put_u32(stmt, key_len == 16 ? 0u : 1u);
        host::put_bytes(stmt, nonce, 12);
        host::put_u32(stmt, counter);
        {
            Hash32 pth = host::blake2s_bytes(plaintext, len), cth = host::blake2s_bytes(ciphertext, len);
            host::put_bytes(stmt, pth.b, 32);
            host::put_bytes(stmt, cth.b, 32);
        }
        ch.mix_u64(key_len == 16 ? 0 : 1);
        for (int i = 0; i < 3; i++) ch.mix_u64(host::load_le32(&stmt[8 + 4 * i]));
        ch.mix_u64(counter);
        for (int i = 0; i < 16; i++) ch.mix_u64(host::load_le32(&stmt[24 + 4 * i]));
    }

    // ---- tree 1: main trace + S-box multiplicities
    DBuf<uint32_t> lde1(ctx, (size_t)C * M), mult(ctx, 256), mult_lde(ctx, 512);
    CB_CUDA(cudaMemcpyAsync(mult.p, mults.data(), 256 * 4, cudaMemcpyHostToDevice, st));
    ctx->sync();
    ctx->stage_begin("trace_lde");
    transform_fast(T.p, C, lde1.p);
    transform(mult.p, 1, 8, mult_lde.p);
    ctx->stage_end();
    trees[1].groups = {{T.p, lde1.p, C, n}, {mult.p, mult_lde.p, 1, 8}};
    commit_tree(trees[1], "trace_merkle");

    // ---- lookup elements, interaction traces (gen_ctr.rs:640-683, gen.rs:438-478)
    uint32_t zf[8];
    ch.draw_base_felts(zf);
    const QM31 z{{zf[0], zf[1], zf[2], zf[3]}}, alpha{{zf[4], zf[5], zf[6], zf[7]}};
    ctx->stage_begin("interaction");
    DBuf<uint32_t> I(ctx, (size_t)NI * N), inter_lde(ctx, (size_t)NI * M), tinter(ctx, 4 * 256), tinter_lde(ctx, 4 * 512);
    QM31 csum, tsum;
    {
        DBuf<int> d_in(ctx, NL), d_out(ctx, NL);
        CB_CUDA(cudaMemcpyAsync(d_in.p, lay.lk_in.data(), NL * 4, cudaMemcpyHostToDevice, st));
        CB_CUDA(cudaMemcpyAsync(d_out.p, lay.lk_out.data(), NL * 4, cudaMemcpyHostToDevice, st));
        CB_CUDA(launch_aes_interaction(st, T.p, N, n, d_in.p, d_out.p, NL, z, alpha, I.p, N));
        ctx->launches++;
        {   // LogupTraceGenerator::finalize_last on the device: scan in coset order, claimed sum = its last element
            DBuf<uint32_t> fscr(ctx, logup_finalize_scratch_words(n));
            uint32_t* d_claimed = nullptr;
            CB_CUDA(launch_logup_finalize_last(st, I.p + (size_t)(NI - 4) * N, N, n, fscr.p, &d_claimed));
            ctx->launches += 3;
            CB_CUDA(cudaMemcpyAsync(csum.v, d_claimed, 16, cudaMemcpyDeviceToHost, st));
            ctx->sync();
        }
        // table side: -mult / combine(i, SBOX[i]) over the 256 rows, one column
        std::vector<uint32_t> tcol(4 * 256);
        for (int i = 0; i < 256; i++) {
            QM31 p = qmul_m(alpha, tables().sbox[i]);
            p.v[0] = add(p.v[0], (uint32_t)i);
            p = qsub(p, z);
            QM31 f = qmul_m(qinv(p), neg(mults[i] % P));
            for (int c = 0; c < 4; c++) tcol[c * 256 + i] = f.v[c];
        }
        tsum = finalize_last(tcol, 8);
        CB_CUDA(cudaMemcpyAsync(tinter.p, tcol.data(), 4 * 256 * 4, cudaMemcpyHostToDevice, st));
        ctx->sync();
    }
    ctx->stage_end();
    {
        QM31 sums[2] = {csum, tsum};
        ch.mix_felts(sums, 2);
    }
    ctx->stage_begin("interaction_lde");
    transform_fast(I.p, NI, inter_lde.p);
    transform(tinter.p, 4, 8, tinter_lde.p);
    ctx->stage_end();
    trees[2].groups = {{I.p, inter_lde.p, NI, n}, {tinter.p, tinter_lde.p, 4, 8}};
    commit_tree(trees[2], "interaction_merkle");
    if (!qeq(qadd(csum, tsum), qzero())) return "LogUp sums don't balance";

    // ---- composition polynomial: CTR component on the 2^(n+1) domain, table component on 2^9, lifted and added
    const QM31 random_coeff = ch.draw_secure_felt();
    ctx->stage_begin("constraints");
    DBuf<uint32_t> apr(ctx, (size_t)(K + 1) * 4), apr_lo(ctx, (size_t)(K + 1) * 4), apr_hi(ctx, (size_t)(K + 1) * 4), acc(ctx, 4 * M),
        acc_t(ctx, 4 * 512), d_den(ctx, 2), d_den8(ctx, 2);
    DBuf<int> d_lk_in(ctx, NL), d_lk_out(ctx, NL);
    CB_CUDA(launch_secure_powers_rev(st, random_coeff, K + 1, apr.p));
    CB_CUDA(launch_split16(st, apr.p, K + 1, apr_lo.p, apr_hi.p));
    CB_CUDA(cudaMemcpyAsync(d_lk_in.p, lay.lk_in.data(), NL * 4, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(d_lk_out.p, lay.lk_out.data(), NL * 4, cudaMemcpyHostToDevice, st));
    auto den_table = [&](int tlog, uint32_t out[2]) {
        for (uint32_t i = 0; i < 2; i++) {
            uint32_t row = i << tlog;
            host::Pt p = host::index_to_point(host::canonic_index_at(tlog + 1, host::bit_reverse(row, tlog + 1)));
            out[i] = inv(host::coset_vanishing_m31(tlog, p));
        }
    };
    uint32_t den_n[2], den_8[2];
    den_table(n, den_n);
    den_table(8, den_8);
    CB_CUDA(cudaMemcpyAsync(d_den.p, den_n, 8, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(d_den8.p, den_8, 8, cudaMemcpyHostToDevice, st));
    ctx->sync();
    auto div_n = [&](const QM31& s, int lg) { return qmul_m(s, inv((uint32_t)(((uint64_t)1 << lg) % P))); };
    {
        AesConsArgs a{};
        a.lde = lde1.p; a.stride = M; a.inter = inter_lde.p; a.i_stride = M;
        a.apr_lo = apr_lo.p; a.apr_hi = apr_hi.p; a.apr = apr.p; a.den_inv = d_den.p;
        a.block = block_air ? 1 : 0;
        a.lk_in = d_lk_in.p; a.lk_out = d_lk_out.p;
        a.z = z; a.alpha = alpha; a.shift = div_n(csum, n);
        a.eval_log = m; a.trace_log = n; a.n_rounds = nr; a.n_lookups = NL;
        a.out = acc.p; a.out_stride = M;
        CB_CUDA(launch_aes_constraints(st, a));
        AesTableArgs t{};
        t.pre_in = pre_lde.p; t.pre_out = pre_lde.p + 512; t.mult = mult_lde.p; t.inter = tinter_lde.p; t.i_stride = 512;
        t.z = z; t.alpha = alpha; t.shift = div_n(tsum, 8); t.apow = qone();  // alpha^0: the last constraint overall
        t.den_inv = d_den8.p; t.eval_log = 9; t.trace_log = 8; t.out = acc_t.p;
        CB_CUDA(launch_aes_table_constraint(st, t));
        CB_CUDA(launch_lift_accumulate(st, acc.p, M, m, acc_t.p, 9));
        ctx->launches += 5;
    }
    ctx->stage_end();
    ctx->stage_begin("composition_commit");
    DBuf<uint32_t> comp_coef(ctx, 4 * M), comp_lde(ctx, 8 * M);
    {
        DBuf<uint32_t> scratch4(ctx, 4 * M);
        ColSrc src{SRC_M31, acc.p, M, 0};
        CB_CUDA(launch_fft(st, src, 4, m, 0, 1 | 2, comp_coef.p, M, nullptr, 0, ctx->tw, scratch4.p, M));
        for (int half = 0; half < 2; half++) {
            ColSrc cs{SRC_M31, comp_coef.p + half * N, M, 0};
            CB_CUDA(launch_fft(st, cs, 4, n, 1, 4, nullptr, 0, comp_lde.p + (size_t)half * 4 * M, M, ctx->tw, nullptr, 0));
        }
        ctx->launches += 6;
        ctx->sync();
    }
    ctx->stage_end();
    // the composition tree's coefficient layout is [left c0..c3 | right c0..c3] with stride M between coordinates
    {
        LeafGroups lg{};
        lg.n = 1;
        lg.g[0] = {comp_lde.p, M, 8, m, nullptr, nullptr, nullptr};
        trees[3].merkle = build_merkle(ctx, lg, m);
        roots.push_back(trees[3].merkle.root);
        ch.mix_root(trees[3].merkle.root);
    }

    // ---- OODS sampling.  Points are defined on the lifting domain; a log-8 column is the lift g(pi^k(p)), k = n - 8, of its
    //      polynomial g, which is evaluated at the k-fold doubled point; the last LogUp column of each component is also
    //      sampled at the previous row z - step_n (mask [-1, 0]).
    host::CirclePointQ zq = host::get_random_point(ch);
    ctx->stage_begin("oods");
    const PtQ Z{zq.x, zq.y};
    const PtQ Zp = pt_add_m(Z, host::index_to_point((0x80000000u - (1u << (31 - n))) & 0x7fffffffu));
    PtQ Z8 = Z, Z8p = Zp;
    for (int i = 0; i < n - 8; i++) { Z8 = pt_double(Z8); Z8p = pt_double(Z8p); }
    auto eval_cols = [&](const uint32_t* coeffs, size_t stride, int ncols, int lg, const PtQ& pt, QM31* out) {
        const size_t nn = (size_t)1 << lg;
        std::vector<QM31> maps(lg);
        maps[0] = pt.y;
        QM31 x = pt.x;
        for (int j = 1; j < lg; j++) { maps[j] = x; x = qsub(qmul_m(qmul(x, x), 2), qone()); }
        DBuf<uint32_t> basis(ctx, 4 * nn), d_out(ctx, (size_t)ncols * 4);
        CB_CUDA(launch_basis(st, basis.p, nn, lg, maps.data()));
        CB_CUDA(launch_oods_dot(st, coeffs, stride, ncols, lg, basis.p, nn, d_out.p));
        ctx->launches += (lg <= 10 ? 1 : lg) + 1;
        CB_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)ncols * 16, cudaMemcpyDeviceToHost, st));
        ctx->sync();
    };
    // columns kept as trace-domain VALUES: f(p) = 2^-lg <values, w(p)>, w(p) = forward butterflies with inverse twiddles applied
    // to basis(p) (the transpose of the inverse transform; kernels_stream.cu fact 2)
    const FftTables tw_t{ctx->tw.IX, ctx->tw.IY, ctx->tw.X, ctx->tw.Y, ctx->tw.max_log};
    DBuf<uint32_t> d_v1;  // the main trace's samples stay on the device (unscaled) for the FRI line coefficients below
    auto eval_vals = [&](const uint32_t* vals, size_t stride, int ncols, int lg, const PtQ& pt, QM31* out, DBuf<uint32_t>* keep = nullptr) {
        const size_t nn = (size_t)1 << lg;
        std::vector<QM31> maps(lg);
        maps[0] = pt.y;
        QM31 x = pt.x;
        for (int j = 1; j < lg; j++) { maps[j] = x; x = qsub(qmul_m(qmul(x, x), 2), qone()); }
        DBuf<uint32_t> basis(ctx, 4 * nn), wt(ctx, 4 * nn), d_out(ctx, (size_t)ncols * 4);
        CB_CUDA(launch_basis(st, basis.p, nn, lg, maps.data()));
        ColSrc bs{SRC_M31, basis.p, nn, 0};
        CB_CUDA(launch_fft(st, bs, 4, lg, 0, 4, nullptr, 0, wt.p, nn, tw_t, nullptr, 0));
        CB_CUDA(launch_oods_dot(st, vals, stride, ncols, lg, wt.p, nn, d_out.p));
        ctx->launches += (lg <= 10 ? 1 : lg) + 3;
        CB_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)ncols * 16, cudaMemcpyDeviceToHost, st));
        ctx->sync();
        const uint32_t inv_n = 1u << (31 - lg);
        for (int j = 0; j < ncols; j++) out[j] = qmul_m(out[j], inv_n);
        if (keep) *keep = std::move(d_out);
    };
    // samples[tree][col] = list of (point, value) in mask order
    struct Sample { PtQ pt; QM31 val; };
    std::vector<std::vector<std::vector<Sample>>> samples(4);
    {
        std::vector<QM31> v0(2), v1(C), vm(1), vi(NI), vip(4), vt(4), vtp(4), vc(8);
        eval_cols(pre.p, 256, 2, 8, Z8, v0.data());
        eval_vals(T.p, N, C, n, Z, v1.data(), &d_v1);
        eval_cols(mult.p, 256, 1, 8, Z8, vm.data());
        eval_vals(I.p, N, NI, n, Z, vi.data());
        eval_vals(I.p + (size_t)(NI - 4) * N, N, 4, n, Zp, vip.data());
        eval_cols(tinter.p, 256, 4, 8, Z8, vt.data());
        eval_cols(tinter.p, 256, 4, 8, Z8p, vtp.data());
        for (int half = 0; half < 2; half++) eval_cols(comp_coef.p + half * N, M, 4, n, Z, vc.data() + 4 * half);
        for (int j = 0; j < 2; j++) samples[0].push_back({{Z, v0[j]}});
        for (int j = 0; j < C; j++) samples[1].push_back({{Z, v1[j]}});
        samples[1].push_back({{Z, vm[0]}});
        for (int j = 0; j < NI; j++) {
            if (j < NI - 4) samples[2].push_back({{Z, vi[j]}});
            else samples[2].push_back({{Zp, vip[j - (NI - 4)]}, {Z, vi[j]}});
        }
        for (int j = 0; j < 4; j++) samples[2].push_back({{Zp, vtp[j]}, {Z, vt[j]}});
        for (int j = 0; j < 8; j++) samples[3].push_back({{Z, vc[j]}});
    }
    ctx->stage_end();
    std::vector<QM31> flat;
    for (auto& t : samples)
        for (auto& c : t)
            for (auto& s : c) flat.push_back(s.val);
    ch.mix_felts(flat.data(), flat.size());

    // ---- FRI quotients: every (column, sample) gets its own power of the random coefficient, alpha^0 first, in tree / column /
    //      sample order; a column with two samples first gets a "periodicity" copy of its offset-0 sample at z + h_k (h_k = point
    //      of order 2^k, k = its lift; z itself for k = 0); samples are grouped by point and the batches summed.
    const QM31 rc = ch.draw_secure_felt();
    ctx->stage_begin("quotients");
    DBuf<uint32_t> quot(ctx, 4 * M);
    {
        // flattened column table: device pointer of the LDE column and its log size
        std::vector<const uint32_t*> col_ptr;
        std::vector<uint8_t> col_log;
        std::vector<int> col_lift;
        auto add_cols = [&](const uint32_t* base, size_t stride, int ncols, int lde_log) {
            for (int j = 0; j < ncols; j++) { col_ptr.push_back(base + (size_t)j * stride); col_log.push_back((uint8_t)lde_log); col_lift.push_back(m - lde_log); }
        };
        add_cols(pre_lde.p, 512, 2, 9);
        add_cols(lde1.p, M, C, m);
        add_cols(mult_lde.p, 512, 1, 9);
        add_cols(inter_lde.p, M, NI, m);
        add_cols(tinter_lde.p, 512, 4, 9);
        add_cols(comp_lde.p, M, 8, m);
        struct Entry { int ci; QM31 val, apow; };
        std::map<std::array<uint32_t, 8>, std::pair<PtQ, std::vector<Entry>>> batches;
        // The C main-trace columns (one sample each, at z, consecutive powers rc^2 .. rc^(C+1)) are handled by a kernel over
        // their device-resident samples (kernels_tail.cu quot_coefs_kernel: coefficients straight into the row-combination
        // table, partial line sums added below); the host walks the remaining ~350 (column, sample) pairs.
        const int big0 = 2, big1 = 2 + C, big2 = 2 + C + 1, big3 = big2 + NI;   // [big0,big1) = main trace, [big2,big3) = CTR interaction
        QM31 ap = qone();
        int ci = 0;
        for (auto& t : samples)
            for (auto& c : t) {
                if (ci >= big0 && ci < big1) {
                    if (ci == big0) ap = qmul(ap, qpow(rc, (uint64_t)C));
                    ci++;
                    continue;
                }
                std::vector<Sample> entries = c;
                if (c.size() > 1) {
                    Sample per = c.back();
                    if (col_lift[ci] > 0) per.pt = pt_add_m(per.pt, host::index_to_point(1u << (31 - col_lift[ci])));
                    entries.insert(entries.begin(), per);
                }
                for (auto& s : entries) {
                    auto& b = batches[pt_key(s.pt)];
                    b.first = s.pt;
                    b.second.push_back({ci, s.val, ap});
                    ap = qmul(ap, rc);
                }
                ci++;
            }
        // Columns of the two big groups (main trace, CTR interaction trace) enter the batch of point z only through
        // G(p) = sum_j coef_j f_j(p), which is the extension of the row-wise combination of their trace-domain values
        // (extension is linear): one pass over the values + 4 column transforms instead of reading their LDE.
        auto is_big = [&](int c) { return (c >= big0 && c < big1) || (c >= big2 && c < big3); };
        const auto zkey = pt_key(Z);
        std::vector<uint32_t> gcoef((size_t)(C + NI) * 4, 0);
        DBuf<uint32_t> g(ctx, 4 * N), g_lde(ctx, 4 * M), d_gcoef(ctx, gcoef.size());
        QM31 lin_main_a = qzero(), lin_main_b = qzero();
        {
            DBuf<uint32_t> d_pw(ctx, (size_t)(C + 2) * 4), d_lin(ctx, 8);
            CB_CUDA(launch_secure_powers_rev(st, rc, C + 2, d_pw.p));   // d_pw[i] = rc^(C+1-i): column j reads index C-1-j = rc^(j+2)
            CB_CUDA(launch_quot_coefs(st, d_v1.p, C, d_pw.p, Z.y, d_gcoef.p, d_lin.p));
            uint32_t lin[8];
            CB_CUDA(cudaMemcpyAsync(lin, d_lin.p, sizeof lin, cudaMemcpyDeviceToHost, st));
            ctx->sync();
            ctx->launches += 2;
            const uint32_t inv_n = 1u << (31 - n);  // the device samples are unscaled; a_j and b_j are linear in the sample
            for (int k = 0; k < 4; k++) { lin_main_a.v[k] = mul(lin[k], inv_n); lin_main_b.v[k] = mul(lin[4 + k], inv_n); }
        }
        std::vector<QuotBatch> qbs;
        std::vector<DBuf<uint32_t>> keep32;
        std::vector<DBuf<const uint32_t*>> keep_ptr;
        std::vector<DBuf<uint8_t>> keep8;
        for (auto& kv : batches) {
            const PtQ& pt = kv.second.first;
            const std::vector<Entry>& es = kv.second.second;
            const bool is_z = kv.first == zkey;
            std::vector<uint32_t> coefs;
            std::vector<const uint32_t*> ptrs;
            std::vector<uint8_t> logs;
            QM31 lin_a = is_z ? lin_main_a : qzero(), lin_b = is_z ? lin_main_b : qzero();
            const QM31 c = qsub(qconj(pt.y), pt.y);
            if (is_z)
                for (int k = 0; k < 4; k++) {  // the 4 coordinate columns of G with unit coefficients
                    for (int q = 0; q < 4; q++) coefs.push_back(q == k ? 1u : 0u);
                    ptrs.push_back(g_lde.p + (size_t)k * M);
                    logs.push_back((uint8_t)m);
                }
            for (size_t j = 0; j < es.size(); j++) {
                const QM31 a = qsub(qconj(es[j].val), es[j].val);
                const QM31 b = qsub(qmul(es[j].val, c), qmul(a, pt.y));
                lin_a = qadd(lin_a, qmul(es[j].apow, a));
                lin_b = qadd(lin_b, qmul(es[j].apow, b));
                const QM31 ac = qmul(es[j].apow, c);
                if (is_z && is_big(es[j].ci)) {
                    const size_t gi = es[j].ci < big1 ? (size_t)(es[j].ci - big0) : (size_t)C + (es[j].ci - big2);
                    for (int k = 0; k < 4; k++) gcoef[gi * 4 + k] = add(gcoef[gi * 4 + k], ac.v[k]);
                    continue;
                }
                for (int k = 0; k < 4; k++) coefs.push_back(ac.v[k]);
                ptrs.push_back(col_ptr[es[j].ci]);
                logs.push_back(col_log[es[j].ci]);
            }
            keep32.emplace_back(ctx, coefs.size());
            keep_ptr.emplace_back(ctx, ptrs.size());
            keep8.emplace_back(ctx, logs.size());
            CB_CUDA(cudaMemcpyAsync(keep32.back().p, coefs.data(), coefs.size() * 4, cudaMemcpyHostToDevice, st));
            CB_CUDA(cudaMemcpyAsync(keep_ptr.back().p, ptrs.data(), ptrs.size() * sizeof(void*), cudaMemcpyHostToDevice, st));
            CB_CUDA(cudaMemcpyAsync(keep8.back().p, logs.data(), logs.size(), cudaMemcpyHostToDevice, st));
            ctx->sync();  // the host vectors go out of scope
            QuotBatch qb{};
            qb.prx = {pt.x.v[0], pt.x.v[1]}; qb.pix = {pt.x.v[2], pt.x.v[3]};
            qb.pry = {pt.y.v[0], pt.y.v[1]}; qb.piy = {pt.y.v[2], pt.y.v[3]};
            qb.lin_a = lin_a; qb.lin_b = lin_b; qb.batch_coeff = qone();  // batches are summed
            qb.coefs = keep32.back().p; qb.col_idx = nullptr; qb.n_cols = (int)ptrs.size();
            qb.col_ptr = keep_ptr.back().p; qb.col_log = keep8.back().p;
            qbs.push_back(qb);
        }
        // (rows [0, C) of the table were written by the kernel above; the host part covers the interaction columns)
        CB_CUDA(cudaMemcpyAsync(d_gcoef.p + (size_t)C * 4, gcoef.data() + (size_t)C * 4, (gcoef.size() - (size_t)C * 4) * 4, cudaMemcpyHostToDevice, st));
        CB_CUDA(launch_rowcomb_m31(st, T.p, N, C, N, d_gcoef.p, g.p, 0));
        CB_CUDA(launch_rowcomb_m31(st, I.p, N, NI, N, d_gcoef.p + (size_t)C * 4, g.p, 1));
        {
            ColSrc gs{SRC_M31, g.p, N, 0};
            CB_CUDA(launch_fft(st, gs, 4, n, 1, 1 | 4, nullptr, 0, g_lde.p, M, ctx->tw, g.p, N));
        }
        ctx->launches += 5;
        DBuf<QuotBatch> d_qb(ctx, qbs.size());
        CB_CUDA(cudaMemcpyAsync(d_qb.p, qbs.data(), qbs.size() * sizeof(QuotBatch), cudaMemcpyHostToDevice, st));
        CB_CUDA(launch_quotients(st, nullptr, 0, 0, nullptr, 0, d_qb.p, (int)qbs.size(), m, ctx->tw, quot.p, M));
        ctx->launches++;
        ctx->sync();
    }
    ctx->stage_end();

    // ---- FRI commit, proof of work, queries
    ctx->stage_begin("fri_commit");
    FriProverState fri = fri_commit(ctx, ch, cfg, std::move(quot), m);
    ctx->stage_end();
    ctx->stage_begin("grind");
    const uint64_t pow_nonce = grind(ctx, ch, cfg.pow_bits);
    ctx->stage_end();
    ch.mix_u64(pow_nonce);
    const std::vector<uint32_t> queries = host::queries_generate(ch, m, cfg.n_queries);
    const int nq = (int)queries.size();

    // ---- decommit
    ctx->stage_begin("decommit");
    std::vector<uint8_t> fri_bytes = fri_decommit(ctx, fri, cfg, queries);
    auto lifted = [&](int lde_log) {
        std::vector<uint32_t> r(nq);
        const int sh = m - lde_log;
        for (int i = 0; i < nq; i++) r[i] = sh ? (((queries[i] >> (sh + 1)) << 1) | (queries[i] & 1)) : queries[i];
        return r;
    };
    auto gather = [&](const uint32_t* base, size_t stride, int ncols, int lde_log, std::vector<uint8_t>& out) {
        std::vector<uint32_t> rows = lifted(lde_log), vals((size_t)ncols * nq);
        DBuf<uint32_t> d_rows(ctx, nq), d_v(ctx, vals.size());
        CB_CUDA(cudaMemcpyAsync(d_rows.p, rows.data(), nq * 4, cudaMemcpyHostToDevice, st));
        CB_CUDA(launch_gather_rows(st, base, stride, ncols, d_rows.p, nq, d_v.p));
        ctx->launches++;
        CB_CUDA(cudaMemcpyAsync(vals.data(), d_v.p, vals.size() * 4, cudaMemcpyDeviceToHost, st));
        ctx->sync();
        for (int j = 0; j < ncols; j++) { host::put_u64(out, nq); host::put_bytes(out, &vals[(size_t)j * nq], 4 * nq); }
    };
    std::vector<std::vector<Hash32>> decs;
    for (int t = 0; t < 4; t++) {
        std::vector<uint32_t> pos = lifted(trees[t].merkle.log_leaves);
        decs.push_back(merkle_decommit(ctx, trees[t].merkle, pos));
    }
    std::vector<uint8_t> qv[4];
    gather(pre_lde.p, 512, 2, 9, qv[0]);
    gather(lde1.p, M, C, m, qv[1]);
    gather(mult_lde.p, 512, 1, 9, qv[1]);
    gather(inter_lde.p, M, NI, m, qv[2]);
    gather(tinter_lde.p, 512, 4, 9, qv[2]);
    gather(comp_lde.p, M, 8, m, qv[3]);
    ctx->stage_end();

    // ---- serialise AESCtrProof{stmt0, stmt1, StarkProof(CommitmentSchemeProof{...})}
    proof.clear();
    host::put_bytes(proof, stmt.data(), stmt.size());
    host::put_qm31(proof, csum);
    host::put_qm31(proof, tsum);
    host::put_u64(proof, NI);
    host::put_u64(proof, 4);
    cfg.serialize(proof);
    host::put_u64(proof, roots.size());
    for (auto& r : roots) host::put_bytes(proof, r.b, 32);
    host::put_u64(proof, 4);
    for (auto& t : samples) {
        host::put_u64(proof, t.size());
        for (auto& c : t) {
            host::put_u64(proof, c.size());
            for (auto& s : c) host::put_qm31(proof, s.val);
        }
    }
    host::put_u64(proof, 4);
    for (auto& d : decs) {
        host::put_u64(proof, d.size());
        for (auto& h : d) host::put_bytes(proof, h.b, 32);
    }
    host::put_u64(proof, 4);
    const size_t ncols_tree[4] = {2, (size_t)C + 1, (size_t)NI + 4, 8};
    for (int t = 0; t < 4; t++) {
        host::put_u64(proof, ncols_tree[t]);
        host::put_bytes(proof, qv[t].data(), qv[t].size());
    }
    host::put_u64(proof, pow_nonce);
    host::put_bytes(proof, fri_bytes.data(), fri_bytes.size());
    ctx->collect_stages();
    return "";
}
