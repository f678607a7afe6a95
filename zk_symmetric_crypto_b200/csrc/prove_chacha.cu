// ChaCha20 stream proof driver: replays the reference's prove flow on the GPU backend.
//
// Mirrors /root/reference/stwo/src/wasm_api.rs:467-602 (generate_chacha20_proof: validation, log_size, lane packing) and
// /root/reference/stwo/src/chacha/bitwise/air_stream.rs:143-234 (prove_stream_with_inputs / prove_stream_internal:
// empty preprocessed tree, witness, statement mixing :66-123, trace commit, prove), followed by upstream
// stwo::prover::prove / CommitmentSchemeProver::prove_values / FriProver::{commit,decommit}.
// The Fiat-Shamir channel (tiny Blake2s calls) stays on the host; every heavy step is a kernel from kernels_*.cu.
// Output bytes = bincode(StreamProof{stmt, stark_proof}) exactly as the reference serialises it
// (air_stream.rs:30-131, wasm_api.rs:588): byte-identical to the reference, checked in tests/.
#include "prover.hpp"

using namespace m31;
using host::Channel;
using host::Hash32;

namespace {

constexpr int N_COLS = 33280;
constexpr int N_WORDS = 1040;
constexpr int N_CONSTRAINTS = 54784;

// ---- host evaluation of the AIR on QM31 mask values (prove()'s closing sanity check; same sequence as the kernel) ----
struct QAcc {
    const std::vector<QM31>& apr;  // apr[k] = alpha^(K-1-k)
    QM31 acc = qzero();
    int k = 0;
    void add(QM31 c) { acc = qadd(acc, qmul(c, apr[k++])); }
};

QM31 eval_constraints_at_mask(const std::vector<QM31>& v, const std::vector<QM31>& apr) {
    QAcc A{apr};
    int col = 0;
    const QM31 one = qone();
    auto boolc = [&](QM31 b) { return qmul(b, qsub(one, b)); };
    using U32 = std::array<QM31, 32>;
    auto next_u32 = [&]() {
        U32 r;
        for (int i = 0; i < 32; i++) { r[i] = v[col++]; A.add(boolc(r[i])); }
        return r;
    };
    auto add_u32 = [&](const U32& a, const U32& b) {
        U32 res = next_u32();
        U32 car;
        for (int i = 0; i < 32; i++) car[i] = v[col++];
        for (int i = 0; i < 32; i++) {
            QM31 cin = i == 0 ? qzero() : car[i - 1];
            A.add(boolc(car[i]));
            A.add(qsub(qsub(qsub(qadd(res[i], qadd(car[i], car[i])), a[i]), b[i]), cin));
        }
        return res;
    };
    auto xor_rotl = [&](const U32& a, const U32& b, int r) {
        U32 res = next_u32();
        for (int i = 0; i < 32; i++) {
            int s = (i + 32 - r) % 32;
            QM31 ab = qmul(a[s], b[s]);
            A.add(qadd(qsub(qsub(res[i], a[s]), b[s]), qadd(ab, ab)));
        }
        return res;
    };
    std::array<U32, 16> init, s;
    for (int i = 0; i < 16; i++) init[i] = next_u32();
    s = init;
    static const int QRS[8][4] = {{0, 4, 8, 12}, {1, 5, 9, 13}, {2, 6, 10, 14}, {3, 7, 11, 15},
                                  {0, 5, 10, 15}, {1, 6, 11, 12}, {2, 7, 8, 13}, {3, 4, 9, 14}};
    for (int rnd = 0; rnd < 10; rnd++)
        for (auto& q : QRS) {
            int a = q[0], b = q[1], c = q[2], d = q[3];
            s[a] = add_u32(s[a], s[b]); s[d] = xor_rotl(s[a], s[d], 16);
            s[c] = add_u32(s[c], s[d]); s[b] = xor_rotl(s[c], s[b], 12);
            s[a] = add_u32(s[a], s[b]); s[d] = xor_rotl(s[a], s[d], 8);
            s[c] = add_u32(s[c], s[d]); s[b] = xor_rotl(s[c], s[b], 7);
        }
    std::array<U32, 16> ks, pt, ct;
    for (int i = 0; i < 16; i++) ks[i] = add_u32(s[i], init[i]);
    for (int i = 0; i < 16; i++) pt[i] = next_u32();
    for (int i = 0; i < 16; i++) ct[i] = next_u32();
    for (int i = 0; i < 16; i++)
        for (int b = 0; b < 32; b++) {
            QM31 kp = qmul(ks[i][b], pt[i][b]);
            A.add(qsub(qsub(qadd(ks[i][b], pt[i][b]), qadd(kp, kp)), ct[i][b]));
        }
    return A.acc;
}

}  // namespace

// Proves ChaCha20 encryption of `len` bytes (multiple of 64).  On success fills proof bytes (bincode StreamProof).
// Returns "" on success, else the reference's error string.
std::string prove_chacha20(cb_ctx* ctx, const uint8_t key[32], const uint8_t nonce[12], uint32_t counter, const uint8_t* plaintext,
                           const uint8_t* ciphertext, size_t len, std::vector<uint8_t>& proof, ProveOptions opt) {
    const PcsConfig cfg;
    const uint32_t num_blocks = (uint32_t)(len / 64);
    int log_size = 4;
    while (((size_t)1 << log_size) < num_blocks) log_size++;
    if (opt.force_log_size > log_size) log_size = opt.force_log_size;
    if (log_size > 24) return "log_size (" + std::to_string(log_size) + ") must be <= MAX_LOG_SIZE (24)";
    const int n = log_size, m = n + cfg.log_blowup;  // trace / LDE domain logs
    const size_t N = (size_t)1 << n, M = (size_t)1 << m;
    const uint32_t rows_needed = (num_blocks + 15) / 16;
    cudaStream_t st = ctx->stream;
    ctx->ensure_twiddles(m);
    ctx->pending_events.clear();

    uint32_t key_w[8], nonce_w[3];
    for (int i = 0; i < 8; i++) key_w[i] = host::load_le32(key + 4 * i);
    for (int i = 0; i < 3; i++) nonce_w[i] = host::load_le32(nonce + 4 * i);

    Channel ch;
    std::vector<Hash32> roots;
    // tree 0: empty preprocessed tree -> root = Blake2s("")
    roots.push_back(host::blake2s_bytes(nullptr, 0));
    ch.mix_root(roots[0]);

    // ---- witness
    ctx->stage_begin("witness");
    DBuf<uint32_t> d_pt, d_ct, W(ctx, (size_t)N_WORDS * N);
    DBuf<int> d_invalid(ctx, 1);
    const uint32_t *pt_d = opt.pt_dev, *ct_d = opt.ct_dev;
    if (!pt_d) {
        d_pt = DBuf<uint32_t>(ctx, len / 4);
        d_ct = DBuf<uint32_t>(ctx, len / 4);
        CB_CUDA(cudaMemcpyAsync(d_pt.p, plaintext, len, cudaMemcpyHostToDevice, st));
        CB_CUDA(cudaMemcpyAsync(d_ct.p, ciphertext, len, cudaMemcpyHostToDevice, st));
        pt_d = d_pt.p;
        ct_d = d_ct.p;
    }
    CB_CUDA(cudaMemsetAsync(d_invalid.p, 0, sizeof(int), st));
    CB_CUDA(launch_chacha_witness(st, key_w, nonce_w, counter, num_blocks, rows_needed * 16, pt_d, ct_d, n, W.p, N, d_invalid.p));
    ctx->launches++;
    ctx->stage_end();
    int invalid = 0;
    CB_CUDA(cudaMemcpyAsync(&invalid, d_invalid.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    ctx->sync();
    if (invalid) return "Ciphertext does not match encryption - invalid witness";

    // ---- statement
    std::vector<uint8_t> stmt;
    host::put_u32(stmt, (uint32_t)log_size);
    host::put_bytes(stmt, nonce, 12);
    host::put_u32(stmt, counter);
    Hash32 pth, cth;
    if (opt.pt_hash) {
        memcpy(pth.b, opt.pt_hash, 32);
        memcpy(cth.b, opt.ct_hash, 32);
    } else if (opt.empty_public_hashes) {
        pth = host::blake2s_bytes(nullptr, 0);
        cth = pth;
    } else {
        pth = host::blake2s_bytes(plaintext, len);
        cth = host::blake2s_bytes(ciphertext, len);
    }
    host::put_bytes(stmt, pth.b, 32);
    host::put_bytes(stmt, cth.b, 32);
    ch.mix_u64((uint64_t)log_size);
    for (int i = 0; i < 3; i++) ch.mix_u64(host::load_le32(&stmt[4 + 4 * i]));
    ch.mix_u64(counter);
    for (int i = 0; i < 16; i++) ch.mix_u64(host::load_le32(&stmt[20 + 4 * i]));

    // ---- tree 1: interpolate + LDE + Merkle
    DBuf<uint32_t> coeffs(ctx, (size_t)N_COLS * N), lde(ctx, (size_t)N_COLS * M);
    {
        ColSrc src{SRC_BITS, W.p, N, 0};
        StageHook hk = ctx->hook();
        CB_CUDA(launch_fft(st, src, N_COLS, n, cfg.log_blowup, 1 | 2 | 4, coeffs.p, N, lde.p, M, ctx->tw, coeffs.p, N, &hk));
        ctx->launches += (m <= 13) ? 1 : 3;
    }
    LeafGroups g1{};
    g1.n = 1;
    g1.g[0] = {lde.p, M, N_COLS, m};
    DevMerkle tree1 = build_merkle(ctx, g1, m, "trace_merkle_leaves");
    roots.push_back(tree1.root);
    ch.mix_root(tree1.root);

    // ---- composition polynomial
    QM31 random_coeff = ch.draw_secure_felt();
    ctx->stage_begin("constraints");
    DBuf<uint32_t> apr(ctx, (size_t)N_CONSTRAINTS * 4), d_den(ctx, (size_t)1 << cfg.log_blowup), acc(ctx, 4 * M);
    CB_CUDA(launch_secure_powers_rev(st, random_coeff, N_CONSTRAINTS, apr.p));
    {
        std::vector<uint32_t> den((size_t)1 << cfg.log_blowup);
        for (uint32_t i = 0; i < den.size(); i++) {
            uint32_t row = i << n;
            host::Pt p = host::index_to_point(host::canonic_index_at(m, host::bit_reverse(row, m)));
            den[i] = inv(host::coset_vanishing_m31(n, p));
        }
        CB_CUDA(cudaMemcpyAsync(d_den.p, den.data(), den.size() * 4, cudaMemcpyHostToDevice, st));
        ctx->sync();
    }
    CB_CUDA(launch_chacha_constraints(st, lde.p, M, m, n, apr.p, d_den.p, acc.p, M, 0));
    ctx->launches += 2;
    ctx->stage_end();
    ctx->stage_begin("composition_commit");
    // interpolate the 4 coordinate columns (log m), split into halves, evaluate each half (log n) on the LDE domain
    DBuf<uint32_t> comp_coef(ctx, 4 * M), comp_lde(ctx, 8 * M), scratch(ctx, 4 * M);
    {
        ColSrc src{SRC_M31, acc.p, M, 0};
        CB_CUDA(launch_fft(st, src, 4, m, 0, 1 | 2, comp_coef.p, M, nullptr, 0, ctx->tw, scratch.p, M));
        for (int half = 0; half < 2; half++) {
            ColSrc cs{SRC_M31, comp_coef.p + half * N, M, 0};
            CB_CUDA(launch_fft(st, cs, 4, n, cfg.log_blowup, 4, nullptr, 0, comp_lde.p + (size_t)half * 4 * M, M, ctx->tw, nullptr, 0));
        }
        ctx->launches += 6;
    }
    LeafGroups g2{};
    g2.n = 1;
    g2.g[0] = {comp_lde.p, M, 8, m};
    DevMerkle tree2 = build_merkle(ctx, g2, m);
    ctx->stage_end();
    roots.push_back(tree2.root);
    ch.mix_root(tree2.root);

    // ---- OODS sampling
    host::CirclePointQ z = host::get_random_point(ch);
    ctx->stage_begin("oods");
    std::vector<QM31> sampled((size_t)N_COLS + 8);
    {
        std::vector<QM31> maps(n);
        maps[0] = z.y;
        QM31 x = z.x;
        for (int j = 1; j < n; j++) { maps[j] = x; x = qsub(qmul_m(qmul(x, x), 2), qone()); }
        DBuf<uint32_t> basis(ctx, 4 * N), d_sampled(ctx, ((size_t)N_COLS + 8) * 4);
        CB_CUDA(launch_basis(st, basis.p, N, n, maps.data()));
        CB_CUDA(launch_oods_dot(st, coeffs.p, N, N_COLS, n, basis.p, N, d_sampled.p));
        for (int half = 0; half < 2; half++)
            CB_CUDA(launch_oods_dot(st, comp_coef.p + half * N, M, 4, n, basis.p, N, d_sampled.p + ((size_t)N_COLS + 4 * half) * 4));
        ctx->launches += n + 3;
        CB_CUDA(cudaMemcpyAsync(sampled.data(), d_sampled.p, sampled.size() * 16, cudaMemcpyDeviceToHost, st));
        ctx->sync();
    }
    ctx->stage_end();
    ch.mix_felts(sampled.data(), sampled.size());

    // ---- FRI quotients
    QM31 rc = ch.draw_secure_felt();
    ctx->stage_begin("quotients");
    DBuf<uint32_t> quot(ctx, 4 * M);
    {
        const size_t nc = sampled.size();
        std::vector<uint32_t> coefs(nc * 4);
        QM31 alpha = qone(), lin_a = qzero(), lin_b = qzero();
        const QM31 c = qsub(qconj(z.y), z.y);
        for (size_t j = 0; j < nc; j++) {
            QM31 v = sampled[j];
            QM31 a = qsub(qconj(v), v);
            QM31 b = qsub(qmul(v, c), qmul(a, z.y));
            lin_a = qadd(lin_a, qmul(alpha, a));
            lin_b = qadd(lin_b, qmul(alpha, b));
            QM31 ac = qmul(alpha, c);
            for (int k = 0; k < 4; k++) coefs[j * 4 + k] = ac.v[k];
            alpha = qmul(alpha, rc);
        }
        DBuf<uint32_t> d_coefs(ctx, nc * 4);
        CB_CUDA(cudaMemcpyAsync(d_coefs.p, coefs.data(), coefs.size() * 4, cudaMemcpyHostToDevice, st));
        QuotBatch qb{};
        qb.prx = {z.x.v[0], z.x.v[1]}; qb.pix = {z.x.v[2], z.x.v[3]};
        qb.pry = {z.y.v[0], z.y.v[1]}; qb.piy = {z.y.v[2], z.y.v[3]};
        qb.lin_a = lin_a; qb.lin_b = lin_b; qb.batch_coeff = qzero();
        qb.coefs = d_coefs.p; qb.col_idx = nullptr; qb.n_cols = (int)nc;
        DBuf<QuotBatch> d_qb(ctx, 1);
        CB_CUDA(cudaMemcpyAsync(d_qb.p, &qb, sizeof(qb), cudaMemcpyHostToDevice, st));
        CB_CUDA(launch_quotients(st, lde.p, M, N_COLS, comp_lde.p, M, d_qb.p, 1, m, ctx->tw, quot.p, M));
        ctx->launches++;
        ctx->sync();
    }
    ctx->stage_end();

    // ---- FRI commit
    ctx->stage_begin("fri_commit");
    FriProverState fri = fri_commit(ctx, ch, cfg, std::move(quot), m);
    ctx->stage_end();

    // ---- proof of work, queries
    ctx->stage_begin("grind");
    uint64_t pow_nonce = grind(ctx, ch, cfg.pow_bits);
    ctx->stage_end();
    ch.mix_u64(pow_nonce);
    std::vector<uint32_t> queries = host::queries_generate(ch, m, cfg.n_queries);

    // ---- decommit
    ctx->stage_begin("decommit");
    std::vector<uint8_t> fri_bytes = fri_decommit(ctx, fri, cfg, queries);
    std::vector<Hash32> dec1 = merkle_decommit(ctx, tree1, queries), dec2 = merkle_decommit(ctx, tree2, queries);
    const int nq = (int)queries.size();
    std::vector<uint32_t> qv1((size_t)N_COLS * nq), qv2((size_t)8 * nq);
    {
        DBuf<uint32_t> d_rows(ctx, nq), d_q1(ctx, qv1.size()), d_q2(ctx, qv2.size());
        CB_CUDA(cudaMemcpyAsync(d_rows.p, queries.data(), nq * 4, cudaMemcpyHostToDevice, st));
        CB_CUDA(launch_gather_rows(st, lde.p, M, N_COLS, d_rows.p, nq, d_q1.p));
        CB_CUDA(launch_gather_rows(st, comp_lde.p, M, 8, d_rows.p, nq, d_q2.p));
        ctx->launches += 2;
        CB_CUDA(cudaMemcpyAsync(qv1.data(), d_q1.p, qv1.size() * 4, cudaMemcpyDeviceToHost, st));
        CB_CUDA(cudaMemcpyAsync(qv2.data(), d_q2.p, qv2.size() * 4, cudaMemcpyDeviceToHost, st));
        ctx->sync();
    }
    ctx->stage_end();

    // ---- prove()'s closing check: composition OODS value == constraints on the sampled mask / Z_H(z)
    {
        std::vector<QM31> aprh(N_CONSTRAINTS);
        QM31 cur = qone();
        for (int e = 0; e < N_CONSTRAINTS; e++) { aprh[N_CONSTRAINTS - 1 - e] = cur; cur = qmul(cur, random_coeff); }
        std::vector<QM31> mask(sampled.begin(), sampled.begin() + N_COLS);
        QM31 num = eval_constraints_at_mask(mask, aprh);
        QM31 zh = coset_vanishing_q(n, z);
        QM31 expect = qmul(num, qinv(zh));
        const QM31 units[4] = {{{1, 0, 0, 0}}, {{0, 1, 0, 0}}, {{0, 0, 1, 0}}, {{0, 0, 0, 1}}};
        QM31 left = qzero(), right = qzero();
        for (int k = 0; k < 4; k++) {
            left = qadd(left, qmul(sampled[N_COLS + k], units[k]));
            right = qadd(right, qmul(sampled[N_COLS + 4 + k], units[k]));
        }
        QM31 pix = z.x;
        for (int i = 0; i < n - 1; i++) pix = qsub(qmul_m(qmul(pix, pix), 2), qone());
        if (!qeq(qadd(left, qmul(pix, right)), expect)) return "Proof generation failed: ConstraintsNotSatisfied";
    }

    // ---- serialise: StreamProof{stmt, StarkProof(CommitmentSchemeProof{config, commitments, sampled_values, decommitments,
    //                 queried_values, proof_of_work, fri_proof})}
    proof.clear();
    proof.reserve(stmt.size() + 64 + sampled.size() * 24 + (size_t)N_COLS * (8 + 4 * nq) + 4096);
    host::put_bytes(proof, stmt.data(), stmt.size());
    cfg.serialize(proof);
    host::put_u64(proof, roots.size());
    for (auto& r : roots) host::put_bytes(proof, r.b, 32);
    host::put_u64(proof, 3);
    host::put_u64(proof, 0);
    host::put_u64(proof, N_COLS);
    for (int j = 0; j < N_COLS; j++) { host::put_u64(proof, 1); host::put_qm31(proof, sampled[j]); }
    host::put_u64(proof, 8);
    for (int j = 0; j < 8; j++) { host::put_u64(proof, 1); host::put_qm31(proof, sampled[N_COLS + j]); }
    host::put_u64(proof, 3);
    host::put_u64(proof, 0);
    host::put_u64(proof, dec1.size());
    for (auto& h : dec1) host::put_bytes(proof, h.b, 32);
    host::put_u64(proof, dec2.size());
    for (auto& h : dec2) host::put_bytes(proof, h.b, 32);
    host::put_u64(proof, 3);
    host::put_u64(proof, 0);
    host::put_u64(proof, N_COLS);
    for (int j = 0; j < N_COLS; j++) { host::put_u64(proof, nq); host::put_bytes(proof, &qv1[(size_t)j * nq], 4 * nq); }
    host::put_u64(proof, 8);
    for (int j = 0; j < 8; j++) { host::put_u64(proof, nq); host::put_bytes(proof, &qv2[(size_t)j * nq], 4 * nq); }
    host::put_u64(proof, pow_nonce);
    host::put_bytes(proof, fri_bytes.data(), fri_bytes.size());
    ctx->collect_stages();
    return "";
}
