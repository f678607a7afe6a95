// C ABI, second part (include/s2c_b200.h): the rest of the backend-trait surface of SURVEY.md 8(b) -- ColumnOps / FieldOps
// helpers, barycentric evaluation, legacy commit_on_layer, coset-parameterised twiddles, multi-batch quotient accumulation and
// the AES-CTR AIR stages (trace generation, LogUp interaction trace, both components' constraint quotients,
// AccumulationOps::{accumulate, lift_and_accumulate}).  Thin, exception-free wrappers over kernels_*.cu, like cb_api.cu.
#include <stdlib.h>
#include "../../include/s2c_b200.h"
#include "prover.hpp"

using namespace m31;

#define CB_TRY(ctx) try {
#define CB_CATCH(ctx)                                  \
    }                                                  \
    catch (const std::exception& e) {                  \
        if (ctx) (ctx)->err = e.what();                \
        return 1;                                      \
    }                                                  \
    return 0;

namespace {
__global__ void add_columns_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = add(dst[i], src[i]);
}
QM31 q4(const uint32_t w[4]) { return QM31{{w[0], w[1], w[2], w[3]}}; }
void need(bool ok, const char* what) {
    if (!ok) throw CbError(what);
}
}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------- ColumnOps / FieldOps
int cb_bit_reverse(cb_ctx* ctx, uint32_t* col, int log_size) {
    CB_TRY(ctx)
    need(log_size >= 0 && log_size <= 30, "cb_bit_reverse: log_size out of range");
    CB_CUDA(launch_bit_reverse(ctx->stream, col, log_size));
    ctx->launches++;
    CB_CATCH(ctx)
}

int cb_batch_inverse_m31(cb_ctx* ctx, const uint32_t* src, uint32_t* dst, size_t n) {
    CB_TRY(ctx)
    CB_CUDA(launch_inverse_m31(ctx->stream, src, dst, n));
    ctx->launches++;
    CB_CATCH(ctx)
}

int cb_batch_inverse_qm31(cb_ctx* ctx, const uint32_t* src, size_t src_stride, uint32_t* dst, size_t dst_stride, size_t n) {
    CB_TRY(ctx)
    CB_CUDA(launch_inverse_qm31(ctx->stream, src, src_stride, dst, dst_stride, n));
    ctx->launches++;
    CB_CATCH(ctx)
}

int cb_col_at(cb_ctx* ctx, const uint32_t* col, size_t index, uint32_t* value_out) {
    CB_TRY(ctx)
    CB_CUDA(cudaMemcpyAsync(value_out, col + index, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    CB_CATCH(ctx)
}

int cb_col_set(cb_ctx* ctx, uint32_t* col, size_t index, uint32_t value) {
    CB_TRY(ctx)
    need(value < P, "cb_col_set: value is not a canonical M31 word");
    CB_CUDA(cudaMemcpyAsync(col + index, &value, 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->sync();
    CB_CATCH(ctx)
}

// ---------------------------------------------------------------------------------------------- PolyOps
int cb_extend(cb_ctx* ctx, const uint32_t* coeffs, size_t stride, int n_cols, int log_size, int log_ext, uint32_t* out, size_t out_stride) {
    CB_TRY(ctx)
    need(log_ext >= 0 && log_size + log_ext <= 30, "cb_extend: log size out of range");
    const size_t n = (size_t)1 << log_size, big = (size_t)1 << (log_size + log_ext);
    for (int c = 0; c < n_cols; c++) {
        CB_CUDA(cudaMemcpyAsync(out + (size_t)c * out_stride, coeffs + (size_t)c * stride, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        if (big > n) CB_CUDA(cudaMemsetAsync(out + (size_t)c * out_stride + n, 0, (big - n) * 4, ctx->stream));
    }
    CB_CATCH(ctx)
}

int cb_barycentric_weights(cb_ctx* ctx, int log_size, const uint32_t pt[8], uint32_t* weights_out) {
    CB_TRY(ctx)
    need(log_size >= 1 && log_size <= 28, "cb_barycentric_weights: log_size out of range");
    ctx->ensure_twiddles(log_size);
    const size_t N = (size_t)1 << log_size;
    QM31 x = q4(pt), y = q4(pt + 4);
    std::vector<QM31> maps(log_size);
    maps[0] = y;
    for (int j = 1; j < log_size; j++) { maps[j] = x; x = qsub(qmul_m(qmul(x, x), 2), qone()); }
    // weights = 2^-n (iFFT)^T basis(point): the transposed inverse transform is the forward butterfly network run with the
    // inverse twiddles (same construction as the streaming prover's out-of-domain sampling, prove_chacha.cu)
    DBuf<uint32_t> basis(ctx, 4 * N);
    CB_CUDA(launch_basis(ctx->stream, basis.p, N, log_size, maps.data()));
    const FftTables tw_t{ctx->tw.IX, ctx->tw.IY, ctx->tw.X, ctx->tw.Y, ctx->tw.max_log};
    ColSrc bs{SRC_M31, basis.p, N, 0};
    CB_CUDA(launch_fft(ctx->stream, bs, 4, log_size, 0, 4, nullptr, 0, weights_out, N, tw_t, nullptr, 0));
    // in place: weights *= 2^-n (the row-scaling kernel of the constraint pass with a one-entry table)
    const uint32_t inv_n = 1u << (31 - log_size);
    DBuf<uint32_t> d_s(ctx, 1);
    CB_CUDA(cudaMemcpyAsync(d_s.p, &inv_n, 4, cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(launch_scale_rows(ctx->stream, weights_out, N, log_size, d_s.p, 0));
    ctx->launches += log_size + 3;
    ctx->sync();
    CB_CATCH(ctx)
}

int cb_barycentric_eval_at_point(cb_ctx* ctx, const uint32_t* evals, size_t stride, int n_cols, int log_size, const uint32_t* weights,
                                 uint32_t* out_host) {
    CB_TRY(ctx)
    const size_t N = (size_t)1 << log_size;
    DBuf<uint32_t> d_out(ctx, (size_t)n_cols * 4);
    CB_CUDA(launch_oods_dot(ctx->stream, evals, stride, n_cols, log_size, weights, N, d_out.p));
    ctx->launches++;
    CB_CUDA(cudaMemcpyAsync(out_host, d_out.p, (size_t)n_cols * 16, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    CB_CATCH(ctx)
}

int cb_precompute_twiddles_coset(cb_ctx* ctx, uint32_t coset_initial_index, int coset_log_size, uint32_t* twiddles_out,
                                 uint32_t* itwiddles_out) {
    CB_TRY(ctx)
    need(coset_log_size >= 1 && coset_log_size <= 28, "cb_precompute_twiddles_coset: log size out of range");
    // upstream slow_precompute_twiddles: for every layer the x coordinates of the (repeatedly doubled) coset's first half in
    // bit-reversed order, then a trailing 1; itwiddles = element-wise inverses
    const size_t n = (size_t)1 << coset_log_size;
    std::vector<uint32_t> tw;
    tw.reserve(n);
    uint32_t initial = coset_initial_index & 0x7fffffffu;
    uint32_t step = 1u << (31 - coset_log_size);
    for (int lg = coset_log_size; lg >= 1; lg--) {
        const size_t half = (size_t)1 << (lg - 1);
        const size_t i0 = tw.size();
        tw.resize(i0 + half);
        for (size_t i = 0; i < half; i++) {
            const uint32_t idx = (uint32_t)(((uint64_t)initial + (uint64_t)step * i) & 0x7fffffffu);
            tw[i0 + host::bit_reverse((uint32_t)i, lg - 1)] = host::index_to_point(idx).x;
        }
        initial = (uint32_t)((2ull * initial) & 0x7fffffffu);
        step = (uint32_t)((2ull * step) & 0x7fffffffu);
    }
    tw.push_back(1);
    std::vector<uint32_t> itw(tw.size());
    for (size_t i = 0; i < tw.size(); i++) itw[i] = inv(tw[i]);
    CB_CUDA(cudaMemcpyAsync(twiddles_out, tw.data(), tw.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(cudaMemcpyAsync(itwiddles_out, itw.data(), itw.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->sync();
    CB_CATCH(ctx)
}

// ---------------------------------------------------------------------------------------------- MerkleOps (legacy)
int cb_commit_on_layer(cb_ctx* ctx, int log_size, const uint32_t* prev_or_null, const uint32_t* const* cols_host, int n_cols,
                       uint32_t* out) {
    CB_TRY(ctx)
    need(log_size >= 0 && log_size <= 30 && n_cols >= 0, "cb_commit_on_layer: bad arguments");
    DBuf<const uint32_t*> d_cols(ctx, n_cols > 0 ? n_cols : 1);
    if (n_cols) CB_CUDA(cudaMemcpyAsync(d_cols.p, cols_host, (size_t)n_cols * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(launch_commit_on_layer(ctx->stream, prev_or_null, d_cols.p, n_cols, 1u << log_size, out));
    ctx->launches++;
    ctx->sync();
    CB_CATCH(ctx)
}

// ---------------------------------------------------------------------------------------------- AccumulationOps
int cb_accumulate(cb_ctx* ctx, uint32_t* dst, const uint32_t* src, size_t n_words) {
    CB_TRY(ctx)
    add_columns_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, ctx->stream>>>(dst, src, n_words);
    CB_CUDA(cudaGetLastError());
    ctx->launches++;
    CB_CATCH(ctx)
}

int cb_lift_and_accumulate(cb_ctx* ctx, uint32_t* big, size_t big_stride, int big_log, const uint32_t* small_cols, int small_log) {
    CB_TRY(ctx)
    need(small_log <= big_log && small_log >= 1, "cb_lift_and_accumulate: the lifted accumulation must not be larger");
    CB_CUDA(launch_lift_accumulate(ctx->stream, big, big_stride, big_log, small_cols, small_log));
    ctx->launches++;
    CB_CATCH(ctx)
}

// ---------------------------------------------------------------------------------------------- QuotientOps, general form
int cb_accumulate_quotients_batches(cb_ctx* ctx, const uint32_t* const* col_ptrs_host, const int* col_logs_host, int n_cols,
                                    int domain_log, int n_batches, const uint32_t* batch_points_host, const int* batch_offsets_host,
                                    const int* entry_col_host, const uint32_t* entry_value_host, const uint32_t* entry_alpha_host,
                                    uint32_t* out, size_t out_stride) {
    CB_TRY(ctx)
    need(n_batches >= 1 && n_cols >= 1, "cb_accumulate_quotients_batches: nothing to accumulate");
    ctx->ensure_twiddles(domain_log);
    cudaStream_t st = ctx->stream;
    std::vector<QuotBatch> qbs;
    std::vector<DBuf<uint32_t>> keep32;
    std::vector<DBuf<const uint32_t*>> keep_ptr;
    std::vector<DBuf<uint8_t>> keep8;
    for (int b = 0; b < n_batches; b++) {
        const QM31 px = q4(batch_points_host + 8 * b), py = q4(batch_points_host + 8 * b + 4);
        const QM31 c = qsub(qconj(py), py);
        std::vector<uint32_t> coefs;
        std::vector<const uint32_t*> ptrs;
        std::vector<uint8_t> logs;
        QM31 lin_a = qzero(), lin_b = qzero();
        for (int e = batch_offsets_host[b]; e < batch_offsets_host[b + 1]; e++) {
            const int ci = entry_col_host[e];
            need(ci >= 0 && ci < n_cols, "cb_accumulate_quotients_batches: column index out of range");
            need(col_logs_host[ci] <= domain_log, "cb_accumulate_quotients_batches: column larger than the domain");
            const QM31 v = q4(entry_value_host + 4 * (size_t)e), ap = q4(entry_alpha_host + 4 * (size_t)e);
            const QM31 a = qsub(qconj(v), v);
            const QM31 bb = qsub(qmul(v, c), qmul(a, py));
            lin_a = qadd(lin_a, qmul(ap, a));
            lin_b = qadd(lin_b, qmul(ap, bb));
            const QM31 ac = qmul(ap, c);
            for (int k = 0; k < 4; k++) coefs.push_back(ac.v[k]);
            ptrs.push_back(col_ptrs_host[ci]);
            logs.push_back((uint8_t)col_logs_host[ci]);
        }
        keep32.emplace_back(ctx, coefs.size() ? coefs.size() : 1);
        keep_ptr.emplace_back(ctx, ptrs.size() ? ptrs.size() : 1);
        keep8.emplace_back(ctx, logs.size() ? logs.size() : 1);
        CB_CUDA(cudaMemcpyAsync(keep32.back().p, coefs.data(), coefs.size() * 4, cudaMemcpyHostToDevice, st));
        CB_CUDA(cudaMemcpyAsync(keep_ptr.back().p, ptrs.data(), ptrs.size() * sizeof(void*), cudaMemcpyHostToDevice, st));
        CB_CUDA(cudaMemcpyAsync(keep8.back().p, logs.data(), logs.size(), cudaMemcpyHostToDevice, st));
        ctx->sync();  // the host vectors go out of scope
        QuotBatch qb{};
        qb.prx = {px.v[0], px.v[1]}; qb.pix = {px.v[2], px.v[3]};
        qb.pry = {py.v[0], py.v[1]}; qb.piy = {py.v[2], py.v[3]};
        qb.lin_a = lin_a; qb.lin_b = lin_b; qb.batch_coeff = qone();  // the pinned stwo rev sums the per-point batches
        qb.coefs = keep32.back().p; qb.col_idx = nullptr; qb.n_cols = (int)ptrs.size();
        qb.col_ptr = keep_ptr.back().p; qb.col_log = keep8.back().p;
        qbs.push_back(qb);
    }
    DBuf<QuotBatch> d_qb(ctx, qbs.size());
    CB_CUDA(cudaMemcpyAsync(d_qb.p, qbs.data(), qbs.size() * sizeof(QuotBatch), cudaMemcpyHostToDevice, st));
    CB_CUDA(launch_quotients(st, nullptr, 0, 0, nullptr, 0, d_qb.p, (int)qbs.size(), domain_log, ctx->tw, out, out_stride));
    ctx->launches++;
    ctx->sync();
    CB_CATCH(ctx)
}

// ---------------------------------------------------------------------------------------------- AES-CTR AIR stages
int cb_aes_ctr_layout(int key_len, int* n_cols, int* n_constraints, int* n_lookups, int* lookup_in_cols, int* lookup_out_cols) {
    if (key_len != 16 && key_len != 32) return 1;
    const AesLayout L = aes_make_layout(key_len == 16 ? 10 : 14);
    if (n_cols) *n_cols = L.n_cols;
    if (n_constraints) *n_constraints = L.n_constraints;
    if (n_lookups) *n_lookups = (int)L.lk_in.size();
    for (size_t i = 0; i < L.lk_in.size(); i++) {
        if (lookup_in_cols) lookup_in_cols[i] = L.lk_in[i];
        if (lookup_out_cols) lookup_out_cols[i] = L.lk_out[i];
    }
    return 0;
}

int cb_gen_trace_aes_ctr(cb_ctx* ctx, int key_len, const uint8_t* key, const uint8_t nonce[12], uint32_t counter, const uint8_t* pt_host,
                         const uint8_t* ct_host, uint32_t n_blocks, int log_size, uint32_t* trace_out, size_t stride,
                         uint32_t mults_out_host[256], int* valid) {
    CB_TRY(ctx)
    need(key_len == 16 || key_len == 32, "cb_gen_trace_aes_ctr: key_len must be 16 or 32");
    need(log_size >= 4 && log_size <= 24 && n_blocks >= 1 && n_blocks <= (1u << log_size), "cb_gen_trace_aes_ctr: bad size");
    cudaStream_t st = ctx->stream;
    const int nr = key_len == 16 ? 10 : 14;
    const size_t len = (size_t)n_blocks * 16;
    CB_CUDA(aes_upload_sbox(aes_sbox()));
    const std::vector<uint8_t> rk = aes_expand_key(key, key_len);
    DBuf<uint8_t> d_pt(ctx, len), d_ct(ctx, len);
    DBuf<unsigned int> d_mults(ctx, 256);
    DBuf<int> d_invalid(ctx, 1);
    CB_CUDA(cudaMemcpyAsync(d_pt.p, pt_host, len, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(d_ct.p, ct_host, len, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemsetAsync(d_mults.p, 0, 256 * 4, st));
    CB_CUDA(cudaMemsetAsync(d_invalid.p, 0, 4, st));
    const uint32_t rows_needed = (n_blocks + 15) / 16;
    CB_CUDA(launch_aes_witness(st, rk.data(), nr, nonce, counter, n_blocks, rows_needed * 16, d_pt.p, d_ct.p, log_size, trace_out, stride,
                               d_mults.p, d_invalid.p));
    ctx->launches++;
    int invalid = 0;
    CB_CUDA(cudaMemcpyAsync(&invalid, d_invalid.p, 4, cudaMemcpyDeviceToHost, st));
    if (mults_out_host) CB_CUDA(cudaMemcpyAsync(mults_out_host, d_mults.p, 256 * 4, cudaMemcpyDeviceToHost, st));
    ctx->sync();
    if (valid) *valid = invalid ? 0 : 1;
    CB_CATCH(ctx)
}

int cb_logup_finalize_last(cb_ctx* ctx, uint32_t* col4, size_t stride, int log_size, uint32_t claimed_sum_out[4]) {
    CB_TRY(ctx)
    need(log_size >= 0 && log_size <= 28, "cb_logup_finalize_last: log_size out of range");
    DBuf<uint32_t> scr(ctx, logup_finalize_scratch_words(log_size));
    uint32_t* d_claimed = nullptr;
    CB_CUDA(launch_logup_finalize_last(ctx->stream, col4, stride, log_size, scr.p, &d_claimed));
    ctx->launches += 3;
    CB_CUDA(cudaMemcpyAsync(claimed_sum_out, d_claimed, 16, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    CB_CATCH(ctx)
}

int cb_gen_logup_interaction_aes_ctr(cb_ctx* ctx, int key_len, const uint32_t* trace, size_t stride, int log_size, const uint32_t z[4],
                                     const uint32_t alpha[4], uint32_t* inter_out, size_t inter_stride, uint32_t claimed_sum_out[4]) {
    CB_TRY(ctx)
    need(key_len == 16 || key_len == 32, "cb_gen_logup_interaction_aes_ctr: key_len must be 16 or 32");
    const AesLayout L = aes_make_layout(key_len == 16 ? 10 : 14);
    const int NL = (int)L.lk_in.size(), NI = 4 * (NL / 2);
    cudaStream_t st = ctx->stream;
    DBuf<int> d_in(ctx, NL), d_out(ctx, NL);
    CB_CUDA(cudaMemcpyAsync(d_in.p, L.lk_in.data(), NL * 4, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(d_out.p, L.lk_out.data(), NL * 4, cudaMemcpyHostToDevice, st));
    CB_CUDA(launch_aes_interaction(st, trace, stride, log_size, d_in.p, d_out.p, NL, q4(z), q4(alpha), inter_out, inter_stride));
    DBuf<uint32_t> scr(ctx, logup_finalize_scratch_words(log_size));
    uint32_t* d_claimed = nullptr;
    CB_CUDA(launch_logup_finalize_last(st, inter_out + (size_t)(NI - 4) * inter_stride, inter_stride, log_size, scr.p, &d_claimed));
    ctx->launches += 4;
    CB_CUDA(cudaMemcpyAsync(claimed_sum_out, d_claimed, 16, cudaMemcpyDeviceToHost, st));
    ctx->sync();
    CB_CATCH(ctx)
}

static void den_table(int tlog, uint32_t out[2]) {
    for (uint32_t i = 0; i < 2; i++) {
        const uint32_t row = i << tlog;
        const host::Pt p = host::index_to_point(host::canonic_index_at(tlog + 1, host::bit_reverse(row, tlog + 1)));
        out[i] = inv(host::coset_vanishing_m31(tlog, p));
    }
}

int cb_eval_constraints_aes_ctr(cb_ctx* ctx, int key_len, const uint32_t* lde, size_t stride, const uint32_t* inter_lde,
                                size_t inter_stride, int trace_log, const uint32_t* alpha_pows_rev, const uint32_t z[4],
                                const uint32_t alpha[4], const uint32_t claimed_sum[4], uint32_t* accum, size_t accum_stride) {
    CB_TRY(ctx)
    need(key_len == 16 || key_len == 32, "cb_eval_constraints_aes_ctr: key_len must be 16 or 32");
    const AesLayout L = aes_make_layout(key_len == 16 ? 10 : 14);
    const int K = L.n_constraints, NL = (int)L.lk_in.size();
    cudaStream_t st = ctx->stream;
    const int n = trace_log, m = n + 1;
    DBuf<uint32_t> apr_lo(ctx, (size_t)K * 4), apr_hi(ctx, (size_t)K * 4), d_den(ctx, 2);
    DBuf<int> d_in(ctx, NL), d_out(ctx, NL);
    CB_CUDA(launch_split16(st, alpha_pows_rev, K, apr_lo.p, apr_hi.p));
    CB_CUDA(cudaMemcpyAsync(d_in.p, L.lk_in.data(), NL * 4, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(d_out.p, L.lk_out.data(), NL * 4, cudaMemcpyHostToDevice, st));
    uint32_t den_n[2];
    den_table(n, den_n);
    CB_CUDA(cudaMemcpyAsync(d_den.p, den_n, 8, cudaMemcpyHostToDevice, st));
    AesConsArgs a{};
    a.lde = lde; a.stride = stride; a.inter = inter_lde; a.i_stride = inter_stride;
    a.apr_lo = apr_lo.p; a.apr_hi = apr_hi.p; a.apr = alpha_pows_rev; a.den_inv = d_den.p;
    a.lk_in = d_in.p; a.lk_out = d_out.p;
    a.z = q4(z); a.alpha = q4(alpha);
    a.shift = qmul_m(q4(claimed_sum), inv((uint32_t)(((uint64_t)1 << n) % P)));
    a.eval_log = m; a.trace_log = n; a.n_rounds = key_len == 16 ? 10 : 14; a.n_lookups = NL;
    a.out = accum; a.out_stride = accum_stride;
    CB_CUDA(launch_aes_constraints(st, a));
    ctx->launches += 2;
    ctx->sync();
    CB_CATCH(ctx)
}

int cb_eval_constraints_sbox_table(cb_ctx* ctx, const uint32_t* pre_in_lde, const uint32_t* pre_out_lde, const uint32_t* mult_lde,
                                   const uint32_t* inter_lde, size_t inter_stride, const uint32_t z[4], const uint32_t alpha[4],
                                   const uint32_t claimed_sum[4], const uint32_t alpha_pow[4], uint32_t* accum) {
    CB_TRY(ctx)
    cudaStream_t st = ctx->stream;
    DBuf<uint32_t> d_den(ctx, 2);
    uint32_t den_8[2];
    den_table(8, den_8);
    CB_CUDA(cudaMemcpyAsync(d_den.p, den_8, 8, cudaMemcpyHostToDevice, st));
    AesTableArgs t{};
    t.pre_in = pre_in_lde; t.pre_out = pre_out_lde; t.mult = mult_lde; t.inter = inter_lde; t.i_stride = inter_stride;
    t.z = q4(z); t.alpha = q4(alpha); t.shift = qmul_m(q4(claimed_sum), inv(256)); t.apow = q4(alpha_pow);
    t.den_inv = d_den.p; t.eval_log = 9; t.trace_log = 8; t.out = accum;
    CB_CUDA(launch_aes_table_constraint(st, t));
    ctx->launches++;
    ctx->sync();
    CB_CATCH(ctx)
}

}  // extern "C"
