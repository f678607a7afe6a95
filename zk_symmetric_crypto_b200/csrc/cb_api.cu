// C ABI (include/s2c_b200.h): thin, exception-free wrappers over the kernels and the proof drivers.
#include <mutex>
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

static std::string g_stage_str;

extern "C" {

int cb_init(int device, cb_ctx** out) {
    if (!out) return 1;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= device) return 2;  // no usable CUDA device: there is no CPU fallback
    cb_ctx* ctx = new cb_ctx();
    try {
        ctx->device = device;
        CB_CUDA(cudaSetDevice(device));
        // the consumer stream outranks the producer stream: when both have blocks waiting (row-sharded mode: the last transform
        // pass of group g+1 is throttled by NVLink while the leaf hashing of group g wants the SMs), the consumer goes first
        int prio_lo = 0, prio_hi = 0;
        CB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CB_CUDA(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi));
        CB_CUDA(cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, prio_lo));
        CB_CUDA(cudaStreamCreateWithPriority(&ctx->stream3, cudaStreamNonBlocking, prio_lo));
        // measured on B200 at log 18/20: no gain (both kernels fill the GPU; the block scheduler runs them back to back), so
        // the second stream is opt-in
        ctx->overlap = getenv("S2C_OVERLAP") && atoi(getenv("S2C_OVERLAP"));
        cudaMemPool_t pool;
        CB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thr = UINT64_MAX;
        CB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        fft_init_attrs();
        fft2_init_attrs();
        chacha_init_attrs();
    } catch (const std::exception& ex) {
        delete ctx;
        return 3;
    }
    *out = ctx;
    return 0;
}

void cb_destroy(cb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->tw_dev) cudaFree(ctx->tw_dev);
    if (ctx->tw_shift_dev) cudaFree(ctx->tw_shift_dev);
    ctx->close_peers();
    ctx->release_arena();
    for (int v = 0; v < 2; v++) {
        if (ctx->chacha_consts[v]) cudaFree(ctx->chacha_consts[v]);
        if (ctx->chacha_cidx[v]) cudaFree(ctx->chacha_cidx[v]);
    }
    if (ctx->small_jobs) cudaFree(ctx->small_jobs);
    if (ctx->pin_buf) cudaFreeHost(ctx->pin_buf);
    if (ctx->hash_stage) cudaFreeHost(ctx->hash_stage);
    try { comm_destroy(ctx->comm); } catch (...) {}
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
    if (ctx->stream3) { cudaStreamSynchronize(ctx->stream3); cudaStreamDestroy(ctx->stream3); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* cb_last_error(cb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int cb_set_stream(cb_ctx* ctx, void* s) {
    CB_TRY(ctx)
    ctx->sync();
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (s) {
        ctx->stream = (cudaStream_t)s;
        ctx->own_stream = false;
    } else {
        CB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    CB_CATCH(ctx)
}

int cb_sync(cb_ctx* ctx) {
    CB_TRY(ctx)
    ctx->sync();
    CB_CATCH(ctx)
}
uint64_t cb_launch_count(cb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int cb_malloc(cb_ctx* ctx, size_t bytes, void** dptr) {
    CB_TRY(ctx)
    CB_CUDA(cudaSetDevice(ctx->device));
    *dptr = ctx->dmalloc(bytes);
    CB_CATCH(ctx)
}
int cb_free(cb_ctx* ctx, void* dptr) {
    CB_TRY(ctx)
    ctx->dfree(dptr);
    CB_CATCH(ctx)
}
int cb_h2d(cb_ctx* ctx, void* d, const void* h, size_t bytes) {
    CB_TRY(ctx)
    CB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->sync();
    CB_CATCH(ctx)
}
int cb_d2h(cb_ctx* ctx, void* h, const void* d, size_t bytes) {
    CB_TRY(ctx)
    CB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    CB_CATCH(ctx)
}
int cb_memset_zero(cb_ctx* ctx, void* d, size_t bytes) {
    CB_TRY(ctx)
    CB_CUDA(cudaMemsetAsync(d, 0, bytes, ctx->stream));
    CB_CATCH(ctx)
}

int cb_precompute_twiddles(cb_ctx* ctx, int max_log) {
    CB_TRY(ctx)
    ctx->ensure_twiddles(max_log);
    CB_CATCH(ctx)
}

int cb_interpolate_columns(cb_ctx* ctx, uint32_t* cols, size_t stride, int n_cols, int log_size) {
    CB_TRY(ctx)
    ctx->ensure_twiddles(log_size);
    ColSrc src{SRC_M31, cols, stride, 0};
    CB_CUDA(launch_fft(ctx->stream, src, n_cols, log_size, 0, 1 | 2, cols, stride, nullptr, 0, ctx->tw, cols, stride));
    ctx->launches += log_size <= 13 ? 1 : 2;
    CB_CATCH(ctx)
}

int cb_evaluate_polynomials(cb_ctx* ctx, const uint32_t* coeffs, size_t stride, int n_cols, int log_size, int log_ext,
                            uint32_t* evals, size_t eval_stride) {
    CB_TRY(ctx)
    ctx->ensure_twiddles(log_size + log_ext);
    ColSrc src{SRC_M31, coeffs, stride, 0};
    CB_CUDA(launch_fft(ctx->stream, src, n_cols, log_size, log_ext, 4, nullptr, 0, evals, eval_stride, ctx->tw, nullptr, 0));
    ctx->launches += log_size + log_ext <= 13 ? 1 : 2;
    CB_CATCH(ctx)
}

int cb_commit_lde(cb_ctx* ctx, int src_kind, const uint32_t* srcp, size_t src_stride, uint32_t first_col, int n_cols, int log_size,
                  int log_ext, uint32_t* coeffs_out, size_t coeff_stride, uint32_t* lde_out, size_t lde_stride) {
    CB_TRY(ctx)
    ctx->ensure_twiddles(log_size + log_ext);
    ColSrc src{src_kind, srcp, src_stride, first_col};
    int mode = 1 | 4 | (coeffs_out ? 2 : 0);
    DBuf<uint32_t> scratch;
    uint32_t* sc = coeffs_out;
    size_t sc_stride = coeff_stride;
    if (!coeffs_out && log_size + log_ext > 13) {
        scratch = DBuf<uint32_t>(ctx, (size_t)n_cols << log_size);
        sc = scratch.p;
        sc_stride = (size_t)1 << log_size;
    }
    CB_CUDA(launch_fft(ctx->stream, src, n_cols, log_size, log_ext, mode, coeffs_out, coeff_stride, lde_out, lde_stride, ctx->tw, sc,
                       sc_stride));
    ctx->launches += log_size + log_ext <= 13 ? 1 : 3;
    CB_CATCH(ctx)
}

int cb_lde_packed(cb_ctx* ctx, int src_kind, const uint32_t* src_words, int n_words, int log_size, uint32_t* tiles_out) {
    CB_TRY(ctx)
    if (src_kind != SRC_BITS && src_kind != SRC_BYTES && src_kind != SRC_M31)
        throw CbError("cb_lde_packed: src_kind must be 0 (4 M31 columns per job), 1 (bits) or 2 (bytes)");
    if (log_size < 1 || log_size > 24) throw CbError("cb_lde_packed: log_size out of range");
    ctx->ensure_twiddles(log_size + 1);
    const int cpj = src_kind == SRC_BITS ? 32 : 4;
    const int batch = n_words < MAX_FFT_JOBS ? n_words : MAX_FFT_JOBS;
    DBuf<uint32_t> scratch(ctx, fft_packed_scratch_words(src_kind, batch, log_size));
    std::vector<const uint32_t*> src(n_words);
    std::vector<uint32_t*> out(n_words);
    for (int w = 0; w < n_words; w++) {
        src[w] = src_words + (((size_t)w * (src_kind == SRC_M31 ? 4 : 1)) << log_size);
        out[w] = tiles_out + (((size_t)w * cpj) << (log_size + 1));
    }
    int nl = 0;
    StageHook hk = ctx->hook();
    ctx->pending_events.clear();
    CB_CUDA(launch_fft_packed(ctx->stream, src_kind, src.data(), out.data(), n_words, log_size, ctx->tw, scratch.p,
                              ctx->profile ? &hk : nullptr, &nl));
    ctx->launches += nl;
    ctx->collect_stages();
    ctx->sync();
    CB_CATCH(ctx)
}

extern int g_force_generic_fft;
int cb_debug_force_generic_fft(int on) {
    g_force_generic_fft = on;
    return 0;
}

int cb_comm_unique_id(uint8_t id_out[128]) {
    try {
        comm_unique_id(id_out);
    } catch (const std::exception&) {
        return 1;
    }
    return 0;
}
int cb_comm_init(cb_ctx* ctx, int rank, int world, const uint8_t id[128]) {
    CB_TRY(ctx)
    CB_CUDA(cudaSetDevice(ctx->device));
    if (world < 1 || (world & (world - 1))) throw CbError("cb_comm_init: world size must be a power of two");
    ctx->close_peers();
    ctx->p2p_state = 0;
    comm_destroy(ctx->comm);
    if (world > 1) comm_init(ctx->comm, rank, world, id);
    CB_CATCH(ctx)
}
int cb_comm_destroy(cb_ctx* ctx) {
    CB_TRY(ctx)
    ctx->sync();
    ctx->close_peers();
    ctx->p2p_state = 0;
    comm_destroy(ctx->comm);
    CB_CATCH(ctx)
}

int cb_set_max_cached_tiles(cb_ctx* ctx, int n_tiles) {
    if (!ctx) return 1;
    ctx->max_cached_tiles = n_tiles;
    return 0;
}

int cb_eval_at_point(cb_ctx* ctx, const uint32_t* coeffs, size_t stride, int n_cols, int log_size, const uint32_t pt[8],
                     uint32_t* out_host) {
    CB_TRY(ctx)
    const size_t N = (size_t)1 << log_size;
    QM31 x{{pt[0], pt[1], pt[2], pt[3]}}, y{{pt[4], pt[5], pt[6], pt[7]}};
    std::vector<QM31> maps(log_size > 0 ? log_size : 1);
    maps[0] = y;
    for (int j = 1; j < log_size; j++) { maps[j] = x; x = qsub(qmul_m(qmul(x, x), 2), qone()); }
    DBuf<uint32_t> basis(ctx, 4 * N), d_out(ctx, (size_t)n_cols * 4);
    CB_CUDA(launch_basis(ctx->stream, basis.p, N, log_size, maps.data()));
    CB_CUDA(launch_oods_dot(ctx->stream, coeffs, stride, n_cols, log_size, basis.p, N, d_out.p));
    ctx->launches += log_size + 1;
    CB_CUDA(cudaMemcpyAsync(out_host, d_out.p, (size_t)n_cols * 16, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    CB_CATCH(ctx)
}

int cb_merkle_build_leaves(cb_ctx* ctx, const uint32_t* const* base, const size_t* stride, const int* ncols, const int* logs,
                           int n_groups, int lifting_log, uint32_t* hashes_out) {
    CB_TRY(ctx)
    if (n_groups > MAX_LEAF_GROUPS) throw CbError("too many column groups");
    LeafGroups g{};
    g.n = n_groups;
    for (int i = 0; i < n_groups; i++) g.g[i] = {base[i], stride[i], ncols[i], logs[i]};
    CB_CUDA(launch_merkle_leaves(ctx->stream, g, lifting_log, nullptr, 0, 1, 1, hashes_out));
    ctx->launches++;
    CB_CATCH(ctx)
}

int cb_merkle_leaves_absorb(cb_ctx* ctx, const uint32_t* cols, size_t stride, int n_cols, int log_size, int lifting_log,
                            uint32_t* state, uint64_t bytes_before, int is_first, int is_final, uint32_t* hashes_out) {
    CB_TRY(ctx)
    if (!is_final && (n_cols % 16)) throw CbError("non-final absorb needs a multiple of 16 columns");
    LeafGroups g{};
    g.n = 1;
    g.g[0] = {cols, stride, n_cols, log_size};
    CB_CUDA(launch_merkle_leaves(ctx->stream, g, lifting_log, state, bytes_before, is_first, is_final, hashes_out));
    ctx->launches++;
    CB_CATCH(ctx)
}

int cb_merkle_next_layer(cb_ctx* ctx, const uint32_t* prev, uint32_t n_parents, uint32_t* out) {
    CB_TRY(ctx)
    CB_CUDA(launch_merkle_nodes(ctx->stream, prev, n_parents, out));
    ctx->launches++;
    CB_CATCH(ctx)
}

int cb_generate_secure_powers_rev(cb_ctx* ctx, const uint32_t a[4], int n, uint32_t* out_dev) {
    CB_TRY(ctx)
    CB_CUDA(launch_secure_powers_rev(ctx->stream, QM31{{a[0], a[1], a[2], a[3]}}, n, out_dev));
    ctx->launches++;
    CB_CATCH(ctx)
}

int cb_eval_constraints_chacha_stream(cb_ctx* ctx, const uint32_t* lde, size_t stride, int eval_log, int trace_log,
                                      const uint32_t* apr, uint32_t* accum, size_t accum_stride, int accumulate) {
    CB_TRY(ctx)
    const int ext = eval_log - trace_log;
    std::vector<uint32_t> den((size_t)1 << ext);
    for (uint32_t i = 0; i < den.size(); i++) {
        uint32_t row = i << trace_log;
        host::Pt p = host::index_to_point(host::canonic_index_at(eval_log, host::bit_reverse(row, eval_log)));
        den[i] = inv(host::coset_vanishing_m31(trace_log, p));
    }
    DBuf<uint32_t> d_den(ctx, den.size());
    CB_CUDA(cudaMemcpyAsync(d_den.p, den.data(), den.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->sync();
    CB_CUDA(launch_chacha_constraints(ctx->stream, lde, stride, eval_log, trace_log, apr, d_den.p, accum, accum_stride, accumulate));
    ctx->launches++;
    ctx->sync();
    CB_CATCH(ctx)
}

int cb_accumulate_quotients(cb_ctx* ctx, const uint32_t* cols, size_t stride, int n_cols, int domain_log, const uint32_t* sampled,
                            const uint32_t pt[8], const uint32_t rcw[4], uint32_t* out, size_t out_stride) {
    CB_TRY(ctx)
    ctx->ensure_twiddles(domain_log);
    QM31 zx{{pt[0], pt[1], pt[2], pt[3]}}, zy{{pt[4], pt[5], pt[6], pt[7]}}, rc{{rcw[0], rcw[1], rcw[2], rcw[3]}};
    std::vector<uint32_t> coefs((size_t)n_cols * 4);
    QM31 alpha = qone(), lin_a = qzero(), lin_b = qzero();
    const QM31 c = qsub(qconj(zy), zy);
    for (int j = 0; j < n_cols; j++) {
        QM31 v{{sampled[4 * j], sampled[4 * j + 1], sampled[4 * j + 2], sampled[4 * j + 3]}};
        QM31 a = qsub(qconj(v), v);
        QM31 b = qsub(qmul(v, c), qmul(a, zy));
        lin_a = qadd(lin_a, qmul(alpha, a));
        lin_b = qadd(lin_b, qmul(alpha, b));
        QM31 ac = qmul(alpha, c);
        for (int k = 0; k < 4; k++) coefs[(size_t)j * 4 + k] = ac.v[k];
        alpha = qmul(alpha, rc);
    }
    DBuf<uint32_t> d_coefs(ctx, coefs.size());
    CB_CUDA(cudaMemcpyAsync(d_coefs.p, coefs.data(), coefs.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    QuotBatch qb{};
    qb.prx = {zx.v[0], zx.v[1]}; qb.pix = {zx.v[2], zx.v[3]};
    qb.pry = {zy.v[0], zy.v[1]}; qb.piy = {zy.v[2], zy.v[3]};
    qb.lin_a = lin_a; qb.lin_b = lin_b; qb.batch_coeff = qzero();
    qb.coefs = d_coefs.p; qb.col_idx = nullptr; qb.n_cols = n_cols;
    DBuf<QuotBatch> d_qb(ctx, 1);
    CB_CUDA(cudaMemcpyAsync(d_qb.p, &qb, sizeof(qb), cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(launch_quotients(ctx->stream, cols, stride, n_cols, nullptr, 0, d_qb.p, 1, domain_log, ctx->tw, out, out_stride));
    ctx->launches++;
    ctx->sync();
    CB_CATCH(ctx)
}

int cb_fold_circle_into_line(cb_ctx* ctx, const uint32_t* src, size_t src_stride, int src_log, const uint32_t a[4], uint32_t* dst,
                             size_t dst_stride, int dst_is_zero) {
    CB_TRY(ctx)
    ctx->ensure_twiddles(src_log);
    CB_CUDA(launch_fold_circle(ctx->stream, src, src_stride, src_log, QM31{{a[0], a[1], a[2], a[3]}}, ctx->tw, dst, dst_stride,
                               dst_is_zero));
    ctx->launches++;
    CB_CATCH(ctx)
}

int cb_fold_line(cb_ctx* ctx, const uint32_t* src, size_t src_stride, int src_log, const uint32_t a[4], uint32_t* dst,
                 size_t dst_stride) {
    CB_TRY(ctx)
    ctx->ensure_twiddles(src_log + 1);
    CB_CUDA(launch_fold_line(ctx->stream, src, src_stride, src_log, QM31{{a[0], a[1], a[2], a[3]}}, ctx->tw, dst, dst_stride));
    ctx->launches++;
    CB_CATCH(ctx)
}

int cb_grind_blake2s(cb_ctx* ctx, const uint8_t pd[32], uint32_t pow_bits, uint64_t* nonce_out) {
    CB_TRY(ctx)
    DBuf<uint32_t> d_pd(ctx, 8);
    DBuf<unsigned long long> d_best(ctx, 1);
    unsigned long long best = ~0ull;
    CB_CUDA(cudaMemcpyAsync(d_pd.p, pd, 32, cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(cudaMemcpyAsync(d_best.p, &best, 8, cudaMemcpyHostToDevice, ctx->stream));
    for (uint64_t base = 0;; base += (1ull << 20)) {
        CB_CUDA(launch_grind(ctx->stream, d_pd.p, pow_bits, base, 1ull << 20, d_best.p));
        ctx->launches++;
        CB_CUDA(cudaMemcpyAsync(&best, d_best.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->sync();
        if (best != ~0ull) break;
    }
    *nonce_out = best;
    CB_CATCH(ctx)
}

int cb_gather_rows(cb_ctx* ctx, const uint32_t* cols, size_t stride, int n_cols, const uint32_t* rows, int n_rows, uint32_t* out) {
    CB_TRY(ctx)
    DBuf<uint32_t> d_rows(ctx, n_rows), d_out(ctx, (size_t)n_cols * n_rows);
    CB_CUDA(cudaMemcpyAsync(d_rows.p, rows, (size_t)n_rows * 4, cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(launch_gather_rows(ctx->stream, cols, stride, n_cols, d_rows.p, n_rows, d_out.p));
    ctx->launches++;
    CB_CUDA(cudaMemcpyAsync(out, d_out.p, (size_t)n_cols * n_rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    CB_CATCH(ctx)
}

int cb_gen_trace_chacha_stream(cb_ctx* ctx, const uint8_t key[32], const uint8_t nonce[12], uint32_t counter, const uint8_t* pt,
                               const uint8_t* ct, uint32_t n_blocks, int log_size, uint32_t* words_out, size_t stride, int* valid) {
    CB_TRY(ctx)
    uint32_t kw[8], nw[3];
    for (int i = 0; i < 8; i++) kw[i] = host::load_le32(key + 4 * i);
    for (int i = 0; i < 3; i++) nw[i] = host::load_le32(nonce + 4 * i);
    size_t len = (size_t)n_blocks * 64;
    DBuf<uint32_t> d_pt(ctx, len / 4), d_ct(ctx, len / 4);
    DBuf<int> d_inv(ctx, 1);
    CB_CUDA(cudaMemcpyAsync(d_pt.p, pt, len, cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(cudaMemcpyAsync(d_ct.p, ct, len, cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(cudaMemsetAsync(d_inv.p, 0, 4, ctx->stream));
    uint32_t rows_needed = (n_blocks + 15) / 16;
    CB_CUDA(launch_chacha_witness(ctx->stream, kw, nw, counter, n_blocks, rows_needed * 16, d_pt.p, d_ct.p, log_size, words_out, stride,
                                  d_inv.p));
    ctx->launches++;
    int invalid = 0;
    CB_CUDA(cudaMemcpyAsync(&invalid, d_inv.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    *valid = !invalid;
    CB_CATCH(ctx)
}

int cb_set_profile(cb_ctx* ctx, int enable) {
    if (!ctx) return 1;
    ctx->profile = enable != 0;
    return 0;
}

const char* cb_stage_times(cb_ctx* ctx) {
    static thread_local std::string s;
    s.clear();
    if (!ctx) return "";
    for (auto& st : ctx->stages) s += st.name + "=" + std::to_string(st.ms) + ";";
    return s.c_str();
}

const char* cb_host_times(cb_ctx* ctx) {
    static thread_local std::string s;
    s.clear();
    if (!ctx) return "";
    for (size_t i = 0; i + 1 < ctx->host_marks.size(); i++)
        s += ctx->host_marks[i].first + "=" + std::to_string((ctx->host_marks[i + 1].second - ctx->host_marks[i].second) * 1e3) + ";";
    return s.c_str();
}

const char* cb_counters(cb_ctx* ctx) {
    static thread_local std::string s;
    s.clear();
    if (!ctx) return "";
    s = "fft_words=" + std::to_string(ctx->fft_words) + ";fft_words_half=" + std::to_string(ctx->fft_words_half) + ";cached_tiles=" + std::to_string(ctx->cached_tiles) +
        ";transient_tiles=" + std::to_string(ctx->transient_tiles) + ";hash_wait_us=" + std::to_string(ctx->hash_wait_us) + ";peer_windows=" +
        std::to_string(ctx->last_p2p ? 1 : 0) + ";";
    return s.c_str();
}

// ---------------------------------------------------------------------------------------------- product level
// Calls made with ctx == NULL share ONE process-wide context (stream, arena, error string, counters): they are serialised by
// holding its mutex for the whole call.  Callers that want concurrency create their own contexts (one per thread).
struct CtxUse {
    cb_ctx* ctx = nullptr;
    std::unique_lock<std::mutex> lk;
    std::string err;
    explicit CtxUse(cb_ctx* given) : ctx(given) {
        if (ctx) return;
        static std::mutex mu;
        static cb_ctx* g = nullptr;
        lk = std::unique_lock<std::mutex>(mu);
        if (!g) {
            int rc = cb_init(0, &g);
            if (rc) {
                err = "no usable CUDA device (cb_init=" + std::to_string(rc) + "); this backend has no CPU fallback";
                g = nullptr;
            }
        }
        ctx = g;
    }
};

// copies the proof into a malloc'd buffer owned by the caller (s2c_free)
static void give_proof(const std::vector<uint8_t>& proof, uint8_t** proof_out, size_t* proof_len) {
    uint8_t* p = (uint8_t*)malloc(proof.size() ? proof.size() : 1);
    if (!p) throw CbError("out of host memory for the proof buffer");
    memcpy(p, proof.data(), proof.size());
    *proof_out = p;
    *proof_len = proof.size();
}

// wasm_api.rs:83-86 / :675-678: the last block's counter must fit a u32
static void check_counter(uint32_t counter, size_t num_blocks) {
    if (num_blocks > 1 && (uint64_t)counter + num_blocks - 1 > 0xFFFFFFFFull)
        throw CbError("Counter overflow: counter " + std::to_string(counter) + " + " + std::to_string(num_blocks) +
                      " blocks would exceed u32::MAX");
}

static int ret_json(const std::string& s, char** out, size_t* len) {
    char* p = (char*)malloc(s.size() + 1);
    if (!p) return 1;
    memcpy(p, s.c_str(), s.size() + 1);
    *out = p;
    if (len) *len = s.size();
    return 0;
}

static std::string json_escape(const std::string& s) {
    std::string o;
    for (char c : s) {
        if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back(c); }
        else if (c == '\n') o += "\\n";
        else o.push_back(c);
    }
    return o;
}
static std::string json_error(const std::string& m) { return "{\"error\":\"" + json_escape(m) + "\"}"; }

int s2c_prove_chacha20_raw(cb_ctx* ctx, const uint8_t key[32], const uint8_t nonce[12], uint32_t counter, const uint8_t* pt,
                           const uint8_t* ct, size_t len, uint8_t** proof_out, size_t* proof_len) {
    CtxUse use(ctx);
    ctx = use.ctx;
    if (!ctx) return 2;
    CB_TRY(ctx)
    CB_CUDA(cudaSetDevice(ctx->device));
    if (len == 0 || len % 64) throw CbError("Plaintext must be non-empty multiple of 64 bytes, got " + std::to_string(len));
    check_counter(counter, len / 64);
    std::vector<uint8_t> proof;
    std::string e = prove_chacha20(ctx, key, nonce, counter, pt, ct, len, proof);
    if (!e.empty()) throw CbError(e);
    give_proof(proof, proof_out, proof_len);
    CB_CATCH(ctx)
}

int s2c_prove_chacha20_dev(cb_ctx* ctx, const uint8_t key[32], const uint8_t nonce[12], uint32_t counter, const void* pt_dev,
                           const void* ct_dev, size_t len, const uint8_t pt_hash[32], const uint8_t ct_hash[32],
                           uint8_t** proof_out, size_t* proof_len) {
    if (!ctx) return 2;
    CB_TRY(ctx)
    CB_CUDA(cudaSetDevice(ctx->device));
    if (len == 0 || len % 64) throw CbError("Plaintext must be non-empty multiple of 64 bytes, got " + std::to_string(len));
    ProveOptions opt;
    opt.pt_dev = (const uint32_t*)pt_dev;
    opt.ct_dev = (const uint32_t*)ct_dev;
    if (pt_hash && ct_hash) {  // both or neither: without them the library reads the buffers back and hashes them itself
        opt.pt_hash = pt_hash;
        opt.ct_hash = ct_hash;
    }
    std::vector<uint8_t> proof;
    check_counter(counter, len / 64);
    std::string e = prove_chacha20(ctx, key, nonce, counter, nullptr, nullptr, len, proof, opt);
    if (!e.empty()) throw CbError(e);
    give_proof(proof, proof_out, proof_len);
    CB_CATCH(ctx)
}

// Validation + proving shared by generate_chacha20_proof and prove_chacha20_encrypt (wasm_api.rs:68-86 == :475-493).
// Returns -1 with the proof bytes in `proof`, or the value the caller must return (the error JSON is already written).
static int chacha_prove_checked(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                                uint32_t counter, const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len,
                                std::vector<uint8_t>& proof, size_t& num_blocks, char** json_out, size_t* json_len) {
    if (key_len != 32) return ret_json(json_error("Key must be 32 bytes, got " + std::to_string(key_len)), json_out, json_len);
    if (nonce_len != 12) return ret_json(json_error("Nonce must be 12 bytes, got " + std::to_string(nonce_len)), json_out, json_len);
    if (pt_len == 0 || pt_len % 64 != 0)
        return ret_json(json_error("Plaintext must be non-empty multiple of 64 bytes, got " + std::to_string(pt_len)), json_out, json_len);
    if (ct_len != pt_len)
        return ret_json(json_error("Ciphertext must be same length as plaintext, got " + std::to_string(ct_len) + " vs " +
                                   std::to_string(pt_len)), json_out, json_len);
    num_blocks = pt_len / 64;
    if (num_blocks > 1 && (uint64_t)counter + num_blocks - 1 > 0xFFFFFFFFull)
        return ret_json(json_error("Counter overflow: counter " + std::to_string(counter) + " + " + std::to_string(num_blocks) +
                                   " blocks would exceed u32::MAX"), json_out, json_len);
    CtxUse use(ctx);
    ctx = use.ctx;
    if (!ctx) return ret_json(json_error(use.err), json_out, json_len) ? 1 : 2;
    std::string e;
    try {
        CB_CUDA(cudaSetDevice(ctx->device));
        e = prove_chacha20(ctx, key, nonce, counter, pt, ct, pt_len, proof);
    } catch (const std::exception& ex) {
        ctx->err = ex.what();
        ret_json(json_error(std::string("backend failure: ") + ex.what()), json_out, json_len);
        return 1;
    }
    if (!e.empty()) return ret_json(json_error(e), json_out, json_len);
    return -1;
}

// StarkProof::size_estimate() of the reference is reported as proof_size_bytes; see estimate in prove driver notes.
int s2c_generate_chacha20_proof(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                                uint32_t counter, const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, char** json_out,
                                size_t* json_len) {
    std::vector<uint8_t> proof;
    size_t num_blocks = 0;
    const int rc = chacha_prove_checked(ctx, key, key_len, nonce, nonce_len, counter, pt, pt_len, ct, ct_len, proof, num_blocks, json_out,
                                        json_len);
    if (rc != -1) return rc;
    try {  // nothing may unwind across the C ABI
        std::string b64 = host::base64_encode(proof.data(), proof.size());
        std::string js = "{\"algorithm\":\"chacha20\",\"blocks\":" + std::to_string(num_blocks) + ",\"proof\":\"" + b64 +
                         "\",\"proof_size_bytes\":" + std::to_string(stark_proof_size_estimate(proof.data(), proof.size(), 84)) +
                         ",\"success\":true}";
        return ret_json(js, json_out, json_len);
    } catch (const std::exception& ex) {
        return ret_json(json_error(std::string("backend failure: ") + ex.what()), json_out, json_len) ? 1 : 1;
    }
}

// Validation + proving shared by generate_aes*_ctr_proof and prove_aes*_ctr_encrypt; same contract as chacha_prove_checked.
static int aes_prove_checked(cb_ctx* ctx, int key_bytes, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                             uint32_t counter, const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len,
                             std::vector<uint8_t>& proof, size_t& num_blocks, char** json_out, size_t* json_len) {
    // validation order and messages: wasm_api.rs:660-678 / 784-802
    if (key_len != (size_t)key_bytes)
        return ret_json(json_error("Key must be " + std::to_string(key_bytes) + " bytes, got " + std::to_string(key_len)), json_out, json_len);
    if (nonce_len != 12) return ret_json(json_error("Nonce must be 12 bytes, got " + std::to_string(nonce_len)), json_out, json_len);
    if (pt_len == 0 || pt_len % 16 != 0)
        return ret_json(json_error("Plaintext must be non-empty multiple of 16 bytes, got " + std::to_string(pt_len)), json_out, json_len);
    if (ct_len != pt_len)
        return ret_json(json_error("Ciphertext must be same length as plaintext, got " + std::to_string(ct_len) + " vs " +
                                   std::to_string(pt_len)), json_out, json_len);
    num_blocks = pt_len / 16;
    if (num_blocks > 1 && (uint64_t)counter + num_blocks - 1 > 0xFFFFFFFFull)
        return ret_json(json_error("Counter overflow: counter " + std::to_string(counter) + " + " + std::to_string(num_blocks) +
                                   " blocks would exceed u32::MAX"), json_out, json_len);
    CtxUse use(ctx);
    ctx = use.ctx;
    if (!ctx) return ret_json(json_error(use.err), json_out, json_len) ? 1 : 2;
    std::string e;
    try {
        CB_CUDA(cudaSetDevice(ctx->device));
        e = prove_aes_ctr(ctx, key_bytes, key, nonce, counter, pt, ct, pt_len, proof);
    } catch (const std::exception& ex) {
        ctx->err = ex.what();
        ret_json(json_error(std::string("backend failure: ") + ex.what()), json_out, json_len);
        return 1;
    }
    if (!e.empty()) return ret_json(json_error(e), json_out, json_len);
    return -1;
}

static int aes_generate(cb_ctx* ctx, int key_bytes, const char* algorithm, const uint8_t* key, size_t key_len, const uint8_t* nonce,
                        size_t nonce_len, uint32_t counter, const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len,
                        char** json_out, size_t* json_len) {
    std::vector<uint8_t> proof;
    size_t num_blocks = 0;
    const int rc = aes_prove_checked(ctx, key_bytes, key, key_len, nonce, nonce_len, counter, pt, pt_len, ct, ct_len, proof, num_blocks,
                                     json_out, json_len);
    if (rc != -1) return rc;
    try {
        std::string b64 = host::base64_encode(proof.data(), proof.size());
        std::string js = std::string("{\"algorithm\":\"") + algorithm + "\",\"blocks\":" + std::to_string(num_blocks) + ",\"proof\":\"" +
                         b64 + "\",\"proof_size_bytes\":" + std::to_string(stark_proof_size_estimate(proof.data(), proof.size(), 136)) +
                         ",\"success\":true}";
        return ret_json(js, json_out, json_len);
    } catch (const std::exception& ex) {
        return ret_json(json_error(std::string("backend failure: ") + ex.what()), json_out, json_len) ? 1 : 1;
    }
}

int s2c_generate_aes128_ctr_proof(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                                  uint32_t counter, const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, char** json_out,
                                  size_t* json_len) {
    return aes_generate(ctx, 16, "aes128-ctr", key, key_len, nonce, nonce_len, counter, pt, pt_len, ct, ct_len, json_out, json_len);
}
int s2c_generate_aes256_ctr_proof(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len,
                                  uint32_t counter, const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, char** json_out,
                                  size_t* json_len) {
    return aes_generate(ctx, 32, "aes256-ctr", key, key_len, nonce, nonce_len, counter, pt, pt_len, ct, ct_len, json_out, json_len);
}

int s2c_prove_aes_ctr_raw(cb_ctx* ctx, int key_len, const uint8_t* key, const uint8_t nonce[12], uint32_t counter, const uint8_t* pt,
                          const uint8_t* ct, size_t len, uint8_t** proof_out, size_t* proof_len) {
    CtxUse use(ctx);
    ctx = use.ctx;
    if (!ctx) return 2;
    CB_TRY(ctx)
    CB_CUDA(cudaSetDevice(ctx->device));
    if (key_len != 16 && key_len != 32) throw CbError("key_len must be 16 or 32");
    if (len == 0 || len % 16) throw CbError("Plaintext must be non-empty multiple of 16 bytes, got " + std::to_string(len));
    check_counter(counter, len / 16);
    std::vector<uint8_t> proof;
    std::string e = prove_aes_ctr(ctx, key_len, key, nonce, counter, pt, ct, len, proof);
    if (!e.empty()) throw CbError(e);
    give_proof(proof, proof_out, proof_len);
    CB_CATCH(ctx)
}

// native ChaCha20 block function (chacha/block.rs:95): out = keystream words of one block
static void chacha_block_words(const uint32_t key[8], const uint32_t nonce[3], uint32_t counter, uint32_t out[16]) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u}, v[16];
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = counter;
    for (int i = 0; i < 3; i++) s[13 + i] = nonce[i];
    memcpy(v, s, sizeof v);
    auto rotl = [](uint32_t x, int r) { return (x << r) | (x >> (32 - r)); };
    auto qr = [&](int a, int b, int c, int d) {
        v[a] += v[b]; v[d] = rotl(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = rotl(v[b] ^ v[c], 12);
        v[a] += v[b]; v[d] = rotl(v[d] ^ v[a], 8);  v[c] += v[d]; v[b] = rotl(v[b] ^ v[c], 7);
    };
    for (int r = 0; r < 10; r++) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) out[i] = v[i] + s[i];
}

int s2c_debug_blake2s(const uint8_t* data, size_t len, uint8_t out[32]) {
    if (!out || (!data && len)) return 1;
    const host::Hash32 h = host::blake2s_bytes(data, len);
    memcpy(out, h.b, 32);
    return 0;
}

int s2c_debug_chacha20_keystream(const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter,
                                 char** json_out, size_t* json_len) {
    // wasm_api.rs:953-990: native block function, hex of the 64 keystream bytes
    if (key_len != 32 || nonce_len != 12) return ret_json(json_error("Invalid key or nonce length"), json_out, json_len);
    uint32_t kw[8], nw[3], ks[16];
    for (int i = 0; i < 8; i++) kw[i] = host::load_le32(key + 4 * i);
    for (int i = 0; i < 3; i++) nw[i] = host::load_le32(nonce + 4 * i);
    chacha_block_words(kw, nw, counter, ks);
    static const char* H = "0123456789abcdef";
    std::string hex;
    for (int i = 0; i < 16; i++) {
        for (int b = 0; b < 4; b++) { uint8_t x = (uint8_t)(ks[i] >> (8 * b)); hex.push_back(H[x >> 4]); hex.push_back(H[x & 15]); }
    }
    std::string js = "{\"counter\":" + std::to_string(counter) + ",\"key_len\":32,\"keystream_hex\":\"" + hex + "\",\"nonce_len\":12}";
    return ret_json(js, json_out, json_len);
}

// prove_stream::<Blake2sMerkleChannel>(log_size, PcsConfig::default()) of the reference (air_stream.rs:237-289): the test-data
// generator behind its `bench_stream` / round-trip tests -- key 00..1f, witness nonce words [0, 0x4a, 0], block r (0-based) has
// counter r + 1 and plaintext word w = r * 16 + w, ciphertext = its ChaCha20 encryption; the STATEMENT binds an all-zero nonce,
// counter 1 and the hashes of empty strings.  Lets anyone with cargo compare bytes with the reference's own generator
// (integration/rust/examples/prove_stream_dump.rs).
int s2c_prove_chacha20_stream_testdata(cb_ctx* ctx, int log_size, uint8_t** proof_out, size_t* proof_len) {
    CtxUse use(ctx);
    ctx = use.ctx;
    if (!ctx) return 2;
    CB_TRY(ctx)
    CB_CUDA(cudaSetDevice(ctx->device));
    if (log_size < 4 || log_size > 24) throw CbError("log_size must be in [4, 24]");
    const size_t n_blocks = (size_t)1 << log_size;
    uint32_t kw[8], nw[3] = {0, 0x4a, 0};
    for (int i = 0; i < 8; i++) kw[i] = 0x03020100u + 0x04040404u * (uint32_t)i;
    std::vector<uint32_t> pt(n_blocks * 16), ct(n_blocks * 16);
    for (size_t r = 0; r < n_blocks; r++) {
        uint32_t ks[16];
        chacha_block_words(kw, nw, (uint32_t)(r + 1), ks);
        for (int w = 0; w < 16; w++) {
            pt[r * 16 + w] = (uint32_t)(r * 16 + w);
            ct[r * 16 + w] = pt[r * 16 + w] ^ ks[w];
        }
    }
    uint8_t key[32], nonce[12];
    static const uint8_t zero_nonce[12] = {0};
    memcpy(key, kw, 32);
    memcpy(nonce, nw, 12);
    ProveOptions opt;
    opt.empty_public_hashes = true;
    opt.stmt_nonce = zero_nonce;
    std::vector<uint8_t> proof;
    std::string e = prove_chacha20(ctx, key, nonce, 1, (const uint8_t*)pt.data(), (const uint8_t*)ct.data(), n_blocks * 64, proof, opt);
    if (!e.empty()) throw CbError(e);
    give_proof(proof, proof_out, proof_len);
    CB_CATCH(ctx)
}

// The reference's block-AIR entry points (chacha/bitwise/air.rs: prove_bitwise / verify_bitwise; reached only from its own tests
// and `bench_bitwise`, not from the product API): the trace is generated from log_size alone - key bytes 00..1f, nonce
// 00 00 00 09 00 00 00 4a 00 00 00 00, block counter = row index.  Proof bytes = u32 log_size || bincode(StarkProof) (the
// reference does not serialise BitwiseProof; the layout follows its other proof containers).
int s2c_prove_chacha20_block(cb_ctx* ctx, int log_size, uint8_t** proof_out, size_t* proof_len) {
    CtxUse use(ctx);
    ctx = use.ctx;
    if (!ctx) return 2;
    CB_TRY(ctx)
    CB_CUDA(cudaSetDevice(ctx->device));
    if (log_size < 4 || log_size > 24) throw CbError("log_size must be in [4, 24]");
    uint32_t kw[8];
    for (int i = 0; i < 8; i++) kw[i] = 0x03020100u + 0x04040404u * (uint32_t)i;
    const uint32_t nw[3] = {0x09000000u, 0x4a000000u, 0};
    uint8_t key[32], nonce[12];
    memcpy(key, kw, 32);
    memcpy(nonce, nw, 12);
    ProveOptions opt;
    opt.block_air = true;
    std::vector<uint8_t> proof;
    std::string e = prove_chacha20(ctx, key, nonce, 0, nullptr, nullptr, (size_t)64 << log_size, proof, opt);
    if (!e.empty()) throw CbError(e);
    give_proof(proof, proof_out, proof_len);
    CB_CATCH(ctx)
}
// 0: verifies; 1: does not (err_out = the reference's rendering of its VerificationError, or "Invalid proof format: ..")
int s2c_verify_chacha20_block(const uint8_t* proof, size_t proof_len, char** err_out, size_t* err_len) {
    std::string e;
    try {
        e = verify_chacha20_block(proof, proof_len);
    } catch (const std::exception& ex) {
        e = std::string("Invalid proof format: ") + ex.what();
    }
    if (err_out) ret_json(e, err_out, err_len);
    return e.empty() ? 0 : 1;
}

// AES-128 block AIR of the reference's tests (aes/lookup/air.rs:139-305 prove_aes_lookup / verify_aes_lookup; constraints.rs,
// gen.rs): the CTR AIR without the counter block, the plaintext / ciphertext columns and the final xor.  The trace is the
// reference generator's: key 00..0f, input byte b of row r = (r + b) & 0xFF.  Proof bytes = u32 log_size || stmt1 (two QM31
// claimed sums, two usize column counts) || bincode(StarkProof) - the field order of AESLookupProof.
int s2c_prove_aes128_block(cb_ctx* ctx, int log_size, uint8_t** proof_out, size_t* proof_len) {
    CtxUse use(ctx);
    ctx = use.ctx;
    if (!ctx) return 2;
    CB_TRY(ctx)
    CB_CUDA(cudaSetDevice(ctx->device));
    if (log_size < 8 || log_size > 19) throw CbError("log_size must be in [8, 19]");
    const size_t n = (size_t)1 << log_size;
    uint8_t key[16];
    for (int i = 0; i < 16; i++) key[i] = (uint8_t)i;
    std::vector<uint8_t> blocks(n * 16);
    for (size_t r = 0; r < n; r++)
        for (int b = 0; b < 16; b++) blocks[r * 16 + b] = (uint8_t)((r + b) & 0xFF);
    const uint8_t nonce[12] = {0};
    std::vector<uint8_t> proof;
    std::string e = prove_aes_ctr(ctx, 16, key, nonce, 0, blocks.data(), nullptr, blocks.size(), proof, true);
    if (!e.empty()) throw CbError(e);
    give_proof(proof, proof_out, proof_len);
    CB_CATCH(ctx)
}
int s2c_verify_aes128_block(const uint8_t* proof, size_t proof_len, char** err_out, size_t* err_len) {
    std::string e;
    try {
        e = verify_aes128_block(proof, proof_len);
    } catch (const std::exception& ex) {
        e = std::string("Invalid proof format: ") + ex.what();
    }
    if (err_out) ret_json(e, err_out, err_len);
    return e.empty() ? 0 : 1;
}

int s2c_get_circuits_info(char** json_out, size_t* json_len) {
    // wasm_api.rs:993-1008 (values confirmed against the reference binary's own get_circuits_info())
    return ret_json("{\"aes128_ctr\":{\"block_bytes\":16,\"cols\":24480,\"constraints\":34464,\"key_bytes\":16},"
                    "\"aes256_ctr\":{\"block_bytes\":16,\"cols\":34784,\"constraints\":49024,\"key_bytes\":32},"
                    "\"chacha20\":{\"block_bytes\":64,\"cols\":33280,\"constraints\":54784,\"key_bytes\":32}}",
                    json_out, json_len);
}

// ---------------------------------------------------------------------------------------------- verify / prove+verify
// verify_chacha20_proof / verify_aes_ctr_proof (wasm_api.rs:609-648, :904-946) and prove_*_encrypt (wasm_api.rs:61-188,
// :210-330, :343-463).  Verification is host work (verify.cu); these entry points need no CUDA context.
static const size_t MAX_PROOF_B64_LEN = 8u * 1024 * 1024;  // wasm_api.rs:27

static std::string verify_dispatch(bool aes, const uint8_t* proof, size_t len, const uint8_t* nonce, uint32_t counter, const uint8_t* pt,
                                   size_t pt_len, const uint8_t* ct, size_t ct_len, std::string& algorithm) {
    static const uint8_t none = 0;
    if (!pt) pt = &none;
    if (!ct) ct = &none;
    if (!aes) {
        algorithm = "chacha20";
        return verify_chacha20(proof, len, nonce, counter, pt, pt_len, ct, ct_len);
    }
    int ks = 0;
    std::string e = verify_aes_ctr(proof, len, nonce, counter, pt, pt_len, ct, ct_len, &ks);
    algorithm = ks == 0 ? "aes128-ctr" : "aes256-ctr";
    return e;
}

static int verify_json(bool aes, const char* proof_b64, size_t b64_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter,
                       const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, char** json_out, size_t* json_len) {
    if (b64_len > MAX_PROOF_B64_LEN) return ret_json(json_error("Proof payload too large"), json_out, json_len);
    if (nonce_len != 12) return ret_json(json_error("Nonce must be 12 bytes, got " + std::to_string(nonce_len)), json_out, json_len);
    std::vector<uint8_t> bytes;
    std::string berr = host::base64_decode(proof_b64, b64_len, bytes);
    if (!berr.empty()) return ret_json(json_error("Invalid base64: " + berr), json_out, json_len);
    std::string algorithm, e;
    try {
        e = verify_dispatch(aes, bytes.data(), bytes.size(), nonce, counter, pt, pt_len, ct, ct_len, algorithm);
    } catch (const VerifyFormatError& ex) {
        return ret_json(json_error(std::string("Invalid proof format: ") + ex.what()), json_out, json_len);
    } catch (const std::exception& ex) {
        return ret_json("{\"error\":\"" + json_escape(ex.what()) + "\",\"valid\":false}", json_out, json_len);
    }
    if (e.empty()) return ret_json("{\"algorithm\":\"" + algorithm + "\",\"valid\":true}", json_out, json_len);
    return ret_json("{\"error\":\"" + json_escape(e) + "\",\"valid\":false}", json_out, json_len);
}

int s2c_verify_chacha20_proof(const char* proof_b64, size_t proof_b64_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter,
                              const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, char** json_out, size_t* json_len) {
    return verify_json(false, proof_b64, proof_b64_len, nonce, nonce_len, counter, pt, pt_len, ct, ct_len, json_out, json_len);
}

int s2c_verify_aes_ctr_proof(const char* proof_b64, size_t proof_b64_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter,
                             const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, char** json_out, size_t* json_len) {
    return verify_json(true, proof_b64, proof_b64_len, nonce, nonce_len, counter, pt, pt_len, ct, ct_len, json_out, json_len);
}

// raw form: bincode bytes in, 0 = valid, 1 = rejected / malformed; *error_out (s2c_free) receives the reference's error rendering
static int verify_raw(bool aes, const uint8_t* proof, size_t len, const uint8_t nonce[12], uint32_t counter, const uint8_t* pt,
                      size_t pt_len, const uint8_t* ct, size_t ct_len, char** error_out) {
    std::string algorithm, e;
    try {
        e = verify_dispatch(aes, proof, len, nonce, counter, pt, pt_len, ct, ct_len, algorithm);
    } catch (const VerifyFormatError& ex) {
        e = std::string("Invalid proof format: ") + ex.what();
    } catch (const std::exception& ex) {
        e = ex.what();
    }
    if (error_out) {
        *error_out = nullptr;
        if (!e.empty()) ret_json(e, error_out, nullptr);
    }
    return e.empty() ? 0 : 1;
}
int s2c_verify_chacha20_raw(const uint8_t* proof, size_t proof_len, const uint8_t nonce[12], uint32_t counter, const uint8_t* pt,
                            size_t pt_len, const uint8_t* ct, size_t ct_len, char** error_out) {
    return verify_raw(false, proof, proof_len, nonce, counter, pt, pt_len, ct, ct_len, error_out);
}
int s2c_verify_aes_ctr_raw(const uint8_t* proof, size_t proof_len, const uint8_t nonce[12], uint32_t counter, const uint8_t* pt,
                           size_t pt_len, const uint8_t* ct, size_t ct_len, char** error_out) {
    return verify_raw(true, proof, proof_len, nonce, counter, pt, pt_len, ct, ct_len, error_out);
}

static int encrypt_finish(bool aes, const char* algorithm, const std::vector<uint8_t>& proof, size_t num_blocks, const uint8_t* nonce,
                          uint32_t counter, const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, char** json_out,
                          size_t* json_len) {
    std::string alg, e;
    try {
        e = verify_dispatch(aes, proof.data(), proof.size(), nonce, counter, pt, pt_len, ct, ct_len, alg);
    } catch (const std::exception& ex) {
        e = ex.what();
    }
    if (!e.empty()) return ret_json(json_error("Verification failed: " + e), json_out, json_len);
    return ret_json(std::string("{\"algorithm\":\"") + algorithm + "\",\"blocks\":" + std::to_string(num_blocks) + ",\"success\":true}",
                    json_out, json_len);
}

int s2c_prove_chacha20_encrypt(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter,
                               const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, char** json_out, size_t* json_len) {
    std::vector<uint8_t> proof;
    size_t num_blocks = 0;
    const int rc = chacha_prove_checked(ctx, key, key_len, nonce, nonce_len, counter, pt, pt_len, ct, ct_len, proof, num_blocks, json_out,
                                        json_len);
    if (rc != -1) return rc;
    return encrypt_finish(false, "chacha20", proof, num_blocks, nonce, counter, pt, pt_len, ct, ct_len, json_out, json_len);
}
int s2c_prove_aes128_ctr_encrypt(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter,
                                 const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, char** json_out, size_t* json_len) {
    std::vector<uint8_t> proof;
    size_t num_blocks = 0;
    const int rc = aes_prove_checked(ctx, 16, key, key_len, nonce, nonce_len, counter, pt, pt_len, ct, ct_len, proof, num_blocks, json_out,
                                     json_len);
    if (rc != -1) return rc;
    return encrypt_finish(true, "aes128-ctr", proof, num_blocks, nonce, counter, pt, pt_len, ct, ct_len, json_out, json_len);
}
int s2c_prove_aes256_ctr_encrypt(cb_ctx* ctx, const uint8_t* key, size_t key_len, const uint8_t* nonce, size_t nonce_len, uint32_t counter,
                                 const uint8_t* pt, size_t pt_len, const uint8_t* ct, size_t ct_len, char** json_out, size_t* json_len) {
    std::vector<uint8_t> proof;
    size_t num_blocks = 0;
    const int rc = aes_prove_checked(ctx, 32, key, key_len, nonce, nonce_len, counter, pt, pt_len, ct, ct_len, proof, num_blocks, json_out,
                                     json_len);
    if (rc != -1) return rc;
    return encrypt_finish(true, "aes256-ctr", proof, num_blocks, nonce, counter, pt, pt_len, ct, ct_len, json_out, json_len);
}

void s2c_free(void* p) { free(p); }

}  // extern "C"
