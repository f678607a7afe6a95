// Shared declarations for the CUDA side of the B200 stwo backend.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "m31.cuh"

// Where a column's trace-domain values come from.
enum { SRC_M31 = 0, SRC_BITS = 1, SRC_BYTES = 2 };
struct ColSrc {
    int kind;              // SRC_M31: base[col*stride + row];  SRC_BITS: bit (c&31) of base[(c>>5)*stride + row];
                           // SRC_BYTES: byte (c&3) of base[(c>>2)*stride + row];  c = col + first_col
    const uint32_t* base;
    size_t stride;         // words between consecutive columns (SRC_M31) or packed words (SRC_BITS/BYTES)
    uint32_t first_col;
};

struct FftTables {
    const uint32_t *X, *Y, *IX, *IY;  // flattened twiddle tables, see kernels_fft.cu
    int max_log;                      // largest canonic domain covered
};

// Equally sized, equally strided columns feeding Merkle leaves (see kernels_merkle.cu)
struct LeafGroup {
    const uint32_t* base;
    size_t stride;
    int ncols;
    int log_size;
    // adder-sum word computed on the fly (streaming prover): when cy != nullptr the group has 32 columns of the lifting size,
    // value(col) = base[col] + b[col] + cy[col-1] - 2 cy[col] (cy[-1] = 0), which is also stored to res[col] (may alias base)
    const uint32_t* b;
    const uint32_t* cy;
    uint32_t* res;
};
#define MAX_LEAF_GROUPS 16
struct LeafGroups {
    LeafGroup g[MAX_LEAF_GROUPS];
    int n;
};

// One sample batch of the FRI quotient accumulation (see kernels_pcs.cu quotients_kernel)
struct QuotBatch {
    m31::CM31 prx, pry, pix, piy;  // sample point P = Pr + u*Pi
    m31::QM31 lin_a, lin_b;        // sum_j alpha_j a_j, sum_j alpha_j b_j
    m31::QM31 batch_coeff;         // multiplier applied to the running row accumulator before adding this batch
    const uint32_t* coefs;         // device [n_cols][4] = alpha_j * c_j
    const uint32_t* col_idx;       // device [n_cols] column indices (nullptr = identity)
    int n_cols;
    // optional per-entry column pointers (device array of n_cols pointers) with the log size of each column, for trees that
    // mix column sizes: entry j reads col_ptr[j][lift(row)] where lift maps the row of the 2^m domain to the smaller column
    const uint32_t* const* col_ptr;
    const uint8_t* col_log;
};

// ---- streaming prover job lists (passed to kernels by value) ----
#define MAX_FFT_JOBS 16
enum { CJ_BOOL = 0, CJ_XOR = 1, CJ_XORN = 2, CJ_ADDX = 3 };
struct ConstraintJob {
    const uint32_t *t0, *t1, *t2;  // tiles [32][M]: CJ_BOOL t0; CJ_XOR* r = t0, a = t1, d = t2; CJ_ADDX x = t0 (or null), a = t1, d = t2
    const uint32_t *t3, *t4;       // CJ_ADDX: second adder operand b = t3, carry word cy = t4
    uint32_t* res;                 // CJ_ADDX: the adder's sum tile, computed here (a + b + cy[-1] - 2cy) and stored
    int kx;                        // index (in the reversed alpha-power table) of the xor constraint of bit 0
    int kb0, kb1, kb2;             // index of the boolean constraint of bit 0 of t0 / t1 (CJ_ADDX: the sum) / t2 (-1 = none)
    int kbc;                       // CJ_ADDX: index of the boolean constraint of carry bit 0 (step 2)
    int arg;                       // CJ_BOOL: constraint index step per bit; others: left rotation
    int type;
};
#define MAX_CONSTRAINT_JOBS 32
struct ConstraintJobs {
    ConstraintJob j[MAX_CONSTRAINT_JOBS];
    int n;
};

// ---- ChaCha stream AIR as a constraint table (prove_chacha.cu build_cons_recs): constraint k on mask values v[.]
//   CR_BOOL: v[c0] (1 - v[c0])                        CR_ADD: v[c0] + 2 v[c1] - v[c2] - v[c3] - v[c4]   (c4 = -1: no carry in)
//   CR_XOR : v[c0] - v[c1] - v[c2] + 2 v[c1] v[c2]    CR_EQ : v[c0] + v[c1] - 2 v[c0] v[c1] - v[c2]
enum { CR_BOOL = 0, CR_ADD = 1, CR_XOR = 2, CR_EQ = 3 };
struct ConsRec {
    int type, c0, c1, c2, c3, c4;
};
HD m31::QM31 cons_rec_eval(const ConsRec& r, const m31::QM31* v) {
    using namespace m31;
    switch (r.type) {
        case CR_BOOL: return qmul(v[r.c0], qsub(qone(), v[r.c0]));
        case CR_ADD: {
            QM31 t = qsub(qsub(qadd(v[r.c0], qadd(v[r.c1], v[r.c1])), v[r.c2]), v[r.c3]);
            return r.c4 >= 0 ? qsub(t, v[r.c4]) : t;
        }
        case CR_XOR: {
            QM31 ab = qmul(v[r.c1], v[r.c2]);
            return qadd(qsub(qsub(v[r.c0], v[r.c1]), v[r.c2]), qadd(ab, ab));
        }
        default: {
            QM31 kp = qmul(v[r.c0], v[r.c1]);
            return qsub(qsub(qadd(v[r.c0], v[r.c1]), qadd(kp, kp)), v[r.c2]);
        }
    }
}
// Tail kernels of the streaming ChaCha prover (kernels_tail.cu): the per-column QM31 loops of prove() that used to run on the host
struct SumComb { int res, a, b, c; };  // adder sum word res = a + b + carry-in(c) - 2 c  (word indices)
// out[4] = sum_k table[k](mask) * apr[k]; mask: [n_cols][4] QM31 values, or (mask_is_base) [n_cols] base-field values
cudaError_t launch_mask_constraints(cudaStream_t st, const ConsRec* table, int K, const uint32_t* mask, int mask_is_base,
                                    const uint32_t* apr, uint32_t* out);
// sampled[res*32+i] = sampled[a*32+i] + sampled[b*32+i] + sampled[c*32+i-1] - 2 sampled[c*32+i], in list order
cudaError_t launch_oods_fill_sums(cudaStream_t st, uint32_t* sampled, const SumComb* combs, int n_combs);
// FRI quotient line coefficients of every sampled column at the OODS point z (upstream pcs/quotients.rs column_line_coeffs +
// the random-coefficient powers): coefs[j] = rc^j c, lin[0..4) = sum_j rc^j a_j, lin[4..8) = sum_j rc^j b_j with
// c = conj(z.y) - z.y, a_j = conj(v_j) - v_j, b_j = v_j c - a_j z.y.  pw_rev[k] = rc^(nc-1-k).
cudaError_t launch_quot_coefs(cudaStream_t st, const uint32_t* sampled, int nc, const uint32_t* pw_rev, m31::QM31 zy, uint32_t* coefs,
                              uint32_t* lin);
// the sum columns' coefficients folded into their operands', reverse list order (fc: [n_cols][4])
cudaError_t launch_quot_fold_sums(cudaStream_t st, uint32_t* fc, const SumComb* combs, int n_combs);

// ---- AES-CTR AIR kernels (kernels_aes.cu); plain-data mirrors of the kernel argument structs
struct AesConsArgs {
    const uint32_t* lde;
    size_t stride;
    const uint32_t* inter;
    size_t i_stride;
    const uint32_t *apr_lo, *apr_hi;
    const uint32_t* apr;
    const uint32_t* den_inv;
    const int *lk_in, *lk_out;
    m31::QM31 z, alpha, shift;
    int eval_log, trace_log, n_rounds, n_lookups;
    uint32_t* out;
    size_t out_stride;
    int block;  // block AIR variant (aes/lookup/constraints.rs aes128_block)
};
struct AesTableArgs {
    const uint32_t *pre_in, *pre_out, *mult;
    const uint32_t* inter;
    size_t i_stride;
    m31::QM31 z, alpha, shift, apow;
    const uint32_t* den_inv;
    int eval_log, trace_log;
    uint32_t* out;
};
cudaError_t aes_upload_sbox(const uint8_t sbox[256]);
cudaError_t launch_aes_witness(cudaStream_t st, const uint8_t* rk, int n_rounds, const uint8_t nonce[12], uint32_t counter,
                               uint32_t num_blocks, uint32_t n_active_rows, const uint8_t* pt, const uint8_t* ct, int log_size,
                               uint32_t* T, size_t stride, unsigned int* mults, int* invalid, int block_air = 0);
cudaError_t launch_aes_interaction(cudaStream_t st, const uint32_t* T, size_t stride, int log_size, const int* lk_in, const int* lk_out,
                                   int n_lookups, m31::QM31 z, m31::QM31 alpha, uint32_t* I, size_t i_stride);
cudaError_t launch_aes_constraints(cudaStream_t st, const AesConsArgs& a);
cudaError_t launch_aes_table_constraint(cudaStream_t st, const AesTableArgs& a);
cudaError_t launch_lift_accumulate(cudaStream_t st, uint32_t* big, size_t big_stride, int big_log, const uint32_t* small, int small_log);

// ---- small backend-trait kernels (kernels_ops.cu)
cudaError_t launch_bit_reverse(cudaStream_t st, uint32_t* col, int log_size);
cudaError_t launch_inverse_m31(cudaStream_t st, const uint32_t* src, uint32_t* dst, size_t n);
cudaError_t launch_inverse_qm31(cudaStream_t st, const uint32_t* src, size_t s_stride, uint32_t* dst, size_t d_stride, size_t n);
size_t logup_finalize_scratch_words(int log);
cudaError_t launch_logup_finalize_last(cudaStream_t st, uint32_t* col, size_t stride, int log, uint32_t* scratch, uint32_t** claimed_dev);
cudaError_t launch_commit_on_layer(cudaStream_t st, const uint32_t* prev, const uint32_t* const* cols_dev, int n_cols, uint32_t n_nodes,
                                   uint32_t* out);

// optional per-kernel profiling callback (begin=1 before a launch, begin=0 after it)
struct StageHook {
    void (*fn)(void* user, const char* name, int begin);
    void* user;
};

cudaError_t launch_fft(cudaStream_t st, const ColSrc& src, int ncols, int log_n, int ext, int mode, uint32_t* coef_out,
                       size_t coef_stride, uint32_t* eval_out, size_t eval_stride, const FftTables& tw, uint32_t* scratch,
                       size_t scratch_stride, const StageHook* hook = nullptr);
cudaError_t launch_fft_packed_list(cudaStream_t st, const int* wlist_dev, int n_words, const uint32_t* src0, size_t src_stride,
                                   uint32_t* out0, size_t out_stride, int log_n, const FftTables& tw);
void fft_init_attrs();
void fft2_init_attrs();
size_t fft_packed_scratch_words(int kind, int njobs, int log_n);
// Peer windows of the row-sharded mode: with `peer` set, launch_fft_packed's last pass stores the transformed columns into the
// tile slots of the ranks that own the rows (word offset off[j] inside every rank's arena, rows dealt in 2G "virtual shards"
// of 2^lv rows: rank = v mod G, local rows [ (v div G) 2^lv, .. ) ), out[j] is then only the local staging tile.
#define MAX_PEERS 8
struct PeerDst {
    uint32_t* base[MAX_PEERS];
    int logG, lv;
    const unsigned long long* off;  // host array, one entry per job
    // optional: launch the last pass on its own stream (it is throttled by NVLink; the passes of the next group then start
    // on the caller's stream meanwhile).  ab_done is recorded behind the second pass and awaited by last_stream.
    cudaStream_t last_stream;
    cudaEvent_t ab_done;
};
cudaError_t launch_fft_packed(cudaStream_t st, int kind, const uint32_t* const* src, uint32_t* const* out, int njobs, int log_n,
                              const FftTables& tw, uint32_t* scratch, const StageHook* hook, int* launches, int shard_log = 0,
                              int first_half_only = 0, const PeerDst* peer = nullptr);
cudaError_t launch_constraints_tiles(cudaStream_t st, const ConstraintJobs& jobs, size_t M, const uint32_t* apr_lo,
                                     const uint32_t* apr_hi, uint32_t* acc, int first, size_t rows = 0);
// FP64-accumulate form: gtab = the launch's alpha table in consumption order (launch_cons_table), jobs.j[k].kx = offset of job k's
// slice in it (in constraints)
cudaError_t launch_constraints_tiles2(cudaStream_t st, const ConstraintJobs& jobs, size_t M, const double* gtab, uint32_t* acc, int first,
                                      size_t rows = 0);
// product-size traces: the whole AIR in one launch (job list in device memory, one block column per job) + a reduction
cudaError_t launch_constraints_jobs(cudaStream_t st, const ConstraintJob* jobs_dev, int n_jobs, size_t M, const double* gtab,
                                    uint32_t* partial, uint32_t* acc, size_t rows);
cudaError_t launch_sum_tiles(cudaStream_t st, uint32_t* arena, size_t tile_words, size_t rows, const SumComb* combs, int n_combs);
cudaError_t launch_merkle_leaves_seq(cudaStream_t st, const uint32_t* arena, size_t tile_words, int n_words, int lifting_log, uint32_t* out);
cudaError_t launch_cons_table(cudaStream_t st, const uint32_t* apr, const int* idx_dev, int n, double* gtab);
cudaError_t launch_split16(cudaStream_t st, const uint32_t* table, int n, uint32_t* lo, uint32_t* hi);
cudaError_t launch_scale_rows(cudaStream_t st, uint32_t* acc, size_t M, int trace_log, const uint32_t* den_inv, size_t row0 = 0);
struct TileRowJobs {
    int n;
    int word[MAX_LEAF_GROUPS];
    const uint32_t* tile[MAX_LEAF_GROUPS];
};
cudaError_t launch_gather_tile_row(cudaStream_t st, const TileRowJobs& jobs, size_t M, size_t row, uint32_t* out);
cudaError_t launch_gather_cached(cudaStream_t st, const uint32_t* arena, size_t tile_words, size_t M, const int* slot_dev, int n_words,
                                 const uint32_t rows[4], int nq, uint32_t* out);
cudaError_t launch_bitcol_dot(cudaStream_t st, const uint32_t* W, size_t N, int n_words, const uint32_t* wt, uint32_t scale,
                              uint32_t* out, const int* words_dev = nullptr);
cudaError_t launch_bitrow_comb(cudaStream_t st, const uint32_t* W, size_t N, int n_words, const uint32_t* coefs, uint32_t* g,
                               const int* words_dev = nullptr);
// g[idx] = sum over p of partial[p][idx] (mod p): partial results over disjoint word sets ([parts][4][N])
cudaError_t launch_bitrow_reduce(cudaStream_t st, const uint32_t* partial, size_t N, int parts, uint32_t* g);
cudaError_t launch_rowcomb_m31(cudaStream_t st, const uint32_t* vals, size_t stride, int ncols, size_t N, const uint32_t* coefs,
                               uint32_t* g, int accumulate);
cudaError_t launch_basis4(cudaStream_t st, uint32_t* basis, size_t stride, int log_n, const uint32_t init[4],
                          const uint32_t (*maps)[4]);
