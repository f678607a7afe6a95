// AES-128/256-CTR AIR on the GPU: witness generation, S-box LogUp interaction trace, constraint-quotient evaluation.
//
// Follows (relative to /root/reference/stwo/src):
//   witness      aes/lookup/gen_ctr.rs:70-145 (append_byte/append_bits/xor_byte_trace/xtime_trace/sbox_trace), :152-195
//                (mix_columns_trace), :198-310 (process_ctr_block), :386-439 (default padding rows), wasm_api.rs:694-746
//                (padding lanes), aes/sbox_table.rs:52-76 (multiplicities)
//   interaction  aes/lookup/gen_ctr.rs:640-683 with upstream LogupTraceGenerator (constraint-framework prover/logup.rs)
//   constraints  aes/lookup/ctr.rs:26-364 (sbox, xor_byte, xtime, mix_columns, aes_block, ctr_block) + finalize_logup_in_pairs,
//                aes/sbox_table.rs:103-120 (table component), driven by upstream FrameworkComponent::
//                evaluate_constraint_quotients_on_domain; component accumulations of different sizes are combined by
//                AccumulationOps::lift_and_accumulate (pinned against the reference binary, aes_api.py in the oracle directory).
// Column order = trace order of the reference: 12 nonce, 4 counter (big-endian), round keys, 16 plaintext, 16 ciphertext,
// then per operation xor_byte = [8 a-bits, 8 b-bits, 8 c-bits, result], xtime = [8 a-bits, 8 r-bits, result], sbox = [out].
// One thread per trace row (witness, interaction) or per evaluation-domain row (constraints); columns are M31 words,
// column-major (column c at base + c*stride).
#include "common.cuh"
#include "m31_dev.cuh"

namespace aesk {
using namespace m31;
using m31d::AccSplit;
using m31d::bool_c;

__constant__ uint8_t c_sbox[256];

__device__ __forceinline__ uint32_t xt(uint32_t a) { return ((a << 1) & 0xff) ^ ((a >> 7) * 0x1b); }

__constant__ int c_shift_rows[16] = {0, 5, 10, 15, 4, 9, 14, 3, 8, 13, 2, 7, 12, 1, 6, 11};

// ------------------------------------------------------------------------------------------------ witness
struct Emit {
    uint32_t* p;      // trace + row
    size_t stride;
    int col;
    __device__ __forceinline__ void byte(uint32_t v) { p[(size_t)(col++) * stride] = v; }
    __device__ __forceinline__ void bits(uint32_t v) {
#pragma unroll
        for (int i = 0; i < 8; i++) p[(size_t)(col++) * stride] = (v >> i) & 1u;
    }
    __device__ __forceinline__ uint32_t xor_byte(uint32_t a, uint32_t b) {
        uint32_t r = a ^ b;
        bits(a); bits(b); bits(r); byte(r);
        return r;
    }
    __device__ __forceinline__ uint32_t xtime(uint32_t a) {
        uint32_t r = xt(a);
        bits(a); bits(r); byte(r);
        return r;
    }
    __device__ __forceinline__ uint32_t mul3(uint32_t a) { return xor_byte(xtime(a), a); }
};

struct WitnessArgs {
    uint8_t rk[15 * 16];
    uint8_t nonce[12];
    uint32_t counter;
    uint32_t num_blocks;     // rows with caller-supplied plaintext/ciphertext
    uint32_t n_active_rows;  // rows covered by provided vec-rows (multiple of 16); rows beyond use the default input
    int n_rounds;            // 10 or 14
    int block;               // block AIR (aes/lookup/{gen,constraints,air}.rs): input block = pt[row], no counter block, no pt/ct columns
};

__global__ void __launch_bounds__(128) witness_kernel(WitnessArgs a, const uint8_t* __restrict__ pt, const uint8_t* __restrict__ ct,
                                                      int log_size, uint32_t* __restrict__ T, size_t stride,
                                                      unsigned int* __restrict__ mults, int* __restrict__ invalid) {
    __shared__ unsigned int hist[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row < (1u << log_size)) {
        const bool active = row < a.n_active_rows, real = row < a.num_blocks;
        const uint32_t ctr = active ? a.counter + row : (row & 15u);
        Emit e{T + row, stride, 0};
        uint32_t blk[16], p[16];
        if (a.block) {
#pragma unroll
            for (int i = 0; i < 16; i++) { blk[i] = pt[(size_t)row * 16 + i]; e.byte(blk[i]); }
        } else {
#pragma unroll
            for (int i = 0; i < 12; i++) { blk[i] = active ? a.nonce[i] : 0u; e.byte(blk[i]); }
#pragma unroll
            for (int i = 0; i < 4; i++) { blk[12 + i] = (ctr >> (8 * (3 - i))) & 0xffu; e.byte(blk[12 + i]); }
        }
        for (int r = 0; r <= a.n_rounds; r++)
            for (int i = 0; i < 16; i++) e.byte(a.rk[16 * r + i]);
        int ct_col = 0;
        if (!a.block) {
#pragma unroll
            for (int i = 0; i < 16; i++) { p[i] = real ? pt[(size_t)row * 16 + i] : 0u; e.byte(p[i]); }
            ct_col = e.col;  // ciphertext columns are written once the keystream is known (padding: ct = keystream)
            e.col += 16;
        }
        uint32_t s[16], t[16];
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = e.xor_byte(blk[i], a.rk[i]);
        for (int rnd = 1; rnd <= a.n_rounds; rnd++) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                atomicAdd(&hist[s[i]], 1u);
                s[i] = c_sbox[s[i]];
                e.byte(s[i]);
            }
#pragma unroll
            for (int i = 0; i < 16; i++) t[i] = s[c_shift_rows[i]];
            if (rnd < a.n_rounds) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t s0 = t[4 * c], s1 = t[4 * c + 1], s2 = t[4 * c + 2], s3 = t[4 * c + 3];
                    uint32_t t0, t1, t2, t3;
                    t0 = e.xtime(s0); t1 = e.mul3(s1); t2 = e.xor_byte(t0, t1); t3 = e.xor_byte(t2, s2); s[4 * c] = e.xor_byte(t3, s3);
                    t0 = e.xtime(s1); t1 = e.mul3(s2); t2 = e.xor_byte(s0, t0); t3 = e.xor_byte(t2, t1); s[4 * c + 1] = e.xor_byte(t3, s3);
                    t0 = e.xtime(s2); t1 = e.mul3(s3); t2 = e.xor_byte(s0, s1); t3 = e.xor_byte(t2, t0); s[4 * c + 2] = e.xor_byte(t3, t1);
                    t0 = e.mul3(s0); t1 = e.xtime(s3); t2 = e.xor_byte(t0, s1); t3 = e.xor_byte(t2, s2); s[4 * c + 3] = e.xor_byte(t3, t1);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) s[i] = t[i];
            }
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = e.xor_byte(s[i], a.rk[16 * rnd + i]);
        }
        bool ok = true;
        if (!a.block)
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t comp = e.xor_byte(s[i], p[i]);
            const uint32_t c = real ? ct[(size_t)row * 16 + i] : comp;
            if (c != comp) ok = false;
            T[(size_t)(ct_col + i) * stride + row] = c;
        }
        if (real && !ok) atomicOr(invalid, 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (hist[i]) atomicAdd(&mults[i], hist[i]);
}

// ------------------------------------------------------------------------------------------------ LogUp interaction trace
// combine([a, b]) = a + alpha*b - z  (relation!(SboxElements, 2), aes/sbox_table.rs:24)
__device__ __forceinline__ QM31 combine(const QM31& z, const QM31& alpha, uint32_t a, uint32_t b) {
    QM31 r = qmul_m(alpha, b);
    r.v[0] = add(r.v[0], a);
    return qsub(r, z);
}

// I[4k+c][row] = sum_{j<=k} (p0_j + p1_j) / (p0_j p1_j), the cumulative LogUp columns of the CTR component (last column
// still unshifted / not prefix-summed: finalize_last is done by the caller)
__global__ void __launch_bounds__(128) interaction_kernel(const uint32_t* __restrict__ T, size_t stride, int log_size,
                                                          const int* __restrict__ lk_in, const int* __restrict__ lk_out, int n_lookups,
                                                          QM31 z, QM31 alpha, uint32_t* __restrict__ I, size_t i_stride) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (1u << log_size)) return;
    QM31 cum = qzero();
    for (int k = 0; k < n_lookups / 2; k++) {
        const QM31 p0 = combine(z, alpha, T[(size_t)lk_in[2 * k] * stride + row], T[(size_t)lk_out[2 * k] * stride + row]);
        const QM31 p1 = combine(z, alpha, T[(size_t)lk_in[2 * k + 1] * stride + row], T[(size_t)lk_out[2 * k + 1] * stride + row]);
        cum = qadd(cum, qmul(qadd(p0, p1), qinv(qmul(p0, p1))));
#pragma unroll
        for (int c = 0; c < 4; c++) I[(size_t)(4 * k + c) * i_stride + row] = cum.v[c];
    }
}

// ------------------------------------------------------------------------------------------------ constraints
struct Eval {
    const uint32_t* lde;  // + row
    size_t stride;
    const uint4 *tlo, *thi;  // reversed alpha powers split in 16-bit halves
    AccSplit acc;
    int col, k;
    __device__ __forceinline__ uint32_t ld(int c) const { return lde[(size_t)c * stride]; }
    __device__ __forceinline__ void emit(uint32_t C) { acc.mac(tlo, thi, k++, C); }
    // boolean constraints of 8 bit values (already loaded), returns sum 2^i b_i
    __device__ __forceinline__ uint32_t bits8(const uint32_t (&b)[8]) {
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            emit(bool_c(b[i], b[i] + b[i]));
            s = m31d::addm(s, m31d::mulm(b[i], 1u << i));
        }
        return s;
    }
    // Every operation first issues ALL of its column loads (27 for a byte xor, 18 for xtime) and only then evaluates its
    // constraints, in the reference's order: with ~150 registers only 12 warps are resident per SM, so the loads in flight
    // per thread are what hides the HBM latency.
    // returns the column index of the result byte
    __device__ __forceinline__ int xor_byte(int a, int b) {
        uint32_t ab[8], bb[8], cb[8];
        const uint32_t* p = lde + (size_t)col * stride;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ab[i] = p[(size_t)i * stride];
            bb[i] = p[(size_t)(8 + i) * stride];
            cb[i] = p[(size_t)(16 + i) * stride];
        }
        const uint32_t vr = p[(size_t)24 * stride], va = ld(a), vb = ld(b);
        asm volatile("" ::: "memory");  // keep the loads together: the scheduler otherwise sinks each next to its use
        const int r = col + 24;
        col += 25;
        const uint32_t sa = bits8(ab), sb = bits8(bb), sc = bits8(cb);
        emit(m31d::subm(va, sa));
        emit(m31d::subm(vb, sb));
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t m = m31d::mulm(ab[i], bb[i]);
            emit(m31d::addm(m31d::subm(m31d::subm(cb[i], ab[i]), bb[i]), m31d::dbl(m)));
        }
        emit(m31d::subm(vr, sc));
        return r;
    }
    __device__ __forceinline__ int xtime(int a) {
        uint32_t ab[8], rb[8];
        const uint32_t* p = lde + (size_t)col * stride;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ab[i] = p[(size_t)i * stride];
            rb[i] = p[(size_t)(8 + i) * stride];
        }
        const uint32_t vr = p[(size_t)16 * stride], va = ld(a);
        asm volatile("" ::: "memory");
        const int r = col + 16;
        col += 17;
        const uint32_t sa = bits8(ab);
        emit(m31d::subm(va, sa));
        const uint32_t sr = bits8(rb);
        const uint32_t h = ab[7];
        auto x2 = [&](int i, int j) {
            const uint32_t m = m31d::mulm(ab[j], h);
            return m31d::addm(m31d::subm(m31d::subm(rb[i], ab[j]), h), m31d::dbl(m));
        };
        emit(m31d::subm(rb[0], h));
        emit(x2(1, 0));
        emit(m31d::subm(rb[2], ab[1]));
        emit(x2(3, 2));
        emit(x2(4, 3));
        emit(m31d::subm(rb[5], ab[4]));
        emit(m31d::subm(rb[6], ab[5]));
        emit(m31d::subm(rb[7], ab[6]));
        emit(m31d::subm(vr, sr));
        return r;
    }
    __device__ __forceinline__ int mul3(int a) { return xor_byte(xtime(a), a); }
};

__device__ __forceinline__ uint32_t prev_row(uint32_t row, int eval_log, int trace_log) {
    // core/utils.rs offset_bit_reversed_circle_domain_index(row, trace_log, eval_log, -1)
    uint32_t nat = __brev(row) >> (32 - eval_log);
    const uint32_t half = 1u << (eval_log - 1);
    const uint32_t step = 1u << (eval_log - trace_log - 1);
    if (nat < half) nat = (nat + half - step) & (half - 1);
    else nat = ((nat - half + step) & (half - 1)) + half;
    return __brev(nat) >> (32 - eval_log);
}

struct ConsArgs {
    const uint32_t* lde;     // main trace LDE, column-major
    size_t stride;
    const uint32_t* inter;   // CTR interaction LDE: 4*(L/2) coordinate columns
    size_t i_stride;
    const uint32_t *apr_lo, *apr_hi;  // split reversed alpha powers (this component's constraints first)
    const uint32_t* apr;     // unsplit table (QM31 words) for the extension-field constraints
    const uint32_t* den_inv; // [2^(eval_log-trace_log)]
    const int *lk_in, *lk_out;
    QM31 z, alpha, shift;    // lookup elements and claimed_sum / N
    int eval_log, trace_log, n_rounds, n_lookups;
    uint32_t* out;           // 4 coordinate columns
    size_t out_stride;
    int block;               // block AIR: no plaintext / ciphertext columns, the walk ends with the last AddRoundKey
};

// One round of MixColumns as a program of 96 byte operations {type (0: xor_byte, 1: xtime), a, b, dst} over 64 slots of column
// indices: 0..15 state, 16..31 state after ShiftRows, 32..36 temporaries, 48..63 the round-key (or plaintext) columns of the
// current AddRoundKey.  The kernel interprets it in a loop, so its code is the two operation bodies once (a few KB) instead of
// 112 inlined copies per round (636 KB: every warp then ran at instruction-fetch speed - 4 ms for a 512-row product-size proof).
// The operations consume trace columns in the reference's evaluation order (ctr.rs:244-281).
__constant__ uchar4 c_mix_prog[96];
static void build_mix_prog(uchar4 (&prog)[96]) {
    int n = 0;
    auto XT = [&](int a, int dst) { prog[n++] = make_uchar4(1, (unsigned char)a, 0, (unsigned char)dst); };
    auto XR = [&](int a, int b, int dst) { prog[n++] = make_uchar4(0, (unsigned char)a, (unsigned char)b, (unsigned char)dst); };
    const int A = 32, B = 33, C = 34, D = 35, E = 36;
    for (int c = 0; c < 4; c++) {
        const int s0 = 16 + 4 * c, s1 = s0 + 1, s2 = s0 + 2, s3 = s0 + 3, o = 4 * c;
        XT(s0, A); XT(s1, E); XR(E, s1, B); XR(A, B, C); XR(C, s2, D); XR(D, s3, o);          // 2 s0 + 3 s1 + s2 + s3
        XT(s1, A); XT(s2, E); XR(E, s2, B); XR(s0, A, C); XR(C, B, D); XR(D, s3, o + 1);      // s0 + 2 s1 + 3 s2 + s3
        XT(s2, A); XT(s3, E); XR(E, s3, B); XR(s0, s1, C); XR(C, A, D); XR(D, B, o + 2);      // s0 + s1 + 2 s2 + 3 s3
        XT(s0, E); XR(E, s0, A); XT(s3, B); XR(A, s1, C); XR(C, s2, D); XR(D, B, o + 3);      // 3 s0 + s1 + s2 + 2 s3
    }
}

__global__ void __launch_bounds__(64) constraints_kernel(ConsArgs A) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (1u << A.eval_log)) return;
    Eval e;
    e.lde = A.lde + row;
    e.stride = A.stride;
    e.tlo = (const uint4*)A.apr_lo;
    e.thi = (const uint4*)A.apr_hi;
    e.acc.init();
    e.col = 0;
    e.k = 0;
    const int nr = A.n_rounds;
    int s[64];
    // nonce || counter occupy columns 0..15, round keys 16.., plaintext, ciphertext
    const int rk0 = 16, pt0 = 16 + 16 * (nr + 1), ct0 = pt0 + 16;
    e.col = A.block ? pt0 : ct0 + 16;
    auto add_round_key = [&](int key_col0) {  // s[i] = xor_byte(s[i], key_col0 + i)
#pragma unroll 1
        for (int i = 0; i < 16; i++) s[i] = e.xor_byte(s[i], key_col0 + i);
    };
    for (int i = 0; i < 16; i++) s[i] = i;
    add_round_key(rk0);
#pragma unroll 1
    for (int rnd = 1; rnd <= nr; rnd++) {
        // S-box outputs are the next 16 columns (their relation entries are consumed below); ShiftRows renames them
        for (int i = 0; i < 16; i++) s[16 + i] = e.col + c_shift_rows[i];
        e.col += 16;
        if (rnd < nr) {
#pragma unroll 1
            for (int j = 0; j < 96; j++) {
                const uchar4 op = c_mix_prog[j];
                s[op.w] = op.x ? e.xtime(s[op.y]) : e.xor_byte(s[op.y], s[op.z]);
            }
        } else {
            for (int i = 0; i < 16; i++) s[i] = s[16 + i];
        }
        add_round_key(rk0 + 16 * rnd);
    }
    if (!A.block) {
        add_round_key(pt0);
        for (int i = 0; i < 16; i++) e.emit(m31d::subm(e.ld(s[i]), e.ld(ct0 + i)));
    }
    // finalize_logup_in_pairs: L/2 extension-field constraints (cur - prev_col [- prev_row + shift]) * den - num
    QM31 ext = qzero(), prev_col = qzero();
    const int nb = A.n_lookups / 2;
    const uint4* apr4 = (const uint4*)A.apr;
    const uint32_t* ip = A.inter + row;
    for (int k = 0; k < nb; k++) {
        const QM31 p0 = combine(A.z, A.alpha, e.ld(A.lk_in[2 * k]), e.ld(A.lk_out[2 * k]));
        const QM31 p1 = combine(A.z, A.alpha, e.ld(A.lk_in[2 * k + 1]), e.ld(A.lk_out[2 * k + 1]));
        QM31 cur;
#pragma unroll
        for (int c = 0; c < 4; c++) cur.v[c] = ip[(size_t)(4 * k + c) * A.i_stride];
        QM31 diff = qsub(cur, prev_col);
        if (k == nb - 1) {
            const uint32_t pr = prev_row(row, A.eval_log, A.trace_log);
            QM31 pv;
#pragma unroll
            for (int c = 0; c < 4; c++) pv.v[c] = A.inter[(size_t)(4 * k + c) * A.i_stride + pr];
            diff = qadd(qsub(diff, pv), A.shift);
        }
        prev_col = cur;
        const QM31 G = qsub(qmul(diff, qmul(p0, p1)), qadd(p0, p1));
        const uint4 al = __ldg(apr4 + e.k++);
        ext = qadd(ext, qmul(G, QM31{{al.x, al.y, al.z, al.w}}));
    }
    const uint32_t d = A.den_inv[row >> A.trace_log];
#pragma unroll
    for (int c = 0; c < 4; c++) A.out[(size_t)c * A.out_stride + row] = m31d::mulm(m31d::addm(e.acc.result(c), ext.v[c]), d);
}

// S-box table component (aes/sbox_table.rs:103-120) on its 2^9-point evaluation domain, multiplied by its vanishing inverse:
// one LogUp entry with multiplicity -mult and numerator/denominator (-mult, combine(in, out)); mask [-1, 0]
struct TableArgs {
    const uint32_t *pre_in, *pre_out, *mult;  // LDE columns of 2^eval_log values
    const uint32_t* inter;                     // 4 coordinate columns
    size_t i_stride;
    QM31 z, alpha, shift, apow;                // apow = the alpha power of this (last) constraint
    const uint32_t* den_inv;
    int eval_log, trace_log;
    uint32_t* out;                             // [4][2^eval_log]
};
__global__ void table_constraint_kernel(TableArgs A) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (1u << A.eval_log)) return;
    const QM31 p = combine(A.z, A.alpha, A.pre_in[row], A.pre_out[row]);
    const uint32_t pr = prev_row(row, A.eval_log, A.trace_log);
    QM31 cur, pv;
#pragma unroll
    for (int c = 0; c < 4; c++) { cur.v[c] = A.inter[c * A.i_stride + row]; pv.v[c] = A.inter[c * A.i_stride + pr]; }
    const QM31 diff = qadd(qsub(cur, pv), A.shift);
    QM31 G = qmul(diff, p);
    G.v[0] = add(G.v[0], A.mult[row]);  // - num, num = -mult
    G = qmul_m(qmul(G, A.apow), A.den_inv[row >> A.trace_log]);
#pragma unroll
    for (int c = 0; c < 4; c++) A.out[(size_t)c * (1u << A.eval_log) + row] = G.v[c];
}

// AccumulationOps::lift_and_accumulate: big[c][i] += small[c][((i >> (s+1)) << 1) | (i & 1)]
__global__ void lift_accumulate_kernel(uint32_t* __restrict__ big, size_t big_stride, int big_log, const uint32_t* __restrict__ small,
                                       int small_log) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << big_log)) return;
    const int sh = big_log - small_log;
    const uint32_t j = sh ? (((i >> (sh + 1)) << 1) | (i & 1)) : i;
#pragma unroll
    for (int c = 0; c < 4; c++) big[c * big_stride + i] = add(big[c * big_stride + i], small[((size_t)c << small_log) + j]);
}

}  // namespace aesk

cudaError_t aes_upload_sbox(const uint8_t sbox[256]) { return cudaMemcpyToSymbol(aesk::c_sbox, sbox, 256); }

cudaError_t launch_aes_witness(cudaStream_t st, const uint8_t* rk, int n_rounds, const uint8_t nonce[12], uint32_t counter,
                               uint32_t num_blocks, uint32_t n_active_rows, const uint8_t* pt, const uint8_t* ct, int log_size,
                               uint32_t* T, size_t stride, unsigned int* mults, int* invalid, int block_air) {
    aesk::WitnessArgs a;
    a.block = block_air;
    memcpy(a.rk, rk, 16 * (n_rounds + 1));
    memcpy(a.nonce, nonce, 12);
    a.counter = counter;
    a.num_blocks = num_blocks;
    a.n_active_rows = n_active_rows;
    a.n_rounds = n_rounds;
    uint32_t n = 1u << log_size;
    aesk::witness_kernel<<<(n + 127) / 128, 128, 0, st>>>(a, pt, ct, log_size, T, stride, mults, invalid);
    return cudaGetLastError();
}

cudaError_t launch_aes_interaction(cudaStream_t st, const uint32_t* T, size_t stride, int log_size, const int* lk_in, const int* lk_out,
                                   int n_lookups, m31::QM31 z, m31::QM31 alpha, uint32_t* I, size_t i_stride) {
    uint32_t n = 1u << log_size;
    aesk::interaction_kernel<<<(n + 127) / 128, 128, 0, st>>>(T, stride, log_size, lk_in, lk_out, n_lookups, z, alpha, I, i_stride);
    return cudaGetLastError();
}

cudaError_t launch_aes_constraints(cudaStream_t st, const AesConsArgs& a) {
    static const cudaError_t prog_ok = [] {  // once per process and device context
        uchar4 prog[96];
        aesk::build_mix_prog(prog);
        return cudaMemcpyToSymbol(aesk::c_mix_prog, prog, sizeof prog);
    }();
    if (prog_ok != cudaSuccess) return prog_ok;
    aesk::ConsArgs A;
    static_assert(sizeof(aesk::ConsArgs) == sizeof(AesConsArgs), "layout");
    memcpy(&A, &a, sizeof A);
    uint32_t rows = 1u << a.eval_log;
    aesk::constraints_kernel<<<(rows + 63) / 64, 64, 0, st>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_aes_table_constraint(cudaStream_t st, const AesTableArgs& a) {
    aesk::TableArgs A;
    static_assert(sizeof(aesk::TableArgs) == sizeof(AesTableArgs), "layout");
    memcpy(&A, &a, sizeof A);
    uint32_t rows = 1u << a.eval_log;
    aesk::table_constraint_kernel<<<(rows + 127) / 128, 128, 0, st>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_lift_accumulate(cudaStream_t st, uint32_t* big, size_t big_stride, int big_log, const uint32_t* small, int small_log) {
    uint32_t n = 1u << big_log;
    aesk::lift_accumulate_kernel<<<(n + 255) / 256, 256, 0, st>>>(big, big_stride, big_log, small, small_log);
    return cudaGetLastError();
}
