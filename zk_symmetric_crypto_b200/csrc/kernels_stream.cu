// Kernels of the streaming (tile-by-tile) prover: everything that consumes LDE tiles or the packed witness directly.
//
// The ChaCha20 AIR (/root/reference/stwo/src/chacha/bitwise/constraints_stream.rs:20-70) has 33,280 one-bit columns; at
// log_n_rows = 20 its LDE is 279 GB, more than one B200 holds.  The prover therefore never materialises the LDE: it keeps
// the packed witness (1,040 words per row) and streams LDE tiles of 32 columns through the consumers below.  Three facts
// about the reference's arithmetic are used (all exact in M31/QM31, so results are bit-identical to the reference's):
//   (1) interpolation and extension are linear.  The sum word of every 32-bit adder satisfies, bit by bit,
//       s_i = a_i + b_i + c_{i-1} - 2 c_i on the trace domain, both sides have degree < N, hence the identity holds on the
//       extended domain too: sum tiles are combined from operand tiles (inside the leaf-hash and constraint kernels) instead of transformed,
//       and the adder constraints (constraints_stream.rs:117-129) vanish identically on the evaluation domain.
//   (2) f(z) for a column given by its trace-domain values is <values, w(z)> with w(z) = (iFFT)^T basis(z): out-of-domain
//       samples and queried LDE values of all bit columns are masked sums over the packed witness (bitcol_dot_kernel).
//   (3) the FRI quotient numerator sum_j alpha_j c f_j(p) is the extension of the row-wise combination
//       sum_j alpha_j c bit_j(row): one pass over the packed witness (bitrow_comb_kernel) + 4 column transforms replace a
//       pass over the whole LDE (upstream prover/pcs/quotient_ops.rs accumulate_quotients).
#include "common.cuh"
#include "m31_dev.cuh"

#ifndef CONS_MIN_BLOCKS
#define CONS_MIN_BLOCKS 6
#endif
#ifndef CONS2_MIN_BLOCKS
#define CONS2_MIN_BLOCKS 5
#endif
namespace strm {
using namespace m31d;

// acc[row] += sum over jobs of sum_i apr[k] * C(row)   (apr[k] = alpha^(K-1-k), 4 coordinates, split in 16-bit halves)
//   CJ_BOOL: C = b(1-b), b = t0[i], k = kb0 + i*step          (constraints_stream.rs:85-101)
//   CJ_ADDX: one 32-bit adder (:104-131) and, optionally, the xor-rotate that consumes its sum (:134-152): the sum word
//            s = a + b + carry_in - 2 carry is computed here (the adder identity itself is identically zero on the extended
//            domain, kernels_stream.cu header) and stored for later groups; constraints: carry booleans at kbc + 2s, sum
//            booleans at kb1 + s, xor x_i - s_j - d_j + 2 s_j d_j at kx + i (j = i - rot mod 32), result booleans at kb0 + i
//   CJ_XORN: C = a + d - 2ad - r  (keystream xor plaintext = ciphertext, :60-68) with the booleans of r (kb0) and d (kb2)
template <bool HAS_X>
__device__ __forceinline__ void addx_job(const ConstraintJob& J, size_t row, size_t M, const uint4* __restrict__ tlo,
                                         const uint4* __restrict__ thi, AccSplit& A) {
    const uint32_t* x = HAS_X ? J.t0 + row : nullptr;
    const uint32_t* a = J.t1 + row;
    const uint32_t* d = HAS_X ? J.t2 + row : nullptr;
    const uint32_t* b = J.t3 + row;
    const uint32_t* cy = J.t4 + row;
    uint32_t* res = J.res + row;
    uint32_t cin = 0;
    for (int s0 = 0; s0 < 32; s0 += 8) {
        uint32_t av[8], bv[8], cv[8], xv[8], dv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {  // all loads of the 8 bits first (the sum store may alias an operand tile)
            const int s = s0 + u;
            av[u] = a[(size_t)s * M];
            bv[u] = b[(size_t)s * M];
            cv[u] = cy[(size_t)s * M];
            if (HAS_X) {
                xv[u] = x[(size_t)((s + J.arg) & 31) * M];
                dv[u] = d[(size_t)s * M];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int s = s0 + u;
            const uint32_t c2 = cv[u] + cv[u];
            const uint32_t sv = subm(redp(redp(av[u] + bv[u]) + cin), redp(c2));
            cin = cv[u];
            res[(size_t)s * M] = sv;
            const uint32_t s2 = sv + sv;
            A.mac(tlo, thi, J.kbc + 2 * s, bool_c(cv[u], c2));
            A.mac(tlo, thi, J.kb1 + s, bool_c(sv, s2));
            if (HAS_X) {
                const int i = (s + J.arg) & 31;
                const uint64_t v = (uint64_t)dv[u] * s2;  // d * (2s) / 2 = s d (mulw form)
                const uint32_t sd = redp((uint32_t)(v >> 32) + (((uint32_t)v) >> 1));
                const uint32_t up = redp(redp(xv[u] + sd) + sd), dn = redp(sv + dv[u]);
                A.mac(tlo, thi, J.kx + i, up + P - dn);
                A.mac(tlo, thi, J.kb0 + i, bool_c(xv[u], xv[u] + xv[u]));
            }
        }
    }
}

// rows: the first `rows` rows of the tiles are evaluated (M = row count of a tile = column stride)
__global__ void __launch_bounds__(128, CONS_MIN_BLOCKS) constraints_tiles_kernel(ConstraintJobs jobs, size_t M, const uint4* __restrict__ tlo,
                                                                const uint4* __restrict__ thi, uint32_t* __restrict__ acc,
                                                                int first, size_t rows) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    AccSplit A;
    A.init();
    for (int j = 0; j < jobs.n; j++) {
        const ConstraintJob& J = jobs.j[j];
        if (J.type == CJ_ADDX) {
            if (J.t0) addx_job<true>(J, row, M, tlo, thi, A);
            else addx_job<false>(J, row, M, tlo, thi, A);
        } else if (J.type == CJ_BOOL) {
            const uint32_t* __restrict__ t = J.t0 + row;
#pragma unroll 8
            for (int i = 0; i < 32; i++) {
                const uint32_t b = t[(size_t)i * M];
                A.mac(tlo, thi, J.kb0 + i * J.arg, bool_c(b, b + b));
            }
        } else {
            const uint32_t* __restrict__ r = J.t0 + row;
            const uint32_t* __restrict__ a = J.t1 + row;
            const uint32_t* __restrict__ d = J.t2 + row;
            const bool neg = J.type == CJ_XORN;
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
                const int s = (i + 32 - J.arg) & 31;
                const uint32_t rv = r[(size_t)i * M], av = a[(size_t)s * M], dv = d[(size_t)s * M];
                const uint32_t ad = mulm(av, dv);
                uint32_t C = addm(subm(subm(rv, av), dv), dbl(ad));
                if (neg) C = subm(0, C);
                A.mac(tlo, thi, J.kx + i, C);
                if (J.kb0 >= 0) A.mac(tlo, thi, J.kb0 + i, bool_c(rv, rv + rv));
                if (J.kb1 >= 0) A.mac(tlo, thi, J.kb1 + s, bool_c(av, av + av));
                if (J.kb2 >= 0) A.mac(tlo, thi, J.kb2 + s, bool_c(dv, dv + dv));
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t v = A.result(c);
        uint32_t* o = acc + (size_t)c * M + row;
        if (!first) v = addm(v, *o);
        *o = v;
    }
}

// ---- v2: the same sum with the multiply-accumulate on the FP64 pipe ---------------------------------------------------------
// Measured on B200 (profiles/int_peak_r02.json): IMAD.WIDE issues at 25 thread-instructions/clk/SM, DFMA at 58-63 on its own
// pipe next to the integer pipes.  A constraint value C (< 2^32) times a 16-bit half of an alpha-power coordinate is < 2^48, so a
// double holds the exact sum of 32 such products (2^53): the 8 IMAD.WIDE of AccSplit::mac become 8 DFMA, the integer pipes keep
// only the constraint arithmetic itself, and the accumulators are folded into integers once per 32 constraints.
// The alpha table of one launch is pre-arranged in consumption order (cons_table_kernel: 8 doubles per constraint = low and high
// halves of the 4 coordinates) and every job's slice (<= 8 KB) is staged in shared memory once per block, so a constraint costs
// four broadcast LDS.128 with immediate offsets instead of two global LDG.128 with computed addresses.
struct AccF64 {
    double a[8];            // lo halves x 4 coordinates, hi halves x 4 coordinates
    uint64_t lo[4], hi[4];  // folded integer sums
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int c = 0; c < 4; c++) { lo[c] = hi[c] = 0; a[c] = a[4 + c] = 0.0; }
    }
    __device__ __forceinline__ void mac(const double* __restrict__ e, uint32_t C) {
        const double c = __uint2double_rn(C);
        const double2 l01 = *(const double2*)(e), l23 = *(const double2*)(e + 2), h01 = *(const double2*)(e + 4), h23 = *(const double2*)(e + 6);
        a[0] = fma(c, l01.x, a[0]); a[1] = fma(c, l01.y, a[1]); a[2] = fma(c, l23.x, a[2]); a[3] = fma(c, l23.y, a[3]);
        a[4] = fma(c, h01.x, a[4]); a[5] = fma(c, h01.y, a[5]); a[6] = fma(c, h23.x, a[6]); a[7] = fma(c, h23.y, a[7]);
    }
    __device__ __forceinline__ void fold() {  // after at most 32 mac() calls
#pragma unroll
        for (int c = 0; c < 4; c++) {
            lo[c] += __double2ull_rz(a[c]);
            hi[c] += __double2ull_rz(a[4 + c]);
            a[c] = a[4 + c] = 0.0;
        }
    }
    __device__ __forceinline__ uint32_t result(int c) const { return addm(mulm(red64(hi[c]), 1u << 16), red64(lo[c])); }
};

// one adder (+ xor-rotate) job; table slice: step s -> [carry boolean, sum boolean(, xor, result boolean)] x 8 doubles.
// The 32 steps run in batches of SB (all loads of a batch are issued before its arithmetic).  CONS2_PREFETCH=1 issues the loads
// of batch k+1 before the arithmetic of batch k (two register sets, ping-pong): measured SLOWER at every (SB, blocks per SM)
// tried (70-78 ms vs 68 ms per proof at log 20, profiles/r02_summary.md) -- the extra registers cost more resident warps than
// the earlier loads win; kept as a build switch.  No tile of a job aliases its sum tile (own slot in the arena).
#ifndef CONS2_SB
#define CONS2_SB 4
#endif
#ifndef CONS2_PREFETCH
#define CONS2_PREFETCH 0
#endif
template <bool HAS_X>
struct AddxBatch {
    uint32_t av[CONS2_SB], bv[CONS2_SB], cv[CONS2_SB], xv[CONS2_SB], dv[CONS2_SB];
    __device__ __forceinline__ void load(const uint32_t* x, const uint32_t* a, const uint32_t* d, const uint32_t* b, const uint32_t* cy,
                                         size_t M, int s0, int rot) {
#pragma unroll
        for (int u = 0; u < CONS2_SB; u++) {
            const int s = s0 + u;
            av[u] = a[(size_t)s * M];
            bv[u] = b[(size_t)s * M];
            cv[u] = cy[(size_t)s * M];
            if (HAS_X) {
                xv[u] = x[(size_t)((s + rot) & 31) * M];
                dv[u] = d[(size_t)s * M];
            }
        }
    }
    __device__ __forceinline__ void compute(uint32_t* res, size_t M, int s0, uint32_t& cin, const double* __restrict__ tab, AccF64& A) const {
        constexpr int NE = HAS_X ? 4 : 2;
        const double* __restrict__ e = tab + (size_t)s0 * NE * 8;
#pragma unroll
        for (int u = 0; u < CONS2_SB; u++) {
            const int s = s0 + u;
            const uint32_t c2 = cv[u] + cv[u];
            const uint32_t sv = subm(redp(redp(av[u] + bv[u]) + cin), redp(c2));
            cin = cv[u];
            res[(size_t)s * M] = sv;
            const uint32_t s2 = sv + sv;
            A.mac(e + (u * NE) * 8, bool_c(cv[u], c2));
            A.mac(e + (u * NE + 1) * 8, bool_c(sv, s2));
            if (HAS_X) {
                const uint64_t v = (uint64_t)dv[u] * s2;  // d * (2s) / 2 = s d (mulw form)
                const uint32_t sd = redp((uint32_t)(v >> 32) + (((uint32_t)v) >> 1));
                const uint32_t up = redp(redp(xv[u] + sd) + sd), dn = redp(sv + dv[u]);
                A.mac(e + (u * NE + 2) * 8, up + P - dn);
                A.mac(e + (u * NE + 3) * 8, bool_c(xv[u], xv[u] + xv[u]));
            }
        }
    }
};

template <bool HAS_X>
__device__ __forceinline__ void addx_job2(const ConstraintJob& J, size_t row, size_t M, const double* __restrict__ tab, AccF64& A) {
    const uint32_t* x = HAS_X ? J.t0 + row : nullptr;
    const uint32_t* a = J.t1 + row;
    const uint32_t* d = HAS_X ? J.t2 + row : nullptr;
    const uint32_t* b = J.t3 + row;
    const uint32_t* cy = J.t4 + row;
    uint32_t* res = J.res + row;
    constexpr int SB = CONS2_SB;
    constexpr int FOLD = (HAS_X ? 8 : 16);  // steps between folds: at most 32 products per accumulator
    uint32_t cin = 0;
#if CONS2_PREFETCH
    AddxBatch<HAS_X> B0, B1;
    B0.load(x, a, d, b, cy, M, 0, J.arg);
#pragma unroll 1
    for (int s0 = 0; s0 < 32; s0 += 2 * SB) {
        B1.load(x, a, d, b, cy, M, s0 + SB, J.arg);
        B0.compute(res, M, s0, cin, tab, A);
        if ((s0 + SB) % FOLD == 0) A.fold();
        if (s0 + 2 * SB < 32) B0.load(x, a, d, b, cy, M, s0 + 2 * SB, J.arg);
        B1.compute(res, M, s0 + SB, cin, tab, A);
        if ((s0 + 2 * SB) % FOLD == 0) A.fold();
    }
#else
#pragma unroll 1
    for (int s0 = 0; s0 < 32; s0 += SB) {
        AddxBatch<HAS_X> B;
        B.load(x, a, d, b, cy, M, s0, J.arg);
        B.compute(res, M, s0, cin, tab, A);
        if ((s0 + SB) % FOLD == 0) A.fold();
    }
#endif
}

// jobs.j[k].kx = offset (in constraints) of job k's slice inside `gtab`; slice sizes: CJ_BOOL 32, CJ_ADDX 128 (64 without the xor
// part), CJ_XOR/CJ_XORN 128 entries of 8 doubles
// one job for one row: stage the job's slice of the alpha table (all threads of the block), then accumulate its constraints
__device__ __forceinline__ void cons_job2(const ConstraintJob& J, size_t row, size_t M, const double* __restrict__ gtab, double* stab,
                                          AccF64& A, bool live) {
    const int n_e = J.type == CJ_BOOL ? 32 : (J.type == CJ_ADDX && !J.t0 ? 64 : 128);
    __syncthreads();
    {
        const double2* __restrict__ src = (const double2*)(gtab + (size_t)J.kx * 8);
        for (int i = threadIdx.x; i < n_e * 4; i += blockDim.x) ((double2*)stab)[i] = __ldg(src + i);
    }
    __syncthreads();
    if (!live) return;
    if (J.type == CJ_ADDX) {
        if (J.t0) addx_job2<true>(J, row, M, stab, A);
        else addx_job2<false>(J, row, M, stab, A);
    } else if (J.type == CJ_BOOL) {
        const uint32_t* __restrict__ t = J.t0 + row;
#pragma unroll 8
        for (int i = 0; i < 32; i++) {
            const uint32_t b = t[(size_t)i * M];
            A.mac(stab + i * 8, bool_c(b, b + b));
        }
        A.fold();
    } else {
        const uint32_t* __restrict__ r = J.t0 + row;
        const uint32_t* __restrict__ a = J.t1 + row;
        const uint32_t* __restrict__ d = J.t2 + row;
        const bool neg = J.type == CJ_XORN;
#pragma unroll 1
        for (int i0 = 0; i0 < 32; i0 += 8) {
#pragma unroll 4
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u, s = (i + 32 - J.arg) & 31;
                const uint32_t rv = r[(size_t)i * M], av = a[(size_t)s * M], dv = d[(size_t)s * M];
                const uint32_t ad = mulm(av, dv);
                uint32_t C = addm(subm(subm(rv, av), dv), dbl(ad));
                if (neg) C = subm(0, C);
                const double* e = stab + (i * 4) * 8;  // absent booleans have an all-zero table entry
                A.mac(e, C);
                A.mac(e + 8, bool_c(rv, rv + rv));
                A.mac(e + 16, bool_c(av, av + av));
                A.mac(e + 24, bool_c(dv, dv + dv));
            }
            A.fold();
        }
    }
}

__global__ void __launch_bounds__(128, CONS2_MIN_BLOCKS) constraints_tiles_kernel2(ConstraintJobs jobs, size_t M, const double* __restrict__ gtab,
                                                                 uint32_t* __restrict__ acc, int first, size_t rows) {
    __shared__ __align__(16) double stab[128 * 8];
    const size_t row0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = row0 < rows;
    const size_t row = live ? row0 : 0;  // idle threads of the last block still help staging (their results are not stored)
    AccF64 A;
    A.init();
    for (int j = 0; j < jobs.n; j++) cons_job2(jobs.j[j], row, M, gtab, stab, A, live);
    if (!live) return;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t v = A.result(c);
        uint32_t* o = acc + (size_t)c * M + row;
        if (!first) v = addm(v, *o);
        *o = v;
    }
}

// Product-size traces (a few dozen to a few thousand rows): one thread per row walking all jobs is a long dependent chain on a
// nearly empty GPU.  Here blockIdx.y = job (the whole AIR in ONE launch, job list in device memory) and every block writes its
// job's partial sums, partial[(job * 4 + c) * rows + row]; cons_partial_reduce_kernel adds them up.
__global__ void __launch_bounds__(128) constraints_jobs_kernel(const ConstraintJob* __restrict__ jobs, size_t M, const double* __restrict__ gtab,
                                                               uint32_t* __restrict__ partial, size_t rows) {
    __shared__ __align__(16) double stab[128 * 8];
    const size_t row0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = row0 < rows;
    const size_t row = live ? row0 : 0;
    const ConstraintJob J = jobs[blockIdx.y];
    AccF64 A;
    A.init();
    cons_job2(J, row, M, gtab, stab, A, live);
    if (!live) return;
#pragma unroll
    for (int c = 0; c < 4; c++) partial[((size_t)blockIdx.y * 4 + c) * rows + row] = A.result(c);
}
__global__ void cons_partial_reduce_kernel(const uint32_t* __restrict__ partial, int n_jobs, size_t rows, size_t M, uint32_t* __restrict__ acc) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 4 * rows) return;
    const size_t c = idx / rows, row = idx % rows;
    uint32_t v = 0;
    for (int j = 0; j < n_jobs; j++) v = addm(v, partial[((size_t)j * 4 + c) * rows + row]);
    acc[c * M + row] = v;
}

// Adder-sum tiles of a fully materialised LDE (product-size traces): tile(res) = tile(a) + tile(b) + carry-in(c) - 2 tile(c),
// in list order; tile(w) = arena + w * tile_words, [32][rows].  thread = (bit, row): it only reads sums it wrote itself.
__global__ void sum_tiles_kernel(uint32_t* __restrict__ arena, size_t tile_words, size_t rows, const SumComb* __restrict__ combs, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 32 * rows) return;
    const size_t i = idx / rows;
    for (int t = 0; t < n; t++) {
        const SumComb cb = combs[t];
        const uint32_t* cy = arena + (size_t)cb.c * tile_words;
        const uint32_t cv = cy[idx], cin = i ? cy[idx - rows] : 0;
        const uint32_t av = arena[(size_t)cb.a * tile_words + idx], bv = arena[(size_t)cb.b * tile_words + idx];
        arena[(size_t)cb.res * tile_words + idx] = subm(addm(addm(av, bv), cin), dbl(cv));
    }
}

// gtab[e][0..3] = low 16 bits, gtab[e][4..7] = high bits of the 4 coordinates of apr[idx[e]] as doubles (idx < 0: zeros)
__global__ void cons_table_kernel(const uint4* __restrict__ apr, const int* __restrict__ idx, int n, double* __restrict__ gtab) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int k = idx[e];
    const uint4 v = k >= 0 ? apr[k] : make_uint4(0, 0, 0, 0);
    double* o = gtab + (size_t)e * 8;
    o[0] = (double)(v.x & 0xffffu); o[1] = (double)(v.y & 0xffffu); o[2] = (double)(v.z & 0xffffu); o[3] = (double)(v.w & 0xffffu);
    o[4] = (double)(v.x >> 16); o[5] = (double)(v.y >> 16); o[6] = (double)(v.z >> 16); o[7] = (double)(v.w >> 16);
}

// split a table of QM31 values (4 words each) into 16-bit halves
__global__ void split16_kernel(const uint4* __restrict__ t, int n, uint4* __restrict__ lo, uint4* __restrict__ hi) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint4 v = t[k];
    lo[k] = make_uint4(v.x & 0xffffu, v.y & 0xffffu, v.z & 0xffffu, v.w & 0xffffu);
    hi[k] = make_uint4(v.x >> 16, v.y >> 16, v.z >> 16, v.w >> 16);
}

// acc[c][row] *= den_inv[row >> trace_log]
__global__ void scale_rows_kernel(uint32_t* __restrict__ acc, size_t M, int trace_log, const uint32_t* __restrict__ den_inv,
                                  size_t row0) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    const uint32_t d = __ldg(den_inv + ((row0 + row) >> trace_log));  // row0: first global row of this rank's row shard
#pragma unroll
    for (int c = 0; c < 4; c++) acc[(size_t)c * M + row] = mulm(acc[(size_t)c * M + row], d);
}

// out[(word*32 + bit)*4 + c] = scale * sum_r bit(W[word][r]) * wt[c][r]     (one block per witness word)
// (masked adds with a one-instruction reduction each; a 64-bit IMAD.WIDE accumulation was measured 40 % slower here:
// 121 registers and IMAD.WIDE issuing at half rate)
// `words` (optional): the witness words to process, one block each (the adder-sum words are skipped by the callers: their
// values follow from their operands' by linearity); output rows are indexed by the word number.
// blockIdx.y = row slice: a block covers rows [slice * rows_per_slice, +rows_per_slice) and, when there are several slices,
// writes an unscaled partial sum to out[slice][1040 * 128] (summed and scaled by bitcol_reduce_kernel).  Slicing keeps all
// SMs busy when few words are visited (62 words for the query values) and trims the last partial wave.
__global__ void __launch_bounds__(256) bitcol_dot_kernel(const uint32_t* __restrict__ W, size_t N, const uint32_t* __restrict__ wt,
                                                         uint32_t scale, uint32_t* __restrict__ out, const int* __restrict__ words,
                                                         size_t rows_per_slice, size_t slice_stride) {
    const int lane = threadIdx.x & 63, bg = threadIdx.x >> 6;
    const size_t word = words ? (size_t)words[blockIdx.x] : (size_t)blockIdx.x;
    const uint32_t* __restrict__ wrow = W + word * N;
    const size_t r_begin = (size_t)blockIdx.y * rows_per_slice;
    const size_t r_end = r_begin + rows_per_slice < N ? r_begin + rows_per_slice : N;
    uint32_t acc[8][4];
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[b][c] = 0;
    for (size_t r = r_begin + lane; r < r_end; r += 64) {
        const uint32_t w = __ldg(wrow + r) >> (8 * bg);
        uint32_t t[4];
#pragma unroll
        for (int c = 0; c < 4; c++) t[c] = __ldg(wt + (size_t)c * N + r);
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const uint32_t mask = 0u - ((w >> b) & 1u);
#pragma unroll
            for (int c = 0; c < 4; c++) acc[b][c] = redp(acc[b][c] + (mask & t[c]));
        }
    }
    __shared__ uint32_t red[8][32];
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t v = acc[b][c];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v = addm(v, __shfl_xor_sync(0xffffffffu, v, o));
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][b * 4 + c] = v;
        }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int g = threadIdx.x >> 5, x = threadIdx.x & 31;  // x = bit*4 + c
        uint32_t v = addm(red[2 * g][x], red[2 * g + 1][x]);
        if (gridDim.y == 1) v = mulm(v, scale);
        out[(size_t)blockIdx.y * slice_stride + (word * 32 + 8 * g) * 4 + x] = v;
    }
}

// The same masked sums accumulated on the FP64 pipe: the weights are < 2^31, so a double holds the exact sum of 2^22 of them
// (a thread adds N / 64 <= 2^18 rows); a set bit costs one predicated DADD per coordinate instead of mask + add + reduce on the
// ALU pipe (15 ALU-pipe instructions per bit before, ~2 now; DADD issues at 63 thread-instructions/clk/SM next to them).
__global__ void __launch_bounds__(256) bitcol_dot_kernel2(const uint32_t* __restrict__ W, size_t N, const uint32_t* __restrict__ wt,
                                                          uint32_t scale, uint32_t* __restrict__ out, const int* __restrict__ words,
                                                          size_t rows_per_slice, size_t slice_stride) {
    const int lane = threadIdx.x & 63, bg = threadIdx.x >> 6;
    const size_t word = words ? (size_t)words[blockIdx.x] : (size_t)blockIdx.x;
    const uint32_t* __restrict__ wrow = W + word * N;
    const size_t r_begin = (size_t)blockIdx.y * rows_per_slice;
    const size_t r_end = r_begin + rows_per_slice < N ? r_begin + rows_per_slice : N;
    double acc[8][4];
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[b][c] = 0.0;
    for (size_t r = r_begin + lane; r < r_end; r += 64) {
        const uint32_t w = __ldg(wrow + r) >> (8 * bg);
        double t[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {  // converted once per row (volatile: ptxas would otherwise sink the conversion into each
            const uint32_t x = __ldg(wt + (size_t)c * N + r);  // predicated add, 8 conversions instead of 1 on the slow XU pipe)
            asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(t[c]) : "r"(x));
        }
#pragma unroll
        for (int b = 0; b < 8; b++) {
            if ((w >> b) & 1u) {
#pragma unroll
                for (int c = 0; c < 4; c++) acc[b][c] += t[c];
            }
        }
    }
    __shared__ uint32_t red[8][32];
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t v = red64(__double2ull_rz(acc[b][c]));
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v = addm(v, __shfl_xor_sync(0xffffffffu, v, o));
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][b * 4 + c] = v;
        }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int g = threadIdx.x >> 5, x = threadIdx.x & 31;  // x = bit*4 + c
        uint32_t v = addm(red[2 * g][x], red[2 * g + 1][x]);
        if (gridDim.y == 1) v = mulm(v, scale);
        out[(size_t)blockIdx.y * slice_stride + (word * 32 + 8 * g) * 4 + x] = v;
    }
}

// v3: the FP64 form with the rows staged through shared memory.  v2 was latency-bound (ncu: 12.7 warps stalled on the long
// scoreboard per issue, 2 blocks of 95 registers per SM): every thread waited for its own five loads before 32 short adds.  Here
// the block loads a tile of 256 rows cooperatively (one row per thread, the four weights converted to double once per row instead
// of once per byte group), the next tile's loads are in flight while the current one is summed, and the adds read shared memory.
__global__ void __launch_bounds__(256) bitcol_dot_kernel3(const uint32_t* __restrict__ W, size_t N, const uint32_t* __restrict__ wt,
                                                          uint32_t scale, uint32_t* __restrict__ out, const int* __restrict__ words,
                                                          size_t rows_per_slice, size_t slice_stride) {
    __shared__ uint32_t s_w[2][256];
    __shared__ double s_t[2][4][256];
    const int lane = threadIdx.x & 63, bg = threadIdx.x >> 6;
    const size_t word = words ? (size_t)words[blockIdx.x] : (size_t)blockIdx.x;
    const uint32_t* __restrict__ wrow = W + word * N;
    const size_t r_begin = (size_t)blockIdx.y * rows_per_slice;
    const size_t r_end = r_begin + rows_per_slice < N ? r_begin + rows_per_slice : N;
    double acc[8][4];
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[b][c] = 0.0;
    uint32_t pw = 0, pt[4] = {0, 0, 0, 0};
    auto fetch = [&](size_t r0) {
        const size_t r = r0 + threadIdx.x;
        if (r < r_end) {
            pw = __ldg(wrow + r);
#pragma unroll
            for (int c = 0; c < 4; c++) pt[c] = __ldg(wt + (size_t)c * N + r);
        } else {
            pw = 0;  // rows beyond the slice contribute nothing
        }
    };
    fetch(r_begin);
    int buf = 0;
    for (size_t r0 = r_begin; r0 < r_end; r0 += 256, buf ^= 1) {
        s_w[buf][threadIdx.x] = pw;
#pragma unroll
        for (int c = 0; c < 4; c++) s_t[buf][c][threadIdx.x] = __uint2double_rn(pt[c]);
        __syncthreads();
        if (r0 + 256 < r_end) fetch(r0 + 256);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int row = lane + 64 * j;
            const uint32_t w = s_w[buf][row] >> (8 * bg);
            const double t0 = s_t[buf][0][row], t1 = s_t[buf][1][row], t2 = s_t[buf][2][row], t3 = s_t[buf][3][row];
#pragma unroll
            for (int b = 0; b < 8; b++) {
                if ((w >> b) & 1u) {
                    acc[b][0] += t0; acc[b][1] += t1; acc[b][2] += t2; acc[b][3] += t3;
                }
            }
        }
    }
    __shared__ uint32_t red[8][32];
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t v = red64(__double2ull_rz(acc[b][c]));
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v = addm(v, __shfl_xor_sync(0xffffffffu, v, o));
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][b * 4 + c] = v;
        }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int g = threadIdx.x >> 5, x = threadIdx.x & 31;  // x = bit*4 + c
        uint32_t v = addm(red[2 * g][x], red[2 * g + 1][x]);
        if (gridDim.y == 1) v = mulm(v, scale);
        out[(size_t)blockIdx.y * slice_stride + (word * 32 + 8 * g) * 4 + x] = v;
    }
}

// out[word][..] = scale * sum over slices of partial[slice][word][..] for the listed words
__global__ void bitcol_reduce_kernel(const uint32_t* __restrict__ partial, size_t slice_stride, int slices, uint32_t scale,
                                     const int* __restrict__ words, int n_words, uint32_t* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_words * 128) return;
    const size_t word = words ? (size_t)words[idx >> 7] : (size_t)(idx >> 7);
    const size_t o = word * 128 + (idx & 127);
    uint32_t v = 0;
    for (int s = 0; s < slices; s++) v = addm(v, partial[(size_t)s * slice_stride + o]);
    out[o] = mulm(v, scale);
}

// g[c][r] = sum over words w, bits b of bit(W[w][r], b) * coefs[w*32+b][c]     block = 64 rows x PARTS word-slices
// (64-bit IMAD.WIDE accumulation as above: at most 33,280 terms)
__global__ void __launch_bounds__(1024) bitrow_comb_kernel(const uint32_t* __restrict__ W, size_t N, int n_words,
                                                           const uint4* __restrict__ coefs, uint32_t* __restrict__ g,
                                                           const int* __restrict__ words) {
    const int rl = threadIdx.x & 63, part = threadIdx.x >> 6, parts = blockDim.x >> 6;
    const size_t r = (size_t)blockIdx.x * 64 + rl;
    uint64_t acc[4] = {0, 0, 0, 0};
    if (r < N) {
        for (int wi = part; wi < n_words; wi += parts) {
            const int w = words ? words[wi] : wi;
            const uint32_t word = __ldg(W + (size_t)w * N + r);
            const uint4* __restrict__ cf = coefs + (size_t)w * 32;
#pragma unroll 8
            for (int b = 0; b < 32; b++) {
                const uint4 c4 = __ldg(cf + b);
                const uint32_t bit = (word >> b) & 1u;
                acc[0] += (uint64_t)bit * c4.x;
                acc[1] += (uint64_t)bit * c4.y;
                acc[2] += (uint64_t)bit * c4.z;
                acc[3] += (uint64_t)bit * c4.w;
            }
        }
    }
    __shared__ uint32_t red[16][4][64];
#pragma unroll
    for (int c = 0; c < 4; c++) red[part][c][rl] = red64(acc[c]);
    __syncthreads();
    if (threadIdx.x < 256) {
        const int c = threadIdx.x >> 6;
        uint32_t v = 0;
        for (int p = 0; p < parts; p++) v = addm(v, red[p][c][rl]);
        if (r < N) g[(size_t)c * N + r] = v;
    }
}

// ---- the same row-wise combination by byte tables ("four Russians"): for every visited word and each of its 4 bytes a
// table of the 256 possible coefficient sums (QM31) is built once; a row then costs 4 shared-memory lookups and 16 64-bit
// additions per word instead of 32 bit extractions and 128 multiply-adds.
// T[(wi*4 + k)*256 + v] = sum over set bits i of v of coefs[word*32 + 8k + i]
__global__ void bitrow_tables_kernel(const uint4* __restrict__ coefs, const int* __restrict__ words, int n_words, uint4* __restrict__ T) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_words * 1024) return;
    const int v = idx & 255, k = (idx >> 8) & 3, wi = idx >> 10;
    const int w = words ? words[wi] : wi;
    const uint4* __restrict__ cf = coefs + (size_t)w * 32 + 8 * k;
    uint32_t a[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if ((v >> i) & 1) {
            const uint4 c4 = __ldg(cf + i);
            a[0] = addm(a[0], c4.x); a[1] = addm(a[1], c4.y); a[2] = addm(a[2], c4.z); a[3] = addm(a[3], c4.w);
        }
    }
    T[idx] = make_uint4(a[0], a[1], a[2], a[3]);
}

// block = 256 threads x BR_ROWS rows each, blockIdx.y = slice of the visited words; out[part][c][r] (reduced mod p)
constexpr int BR_ROWS = 4;
__global__ void __launch_bounds__(256) bitrow_lookup_kernel(const uint32_t* __restrict__ W, size_t N, const int* __restrict__ words,
                                                            int n_words, int words_per_part, const uint4* __restrict__ T,
                                                            uint32_t* __restrict__ out) {
    __shared__ uint4 tab[1024];
    const size_t r0 = (size_t)blockIdx.x * (256 * BR_ROWS) + threadIdx.x;
    const int w_begin = blockIdx.y * words_per_part;
    const int w_end = w_begin + words_per_part < n_words ? w_begin + words_per_part : n_words;
    uint64_t acc[BR_ROWS][4];
#pragma unroll
    for (int j = 0; j < BR_ROWS; j++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[j][c] = 0;
    // the next word's table (16 KB from L2) is fetched into registers while the current word's rows are processed
    uint4 nxt[4];
    if (w_begin < w_end) {
#pragma unroll
        for (int e = 0; e < 4; e++) nxt[e] = __ldg(T + (size_t)w_begin * 1024 + e * 256 + threadIdx.x);
    }
    for (int wi = w_begin; wi < w_end; wi++) {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; e++) tab[e * 256 + threadIdx.x] = nxt[e];
        __syncthreads();
        if (wi + 1 < w_end) {
#pragma unroll
            for (int e = 0; e < 4; e++) nxt[e] = __ldg(T + (size_t)(wi + 1) * 1024 + e * 256 + threadIdx.x);
        }
        const uint32_t* __restrict__ wrow = W + (size_t)(words ? words[wi] : wi) * N;
#pragma unroll
        for (int j = 0; j < BR_ROWS; j++) {
            const size_t r = r0 + (size_t)j * 256;
            if (r < N) {
                const uint32_t word = __ldg(wrow + r);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint4 e4 = tab[k * 256 + ((word >> (8 * k)) & 255u)];
                    acc[j][0] += e4.x; acc[j][1] += e4.y; acc[j][2] += e4.z; acc[j][3] += e4.w;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < BR_ROWS; j++) {
        const size_t r = r0 + (size_t)j * 256;
        if (r < N)
#pragma unroll
            for (int c = 0; c < 4; c++) out[((size_t)blockIdx.y * 4 + c) * N + r] = red64(acc[j][c]);
    }
}

// g[c][r] = sum over parts of partial[p][c][r]
__global__ void bitrow_reduce_kernel(const uint32_t* __restrict__ partial, size_t N, int parts, uint32_t* __restrict__ g) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 4 * N) return;
    uint32_t v = 0;
    for (int p = 0; p < parts; p++) v = addm(v, partial[(size_t)p * 4 * N + idx]);
    g[idx] = v;
}

// g[c][r] (+)= sum_j vals[j][r] * coefs[j][c] for plain M31 columns (row-wise QM31 combination of trace-domain values: the FRI
// quotient numerator of a group of columns is the extension of this combination).  block = 64 rows x PARTS column slices.
__global__ void __launch_bounds__(1024) rowcomb_m31_kernel(const uint32_t* __restrict__ vals, size_t stride, int ncols, size_t N,
                                                           const uint4* __restrict__ coefs, uint32_t* __restrict__ g, int accumulate) {
    const int rl = threadIdx.x & 63, part = threadIdx.x >> 6, parts = blockDim.x >> 6;
    const size_t r = (size_t)blockIdx.x * 64 + rl;
    uint64_t acc[4] = {0, 0, 0, 0};
    if (r < N) {
        int pend = 0;
        for (int j = part; j < ncols; j += parts) {
            const uint32_t v = __ldg(vals + (size_t)j * stride + r);
            const uint4 c4 = __ldg(coefs + j);
            acc[0] += (uint64_t)v * c4.x; acc[1] += (uint64_t)v * c4.y; acc[2] += (uint64_t)v * c4.z; acc[3] += (uint64_t)v * c4.w;
            if (++pend == 4) {
#pragma unroll
                for (int c = 0; c < 4; c++) acc[c] = (acc[c] & P) + (acc[c] >> 31);
                pend = 0;
            }
        }
    }
    __shared__ uint32_t red[16][4][64];
#pragma unroll
    for (int c = 0; c < 4; c++) red[part][c][rl] = red64(acc[c]);
    __syncthreads();
    if (threadIdx.x < 256) {
        const int c = threadIdx.x >> 6;
        uint32_t v = 0;
        for (int p = 0; p < parts; p++) v = addm(v, red[p][c][rl]);
        if (r < N) {
            if (accumulate) v = addm(v, g[(size_t)c * N + r]);
            g[(size_t)c * N + r] = v;
        }
    }
}

// out[(w*32 + i)*4 + c] = tile(w)[i][rows[c]] for the words whose LDE tile ([32][M]) is still cached in the arena
__global__ void gather_cached_kernel(const uint32_t* __restrict__ arena, size_t tile_words, size_t M, const int* __restrict__ slot,
                                     int n_words, uint4 rows, int nq, uint32_t* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_words * 128) return;
    const int c = idx & 3, i = (idx >> 2) & 31, w = idx >> 7;
    const int sl = slot[w];
    if (sl < 0 || c >= nq) return;
    const uint32_t r = c == 0 ? rows.x : c == 1 ? rows.y : c == 2 ? rows.z : rows.w;
    if (r == 0xffffffffu) return;  // row-sharded tiles: that row lives on another rank
    out[idx] = arena[(size_t)sl * tile_words + (size_t)i * M + r];
}

// out[word[t] * 32 + i] = tile_t[i][row] for up to MAX_LEAF_GROUPS tiles ([32][M] each)
__global__ void gather_tile_row_kernel(TileRowJobs jobs, size_t M, size_t row, uint32_t* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= jobs.n * 32) return;
    const int t = idx >> 5, i = idx & 31;
    out[(size_t)jobs.word[t] * 32 + i] = jobs.tile[t][(size_t)i * M + row];
}

// component-wise basis doubling for up to 4 base-field points at once: b[c][half+k] = b[c][k] * f[c]
__global__ void basis4_step_kernel(uint32_t* b, size_t stride, uint32_t half, uint4 f) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= half) return;
    b[half + k] = mulm(b[k], f.x);
    b[stride + half + k] = mulm(b[stride + k], f.y);
    b[2 * stride + half + k] = mulm(b[2 * stride + k], f.z);
    b[3 * stride + half + k] = mulm(b[3 * stride + k], f.w);
}

}  // namespace strm

// apr_lo / apr_hi: the reversed alpha-power table split by launch_split16
cudaError_t launch_constraints_tiles(cudaStream_t st, const ConstraintJobs& jobs, size_t M, const uint32_t* apr_lo,
                                     const uint32_t* apr_hi, uint32_t* acc, int first, size_t rows) {
    if (rows == 0 || rows > M) rows = M;
    int threads = rows >= 128 * 148 ? 128 : 32;
    strm::constraints_tiles_kernel<<<(unsigned)((rows + threads - 1) / threads), threads, 0, st>>>(
        jobs, M, (const uint4*)apr_lo, (const uint4*)apr_hi, acc, first, rows);
    return cudaGetLastError();
}

cudaError_t launch_constraints_tiles2(cudaStream_t st, const ConstraintJobs& jobs, size_t M, const double* gtab, uint32_t* acc, int first,
                                      size_t rows) {
    if (rows == 0 || rows > M) rows = M;
    int threads = rows >= 128 * 148 ? 128 : 32;
    strm::constraints_tiles_kernel2<<<(unsigned)((rows + threads - 1) / threads), threads, 0, st>>>(jobs, M, gtab, acc, first, rows);
    return cudaGetLastError();
}

cudaError_t launch_constraints_jobs(cudaStream_t st, const ConstraintJob* jobs_dev, int n_jobs, size_t M, const double* gtab,
                                    uint32_t* partial, uint32_t* acc, size_t rows) {
    const int threads = rows >= 128 ? 128 : 32;
    strm::constraints_jobs_kernel<<<dim3((unsigned)((rows + threads - 1) / threads), n_jobs), threads, 0, st>>>(jobs_dev, M, gtab, partial, rows);
    strm::cons_partial_reduce_kernel<<<(unsigned)((4 * rows + 127) / 128), 128, 0, st>>>(partial, n_jobs, rows, M, acc);
    return cudaGetLastError();
}
cudaError_t launch_sum_tiles(cudaStream_t st, uint32_t* arena, size_t tile_words, size_t rows, const SumComb* combs, int n_combs) {
    strm::sum_tiles_kernel<<<(unsigned)((32 * rows + 127) / 128), 128, 0, st>>>(arena, tile_words, rows, combs, n_combs);
    return cudaGetLastError();
}

cudaError_t launch_cons_table(cudaStream_t st, const uint32_t* apr, const int* idx_dev, int n, double* gtab) {
    strm::cons_table_kernel<<<(n + 255) / 256, 256, 0, st>>>((const uint4*)apr, idx_dev, n, gtab);
    return cudaGetLastError();
}

cudaError_t launch_split16(cudaStream_t st, const uint32_t* table, int n, uint32_t* lo, uint32_t* hi) {
    strm::split16_kernel<<<(n + 255) / 256, 256, 0, st>>>((const uint4*)table, n, (uint4*)lo, (uint4*)hi);
    return cudaGetLastError();
}

cudaError_t launch_scale_rows(cudaStream_t st, uint32_t* acc, size_t M, int trace_log, const uint32_t* den_inv, size_t row0) {
    strm::scale_rows_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(acc, M, trace_log, den_inv, row0);
    return cudaGetLastError();
}

cudaError_t launch_gather_tile_row(cudaStream_t st, const TileRowJobs& jobs, size_t M, size_t row, uint32_t* out) {
    if (jobs.n <= 0) return cudaSuccess;
    strm::gather_tile_row_kernel<<<(jobs.n * 32 + 127) / 128, 128, 0, st>>>(jobs, M, row, out);
    return cudaGetLastError();
}

cudaError_t launch_gather_cached(cudaStream_t st, const uint32_t* arena, size_t tile_words, size_t M, const int* slot_dev, int n_words,
                                 const uint32_t rows[4], int nq, uint32_t* out) {
    if (n_words <= 0) return cudaSuccess;
    strm::gather_cached_kernel<<<(n_words * 128 + 255) / 256, 256, 0, st>>>(arena, tile_words, M, slot_dev, n_words,
                                                                            make_uint4(rows[0], rows[1], rows[2], rows[3]), nq, out);
    return cudaGetLastError();
}

cudaError_t launch_bitcol_dot(cudaStream_t st, const uint32_t* W, size_t N, int n_words, const uint32_t* wt, uint32_t scale,
                              uint32_t* out, const int* words_dev) {
    if (n_words <= 0) return cudaSuccess;
    const size_t slice_stride = (size_t)1040 * 128;  // outputs are indexed by word number (< 1040 witness words)
    // (a byte-bucket variant - shared-memory atomics on 16-bit halves, 32 per row - measured 31 vs 18 ms; not kept)
    int slices = 1;
    while (n_words * slices < 2368 && (N / (2 * slices)) >= 8192) slices *= 2;  // >= 4 waves of 592 resident blocks
    // A/B switch: 1 = integer accumulation (masked add + reduce), 2 = FP64 accumulation straight from global memory,
    // 3 (default) = FP64 accumulation over shared-memory tiles
    static const int ver = getenv("S2C_BITCOL_V") ? atoi(getenv("S2C_BITCOL_V")) : 3;
    if (slices == 1) {
        if (ver == 1) strm::bitcol_dot_kernel<<<n_words, 256, 0, st>>>(W, N, wt, scale, out, words_dev, N, 0);
        else if (ver == 2) strm::bitcol_dot_kernel2<<<n_words, 256, 0, st>>>(W, N, wt, scale, out, words_dev, N, 0);
        else strm::bitcol_dot_kernel3<<<n_words, 256, 0, st>>>(W, N, wt, scale, out, words_dev, N, 0);
        return cudaGetLastError();
    }
    uint32_t* partial = nullptr;
    cudaError_t e = cudaMallocAsync(&partial, slice_stride * slices * 4, st);
    if (e != cudaSuccess) return e;
    if (ver == 1) strm::bitcol_dot_kernel<<<dim3(n_words, slices), 256, 0, st>>>(W, N, wt, scale, partial, words_dev, N / slices, slice_stride);
    else if (ver == 2) strm::bitcol_dot_kernel2<<<dim3(n_words, slices), 256, 0, st>>>(W, N, wt, scale, partial, words_dev, N / slices, slice_stride);
    else strm::bitcol_dot_kernel3<<<dim3(n_words, slices), 256, 0, st>>>(W, N, wt, scale, partial, words_dev, N / slices, slice_stride);
    strm::bitcol_reduce_kernel<<<(n_words * 128 + 255) / 256, 256, 0, st>>>(partial, slice_stride, slices, scale, words_dev, n_words, out);
    e = cudaGetLastError();
    cudaFreeAsync(partial, st);
    return e;
}

cudaError_t launch_bitrow_comb(cudaStream_t st, const uint32_t* W, size_t N, int n_words, const uint32_t* coefs, uint32_t* g,
                               const int* words_dev) {
    static const int direct = getenv("S2C_BITROW_DIRECT") ? 1 : 0;  // A/B switch
    if (N < 4096 || direct) {  // tiny traces: the direct kernel (building 16 KB of tables per word would dominate)
        int parts = N >= 8192 ? 4 : 16;
        strm::bitrow_comb_kernel<<<(unsigned)((N + 63) / 64), 64 * parts, 0, st>>>(W, N, n_words, (const uint4*)coefs, g, words_dev);
        return cudaGetLastError();
    }
    const unsigned row_blocks = (unsigned)((N + 256 * strm::BR_ROWS - 1) / (256 * strm::BR_ROWS));
    int parts = 1;
    while ((size_t)row_blocks * parts < 1776 && parts < 16) parts *= 2;
    const int wpp = (n_words + parts - 1) / parts;
    uint4* T = nullptr;
    uint32_t* partial = nullptr;
    cudaError_t e = cudaMallocAsync(&T, (size_t)n_words * 1024 * sizeof(uint4), st);
    if (e != cudaSuccess) return e;
    if (parts > 1) {
        e = cudaMallocAsync(&partial, (size_t)parts * 4 * N * 4, st);
        if (e != cudaSuccess) return e;
    }
    strm::bitrow_tables_kernel<<<(n_words * 1024 + 255) / 256, 256, 0, st>>>((const uint4*)coefs, words_dev, n_words, T);
    strm::bitrow_lookup_kernel<<<dim3(row_blocks, parts), 256, 0, st>>>(W, N, words_dev, n_words, wpp, T, parts > 1 ? partial : g);
    if (parts > 1) strm::bitrow_reduce_kernel<<<(unsigned)((4 * N + 255) / 256), 256, 0, st>>>(partial, N, parts, g);
    e = cudaGetLastError();
    cudaFreeAsync(T, st);
    if (partial) cudaFreeAsync(partial, st);
    return e;
}

cudaError_t launch_bitrow_reduce(cudaStream_t st, const uint32_t* partial, size_t N, int parts, uint32_t* g) {
    strm::bitrow_reduce_kernel<<<(unsigned)((4 * N + 255) / 256), 256, 0, st>>>(partial, N, parts, g);
    return cudaGetLastError();
}

cudaError_t launch_rowcomb_m31(cudaStream_t st, const uint32_t* vals, size_t stride, int ncols, size_t N, const uint32_t* coefs,
                               uint32_t* g, int accumulate) {
    int parts = N >= 65536 ? 4 : 16;
    strm::rowcomb_m31_kernel<<<(unsigned)((N + 63) / 64), 64 * parts, 0, st>>>(vals, stride, ncols, N, (const uint4*)coefs, g,
                                                                             accumulate);
    return cudaGetLastError();
}

// basis[c][k] = init[c] * prod over set bits j of k of maps[j][c]   (4 independent base-field coordinates)
namespace strm {
struct Basis4Maps { uint4 f[10]; };
__global__ void __launch_bounds__(512) basis4_all_kernel(uint32_t* b, size_t stride, int log_n, uint4 init, Basis4Maps maps) {
    if (threadIdx.x == 0) { b[0] = init.x; b[stride] = init.y; b[2 * stride] = init.z; b[3 * stride] = init.w; }
    __syncthreads();
    for (int j = 0; j < log_n; j++) {
        const uint32_t half = 1u << j;
        const uint4 f = maps.f[j];
        for (uint32_t k = threadIdx.x; k < half; k += blockDim.x) {
            b[half + k] = mulm(b[k], f.x);
            b[stride + half + k] = mulm(b[stride + k], f.y);
            b[2 * stride + half + k] = mulm(b[2 * stride + k], f.z);
            b[3 * stride + half + k] = mulm(b[3 * stride + k], f.w);
        }
        __syncthreads();
    }
}
}  // namespace strm

cudaError_t launch_basis4(cudaStream_t st, uint32_t* basis, size_t stride, int log_n, const uint32_t init[4],
                          const uint32_t (*maps)[4]) {
    if (log_n <= 10) {  // product-size proofs: one launch
        strm::Basis4Maps bm{};
        for (int j = 0; j < log_n; j++) bm.f[j] = make_uint4(maps[j][0], maps[j][1], maps[j][2], maps[j][3]);
        strm::basis4_all_kernel<<<1, 512, 0, st>>>(basis, stride, log_n, make_uint4(init[0], init[1], init[2], init[3]), bm);
        return cudaGetLastError();
    }
    for (int c = 0; c < 4; c++) cudaMemcpyAsync(basis + c * stride, &init[c], 4, cudaMemcpyHostToDevice, st);
    for (int j = 0; j < log_n; j++) {
        uint32_t half = 1u << j;
        uint4 f = make_uint4(maps[j][0], maps[j][1], maps[j][2], maps[j][3]);
        strm::basis4_step_kernel<<<(half + 255) / 256, 256, 0, st>>>(basis, stride, half, f);
    }
    return cudaGetLastError();
}
