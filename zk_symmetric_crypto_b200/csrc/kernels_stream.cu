// Kernels of the streaming (tile-by-tile) prover: everything that consumes LDE tiles or the packed witness directly.
//
// The ChaCha20 AIR (/root/reference/stwo/src/chacha/bitwise/constraints_stream.rs:20-70) has 33,280 one-bit columns; at
// log_n_rows = 20 its LDE is 279 GB, more than one B200 holds.  The prover therefore never materialises the LDE: it keeps
// the packed witness (1,040 words per row) and streams LDE tiles of 32 columns through the consumers below.  Three facts
// about the reference's arithmetic are used (all exact in M31/QM31, so results are bit-identical to the reference's):
//   (1) interpolation and extension are linear.  The sum word of every 32-bit adder satisfies, bit by bit,
//       s_i = a_i + b_i + c_{i-1} - 2 c_i on the trace domain, both sides have degree < N, hence the identity holds on the
//       extended domain too: sum tiles are combined from operand tiles (combine_add_kernel) instead of transformed,
//       and the adder constraints (constraints_stream.rs:117-129) vanish identically on the evaluation domain.
//   (2) f(z) for a column given by its trace-domain values is <values, w(z)> with w(z) = (iFFT)^T basis(z): out-of-domain
//       samples and queried LDE values of all bit columns are masked sums over the packed witness (bitcol_dot_kernel).
//   (3) the FRI quotient numerator sum_j alpha_j c f_j(p) is the extension of the row-wise combination
//       sum_j alpha_j c bit_j(row): one pass over the packed witness (bitrow_comb_kernel) + 4 column transforms replace a
//       pass over the whole LDE (upstream prover/pcs/quotient_ops.rs accumulate_quotients).
#include "common.cuh"
#include "m31_dev.cuh"

namespace strm {
using namespace m31d;

// res_i = a_i + b_i + c_{i-1} - 2 c_i for the 32 bit-columns of a word tile [32][M]; jobs run in order inside one thread
// (a later job may read an earlier job's result of the same row).
__global__ void __launch_bounds__(256) combine_add_kernel(CombineJobs jobs, size_t M) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    for (int j = 0; j < jobs.n; j++) {
        const uint32_t* __restrict__ a = jobs.j[j].a + row;
        const uint32_t* __restrict__ b = jobs.j[j].b + row;
        const uint32_t* __restrict__ c = jobs.j[j].c + row;
        uint32_t* __restrict__ r = jobs.j[j].res + row;
        uint32_t cin = 0;
#pragma unroll 8
        for (int i = 0; i < 32; i++) {
            uint32_t av = a[(size_t)i * M], bv = b[(size_t)i * M], cv = c[(size_t)i * M];
            uint32_t t = addm(addm(av, bv), cin);
            r[(size_t)i * M] = subm(t, dbl(cv));
            cin = cv;
        }
    }
}

struct Acc4 {
    uint64_t a[4];
    int pending;
    __device__ __forceinline__ void init() { a[0] = a[1] = a[2] = a[3] = 0; pending = 0; }
    __device__ __forceinline__ void fold() {
#pragma unroll
        for (int c = 0; c < 4; c++) a[c] = (a[c] & P) + (a[c] >> 31);
        pending = 0;
    }
    __device__ __forceinline__ void mac(uint4 al, uint32_t C) {
        a[0] += (uint64_t)C * al.x; a[1] += (uint64_t)C * al.y; a[2] += (uint64_t)C * al.z; a[3] += (uint64_t)C * al.w;
        if (++pending == 4) fold();  // 4*(p-1)^2 + fold residual < 2^64
    }
};

// acc[row] += sum over jobs of sum_i apr[k] * C(row)   (apr[k] = alpha^(K-1-k), 4 coordinates)
//   CJ_BOOL: C = b(1-b), b = t0[i], k = kb0 + i*step          (constraints_stream.rs:85-101 and the carry booleans :117-120)
//   CJ_XOR : C = r - a - d + 2ad, r = t0[i], a = t1[s], d = t2[s], s = (i-rot) mod 32, k = kx + i   (:134-152), plus the
//            boolean constraints of the operands the job has loaded anyway: r at kb0+i, a at kb1+s, d at kb2+s
//   CJ_XORN: C = a + d - 2ad - r  (keystream xor plaintext = ciphertext, :60-68)
__global__ void __launch_bounds__(128) constraints_tiles_kernel(ConstraintJobs jobs, size_t M, const uint4* __restrict__ apr,
                                                                uint32_t* __restrict__ acc, int first) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    Acc4 A;
    A.init();
    for (int j = 0; j < jobs.n; j++) {
        const ConstraintJob J = jobs.j[j];
        if (J.type == CJ_BOOL) {
            const uint32_t* __restrict__ t = J.t0 + row;
            const uint4* __restrict__ al = apr + J.kb0;
#pragma unroll 8
            for (int i = 0; i < 32; i++) {
                uint32_t b = t[(size_t)i * M];
                A.mac(__ldg(al + i * J.arg), mulm(b, subm(1, b)));
            }
        } else if (J.type == CJ_ADDX) {
            // adder with the sum computed on the fly (and stored for later groups): carry booleans, sum booleans and, when the
            // sum feeds a xor-rotate, the xor constraints and the result booleans.  Operands are staged 8 bits at a time so
            // the loads are in flight together (the sum store may alias an operand tile).
            const uint32_t* x = J.t0 ? J.t0 + row : nullptr;
            const uint32_t* a = J.t1 + row;
            const uint32_t* d = J.t2 ? J.t2 + row : nullptr;
            const uint32_t* b = J.t3 + row;
            const uint32_t* cy = J.t4 + row;
            uint32_t* res = J.res + row;
            uint32_t cin = 0;
            for (int s0 = 0; s0 < 32; s0 += 8) {
                uint32_t av[8], bv[8], cv[8], xv[8], dv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int s = s0 + u;
                    av[u] = a[(size_t)s * M];
                    bv[u] = b[(size_t)s * M];
                    cv[u] = cy[(size_t)s * M];
                    if (x) {
                        xv[u] = x[(size_t)((s + J.arg) & 31) * M];
                        dv[u] = d[(size_t)s * M];
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int s = s0 + u;
                    const uint32_t sv = subm(addm(addm(av[u], bv[u]), cin), dbl(cv[u]));
                    cin = cv[u];
                    res[(size_t)s * M] = sv;
                    A.mac(__ldg(apr + J.kbc + 2 * s), mulm(cv[u], subm(1, cv[u])));
                    A.mac(__ldg(apr + J.kb1 + s), mulm(sv, subm(1, sv)));
                    if (x) {
                        const int i = (s + J.arg) & 31;
                        const uint32_t sd = mulm(sv, dv[u]);
                        A.mac(__ldg(apr + J.kx + i), addm(subm(subm(xv[u], sv), dv[u]), dbl(sd)));
                        A.mac(__ldg(apr + J.kb0 + i), mulm(xv[u], subm(1, xv[u])));
                    }
                }
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
                A.mac(__ldg(apr + J.kx + i), C);
                if (J.kb0 >= 0) A.mac(__ldg(apr + J.kb0 + i), mulm(rv, subm(1, rv)));
                if (J.kb1 >= 0) A.mac(__ldg(apr + J.kb1 + s), mulm(av, subm(1, av)));
                if (J.kb2 >= 0) A.mac(__ldg(apr + J.kb2 + s), mulm(dv, subm(1, dv)));
            }
        }
    }
    A.fold();
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t v = red64(A.a[c]);
        uint32_t* o = acc + (size_t)c * M + row;
        if (!first) v = addm(v, *o);
        *o = v;
    }
}

// acc[c][row] *= den_inv[row >> trace_log]
__global__ void scale_rows_kernel(uint32_t* __restrict__ acc, size_t M, int trace_log, const uint32_t* __restrict__ den_inv) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    const uint32_t d = __ldg(den_inv + (row >> trace_log));
#pragma unroll
    for (int c = 0; c < 4; c++) acc[(size_t)c * M + row] = mulm(acc[(size_t)c * M + row], d);
}

// out[(word*32 + bit)*4 + c] = sum_r bit(W[word][r]) * wt[c][r]     (one block per witness word)
__global__ void __launch_bounds__(256) bitcol_dot_kernel(const uint32_t* __restrict__ W, size_t N, const uint32_t* __restrict__ wt,
                                                         uint32_t scale, uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 63, bg = threadIdx.x >> 6;
    const uint32_t* __restrict__ wrow = W + (size_t)blockIdx.x * N;
    uint32_t acc[8][4];
#pragma unroll
    for (int b = 0; b < 8; b++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[b][c] = 0;
    for (size_t r = lane; r < N; r += 64) {
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
        uint32_t v = mulm(addm(red[2 * g][x], red[2 * g + 1][x]), scale);
        out[((size_t)blockIdx.x * 32 + 8 * g) * 4 + x] = v;
    }
}

// g[c][r] = sum over words w, bits b of bit(W[w][r], b) * coefs[w*32+b][c]     block = 64 rows x PARTS word-slices
__global__ void __launch_bounds__(1024) bitrow_comb_kernel(const uint32_t* __restrict__ W, size_t N, int n_words,
                                                           const uint4* __restrict__ coefs, uint32_t* __restrict__ g) {
    const int rl = threadIdx.x & 63, part = threadIdx.x >> 6, parts = blockDim.x >> 6;
    const size_t r = (size_t)blockIdx.x * 64 + rl;
    uint32_t acc[4] = {0, 0, 0, 0};
    if (r < N) {
        for (int w = part; w < n_words; w += parts) {
            const uint32_t word = __ldg(W + (size_t)w * N + r);
            const uint4* __restrict__ cf = coefs + (size_t)w * 32;
#pragma unroll 8
            for (int b = 0; b < 32; b++) {
                const uint4 c4 = __ldg(cf + b);
                const uint32_t mask = 0u - ((word >> b) & 1u);
                acc[0] = redp(acc[0] + (mask & c4.x));
                acc[1] = redp(acc[1] + (mask & c4.y));
                acc[2] = redp(acc[2] + (mask & c4.z));
                acc[3] = redp(acc[3] + (mask & c4.w));
            }
        }
    }
    __shared__ uint32_t red[16][4][64];
#pragma unroll
    for (int c = 0; c < 4; c++) red[part][c][rl] = acc[c];
    __syncthreads();
    if (threadIdx.x < 256) {
        const int c = threadIdx.x >> 6;
        uint32_t v = 0;
        for (int p = 0; p < parts; p++) v = addm(v, red[p][c][rl]);
        if (r < N) g[(size_t)c * N + r] = v;
    }
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

cudaError_t launch_combine_add(cudaStream_t st, const CombineJobs& jobs, size_t M) {
    if (jobs.n == 0) return cudaSuccess;
    int threads = M >= 256 ? 256 : 32;
    strm::combine_add_kernel<<<(unsigned)((M + threads - 1) / threads), threads, 0, st>>>(jobs, M);
    return cudaGetLastError();
}

cudaError_t launch_constraints_tiles(cudaStream_t st, const ConstraintJobs& jobs, size_t M, const uint32_t* apr, uint32_t* acc,
                                     int first) {
    int threads = M >= 128 * 148 ? 128 : 32;
    strm::constraints_tiles_kernel<<<(unsigned)((M + threads - 1) / threads), threads, 0, st>>>(jobs, M, (const uint4*)apr, acc,
                                                                                               first);
    return cudaGetLastError();
}

cudaError_t launch_scale_rows(cudaStream_t st, uint32_t* acc, size_t M, int trace_log, const uint32_t* den_inv) {
    strm::scale_rows_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(acc, M, trace_log, den_inv);
    return cudaGetLastError();
}

cudaError_t launch_bitcol_dot(cudaStream_t st, const uint32_t* W, size_t N, int n_words, const uint32_t* wt, uint32_t scale,
                              uint32_t* out) {
    strm::bitcol_dot_kernel<<<n_words, 256, 0, st>>>(W, N, wt, scale, out);
    return cudaGetLastError();
}

cudaError_t launch_bitrow_comb(cudaStream_t st, const uint32_t* W, size_t N, int n_words, const uint32_t* coefs, uint32_t* g) {
    int parts = N >= 8192 ? 4 : 16;
    strm::bitrow_comb_kernel<<<(unsigned)((N + 63) / 64), 64 * parts, 0, st>>>(W, N, n_words, (const uint4*)coefs, g);
    return cudaGetLastError();
}

// basis[c][k] = init[c] * prod over set bits j of k of maps[j][c]   (4 independent base-field coordinates)
cudaError_t launch_basis4(cudaStream_t st, uint32_t* basis, size_t stride, int log_n, const uint32_t init[4],
                          const uint32_t (*maps)[4]) {
    for (int c = 0; c < 4; c++) cudaMemcpyAsync(basis + c * stride, &init[c], 4, cudaMemcpyHostToDevice, st);
    for (int j = 0; j < log_n; j++) {
        uint32_t half = 1u << j;
        uint4 f = make_uint4(maps[j][0], maps[j][1], maps[j][2], maps[j][3]);
        strm::basis4_step_kernel<<<(half + 255) / 256, 256, 0, st>>>(basis, stride, half, f);
    }
    return cudaGetLastError();
}
