// PCS-side kernels: out-of-domain evaluation, FRI quotient accumulation, FRI folds, proof-of-work grind, gathers.
//
// Replace upstream stwo (un-vendored, rev f117d487) behind the C ABI:
//   PolyOps::eval_at_point               prover/backend/{cpu,simd}/circle.rs            -> oods_dot_kernel
//   QuotientOps::accumulate_quotients    prover/pcs/quotient_ops.rs, simd/quotients.rs  -> quotients_kernel
//   FriOps::{fold_circle_into_line,fold_line}  prover/backend/simd/fri.rs               -> fold_*_kernel
//   GrindOps::grind                      prover/backend/simd/grind.rs                   -> grind_kernel
// reached from the reference through stwo::prover::prove (/root/reference/stwo/src/chacha/bitwise/air_stream.rs:226).
#include "common.cuh"
#include "blake2s.cuh"

namespace pcs {
using namespace m31;

__device__ __forceinline__ void fold4(uint64_t a[4]) {
#pragma unroll
    for (int c = 0; c < 4; c++) a[c] = (a[c] & P) + (a[c] >> 31);
}

// basis[k] (4 coordinate arrays of n) = prod over set bits j of k of maps[j]; maps[0]=z.y, maps[1]=z.x, maps[2]=pi(z.x)...
// built by doubling: one launch per bit.
__global__ void basis_step_kernel(uint32_t* b, size_t n_stride, uint32_t half, QM31 f) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= half) return;
    QM31 v{{b[k], b[n_stride + k], b[2 * n_stride + k], b[3 * n_stride + k]}};
    QM31 r = qmul(v, f);
#pragma unroll
    for (int c = 0; c < 4; c++) b[c * n_stride + half + k] = r.v[c];
}

// the same table for n <= 2^10 entries in one launch (product-size proofs: launch count matters more than parallelism)
struct BasisMaps { QM31 f[10]; };
__global__ void __launch_bounds__(512) basis_all_kernel(uint32_t* b, size_t n_stride, int log_n, BasisMaps maps) {
    if (threadIdx.x == 0) { b[0] = 1; b[n_stride] = 0; b[2 * n_stride] = 0; b[3 * n_stride] = 0; }
    __syncthreads();
    for (int j = 0; j < log_n; j++) {
        const uint32_t half = 1u << j;
        for (uint32_t k = threadIdx.x; k < half; k += blockDim.x) {
            QM31 v{{b[k], b[n_stride + k], b[2 * n_stride + k], b[3 * n_stride + k]}};
            QM31 r = qmul(v, maps.f[j]);
#pragma unroll
            for (int c = 0; c < 4; c++) b[c * n_stride + half + k] = r.v[c];
        }
        __syncthreads();
    }
}

// out[col] = sum_k coeffs[col][k] * basis[k]   (M31 x QM31 dot product).  One block per (column, slice): with only a handful
// of columns (the 4 + 4 composition coefficient columns) one block per column left 144 SMs idle (2.2 ms per launch at n = 20).
__global__ void __launch_bounds__(256) oods_dot_kernel(const uint32_t* __restrict__ coeffs, size_t stride, uint32_t n,
                                                       const uint32_t* __restrict__ basis, size_t b_stride,
                                                       uint32_t* __restrict__ out, int slices) {
    const uint32_t col = blockIdx.x / slices, sl = blockIdx.x % slices;
    const uint32_t len = n / slices, k0 = sl * len;
    const uint32_t* c = coeffs + (size_t)col * stride;
    uint64_t a[4] = {0, 0, 0, 0};
    int pend = 0;
    for (uint32_t k = k0 + threadIdx.x; k < k0 + len; k += blockDim.x) {
        uint32_t v = __ldg(c + k);
#pragma unroll
        for (int q = 0; q < 4; q++) a[q] += (uint64_t)v * __ldg(basis + q * b_stride + k);
        if (++pend == 4) { fold4(a); pend = 0; }
    }
    __shared__ uint32_t red[4][256];
#pragma unroll
    for (int q = 0; q < 4; q++) red[q][threadIdx.x] = reduce64_full(a[q]);
    __syncthreads();
    if (threadIdx.x < 4) {
        uint64_t s = 0;
        for (int t = 0; t < 256; t++) s += red[threadIdx.x][t];
        out[(size_t)blockIdx.x * 4 + threadIdx.x] = reduce64_full(s);
    }
}

// out[col][q] = sum over slices of partial[col][slice][q]
__global__ void oods_reduce_kernel(const uint32_t* __restrict__ partial, int n_cols, int slices, uint32_t* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_cols * 4) return;
    const int col = idx >> 2, q = idx & 3;
    uint64_t s = 0;
    for (int t = 0; t < slices; t++) s += partial[((size_t)col * slices + t) * 4 + q];
    out[idx] = reduce64_full(s);
}

// canonic circle-domain point of storage row r (domain log m >= 2) from the twiddle tables
__device__ __forceinline__ void domain_point(const FftTables& tw, int m, uint32_t r, uint32_t& x, uint32_t& y) {
    uint32_t h = r >> 1;
    y = __ldg(tw.Y + (1u << (m - 1)) + h);
    if (r & 1) y = neg(y);
    x = __ldg(tw.X + (1u << (m - 2)) + (h >> 1));
    if (h & 1) x = neg(x);
}


// Single-size version: all columns have the domain's size.  cols: column j at cols + j*stride (uniform matrix) for the first
// n_main columns, then extra columns at extra + (j-n_main)*extra_stride.
__global__ void __launch_bounds__(256) quotients_kernel(const uint32_t* __restrict__ cols, size_t stride, int n_main,
                                                        const uint32_t* __restrict__ extra, size_t extra_stride,
                                                        const QuotBatch* __restrict__ batches, int n_batches, int m, FftTables tw,
                                                        uint32_t* __restrict__ out, size_t out_stride) {
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (1u << m)) return;
    uint32_t x, y;
    domain_point(tw, m, row, x, y);
    QM31 racc = qzero();
    for (int bi = 0; bi < n_batches; bi++) {
        const QuotBatch B = batches[bi];
        uint64_t a[4] = {0, 0, 0, 0};
        int pend = 0;
        const uint4* cf = (const uint4*)B.coefs;
        for (int j = 0; j < B.n_cols; j++) {
            uint32_t v;
            if (B.col_ptr) {
                const int sh = m - (int)B.col_log[j];
                const uint32_t r = sh ? (((row >> (sh + 1)) << 1) | (row & 1)) : row;
                v = __ldg(B.col_ptr[j] + r);
            } else {
                int ci = B.col_idx ? B.col_idx[j] : j;
                v = ci < n_main ? __ldg(cols + (size_t)ci * stride + row) : __ldg(extra + (size_t)(ci - n_main) * extra_stride + row);
            }
            uint4 c4 = __ldg(cf + j);
            a[0] += (uint64_t)v * c4.x; a[1] += (uint64_t)v * c4.y; a[2] += (uint64_t)v * c4.z; a[3] += (uint64_t)v * c4.w;
            if (++pend == 4) { fold4(a); pend = 0; }
        }
        QM31 num{{reduce64_full(a[0]), reduce64_full(a[1]), reduce64_full(a[2]), reduce64_full(a[3])}};
        num = qsub(num, qadd(qmul_m(B.lin_a, y), B.lin_b));
        CM31 den = csub(cmul(csub(B.prx, CM31{x, 0}), B.piy), cmul(csub(B.pry, CM31{y, 0}), B.pix));
        QM31 q = qmul_c(num, cinv(den));
        racc = qadd(qmul(racc, B.batch_coeff), q);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) out[c * out_stride + row] = racc.v[c];
}

// dst[i] = dst[i]*alpha^2 + f0 + alpha*f1, (f0,f1) = ibutterfly(src[2i], src[2i+1], 1/y_i); src on canonic domain log m
__global__ void fold_circle_kernel(const uint32_t* __restrict__ src, size_t s_stride, int m, QM31 alpha, QM31 alpha_sq,
                                   FftTables tw, uint32_t* __restrict__ dst, size_t d_stride, int dst_is_zero) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << (m - 1))) return;
    QM31 a, b;
#pragma unroll
    for (int c = 0; c < 4; c++) { a.v[c] = src[c * s_stride + 2 * i]; b.v[c] = src[c * s_stride + 2 * i + 1]; }
    uint32_t itw = __ldg(tw.IY + (1u << (m - 1)) + i);
    QM31 f0 = qadd(a, b), f1 = qmul_m(qsub(a, b), itw);
    QM31 r = qadd(qmul(f1, alpha), f0);
    if (!dst_is_zero) {
        QM31 d{{dst[i], dst[d_stride + i], dst[2 * d_stride + i], dst[3 * d_stride + i]}};
        r = qadd(qmul(d, alpha_sq), r);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) dst[c * d_stride + i] = r.v[c];
}

// line evaluation on LineDomain(half_odds(L)) (bit-reversed) -> folded evaluation on half_odds(L-1)
__global__ void fold_line_kernel(const uint32_t* __restrict__ src, size_t s_stride, int L, QM31 alpha, FftTables tw,
                                 uint32_t* __restrict__ dst, size_t d_stride) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << (L - 1))) return;
    QM31 a, b;
#pragma unroll
    for (int c = 0; c < 4; c++) { a.v[c] = src[c * s_stride + 2 * i]; b.v[c] = src[c * s_stride + 2 * i + 1]; }
    uint32_t itw = __ldg(tw.IX + (1u << (L - 1)) + i);
    QM31 f0 = qadd(a, b), f1 = qmul_m(qsub(a, b), itw);
    QM31 r = qadd(f0, qmul(f1, alpha));
#pragma unroll
    for (int c = 0; c < 4; c++) dst[c * d_stride + i] = r.v[c];
}

// smallest nonce in [base, base+count) with >= pow_bits trailing zeros of LE-u128(Blake2s(prefixed_digest || nonce_le64))
__global__ void grind_kernel(const uint32_t* __restrict__ prefixed_digest, uint32_t pow_bits, uint64_t base, uint64_t count,
                             unsigned long long* __restrict__ best) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t nonce = base + i;
    uint32_t m[16];
#pragma unroll
    for (int w = 0; w < 8; w++) m[w] = prefixed_digest[w];
    m[8] = (uint32_t)nonce; m[9] = (uint32_t)(nonce >> 32);
#pragma unroll
    for (int w = 10; w < 16; w++) m[w] = 0;
    uint32_t h[8];
    blake2s::init(h);
    blake2s::compress(h, m, 40, true);
    // trailing zeros of the first 16 bytes as a little-endian u128
    uint32_t tz = 0;
    if (h[0]) tz = __ffs(h[0]) - 1;
    else if (h[1]) tz = 32 + __ffs(h[1]) - 1;
    else if (h[2]) tz = 64 + __ffs(h[2]) - 1;
    else if (h[3]) tz = 96 + __ffs(h[3]) - 1;
    else tz = 128;
    if (tz >= pow_bits) atomicMin(best, (unsigned long long)nonce);
}

// out[j*n_rows + q] = col_j[rows[q]]  for uniformly strided columns
__global__ void gather_rows_kernel(const uint32_t* __restrict__ cols, size_t stride, int n_cols, const uint32_t* __restrict__ rows,
                                   int n_rows, uint32_t* __restrict__ out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_cols * n_rows) return;
    int j = idx / n_rows, q = idx % n_rows;
    out[idx] = cols[(size_t)j * stride + rows[q]];
}

// out[q] (8 words) = hashes[idx[q]]
__global__ void gather_hashes_kernel(const uint32_t* __restrict__ hashes, const uint32_t* __restrict__ idx, int n, uint32_t* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 8) return;
    out[t] = hashes[(size_t)idx[t >> 3] * 8 + (t & 7)];
}

// coordinate-wise secure powers table apr[k] = alpha^(K-1-k), one thread per 256-chunk (sequential inside the chunk)
__global__ void secure_powers_rev_kernel(QM31 alpha, QM31 alpha_chunk /* alpha^256 */, int K, uint32_t* __restrict__ apr) {
    int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    int e0 = chunk * 256;
    if (e0 >= K) return;
    QM31 cur = qpow(alpha_chunk, chunk);  // alpha^(256*chunk)
    for (int e = e0; e < min(K, e0 + 256); e++) {
        int k = K - 1 - e;
#pragma unroll
        for (int c = 0; c < 4; c++) apr[(size_t)k * 4 + c] = cur.v[c];
        cur = qmul(cur, alpha);
    }
}

}  // namespace pcs

using m31::QM31;

cudaError_t launch_basis(cudaStream_t st, uint32_t* basis, size_t stride, int log_n, const QM31* maps /*host, log_n entries*/) {
    if (log_n <= 10) {
        pcs::BasisMaps bm{};
        for (int j = 0; j < log_n; j++) bm.f[j] = maps[j];
        pcs::basis_all_kernel<<<1, 512, 0, st>>>(basis, stride, log_n, bm);
        return cudaGetLastError();
    }
    // basis[0] = 1
    uint32_t one[4] = {1, 0, 0, 0};
    for (int c = 0; c < 4; c++) cudaMemcpyAsync(basis + c * stride, &one[c], 4, cudaMemcpyHostToDevice, st);
    for (int j = 0; j < log_n; j++) {
        uint32_t half = 1u << j;
        pcs::basis_step_kernel<<<(half + 255) / 256, 256, 0, st>>>(basis, stride, half, maps[j]);
    }
    return cudaGetLastError();
}

cudaError_t launch_oods_dot(cudaStream_t st, const uint32_t* coeffs, size_t stride, int n_cols, int log_n, const uint32_t* basis,
                            size_t b_stride, uint32_t* out) {
    if (n_cols == 0) return cudaSuccess;
    const uint32_t n = 1u << log_n;
    int slices = 1;
    while (n_cols * slices < 592 && (n / (2 * slices)) >= 4096) slices *= 2;
    if (slices == 1) {
        pcs::oods_dot_kernel<<<n_cols, 256, 0, st>>>(coeffs, stride, n, basis, b_stride, out, 1);
        return cudaGetLastError();
    }
    uint32_t* partial = nullptr;
    cudaError_t e = cudaMallocAsync(&partial, (size_t)n_cols * slices * 16, st);
    if (e != cudaSuccess) return e;
    pcs::oods_dot_kernel<<<n_cols * slices, 256, 0, st>>>(coeffs, stride, n, basis, b_stride, partial, slices);
    pcs::oods_reduce_kernel<<<(n_cols * 4 + 127) / 128, 128, 0, st>>>(partial, n_cols, slices, out);
    e = cudaGetLastError();
    cudaFreeAsync(partial, st);
    return e;
}

cudaError_t launch_quotients(cudaStream_t st, const uint32_t* cols, size_t stride, int n_main, const uint32_t* extra,
                             size_t extra_stride, const void* batches_dev, int n_batches, int m, const FftTables& tw,
                             uint32_t* out, size_t out_stride) {
    uint32_t rows = 1u << m;
    int threads = rows >= 148 * 256 ? 256 : 64;
    pcs::quotients_kernel<<<(rows + threads - 1) / threads, threads, 0, st>>>(cols, stride, n_main, extra, extra_stride,
                                                                             (const QuotBatch*)batches_dev, n_batches, m, tw,
                                                                             out, out_stride);
    return cudaGetLastError();
}

cudaError_t launch_fold_circle(cudaStream_t st, const uint32_t* src, size_t s_stride, int m, QM31 alpha, const FftTables& tw,
                               uint32_t* dst, size_t d_stride, int dst_is_zero) {
    uint32_t n = 1u << (m - 1);
    pcs::fold_circle_kernel<<<(n + 127) / 128, 128, 0, st>>>(src, s_stride, m, alpha, m31::qmul(alpha, alpha), tw, dst, d_stride,
                                                             dst_is_zero);
    return cudaGetLastError();
}

cudaError_t launch_fold_line(cudaStream_t st, const uint32_t* src, size_t s_stride, int L, QM31 alpha, const FftTables& tw,
                             uint32_t* dst, size_t d_stride) {
    uint32_t n = 1u << (L - 1);
    pcs::fold_line_kernel<<<(n + 127) / 128, 128, 0, st>>>(src, s_stride, L, alpha, tw, dst, d_stride);
    return cudaGetLastError();
}

cudaError_t launch_grind(cudaStream_t st, const uint32_t* prefixed_digest_dev, uint32_t pow_bits, uint64_t base, uint64_t count,
                         unsigned long long* best_dev) {
    pcs::grind_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(prefixed_digest_dev, pow_bits, base, count, best_dev);
    return cudaGetLastError();
}

cudaError_t launch_gather_rows(cudaStream_t st, const uint32_t* cols, size_t stride, int n_cols, const uint32_t* rows_dev,
                               int n_rows, uint32_t* out) {
    int total = n_cols * n_rows;
    if (total == 0) return cudaSuccess;
    pcs::gather_rows_kernel<<<(total + 255) / 256, 256, 0, st>>>(cols, stride, n_cols, rows_dev, n_rows, out);
    return cudaGetLastError();
}

cudaError_t launch_gather_hashes(cudaStream_t st, const uint32_t* hashes, const uint32_t* idx_dev, int n, uint32_t* out) {
    if (n == 0) return cudaSuccess;
    pcs::gather_hashes_kernel<<<(n * 8 + 255) / 256, 256, 0, st>>>(hashes, idx_dev, n, out);
    return cudaGetLastError();
}

cudaError_t launch_secure_powers_rev(cudaStream_t st, QM31 alpha, int K, uint32_t* apr) {
    QM31 a256 = m31::qpow(alpha, 256);
    int chunks = (K + 255) / 256;
    pcs::secure_powers_rev_kernel<<<(chunks + 63) / 64, 64, 0, st>>>(alpha, a256, K, apr);
    return cudaGetLastError();
}
