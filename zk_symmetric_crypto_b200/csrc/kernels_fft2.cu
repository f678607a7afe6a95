// Packed-witness circle FFT for sm_100a: bit/byte-packed witness words -> low-degree extension tiles.
//
// One "job" = one packed witness word row (2^n u32 words, one per trace row) -> 32 (bits) or 4 (bytes) LDE columns of
// 2^(n+1) values on CanonicCoset(n+1).circle_domain(), written as a tile [cols][2^(n+1)].  This is upstream stwo's
// `PolyOps::interpolate_columns` + `evaluate_polynomials` (SimdBackend: prover/backend/simd/{circle.rs,fft/ifft.rs,fft/rfft.rs})
// as the reference reaches them through `TreeBuilder::extend_evals` + `commit`
// (/root/reference/stwo/src/chacha/bitwise/air_stream.rs:210-212), fused: the interpolated coefficients never leave the chip
// for n <= 12 and cross HBM twice (instead of being stored, re-read, zero-extended and re-written) above that.
//
// Schedule (exact field arithmetic => any butterfly order gives the reference's values):
//   * every butterfly layer is applied in registers, 2^R points per thread for R <= 4 consecutive layers
//     ("radix-16 steps"), data exchanged through shared memory between steps (one LDS + one STS per point per step);
//   * twiddles are fetched once per step and thread and reused over all columns of the tile (contiguous kernels) or are
//     warp-uniform broadcast loads (strided kernel);
//   * the 1/2^n scale of the inverse transform is folded into the bit expansion (bit ? 2^-n : 0);
//   * n <= 12: one kernel, column group resident in shared memory;
//     n >= 13: A) inverse layers [0,k1) on contiguous 2^k1 chunks, B) inverse layers [k1,n) + forward layers [n-1..k1] of both
//     halves of the extended domain on strided tiles with >= 32..128-byte segments, C) forward layers [k1-1..0] in place.
// Twiddle tables: see kernels_fft.cu (layer 0 uses Y[2^(m-1)+h], layer i>=1 uses X[2^(m-i-1)+h], h = index >> (i+1)).
#include "common.cuh"
#include "m31_dev.cuh"

namespace fft2 {
using namespace m31d;

__device__ __forceinline__ int padi(int a) { return a + (a >> 4); }

// One register-radix step over shared memory: local bits [b, b+R) of the j index of a tile [ncol][2^jbits][2^qbits].
// Local bit b is global butterfly layer i0+b of a domain of log size m; tile_hi = global index bits above the tile's j bits.
template <int R, bool INV, bool PAD>
__device__ __forceinline__ void radix_step(uint32_t* __restrict__ s, int ncol, int colstride, int jbits, int qbits, int b,
                                           const uint32_t* __restrict__ tabX, const uint32_t* __restrict__ tabY, int m, int i0,
                                           uint32_t tile_hi) {
    constexpr int RR = 1 << R;
    const int n_items = 1 << (jbits - R + qbits);
    const int nthr = blockDim.x;
    const int IT = n_items < nthr ? n_items : nthr;
    const int CG = nthr / IT;  // column groups processed in parallel when a column has fewer items than threads
    const int cg = threadIdx.x / IT;
    for (int item = threadIdx.x % IT; item < n_items; item += IT) {
        const int q = item & ((1 << qbits) - 1);
        const int p = item >> qbits;
        const int j0 = ((p >> b) << (b + R)) | (p & ((1 << b) - 1));
        uint32_t tw[RR];  // slot (c-1)+t for layer l with c = 2^(R-1-l) twiddles; pre-doubled
        const uint32_t jg = (tile_hi << jbits) | (uint32_t)j0;
#pragma unroll
        for (int l = 0; l < R; l++) {
            const int gi = i0 + b + l;
            const uint32_t* tab = (gi == 0) ? tabY + (1u << (m - 1)) : tabX + (1u << (m - gi - 1));
            const uint32_t hb = jg >> (b + l + 1);
            constexpr int dummy = 0;
            (void)dummy;
#pragma unroll
            for (int t = 0; t < (1 << (R - 1 - l)); t++) tw[(1 << (R - 1 - l)) - 1 + t] = __ldg(tab + hb + t) << 1;
        }
        for (int col = cg; col < ncol; col += CG) {
            uint32_t* base = s + col * colstride;
            uint32_t v[RR];
#pragma unroll
            for (int k = 0; k < RR; k++) {
                int a = ((j0 + (k << b)) << qbits) | q;
                v[k] = base[PAD ? padi(a) : a];
            }
            if (INV) {
#pragma unroll
                for (int l = 0; l < R; l++) {
#pragma unroll
                    for (int k = 0; k < RR; k++) {
                        if (k & (1 << l)) continue;
                        const uint32_t w2 = tw[(1 << (R - 1 - l)) - 1 + (k >> (l + 1))];
                        uint32_t v0 = v[k], v1 = v[k | (1 << l)];
                        v[k] = addm(v0, v1);
                        v[k | (1 << l)] = mulw(subm(v0, v1), w2);
                    }
                }
            } else {
#pragma unroll
                for (int l = R - 1; l >= 0; l--) {
#pragma unroll
                    for (int k = 0; k < RR; k++) {
                        if (k & (1 << l)) continue;
                        const uint32_t w2 = tw[(1 << (R - 1 - l)) - 1 + (k >> (l + 1))];
                        uint32_t v0 = v[k], t = mulw(v[k | (1 << l)], w2);
                        v[k] = addm(v0, t);
                        v[k | (1 << l)] = subm(v0, t);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < RR; k++) {
                int a = ((j0 + (k << b)) << qbits) | q;
                base[PAD ? padi(a) : a] = v[k];
            }
        }
    }
}

template <bool INV, bool PAD>
__device__ __forceinline__ void step_dispatch(int R, uint32_t* s, int ncol, int colstride, int jbits, int qbits, int b,
                                              const uint32_t* tabX, const uint32_t* tabY, int m, int i0, uint32_t tile_hi) {
    switch (R) {
        case 4: radix_step<4, INV, PAD>(s, ncol, colstride, jbits, qbits, b, tabX, tabY, m, i0, tile_hi); break;
        case 3: radix_step<3, INV, PAD>(s, ncol, colstride, jbits, qbits, b, tabX, tabY, m, i0, tile_hi); break;
        case 2: radix_step<2, INV, PAD>(s, ncol, colstride, jbits, qbits, b, tabX, tabY, m, i0, tile_hi); break;
        default: radix_step<1, INV, PAD>(s, ncol, colstride, jbits, qbits, b, tabX, tabY, m, i0, tile_hi); break;
    }
}

// Apply local layers [b_lo, b_hi): ascending for the inverse transform, descending for the forward one, in balanced
// register-radix steps of at most 4 layers.  Ends with a __syncthreads().
template <bool INV, bool PAD>
__device__ __forceinline__ void apply_layers(uint32_t* s, int ncol, int colstride, int jbits, int qbits, int b_lo, int b_hi,
                                             const uint32_t* tabX, const uint32_t* tabY, int m, int i0, uint32_t tile_hi) {
    int cnt = b_hi - b_lo;
    int steps = (cnt + 3) >> 2;
    int pos = INV ? b_lo : b_hi;
    while (cnt > 0) {
        int R = (cnt + steps - 1) / steps;
        if (INV) {
            step_dispatch<INV, PAD>(R, s, ncol, colstride, jbits, qbits, pos, tabX, tabY, m, i0, tile_hi);
            pos += R;
        } else {
            pos -= R;
            step_dispatch<INV, PAD>(R, s, ncol, colstride, jbits, qbits, pos, tabX, tabY, m, i0, tile_hi);
        }
        __syncthreads();
        cnt -= R;
        steps--;
    }
}

struct Jobs {
    int n;                    // jobs in this launch (<= MAX_FFT_JOBS)
    const uint32_t* src[MAX_FFT_JOBS];  // packed witness word row (2^n words)
    uint32_t* out[MAX_FFT_JOBS];        // LDE tile [cols_per_job][2^(n+1)]
};

// value of column c of a packed word, times `scale`
template <int KIND>
__device__ __forceinline__ uint32_t unpack(uint32_t word, int c, uint32_t scale, uint32_t scale2) {
    if (KIND == SRC_BITS) return (0u - ((word >> c) & 1u)) & scale;
    return mulw((word >> (8 * c)) & 0xffu, scale2);
}

// ---- n <= 12: whole columns in shared memory --------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) small_kernel(Jobs jobs, int log_n, int nc, int groups_per_job, FftTables tw) {
    extern __shared__ uint32_t s[];
    const int n = 1 << log_n, m = log_n + 1, big = 2 << log_n;
    const int colstride = padi(big);
    const int job = blockIdx.x / groups_per_job, c0 = (blockIdx.x % groups_per_job) * nc;
    const uint32_t* __restrict__ src = jobs.src[job];
    const uint32_t scale = 1u << (31 - log_n), scale2 = scale << 1;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        uint32_t w = __ldg(src + r);
        for (int c = 0; c < nc; c++) s[c * colstride + padi(r)] = unpack<KIND>(w, c0 + c, scale, scale2);
    }
    __syncthreads();
    apply_layers<true, true>(s, nc, colstride, log_n, 0, 0, log_n, tw.IX, tw.IY, log_n, 0, 0);
    for (int idx = threadIdx.x; idx < nc * n; idx += blockDim.x) {
        int c = idx >> log_n, r = idx & (n - 1);
        s[c * colstride + padi(n + r)] = s[c * colstride + padi(r)];
    }
    __syncthreads();
    apply_layers<false, true>(s, nc, colstride, m, 0, 0, log_n, tw.X, tw.Y, m, 0, 0);
    uint32_t* __restrict__ out = jobs.out[job];
    for (int idx = threadIdx.x; idx < nc * big; idx += blockDim.x) {
        int c = idx >> m, r = idx & (big - 1);
        out[(size_t)(c0 + c) * big + r] = s[c * colstride + padi(r)];
    }
}

// ---- n >= 13, pass A: expand + inverse layers [0,k1) on a contiguous chunk of 2^k1 rows ----------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) ifft_low_kernel(Jobs jobs, int log_n, int k1, int nc, int groups_per_job, int cols_per_job,
                                                       uint32_t* __restrict__ scratch, FftTables tw) {
    extern __shared__ uint32_t s[];
    const int T = 1 << k1;
    const int colstride = padi(T);
    const uint32_t chunk = blockIdx.x;
    const int job = blockIdx.y / groups_per_job, c0 = (blockIdx.y % groups_per_job) * nc;
    const uint32_t* __restrict__ src = jobs.src[job] + (size_t)chunk * T;
    const uint32_t scale = 1u << (31 - log_n), scale2 = scale << 1;
    for (int r = threadIdx.x; r < T; r += blockDim.x) {
        uint32_t w = __ldg(src + r);
        for (int c = 0; c < nc; c++) s[c * colstride + padi(r)] = unpack<KIND>(w, c0 + c, scale, scale2);
    }
    __syncthreads();
    apply_layers<true, true>(s, nc, colstride, k1, 0, 0, k1, tw.IX, tw.IY, log_n, 0, chunk);
    uint32_t* __restrict__ out = scratch + ((size_t)(job * cols_per_job + c0) << log_n) + (size_t)chunk * T;
    for (int idx = threadIdx.x; idx < nc * T; idx += blockDim.x) {
        int c = idx >> k1, r = idx & (T - 1);
        out[((size_t)c << log_n) + r] = s[c * colstride + padi(r)];
    }
}

// ---- pass B: inverse layers [k1,n) then forward layers [n-1..k1] of both halves, tile = 2^(n-k1) x Q ----------------------------
template <bool PAD>
__global__ void __launch_bounds__(256) mid_kernel(Jobs jobs, int log_n, int k1, int qbits, int cols_per_job,
                                                  const uint32_t* __restrict__ scratch, FftTables tw) {
    extern __shared__ uint32_t s[];
    const int jb = log_n - k1, J = 1 << jb, Q = 1 << qbits;
    const int tile = J * Q;
    const int tstride = PAD ? padi(tile) : tile;
    const uint32_t q0 = blockIdx.x * Q;
    const int job = blockIdx.y / cols_per_job, col = blockIdx.y % cols_per_job;
    const uint32_t* __restrict__ in = scratch + ((size_t)blockIdx.y << log_n);
    uint32_t* s0 = s;
    uint32_t* s1 = s + tstride;
    for (int idx = threadIdx.x; idx < tile; idx += blockDim.x) {
        int q = idx & (Q - 1), j = idx >> qbits;
        s0[PAD ? padi(idx) : idx] = in[((size_t)j << k1) + q0 + q];
    }
    __syncthreads();
    apply_layers<true, PAD>(s0, 1, 0, jb, qbits, 0, jb, tw.IX, tw.IY, log_n, k1, 0);
    for (int idx = threadIdx.x; idx < tile; idx += blockDim.x) {
        int a = PAD ? padi(idx) : idx;
        s1[a] = s0[a];
    }
    __syncthreads();
    apply_layers<false, PAD>(s0, 1, 0, jb, qbits, 0, jb, tw.X, tw.Y, log_n + 1, k1, 0);
    apply_layers<false, PAD>(s1, 1, 0, jb, qbits, 0, jb, tw.X, tw.Y, log_n + 1, k1, 1);
    uint32_t* __restrict__ out = jobs.out[job] + ((size_t)col << (log_n + 1));
    for (int idx = threadIdx.x; idx < tile; idx += blockDim.x) {
        int q = idx & (Q - 1), j = idx >> qbits;
        int a = PAD ? padi(idx) : idx;
        size_t o = ((size_t)j << k1) + q0 + q;
        out[o] = s0[a];
        out[o + ((size_t)1 << log_n)] = s1[a];
    }
}

// ---- pass C: forward layers [k1-1..0] on contiguous chunks of the extended column, in place ---------------------------------
__global__ void __launch_bounds__(256) fft_low_kernel(Jobs jobs, int log_n, int k1, int nc, int groups_per_job, FftTables tw) {
    extern __shared__ uint32_t s[];
    const int T = 1 << k1, m = log_n + 1;
    const int colstride = padi(T);
    const uint32_t chunk = blockIdx.x;
    const int job = blockIdx.y / groups_per_job, c0 = (blockIdx.y % groups_per_job) * nc;
    uint32_t* __restrict__ data = jobs.out[job] + ((size_t)c0 << m) + (size_t)chunk * T;
    for (int idx = threadIdx.x; idx < nc * T; idx += blockDim.x) {
        int c = idx >> k1, r = idx & (T - 1);
        s[c * colstride + padi(r)] = data[((size_t)c << m) + r];
    }
    __syncthreads();
    apply_layers<false, true>(s, nc, colstride, k1, 0, 0, k1, tw.X, tw.Y, m, 0, chunk);
    for (int idx = threadIdx.x; idx < nc * T; idx += blockDim.x) {
        int c = idx >> k1, r = idx & (T - 1);
        data[((size_t)c << m) + r] = s[c * colstride + padi(r)];
    }
}

constexpr int SMEM_MAX = 72 * 1024;

}  // namespace fft2

void fft2_init_attrs() {
    using namespace fft2;
    cudaFuncSetAttribute(small_kernel<SRC_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(small_kernel<SRC_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(ifft_low_kernel<SRC_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(ifft_low_kernel<SRC_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(mid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(mid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(fft_low_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
}

// words of scratch launch_fft_packed needs for `njobs` jobs at log size n
size_t fft_packed_scratch_words(int kind, int njobs, int log_n) {
    if (log_n <= 12) return 0;
    return ((size_t)njobs * (kind == SRC_BITS ? 32 : 4)) << log_n;
}

// kind: SRC_BITS (32 columns per job) or SRC_BYTES (4 columns per job).  src[j]: packed word row of job j (2^log_n words);
// out[j]: tile [cols][2^(log_n+1)].  Returns the number of kernels launched through *launches.
cudaError_t launch_fft_packed(cudaStream_t st, int kind, const uint32_t* const* src, uint32_t* const* out, int njobs, int log_n,
                              const FftTables& tw, uint32_t* scratch, const StageHook* hook, int* launches) {
    using namespace fft2;
#define HOOK(name, b) do { if (hook) hook->fn(hook->user, name, b); } while (0)
    const int cpj = kind == SRC_BITS ? 32 : 4;
    int nl = 0;
    for (int j0 = 0; j0 < njobs; j0 += MAX_FFT_JOBS) {
        Jobs jobs;
        jobs.n = njobs - j0 < MAX_FFT_JOBS ? njobs - j0 : MAX_FFT_JOBS;
        for (int j = 0; j < jobs.n; j++) { jobs.src[j] = src[j0 + j]; jobs.out[j] = out[j0 + j]; }
        if (log_n <= 12) {
            const int big = 2 << log_n;
            int nc = 8192 / big;
            if (nc > cpj) nc = cpj;
            if (nc < 1) nc = 1;
            const int gpj = cpj / nc;
            size_t smem = (size_t)nc * (big + (big >> 4)) * 4;
            int threads = (nc * big / 16) < 256 ? ((nc * big / 16) < 32 ? 32 : nc * big / 16) : 256;
            HOOK("fft_small", 1);
            if (kind == SRC_BITS) small_kernel<SRC_BITS><<<jobs.n * gpj, threads, smem, st>>>(jobs, log_n, nc, gpj, tw);
            else small_kernel<SRC_BYTES><<<jobs.n * gpj, threads, smem, st>>>(jobs, log_n, nc, gpj, tw);
            HOOK("fft_small", 0);
            nl += 1;
            continue;
        }
        int k1 = (log_n + 1) / 2;
        if (log_n - 8 > k1) k1 = log_n - 8;
        if (k1 > 13) k1 = 13;
        const int jb = log_n - k1;
        const int T = 1 << k1;
        int nc = 16384 / T;
        if (nc > cpj) nc = cpj;
        const int gpj = cpj / nc;
        const size_t smem_ac = (size_t)nc * (T + (T >> 4)) * 4;
        int qbits = 13 - jb;  // 2 halves x 2^jb x Q words <= 64 KB
        if (qbits > 5) qbits = 5;
        const int Q = 1 << qbits;
        uint32_t* scr = scratch;  // [jobs.n * cpj][2^n]
        dim3 gA((1u << log_n) / T, jobs.n * gpj);
        HOOK("ifft_low", 1);
        if (kind == SRC_BITS) ifft_low_kernel<SRC_BITS><<<gA, 256, smem_ac, st>>>(jobs, log_n, k1, nc, gpj, cpj, scr, tw);
        else ifft_low_kernel<SRC_BYTES><<<gA, 256, smem_ac, st>>>(jobs, log_n, k1, nc, gpj, cpj, scr, tw);
        HOOK("ifft_low", 0);
        dim3 gB(T / Q, jobs.n * cpj);
        const size_t tile = (size_t)1 << (jb + qbits);
        HOOK("fft_mid", 1);
        if (qbits < 5) mid_kernel<true><<<gB, 256, 2 * (tile + (tile >> 4)) * 4, st>>>(jobs, log_n, k1, qbits, cpj, scr, tw);
        else mid_kernel<false><<<gB, 256, 2 * tile * 4, st>>>(jobs, log_n, k1, qbits, cpj, scr, tw);
        HOOK("fft_mid", 0);
        dim3 gC((2u << log_n) / T, jobs.n * gpj);
        HOOK("fft_low", 1);
        fft_low_kernel<<<gC, 256, smem_ac, st>>>(jobs, log_n, k1, nc, gpj, tw);
        HOOK("fft_low", 0);
        nl += 3;
    }
    if (launches) *launches += nl;
    return cudaGetLastError();
#undef HOOK
}
