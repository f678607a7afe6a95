// Circle FFT over M31 for sm_100a: interpolate (values -> coefficients) and evaluate / extend
// (coefficients -> values on a 2^ext times larger canonic domain), batched over columns.
//
// Replaces, behind the C ABI, upstream stwo `PolyOps::{interpolate_columns, evaluate_polynomials, extend}` for
// SimdBackend (prover/backend/simd/{circle.rs,fft/ifft.rs,fft/rfft.rs}), which the reference reaches through
// `TreeBuilder::{extend_evals,commit}` (/root/reference/stwo/src/chacha/bitwise/air_stream.rs:210-212).
// Any butterfly schedule yields identical results (exact field arithmetic), so the schedule here is chosen for the
// GPU: radix-2^k passes staged through shared memory; low layers on contiguous tiles, high layers on strided tiles
// with >=64-byte coalesced segments; the inverse high layers and the forward high layers are fused in one pass
// (the zero-extension layers of the forward transform are plain duplication).
//
// Memory layout: column-major, one column = contiguous 2^log_size words, values in storage (bit-reversed
// circle-domain) order.  Twiddle tables (flattened, see host/twiddles.hpp):
//   Y[2^k + j]  = y of half_odds(k).at(bitrev_k(j)),            k = 0..M-1
//   X[2^(k-1)+j]= x of half_odds(k).at(bitrev_{k-1}(j)), j<2^(k-1), k = 1..M-1
//   IY / IX     = their inverses.
// canonic domain of log size m: layer 0 uses Y[2^(m-1) + h], layer i>=1 uses X[2^(m-i-1) + h].
#include "common.cuh"

namespace fftk {
using namespace m31;

// one radix-2 layer over a shared-memory tile [ncols][J][Q] (Q fastest); local bit b <-> global layer i
template <bool INV>
__device__ __forceinline__ void layer(uint32_t* s, int ncols, int colstride, int jbits, int qbits, int b, int i, int m,
                                      uint32_t hi_h, const uint32_t* __restrict__ tX, const uint32_t* __restrict__ tY) {
    const int halfJ = 1 << (jbits - 1);
    const int Q = 1 << qbits;
    const int total = ncols * halfJ * Q;
    const uint32_t* tw = (i == 0 ? tY : tX) + (1u << (m - i - 1)) + hi_h;
    const int lowmask = (1 << b) - 1;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        int q = idx & (Q - 1);
        int t = idx >> qbits;
        int p = t & (halfJ - 1);
        int col = t >> (jbits - 1);
        int j0 = ((p >> b) << (b + 1)) | (p & lowmask);
        int j1 = j0 | (1 << b);
        uint32_t w = __ldg(tw + (p >> b));
        uint32_t* base = s + col * colstride;
        uint32_t v0 = base[(j0 << qbits) + q], v1 = base[(j1 << qbits) + q];
        if (INV) {
            base[(j0 << qbits) + q] = add(v0, v1);
            base[(j1 << qbits) + q] = mul(sub(v0, v1), w);
        } else {
            uint32_t t1 = mul(v1, w);
            base[(j0 << qbits) + q] = add(v0, t1);
            base[(j1 << qbits) + q] = sub(v0, t1);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ uint32_t load_src(const ColSrc& src, int col, uint32_t row) {
    if (src.kind == SRC_M31) return src.base[(size_t)col * src.stride + row];
    if (src.kind == SRC_BITS) {
        uint32_t c = col + src.first_col;
        return (src.base[(size_t)(c >> 5) * src.stride + row] >> (c & 31)) & 1u;
    }
    // SRC_BYTES: packed little-endian bytes, 4 columns per word
    uint32_t c = col + src.first_col;
    return (src.base[(size_t)(c >> 2) * src.stride + row] >> ((c & 3) * 8)) & 0xffu;
}

// ---- whole column(s) in shared memory: interpolate [+ store coeffs] [+ extend & evaluate] -------------------------
// mode bits: 1 = do inverse transform first (input are values), 2 = store coefficients, 4 = evaluate to log_n+ext
__global__ void __launch_bounds__(256) fft_small_kernel(ColSrc src, int ncols, int log_n, int ext, int mode, int cpb,
                                                        uint32_t* __restrict__ coef_out, size_t coef_stride,
                                                        uint32_t* __restrict__ eval_out, size_t eval_stride,
                                                        FftTables tw) {
    extern __shared__ uint32_t s[];
    const int n = 1 << log_n;
    const int m = log_n + ext;
    const int big = 1 << m;
    const int col0 = blockIdx.x * cpb;
    const int nc = min(cpb, ncols - col0);
    if (nc <= 0) return;
    for (int idx = threadIdx.x; idx < nc * n; idx += blockDim.x) {
        int c = idx >> log_n, r = idx & (n - 1);
        s[c * big + r] = load_src(src, col0 + c, r);
    }
    __syncthreads();
    if (mode & 1) {
        for (int i = 0; i < log_n; i++) layer<true>(s, nc, big, log_n, 0, i, i, log_n, 0, tw.IX, tw.IY);
        const uint32_t inv_n = 1u << (31 - log_n);  // 2^-log_n = 2^(31-log_n) mod p
        for (int idx = threadIdx.x; idx < nc * n; idx += blockDim.x) {
            int c = idx >> log_n, r = idx & (n - 1);
            uint32_t v = mul(s[c * big + r], inv_n);
            s[c * big + r] = v;
            if (mode & 2) coef_out[(size_t)(col0 + c) * coef_stride + r] = v;
        }
        __syncthreads();
    }
    if (!(mode & 4)) return;
    // zero-extension layers of the forward transform: butterfly(v, 0) = (v, v)
    for (int idx = threadIdx.x; idx < nc * (big - n); idx += blockDim.x) {
        int c = idx / (big - n), r = n + idx % (big - n);
        s[c * big + r] = s[c * big + (r & (n - 1))];
    }
    __syncthreads();
    for (int i = log_n - 1; i >= 0; i--) layer<false>(s, nc, big, m, 0, i, i, m, 0, tw.X, tw.Y);
    for (int idx = threadIdx.x; idx < nc * big; idx += blockDim.x) {
        int c = idx >> m, r = idx & (big - 1);
        eval_out[(size_t)(col0 + c) * eval_stride + r] = s[c * big + r];
    }
}

// ---- large columns, pass 1: inverse layers [0,k1) on contiguous 2^k1 tiles -----------------------------------------
__global__ void __launch_bounds__(256) ifft_low_kernel(ColSrc src, int ncols, int log_n, int k1, int cpb,
                                                       uint32_t* __restrict__ out, size_t out_stride, FftTables tw) {
    extern __shared__ uint32_t s[];
    const int T = 1 << k1;
    const uint32_t tile = blockIdx.x;  // tile index within the column
    const int col0 = blockIdx.y * cpb;
    const int nc = min(cpb, ncols - col0);
    for (int idx = threadIdx.x; idx < nc * T; idx += blockDim.x) {
        int c = idx >> k1, r = idx & (T - 1);
        s[idx] = load_src(src, col0 + c, tile * T + r);
    }
    __syncthreads();
    for (int i = 0; i < k1; i++) layer<true>(s, nc, T, k1, 0, i, i, log_n, tile << (k1 - i - 1), tw.IX, tw.IY);
    for (int idx = threadIdx.x; idx < nc * T; idx += blockDim.x) {
        int c = idx >> k1, r = idx & (T - 1);
        out[(size_t)(col0 + c) * out_stride + tile * T + r] = s[idx];
    }
}

// ---- large columns, pass 2: inverse layers [k1,log_n), scale, store coefficients, extend, forward layers down to k1 ----
// tile: all 2^(log_n-k1) values of the high bits x Q=2^qbits consecutive low indices, one column per block.
__global__ void __launch_bounds__(512) fft_mid_kernel(const uint32_t* __restrict__ in, size_t in_stride, int ncols, int log_n,
                                                      int k1, int ext, int qbits, int mode, uint32_t* __restrict__ coef_out,
                                                      size_t coef_stride, uint32_t* __restrict__ eval_out, size_t eval_stride,
                                                      FftTables tw) {
    extern __shared__ uint32_t s[];
    const int Q = 1 << qbits;
    const int jb = log_n - k1;           // local bits of the inverse part
    const int J = 1 << jb;
    const int m = log_n + ext;
    const int jbF = m - k1;              // local bits of the forward part
    const uint32_t low0 = blockIdx.x * Q;  // first low index (low index < 2^k1)
    const int col = blockIdx.y;
    const uint32_t* cin = in + (size_t)col * in_stride;
    for (int idx = threadIdx.x; idx < J * Q; idx += blockDim.x) {
        int q = idx & (Q - 1), j = idx >> qbits;
        s[idx] = cin[((size_t)j << k1) + low0 + q];
    }
    __syncthreads();
    if (mode & 1) {
        for (int i = k1; i < log_n; i++) layer<true>(s, 1, 0, jb, qbits, i - k1, i, log_n, 0, tw.IX, tw.IY);
        const uint32_t inv_n = 1u << (31 - log_n);
        for (int idx = threadIdx.x; idx < J * Q; idx += blockDim.x) {
            int q = idx & (Q - 1), j = idx >> qbits;
            uint32_t v = mul(s[idx], inv_n);
            s[idx] = v;
            if (mode & 2) coef_out[(size_t)col * coef_stride + ((size_t)j << k1) + low0 + q] = v;
        }
        __syncthreads();
    }
    if (!(mode & 4)) return;
    for (int idx = threadIdx.x + J * Q; idx < (J << ext) * Q; idx += blockDim.x) s[idx] = s[idx & (J * Q - 1)];
    __syncthreads();
    for (int i = log_n - 1; i >= k1; i--) layer<false>(s, 1, 0, jbF, qbits, i - k1, i, m, 0, tw.X, tw.Y);
    uint32_t* cout = eval_out + (size_t)col * eval_stride;
    for (int idx = threadIdx.x; idx < (J << ext) * Q; idx += blockDim.x) {
        int q = idx & (Q - 1), j = idx >> qbits;
        cout[((size_t)j << k1) + low0 + q] = s[idx];
    }
}

// ---- large columns, pass 3: forward layers [k1-1 .. 0] on contiguous 2^k1 tiles, in place -------------------------------
__global__ void __launch_bounds__(256) fft_low_kernel(uint32_t* __restrict__ data, size_t stride, int ncols, int m, int k1,
                                                      int cpb, FftTables tw) {
    extern __shared__ uint32_t s[];
    const int T = 1 << k1;
    const uint32_t tile = blockIdx.x;
    const int col0 = blockIdx.y * cpb;
    const int nc = min(cpb, ncols - col0);
    for (int idx = threadIdx.x; idx < nc * T; idx += blockDim.x) {
        int c = idx >> k1, r = idx & (T - 1);
        s[idx] = data[(size_t)(col0 + c) * stride + tile * T + r];
    }
    __syncthreads();
    for (int i = k1 - 1; i >= 0; i--) layer<false>(s, nc, T, k1, 0, i, i, m, tile << (k1 - i - 1), tw.X, tw.Y);
    for (int idx = threadIdx.x; idx < nc * T; idx += blockDim.x) {
        int c = idx >> k1, r = idx & (T - 1);
        data[(size_t)(col0 + c) * stride + tile * T + r] = s[idx];
    }
}

}  // namespace fftk

void fft_init_attrs() {
    using namespace fftk;
    cudaFuncSetAttribute(fft_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ifft_low_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(fft_low_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(fft_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
}

// Host-side launcher.  mode: 1 = interpolate first, 2 = store coefficients, 4 = evaluate on 2^(log_n+ext).
// When mode&1 is clear the source must be an M31 array of coefficients.
cudaError_t launch_fft(cudaStream_t st, const ColSrc& src, int ncols, int log_n, int ext, int mode, uint32_t* coef_out,
                       size_t coef_stride, uint32_t* eval_out, size_t eval_stride, const FftTables& tw, uint32_t* scratch,
                       size_t scratch_stride, const StageHook* hook) {
    using namespace fftk;
#define HOOK(name, b) do { if (hook) hook->fn(hook->user, name, b); } while (0)
    if (ncols <= 0) return cudaSuccess;
    const int m = log_n + ext;
    if (m <= 13) {
        int big = 1 << m;
        int cpb = max(1, 4096 / big);
        size_t smem = (size_t)cpb * big * 4;
        HOOK("fft_small", 1);
        fft_small_kernel<<<(ncols + cpb - 1) / cpb, 256, smem, st>>>(src, ncols, log_n, ext, mode, cpb, coef_out, coef_stride,
                                                                       eval_out, eval_stride, tw);
        HOOK("fft_small", 0);
        return cudaGetLastError();
    }
    // multi-pass.  k1 = contiguous tile bits; strided pass holds 2^(m-k1) x Q words.
    int k1 = (log_n + 1) / 2;
    if (k1 > 12) k1 = 12;
    while (m - k1 > 11) k1++;  // 2^(m-k1) * Q(16) * 4B <= 128 KB
    const int qbits = 4;
    const int T = 1 << k1;
    int cpb = max(1, 4096 / T);
    const uint32_t* mid_in;
    size_t mid_stride;
    if (mode & 1) {
        // pass 1 writes to scratch (may alias coef_out; every element is read and written by the same block)
        dim3 g1((1u << log_n) / T, (ncols + cpb - 1) / cpb);
        HOOK("ifft_low", 1);
        ifft_low_kernel<<<g1, 256, (size_t)cpb * T * 4, st>>>(src, ncols, log_n, k1, cpb, scratch, scratch_stride, tw);
        HOOK("ifft_low", 0);
        mid_in = scratch;
        mid_stride = scratch_stride;
    } else {
        mid_in = src.base;
        mid_stride = src.stride;
    }
    dim3 g2(T >> qbits, ncols);
    size_t smem2 = ((size_t)1 << (m - k1 + qbits)) * 4;
    HOOK("fft_mid", 1);
    fft_mid_kernel<<<g2, 512, smem2, st>>>(mid_in, mid_stride, ncols, log_n, k1, ext, qbits, mode, coef_out, coef_stride,
                                           eval_out, eval_stride, tw);
    HOOK("fft_mid", 0);
    if (mode & 4) {
        dim3 g3((1u << m) / T, (ncols + cpb - 1) / cpb);
        HOOK("fft_low", 1);
        fft_low_kernel<<<g3, 256, (size_t)cpb * T * 4, st>>>(eval_out, eval_stride, ncols, m, k1, cpb, tw);
        HOOK("fft_low", 0);
    }
    return cudaGetLastError();
}
