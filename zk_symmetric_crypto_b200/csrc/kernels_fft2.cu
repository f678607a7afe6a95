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

// word offset of (column c, row r) in an output tile of `cpj` columns stored as [row shards][cpj][2^lr rows]
__device__ __forceinline__ size_t soff(int lr, int cpj, int c, size_t r) {
    return ((((r >> lr) * cpj) + c) << lr) | (r & (((size_t)1 << lr) - 1));
}

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
    uint32_t one, mone;       // run-time 1 and -1: x*one+y compiles to IMAD, moving butterfly additions to the FMA pipe
    int halves;               // 2: the whole extended column; 1: only storage rows [0, 2^n) (all the constraint pass reads)
    int shard_log;            // log2 of the rows per row-shard of the output tile (= log_n+1 when the tile is not sharded):
                              // tile layout [shards][cols][2^shard_log], so that a rank's row range of all columns is contiguous
    const uint32_t* src[MAX_FFT_JOBS];  // packed witness word row (2^n words)
    uint32_t* out[MAX_FFT_JOBS];        // LDE tile [cols_per_job][2^(n+1)]
    // Row-sharded proving over peer windows (PeerDst): the last pass stores every contiguous chunk of the extended column
    // straight into the tile slot of the rank that owns those rows, over NVLink, instead of writing it back in place; the
    // column transform and the all-to-all that transposes columns into row shards are one kernel.
    uint32_t* peer_base[MAX_PEERS];             // arena of rank r as mapped in this process; used when peer_on
    unsigned long long peer_off[MAX_FFT_JOBS];  // word offset of job j's destination slot inside every rank's arena
    int peer_on, peer_logG, peer_lv;
    // product-size traces (small_kernel only): any number of jobs addressed through a device word list,
    // job j: src = src0 + wlist[j] * src_stride, out = out0 + wlist[j] * out_stride
    const int* wlist;
    const uint32_t* src0;
    uint32_t* out0;
    size_t src_stride, out_stride;
};

// destination of the chunk starting at global row `row0` of column `c0` of job `job` (column stride 2^(lv+1) words)
__device__ __forceinline__ uint32_t* peer_dst(const Jobs& jobs, int job, int c0, size_t row0) {
    const uint32_t v = (uint32_t)(row0 >> jobs.peer_lv);  // virtual shard; owner = v mod G, local half = v div G
    const size_t lrow = ((size_t)(v >> jobs.peer_logG) << jobs.peer_lv) | (row0 & (((size_t)1 << jobs.peer_lv) - 1));
    return jobs.peer_base[v & ((1u << jobs.peer_logG) - 1)] + jobs.peer_off[job] + ((size_t)c0 << (jobs.peer_lv + 1)) + lrow;
}

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
    const int wj = jobs.wlist ? jobs.wlist[job] : 0;
    const uint32_t* __restrict__ src = jobs.wlist ? jobs.src0 + (size_t)wj * jobs.src_stride : jobs.src[job];
    const uint32_t scale = 1u << (31 - log_n), scale2 = scale << 1;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        if (KIND == SRC_M31) {  // job = cols_per_job plain M31 columns, n words apart
            for (int c = 0; c < nc; c++) s[c * colstride + padi(r)] = mulw(__ldg(src + (size_t)(c0 + c) * n + r), scale2);
        } else {
            uint32_t w = __ldg(src + r);
            for (int c = 0; c < nc; c++) s[c * colstride + padi(r)] = unpack<KIND>(w, c0 + c, scale, scale2);
        }
    }
    __syncthreads();
    apply_layers<true, true>(s, nc, colstride, log_n, 0, 0, log_n, tw.IX, tw.IY, log_n, 0, 0);
    for (int idx = threadIdx.x; idx < nc * n; idx += blockDim.x) {
        int c = idx >> log_n, r = idx & (n - 1);
        s[c * colstride + padi(n + r)] = s[c * colstride + padi(r)];
    }
    __syncthreads();
    apply_layers<false, true>(s, nc, colstride, m, 0, 0, log_n, tw.X, tw.Y, m, 0, 0);
    uint32_t* __restrict__ out = jobs.wlist ? jobs.out0 + (size_t)wj * jobs.out_stride : jobs.out[job];
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
        if (KIND == SRC_M31) {
            for (int c = 0; c < nc; c++) s[c * colstride + padi(r)] = mulw(__ldg(src + ((size_t)(c0 + c) << log_n) + r), scale2);
        } else {
            uint32_t w = __ldg(src + r);
            for (int c = 0; c < nc; c++) s[c * colstride + padi(r)] = unpack<KIND>(w, c0 + c, scale, scale2);
        }
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
    uint32_t* __restrict__ out = jobs.out[job];
    for (int idx = threadIdx.x; idx < tile; idx += blockDim.x) {
        int q = idx & (Q - 1), j = idx >> qbits;
        int a = PAD ? padi(idx) : idx;
        size_t o = ((size_t)j << k1) + q0 + q;
        out[soff(jobs.shard_log, cols_per_job, col, o)] = s0[a];
        out[soff(jobs.shard_log, cols_per_job, col, o + ((size_t)1 << log_n))] = s1[a];
    }
}

// ---- pass C: forward layers [k1-1..0] on contiguous chunks of the extended column, in place ---------------------------------
__global__ void __launch_bounds__(256) fft_low_kernel(Jobs jobs, int log_n, int k1, int nc, int groups_per_job, FftTables tw) {
    extern __shared__ uint32_t s[];
    const int T = 1 << k1, m = log_n + 1;
    const int colstride = padi(T);
    const uint32_t chunk = blockIdx.x;
    const int job = blockIdx.y / groups_per_job, c0 = (blockIdx.y % groups_per_job) * nc;
    uint32_t* __restrict__ data = jobs.out[job];
    const int cpj = groups_per_job * nc;
    (void)m;
    for (int idx = threadIdx.x; idx < nc * T; idx += blockDim.x) {
        int c = idx >> k1, r = idx & (T - 1);
        s[c * colstride + padi(r)] = data[soff(jobs.shard_log, cpj, c0 + c, (size_t)chunk * T + r)];
    }
    __syncthreads();
    apply_layers<false, true>(s, nc, colstride, k1, 0, 0, k1, tw.X, tw.Y, m, 0, chunk);
    if (jobs.peer_on) {
        uint32_t* __restrict__ dst = peer_dst(jobs, job, c0, (size_t)chunk * T);
        for (int idx = threadIdx.x; idx < nc * T; idx += blockDim.x) {
            int c = idx >> k1, r = idx & (T - 1);
            dst[((size_t)c << (jobs.peer_lv + 1)) + r] = s[c * colstride + padi(r)];
        }
        return;
    }
    for (int idx = threadIdx.x; idx < nc * T; idx += blockDim.x) {
        int c = idx >> k1, r = idx & (T - 1);
        data[soff(jobs.shard_log, cpj, c0 + c, (size_t)chunk * T + r)] = s[c * colstride + padi(r)];
    }
}


// ===================================================================================================================
// v2 kernels for 13 <= n <= 20: contiguous chunk fixed at 2^12 rows, every step's layer range is a template parameter
// (shared-memory offsets become immediates, no index arithmetic or integer divisions in the loops), the first and the
// last register step of every kernel read from / write to global memory directly.
// Shared-memory layout of a 4096-word column chunk: +4 words per 32 (keeps 16-byte vectors intact) and bank bit 4
// flipped for odd 256-blocks, which makes all three step patterns (stride 256, stride 16, 16 consecutive as 4 x 128-bit)
// conflict-free.
// ===================================================================================================================
constexpr int K1 = 12, T2 = 4096, COLW = 4608;
__device__ __forceinline__ int phys(int a) { return (a + ((a >> 5) << 2)) ^ ((a >> 4) & 16); }

// twiddles of R consecutive layers starting at global layer gi0 for the 2^R-point block whose first point has (tile-local)
// index j0; jg = (tile_hi << jbits) | j0, b = local bit of gi0.  slot (c-1)+t, c = 2^(R-1-l); values pre-doubled.
template <int R>
__device__ __forceinline__ void load_tw(uint32_t (&tw)[1 << R], const uint32_t* __restrict__ tabX, const uint32_t* __restrict__ tabY,
                                        int m, int gi0, uint32_t jg, int b) {
#pragma unroll
    for (int l = 0; l < R; l++) {
        const int gi = gi0 + l;
        const uint32_t* tab = (gi == 0) ? tabY + (1u << (m - 1)) : tabX + (1u << (m - gi - 1));
        const uint32_t hb = jg >> (b + l + 1);
#pragma unroll
        for (int t = 0; t < (1 << (R - 1 - l)); t++) tw[(1 << (R - 1 - l)) - 1 + t] = __ldg(tab + hb + t) << 1;
    }
}

// a + b and a - b issued as IMAD (FMA pipe) instead of IADD3 (ALU pipe): the butterflies are ALU-pipe bound otherwise
__device__ __forceinline__ uint32_t addf(uint32_t a, uint32_t b, uint32_t one) { return redp(a * one + b); }
__device__ __forceinline__ uint32_t subf(uint32_t a, uint32_t b, uint32_t mone) {
    uint32_t d = b * mone + a;
    return __viaddmin_u32(d, P, d);
}

template <int R>
__device__ __forceinline__ void inv_block(uint32_t (&v)[1 << R], const uint32_t (&tw)[1 << R], uint32_t one, uint32_t mone) {
#pragma unroll
    for (int l = 0; l < R; l++) {
#pragma unroll
        for (int k = 0; k < (1 << R); k++) {
            if (k & (1 << l)) continue;
            const uint32_t w2 = tw[(1 << (R - 1 - l)) - 1 + (k >> (l + 1))];
            uint32_t v0 = v[k], v1 = v[k | (1 << l)];
            v[k] = addf(v0, v1, one);
            v[k | (1 << l)] = mulw(subf(v0, v1, mone), w2);
        }
    }
}

template <int R>
__device__ __forceinline__ void fwd_block(uint32_t (&v)[1 << R], const uint32_t (&tw)[1 << R], uint32_t one, uint32_t mone) {
#pragma unroll
    for (int l = R - 1; l >= 0; l--) {
#pragma unroll
        for (int k = 0; k < (1 << R); k++) {
            if (k & (1 << l)) continue;
            const uint32_t w2 = tw[(1 << (R - 1 - l)) - 1 + (k >> (l + 1))];
            uint32_t v0 = v[k], t = mulw(v[k | (1 << l)], w2);
            v[k] = addf(v0, t, one);
            v[k | (1 << l)] = subf(v0, t, mone);
        }
    }
}

// ---- pass A: expand packed words, inverse layers [0,12) on a 4096-row chunk, NC columns per block ---------------------------
template <int KIND, int NC>
__global__ void __launch_bounds__(256) ifft_low12_kernel(Jobs jobs, int log_n, int groups_per_job, int cols_per_job,
                                                         uint32_t* __restrict__ scratch, FftTables tw) {
    extern __shared__ uint32_t s[];
    const uint32_t chunk = blockIdx.x;
    const int job = blockIdx.y / groups_per_job, c0 = (blockIdx.y % groups_per_job) * NC;
    const int p = threadIdx.x;
    const uint32_t scale = 1u << (31 - log_n), scale2 = scale << 1;
    uint32_t w[16], twr[16], v[16];
    if (KIND != SRC_M31) {
        const uint4* __restrict__ src = (const uint4*)(jobs.src[job] + (size_t)chunk * T2 + 16 * p);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint4 x = __ldg(src + i);
            w[4 * i] = x.x; w[4 * i + 1] = x.y; w[4 * i + 2] = x.z; w[4 * i + 3] = x.w;
        }
    }
    // step (4,0): from registers, 16 consecutive points per thread
    load_tw<4>(twr, tw.IX, tw.IY, log_n, 0, (chunk << K1) | (uint32_t)(16 * p), 0);
    int va[4];
#pragma unroll
    for (int i = 0; i < 4; i++) va[i] = phys(16 * p + 4 * i);
#pragma unroll
    for (int c = 0; c < NC; c++) {
#pragma unroll
        if (KIND == SRC_M31) {  // plain M31 columns: 16 consecutive values of column c0+c
            const uint4* __restrict__ src = (const uint4*)(jobs.src[job] + ((size_t)(c0 + c) << log_n) + (size_t)chunk * T2 + 16 * p);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 x = __ldg(src + i);
                v[4 * i] = mulw(x.x, scale2); v[4 * i + 1] = mulw(x.y, scale2); v[4 * i + 2] = mulw(x.z, scale2); v[4 * i + 3] = mulw(x.w, scale2);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = unpack<KIND>(w[k], c0 + c, scale, scale2);
        }
        inv_block<4>(v, twr, jobs.one, jobs.mone);
#pragma unroll
        for (int i = 0; i < 4; i++) *(uint4*)(s + c * COLW + va[i]) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
    __syncthreads();
    // step (4,4)
    {
        const int j0 = ((p >> 4) << 8) | (p & 15);
        load_tw<4>(twr, tw.IX, tw.IY, log_n, 4, (chunk << K1) | (uint32_t)j0, 4);
        int ad[16];
#pragma unroll
        for (int k = 0; k < 16; k++) ad[k] = phys(j0 + 16 * k);
#pragma unroll
        for (int c = 0; c < NC; c++) {
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = s[c * COLW + ad[k]];
            inv_block<4>(v, twr, jobs.one, jobs.mone);
#pragma unroll
            for (int k = 0; k < 16; k++) s[c * COLW + ad[k]] = v[k];
        }
    }
    __syncthreads();
    // step (4,8): to global (scratch), 128-byte coalesced per k
    {
        load_tw<4>(twr, tw.IX, tw.IY, log_n, 8, (chunk << K1) | (uint32_t)p, 8);
        int ad[16];
#pragma unroll
        for (int k = 0; k < 16; k++) ad[k] = phys(p + 256 * k);
        uint32_t* __restrict__ out = scratch + ((size_t)(job * cols_per_job + c0) << log_n) + (size_t)chunk * T2 + p;
#pragma unroll
        for (int c = 0; c < NC; c++) {
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = s[c * COLW + ad[k]];
            inv_block<4>(v, twr, jobs.one, jobs.mone);
#pragma unroll
            for (int k = 0; k < 16; k++) out[((size_t)c << log_n) + 256 * k] = v[k];
        }
    }
}

// ---- pass C: forward layers [11..0] on 4096-point chunks of the extended column, in place -------------------------------------
template <int NC, bool P2P>
__global__ void __launch_bounds__(256) fft_low12_kernel(Jobs jobs, int log_n, int groups_per_job, FftTables tw) {
    extern __shared__ uint32_t s[];
    const int m = log_n + 1;
    const uint32_t chunk = blockIdx.x;
    const int job = blockIdx.y / groups_per_job, c0 = (blockIdx.y % groups_per_job) * NC;
    // a 4096-row chunk lies inside one row shard (shards hold >= 4096 rows); column stride inside a shard = 2^lr
    const int lr = jobs.shard_log, cpj = groups_per_job * NC;
    uint32_t* __restrict__ data = jobs.out[job] + soff(lr, cpj, c0, (size_t)chunk * T2);
    const int p = threadIdx.x;
    uint32_t twr[16], v[16];
    {
        load_tw<4>(twr, tw.X, tw.Y, m, 8, (chunk << K1) | (uint32_t)p, 8);
        int ad[16];
#pragma unroll
        for (int k = 0; k < 16; k++) ad[k] = phys(p + 256 * k);
#pragma unroll
        for (int c = 0; c < NC; c++) {
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = data[((size_t)c << lr) + p + 256 * k];
            fwd_block<4>(v, twr, jobs.one, jobs.mone);
#pragma unroll
            for (int k = 0; k < 16; k++) s[c * COLW + ad[k]] = v[k];
        }
    }
    __syncthreads();
    {
        const int j0 = ((p >> 4) << 8) | (p & 15);
        load_tw<4>(twr, tw.X, tw.Y, m, 4, (chunk << K1) | (uint32_t)j0, 4);
        int ad[16];
#pragma unroll
        for (int k = 0; k < 16; k++) ad[k] = phys(j0 + 16 * k);
#pragma unroll
        for (int c = 0; c < NC; c++) {
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = s[c * COLW + ad[k]];
            fwd_block<4>(v, twr, jobs.one, jobs.mone);
#pragma unroll
            for (int k = 0; k < 16; k++) s[c * COLW + ad[k]] = v[k];
        }
    }
    __syncthreads();
    {
        load_tw<4>(twr, tw.X, tw.Y, m, 0, (chunk << K1) | (uint32_t)(16 * p), 0);
        uint32_t* __restrict__ dst = P2P ? peer_dst(jobs, job, c0, (size_t)chunk * T2) : nullptr;
        int va[4];
#pragma unroll
        for (int i = 0; i < 4; i++) va[i] = phys(16 * p + 4 * i);
#pragma unroll
        for (int c = 0; c < NC; c++) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 x = *(const uint4*)(s + c * COLW + va[i]);
                v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
            }
            fwd_block<4>(v, twr, jobs.one, jobs.mone);
            if (P2P) {
                // Peer stores are not merged on the way (no local L2 in between): a thread's own 64 bytes would cross NVLink as
                // four 16-byte writes.  The warp's 512 consecutive words go back through its shared-memory rows (warp-private,
                // so a warp barrier is enough) and leave as 4 x 512 contiguous bytes per warp store.
#pragma unroll
                for (int i = 0; i < 4; i++) *(uint4*)(s + c * COLW + va[i]) = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                __syncwarp();
                const int wbase = (p & ~31) * 16, lane = p & 31;
                uint4* o = (uint4*)(dst + ((size_t)c << (jobs.peer_lv + 1)) + wbase);
#pragma unroll
                for (int i = 0; i < 4; i++) o[i * 32 + lane] = *(const uint4*)(s + c * COLW + phys(wbase + i * 128 + 4 * lane));
            } else {
                uint4* o = (uint4*)(data + ((size_t)c << lr) + 16 * p);
#pragma unroll
                for (int i = 0; i < 4; i++) o[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
        }
    }
}

// ---- pass B: inverse layers [12,n) + forward layers [n-1..12] of both halves; tile = 2^JB x 32, JB = n-12 in [1,8] ----------
template <int JB>
__global__ void __launch_bounds__(256) mid12_kernel(Jobs jobs, int cols_per_job, const uint32_t* __restrict__ scratch, FftTables tw) {
    extern __shared__ uint32_t s[];
    constexpr int log_n = K1 + JB, m = log_n + 1;
    constexpr int RA = JB <= 4 ? JB : (JB + 1) / 2;  // low local bits (first inverse step / last forward step)
    constexpr int RB = JB - RA;                      // high local bits
    constexpr int J = 1 << JB;
    const int q = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t q0 = blockIdx.x * 32;
    const int job = blockIdx.y / cols_per_job, col = blockIdx.y % cols_per_job;
    const uint32_t* __restrict__ in = scratch + ((size_t)blockIdx.y << log_n) + q0 + q;
    uint32_t* __restrict__ out = jobs.out[job];
    const int lr = jobs.shard_log;
    // rq < 2^12 <= shard rows: the shard / column part of a store offset is uniform over the block (kept off the per-thread
    // integer pipe), only rq is per thread
    const size_t rq = (size_t)q0 + q;
    const bool sharded = lr != m;  // unsharded tiles ([cols][2^m]) keep the compile-time store offsets
    uint32_t* __restrict__ outu = out + ((size_t)col << m) + rq;
    uint32_t twr[1 << RA], v[1 << RA];
    if (RB == 0) {
        load_tw<RA>(twr, tw.IX, tw.IY, log_n, K1, 0, 0);
#pragma unroll
        for (int k = 0; k < J; k++) v[k] = in[(size_t)k << K1];
        inv_block<RA>(v, twr, jobs.one, jobs.mone);
        for (int h = 0; h < jobs.halves; h++) {
            uint32_t u[1 << RA];
#pragma unroll
            for (int k = 0; k < J; k++) u[k] = v[k];
            load_tw<RA>(twr, tw.X, tw.Y, m, K1, (uint32_t)h << JB, 0);
            fwd_block<RA>(u, twr, jobs.one, jobs.mone);
            if (!sharded) {
#pragma unroll
                for (int k = 0; k < J; k++) outu[((size_t)h << log_n) + ((size_t)k << K1)] = u[k];
            } else {
#pragma unroll
                for (int k = 0; k < J; k++) out[soff(lr, cols_per_job, col, ((size_t)h << log_n) + ((size_t)k << K1)) + rq] = u[k];
            }
        }
        return;
    }
    uint32_t* s0 = s;
    uint32_t* s1 = s + J * 32;
    // inverse step A: local bits [0,RA), from global
    for (int pa = wid; pa < (J >> RA); pa += nw) {
        const int j0 = pa << RA;
        load_tw<RA>(twr, tw.IX, tw.IY, log_n, K1, (uint32_t)j0, 0);
#pragma unroll
        for (int k = 0; k < (1 << RA); k++) v[k] = in[(size_t)(j0 + k) << K1];
        inv_block<RA>(v, twr, jobs.one, jobs.mone);
#pragma unroll
        for (int k = 0; k < (1 << RA); k++) s0[((j0 + k) << 5) | q] = v[k];
    }
    __syncthreads();
    // inverse step B: local bits [RA,JB) -> coefficients; forward step B of both halves
    constexpr int RBs = RB > 0 ? RB : 1;
    for (int pb = wid; pb < (J >> RBs); pb += nw) {
        uint32_t twb[1 << RBs], c[1 << RBs];
        load_tw<RBs>(twb, tw.IX, tw.IY, log_n, K1 + RA, (uint32_t)pb, RA);
#pragma unroll
        for (int k = 0; k < (1 << RBs); k++) c[k] = s0[((pb + (k << RA)) << 5) | q];
        inv_block<RBs>(c, twb, jobs.one, jobs.mone);
        for (int h = 0; h < jobs.halves; h++) {
            uint32_t u[1 << RBs];
#pragma unroll
            for (int k = 0; k < (1 << RBs); k++) u[k] = c[k];
            load_tw<RBs>(twb, tw.X, tw.Y, m, K1 + RA, ((uint32_t)h << JB) | (uint32_t)pb, RA);
            fwd_block<RBs>(u, twb, jobs.one, jobs.mone);
            uint32_t* sh = h ? s1 : s0;
#pragma unroll
            for (int k = 0; k < (1 << RBs); k++) sh[((pb + (k << RA)) << 5) | q] = u[k];
        }
    }
    __syncthreads();
    // forward step A: local bits [0,RA), to global
    for (int pa = wid; pa < jobs.halves * (J >> RA); pa += nw) {
        const int h = pa >= (J >> RA);
        const int j0 = (pa - h * (J >> RA)) << RA;
        const uint32_t* sh = h ? s1 : s0;
        load_tw<RA>(twr, tw.X, tw.Y, m, K1, ((uint32_t)h << JB) | (uint32_t)j0, 0);
#pragma unroll
        for (int k = 0; k < (1 << RA); k++) v[k] = sh[((j0 + k) << 5) | q];
        fwd_block<RA>(v, twr, jobs.one, jobs.mone);
        if (!sharded) {  // immediate store offsets from one base pointer
            uint32_t* __restrict__ o = outu + ((size_t)h << log_n) + ((size_t)j0 << K1);
#pragma unroll
            for (int k = 0; k < (1 << RA); k++) o[(size_t)k << K1] = v[k];
        } else {
#pragma unroll
            for (int k = 0; k < (1 << RA); k++) out[soff(lr, cols_per_job, col, ((size_t)h << log_n) + ((size_t)(j0 + k) << K1)) + rq] = v[k];
        }
    }
}

template <int JB>
static void launch_mid12(cudaStream_t st, const Jobs& jobs, int cpj, const uint32_t* scratch, const FftTables& tw) {
    constexpr int RA = JB <= 4 ? JB : (JB + 1) / 2;
    constexpr int RB = JB - RA;
    constexpr int items = (1 << (JB - RA)) * 32;       // step A items per tile
    const int threads = RB == 0 ? 32 : (items < 256 ? items : 256);
    const size_t smem = RB == 0 ? 0 : (size_t)2 * (1 << JB) * 32 * 4;
    dim3 g(T2 / 32, jobs.n * cpj);
    mid12_kernel<JB><<<g, threads, smem, st>>>(jobs, cpj, scratch, tw);
}

constexpr int SMEM_MAX = 72 * 1024;

}  // namespace fft2

int g_force_generic_fft = 0;
int g_fft_fma_adds = 1;  // tests: exercise the generic (runtime-schedule) kernels at sizes the v2 kernels cover

void fft2_init_attrs() {
    using namespace fft2;
    cudaFuncSetAttribute(small_kernel<SRC_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(small_kernel<SRC_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(small_kernel<SRC_M31>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(ifft_low_kernel<SRC_M31>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(ifft_low12_kernel<SRC_M31, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * COLW * 4);
    cudaFuncSetAttribute(ifft_low_kernel<SRC_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(ifft_low_kernel<SRC_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(mid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(mid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(fft_low_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    cudaFuncSetAttribute(ifft_low12_kernel<SRC_BITS, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * COLW * 4);
    cudaFuncSetAttribute(ifft_low12_kernel<SRC_BYTES, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * COLW * 4);
    cudaFuncSetAttribute(fft_low12_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * COLW * 4);
    cudaFuncSetAttribute(fft_low12_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * COLW * 4);
    cudaFuncSetAttribute(mid12_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * 32 * 4);
    cudaFuncSetAttribute(mid12_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * 32 * 4);
}

// Product-size traces (log_n <= 12): ONE launch transforms the packed words listed in wlist_dev (n_words of them), word w from
// src0 + w * src_stride into the tile out0 + w * out_stride.
cudaError_t launch_fft_packed_list(cudaStream_t st, const int* wlist_dev, int n_words, const uint32_t* src0, size_t src_stride,
                                   uint32_t* out0, size_t out_stride, int log_n, const FftTables& tw) {
    using namespace fft2;
    if (log_n > 12 || n_words <= 0) return cudaErrorInvalidValue;
    Jobs jobs{};
    jobs.n = n_words;
    jobs.one = 1u;
    jobs.mone = 0xffffffffu;
    jobs.halves = 2;
    jobs.shard_log = log_n + 1;
    jobs.wlist = wlist_dev;
    jobs.src0 = src0;
    jobs.out0 = out0;
    jobs.src_stride = src_stride;
    jobs.out_stride = out_stride;
    const int cpj = 32, big = 2 << log_n;
    int nc = 8192 / big;
    if (nc > cpj) nc = cpj;
    if (nc < 1) nc = 1;
    const int gpj = cpj / nc;
    const size_t smem = (size_t)nc * (big + (big >> 4)) * 4;
    const int threads = (nc * big / 16) < 256 ? ((nc * big / 16) < 32 ? 32 : nc * big / 16) : 256;
    small_kernel<SRC_BITS><<<n_words * gpj, threads, smem, st>>>(jobs, log_n, nc, gpj, tw);
    return cudaGetLastError();
}

// words of scratch launch_fft_packed needs for `njobs` jobs at log size n
size_t fft_packed_scratch_words(int kind, int njobs, int log_n) {
    if (log_n <= 12) return 0;
    return ((size_t)njobs * (kind == SRC_BITS ? 32 : 4)) << log_n;
}

// kind: SRC_BITS (32 columns per job), SRC_BYTES (4 columns per job) or SRC_M31 (4 plain M31 columns per job, 2^log_n words
// apart, src[j] -> the first of them).  src[j]: packed word row of job j (2^log_n words);
// out[j]: tile [cols][2^(log_n+1)].  Returns the number of kernels launched through *launches.
cudaError_t launch_fft_packed(cudaStream_t st, int kind, const uint32_t* const* src, uint32_t* const* out, int njobs, int log_n,
                              const FftTables& tw, uint32_t* scratch, const StageHook* hook, int* launches, int shard_log, int first_half_only,
                              const PeerDst* peer) {
    using namespace fft2;
#define HOOK(name, b) do { if (hook) hook->fn(hook->user, name, b); } while (0)
    const int cpj = kind == SRC_BITS ? 32 : 4;
    int nl = 0;
    for (int j0 = 0; j0 < njobs; j0 += MAX_FFT_JOBS) {
        Jobs jobs;
        jobs.n = njobs - j0 < MAX_FFT_JOBS ? njobs - j0 : MAX_FFT_JOBS;
        jobs.one = 1u;
        jobs.shard_log = (shard_log > 0 && shard_log < log_n + 1) ? shard_log : log_n + 1;
        jobs.mone = 0xffffffffu;
        jobs.halves = (first_half_only && log_n > 12 && log_n <= 20 && !g_force_generic_fft && jobs.shard_log == log_n + 1) ? 1 : 2;
        for (int j = 0; j < jobs.n; j++) { jobs.src[j] = src[j0 + j]; jobs.out[j] = out[j0 + j]; }
        jobs.peer_on = 0;
        jobs.peer_logG = jobs.peer_lv = 0;
        jobs.wlist = nullptr;
        if (peer) {  // chunks of the last pass (2^12 or 2^k1 <= 2^13 rows) must lie inside one virtual shard
            if (log_n <= 12 || peer->lv < 13 || peer->logG > 3) return cudaErrorInvalidValue;
            jobs.peer_on = 1;
            jobs.peer_logG = peer->logG;
            jobs.peer_lv = peer->lv;
            for (int r = 0; r < MAX_PEERS; r++) jobs.peer_base[r] = r < (1 << peer->logG) ? peer->base[r] : nullptr;
            for (int j = 0; j < jobs.n; j++) jobs.peer_off[j] = peer->off[j0 + j];
        }
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
            else if (kind == SRC_BYTES) small_kernel<SRC_BYTES><<<jobs.n * gpj, threads, smem, st>>>(jobs, log_n, nc, gpj, tw);
            else small_kernel<SRC_M31><<<jobs.n * gpj, threads, smem, st>>>(jobs, log_n, nc, gpj, tw);
            HOOK("fft_small", 0);
            nl += 1;
            continue;
        }
        if (log_n <= 20 && !g_force_generic_fft) {
            constexpr int NC = 4;
            const int gpj2 = cpj / NC;
            uint32_t* scr2 = scratch;
            dim3 gA((1u << log_n) / T2, jobs.n * gpj2);
            HOOK("ifft_low", 1);
            if (kind == SRC_BITS) ifft_low12_kernel<SRC_BITS, NC><<<gA, 256, NC * COLW * 4, st>>>(jobs, log_n, gpj2, cpj, scr2, tw);
            else if (kind == SRC_BYTES) ifft_low12_kernel<SRC_BYTES, NC><<<gA, 256, NC * COLW * 4, st>>>(jobs, log_n, gpj2, cpj, scr2, tw);
            else ifft_low12_kernel<SRC_M31, NC><<<gA, 256, NC * COLW * 4, st>>>(jobs, log_n, gpj2, cpj, scr2, tw);
            HOOK("ifft_low", 0);
            HOOK("fft_mid", 1);
            switch (log_n - K1) {
                case 1: launch_mid12<1>(st, jobs, cpj, scr2, tw); break;
                case 2: launch_mid12<2>(st, jobs, cpj, scr2, tw); break;
                case 3: launch_mid12<3>(st, jobs, cpj, scr2, tw); break;
                case 4: launch_mid12<4>(st, jobs, cpj, scr2, tw); break;
                case 5: launch_mid12<5>(st, jobs, cpj, scr2, tw); break;
                case 6: launch_mid12<6>(st, jobs, cpj, scr2, tw); break;
                case 7: launch_mid12<7>(st, jobs, cpj, scr2, tw); break;
                default: launch_mid12<8>(st, jobs, cpj, scr2, tw); break;
            }
            HOOK("fft_mid", 0);
            dim3 gC(((unsigned)jobs.halves << log_n) / T2, jobs.n * gpj2);
            HOOK("fft_low", 1);
            cudaStream_t stC = st;
            if (peer && peer->last_stream) {
                cudaEventRecord(peer->ab_done, st);
                cudaStreamWaitEvent(peer->last_stream, peer->ab_done, 0);
                stC = peer->last_stream;
            }
            // (a persistent form of this pass with 1-2 resident blocks per SM, meant to leave room for the next group's passes
            // while the stores wait on NVLink, measured slower at 4 GPUs: 225 / 201 ms vs 189 ms)
            if (jobs.peer_on) fft_low12_kernel<NC, true><<<gC, 256, NC * COLW * 4, stC>>>(jobs, log_n, gpj2, tw);
            else fft_low12_kernel<NC, false><<<gC, 256, NC * COLW * 4, st>>>(jobs, log_n, gpj2, tw);
            HOOK("fft_low", 0);
            nl += 3;
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
        else if (kind == SRC_BYTES) ifft_low_kernel<SRC_BYTES><<<gA, 256, smem_ac, st>>>(jobs, log_n, k1, nc, gpj, cpj, scr, tw);
        else ifft_low_kernel<SRC_M31><<<gA, 256, smem_ac, st>>>(jobs, log_n, k1, nc, gpj, cpj, scr, tw);
        HOOK("ifft_low", 0);
        dim3 gB(T / Q, jobs.n * cpj);
        const size_t tile = (size_t)1 << (jb + qbits);
        HOOK("fft_mid", 1);
        if (qbits < 5) mid_kernel<true><<<gB, 256, 2 * (tile + (tile >> 4)) * 4, st>>>(jobs, log_n, k1, qbits, cpj, scr, tw);
        else mid_kernel<false><<<gB, 256, 2 * tile * 4, st>>>(jobs, log_n, k1, qbits, cpj, scr, tw);
        HOOK("fft_mid", 0);
        dim3 gC((2u << log_n) / T, jobs.n * gpj);
        HOOK("fft_low", 1);
        cudaStream_t stC = st;
        if (peer && peer->last_stream) {
            cudaEventRecord(peer->ab_done, st);
            cudaStreamWaitEvent(peer->last_stream, peer->ab_done, 0);
            stC = peer->last_stream;
        }
        fft_low_kernel<<<gC, 256, smem_ac, stC>>>(jobs, log_n, k1, nc, gpj, tw);
        HOOK("fft_low", 0);
        nl += 3;
    }
    if (launches) *launches += nl;
    return cudaGetLastError();
#undef HOOK
}
