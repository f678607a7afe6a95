// Small backend-trait kernels of the drop-in boundary: ColumnOps::bit_reverse_column, FieldOps::batch_inverse,
// the LogUp column finalisation (prefix sum in coset order), legacy MerkleOps::commit_on_layer.
//
// Reference call sites: upstream `LogupTraceGenerator::{finalize_col, finalize_last}` reached from
// /root/reference/stwo/src/aes/lookup/gen_ctr.rs:648-682 and aes/lookup/gen.rs:448-477 (batch inverse of the fraction
// denominators, running sums, claimed sum, prefix sum of the last column in circle-domain coset order);
// `ColumnOps::bit_reverse_column` / `FieldOps::batch_inverse` are the `Backend` bounds of upstream stwo (SimdBackend:
// prover/backend/simd/{bit_reverse,m31,qm31,prefix_sum}.rs).
#include "common.cuh"
#include "blake2s.cuh"
#include "m31_dev.cuh"

namespace ops {
using namespace m31;

// ------------------------------------------------------------------------------------------------ bit reverse
// in place: element i <-> element bitrev(i); one thread per pair with i < bitrev(i).  elem_words = 1 (M31 / one coordinate column)
__global__ void bit_reverse_kernel(uint32_t* __restrict__ col, int log_size) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << log_size)) return;
    const uint32_t j = __brev(i) >> (32 - log_size);
    if (i < j) {
        const uint32_t a = col[i], b = col[j];
        col[i] = b;
        col[j] = a;
    }
}

// ------------------------------------------------------------------------------------------------ inverses
// dst[i] = src[i]^-1 (0 -> 0, like Fermat exponentiation of upstream's `inverse` on zero would assert; callers never pass zero)
__global__ void inverse_m31_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = inv(src[i]);
}
// QM31 elements as 4 coordinate columns `stride` words apart
__global__ void inverse_qm31_kernel(const uint32_t* __restrict__ src, size_t s_stride, uint32_t* __restrict__ dst, size_t d_stride,
                                    size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    QM31 x{{src[i], src[s_stride + i], src[2 * s_stride + i], src[3 * s_stride + i]}};
    const QM31 r = qinv(x);
#pragma unroll
    for (int c = 0; c < 4; c++) dst[(size_t)c * d_stride + i] = r.v[c];
}

// ------------------------------------------------------------------------------------------------ LogUp finalize_last
// A QM31 column (4 coordinate columns in bit-reversed circle-domain storage order) is replaced by the inclusive prefix sum, in
// the order of the trace-domain coset, of (value - claimed_sum / N).  prefix(v - s)[i] = prefix(v)[i] - (i + 1) s, so one scan of
// the raw values gives both the claimed sum (its last element) and the result.
// storage index of the i-th coset point: circle-domain index cd = i/2 (i even) or N-1-i/2 (i odd), bit-reversed.
__device__ __forceinline__ uint32_t coset_to_storage(uint32_t i, int log) {
    const uint32_t n = 1u << log;
    const uint32_t cd = (i & 1u) ? n - 1 - (i >> 1) : (i >> 1);
    return log ? (__brev(cd) >> (32 - log)) : 0;
}

constexpr int SCAN_T = 256, SCAN_PER = 8, SCAN_CHUNK = SCAN_T * SCAN_PER;  // elements of one coordinate per block

// phase A: block-local inclusive scan in coset order -> tmp[c][i] (coset order), block totals -> btot[c][block]
__global__ void __launch_bounds__(SCAN_T) logup_scan_local_kernel(const uint32_t* __restrict__ col, size_t stride, int log,
                                                                  uint32_t* __restrict__ tmp, uint32_t* __restrict__ btot) {
    __shared__ uint32_t wsum[SCAN_T / 32];
    const int c = blockIdx.y;
    const size_t n = (size_t)1 << log;
    const size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_PER;
    uint32_t v[SCAN_PER];
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER; k++) {
        const size_t i = base + k;
        const uint32_t x = i < n ? col[(size_t)c * stride + coset_to_storage((uint32_t)i, log)] : 0;
        run = add(run, x);
        v[k] = run;
    }
    // exclusive scan of the per-thread totals across the block
    uint32_t incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl = add(incl, y);
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff = add(woff, wsum[w]);
    const uint32_t excl = add(woff, sub(incl, run));
#pragma unroll
    for (int k = 0; k < SCAN_PER; k++) {
        const size_t i = base + k;
        if (i < n) tmp[(size_t)c * n + i] = add(v[k], excl);
    }
    if (threadIdx.x == SCAN_T - 1) btot[(size_t)c * gridDim.x + blockIdx.x] = add(excl, run);
}

// phase B (one block per coordinate): exclusive scan of the block totals in place, total -> claimed[c]
__global__ void logup_scan_totals_kernel(uint32_t* __restrict__ btot, int n_blocks, uint32_t* __restrict__ claimed) {
    if (threadIdx.x != 0) return;
    const int c = blockIdx.x;
    uint32_t run = 0;
    for (int b = 0; b < n_blocks; b++) {
        const uint32_t t = btot[(size_t)c * n_blocks + b];
        btot[(size_t)c * n_blocks + b] = run;
        run = add(run, t);
    }
    claimed[c] = run;
}

// phase C: col[storage(i)] = tmp[i] + block offset - (i + 1) * claimed / N
__global__ void __launch_bounds__(SCAN_T) logup_scan_apply_kernel(uint32_t* __restrict__ col, size_t stride, int log,
                                                                  const uint32_t* __restrict__ tmp, const uint32_t* __restrict__ btot,
                                                                  const uint32_t* __restrict__ claimed, uint32_t n_inv) {
    const int c = blockIdx.y;
    const size_t n = (size_t)1 << log;
    const uint32_t shift = mul(claimed[c], n_inv);
    const uint32_t boff = btot[(size_t)c * gridDim.x + blockIdx.x];
    const size_t base = (size_t)blockIdx.x * SCAN_CHUNK;
    for (int k = threadIdx.x; k < SCAN_CHUNK; k += SCAN_T) {
        const size_t i = base + k;
        if (i >= n) break;
        const uint32_t idx1 = (uint32_t)((i + 1) % P);
        col[(size_t)c * stride + coset_to_storage((uint32_t)i, log)] = sub(add(tmp[(size_t)c * n + i], boff), mul(idx1, shift));
    }
}

// ------------------------------------------------------------------------------------------------ legacy commit_on_layer
// out[i] = Blake2s( prev[2i] || prev[2i+1] || col_0[i] || col_1[i] || ... )   (children only when prev != nullptr; LE u32 values)
// upstream MerkleOps::commit_on_layer (non-lifted VCS).  The pinned reference commits through the lifted VCS only, so this entry
// point has no reference call site; it is the RFC 7693 hash of the concatenation described above.
__global__ void commit_on_layer_kernel(const uint32_t* __restrict__ prev, const uint32_t* const* __restrict__ cols, int n_cols,
                                       uint32_t n_nodes, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    uint32_t h[8], m[16];
    blake2s::init(h);
    uint64_t t = 0;
    int k = 0;
    const int total_words = (prev ? 16 : 0) + n_cols;
    int done = 0;
    auto push = [&](uint32_t w) {
        m[k++] = w;
        done++;
        if (k == 16 && done < total_words) {  // a full block that is not the last one
            t += 64;
            blake2s::compress(h, m, t, false);
            k = 0;
        }
    };
    if (prev) {
        for (int w = 0; w < 16; w++) push(prev[(size_t)i * 16 + w]);
    }
    for (int c = 0; c < n_cols; c++) push(cols[c][i]);
    for (int w = k; w < 16; w++) m[w] = 0;
    t += 4 * (uint64_t)k;
    blake2s::compress(h, m, t, true);
#pragma unroll
    for (int w = 0; w < 8; w++) out[(size_t)i * 8 + w] = h[w];
}

}  // namespace ops

cudaError_t launch_bit_reverse(cudaStream_t st, uint32_t* col, int log_size) {
    const uint32_t n = 1u << log_size;
    ops::bit_reverse_kernel<<<(n + 255) / 256, 256, 0, st>>>(col, log_size);
    return cudaGetLastError();
}

cudaError_t launch_inverse_m31(cudaStream_t st, const uint32_t* src, uint32_t* dst, size_t n) {
    ops::inverse_m31_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(src, dst, n);
    return cudaGetLastError();
}

cudaError_t launch_inverse_qm31(cudaStream_t st, const uint32_t* src, size_t s_stride, uint32_t* dst, size_t d_stride, size_t n) {
    ops::inverse_qm31_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(src, s_stride, dst, d_stride, n);
    return cudaGetLastError();
}

size_t logup_finalize_scratch_words(int log) {
    const size_t n = (size_t)1 << log;
    const size_t blocks = (n + ops::SCAN_CHUNK - 1) / ops::SCAN_CHUNK;
    return 4 * n + 4 * blocks + 4;
}

// col: 4 coordinate columns `stride` apart (device); scratch: logup_finalize_scratch_words(log) words; the claimed sum is left
// in scratch[4n + 4 blocks .. +4) (returned pointer) for the caller to copy out
cudaError_t launch_logup_finalize_last(cudaStream_t st, uint32_t* col, size_t stride, int log, uint32_t* scratch,
                                       uint32_t** claimed_dev) {
    const size_t n = (size_t)1 << log;
    const unsigned blocks = (unsigned)((n + ops::SCAN_CHUNK - 1) / ops::SCAN_CHUNK);
    uint32_t* tmp = scratch;
    uint32_t* btot = scratch + 4 * n;
    uint32_t* claimed = btot + 4 * (size_t)blocks;
    const uint32_t n_inv = m31::inv((uint32_t)(((uint64_t)1 << log) % m31::P));
    ops::logup_scan_local_kernel<<<dim3(blocks, 4), ops::SCAN_T, 0, st>>>(col, stride, log, tmp, btot);
    ops::logup_scan_totals_kernel<<<4, 32, 0, st>>>(btot, (int)blocks, claimed);
    ops::logup_scan_apply_kernel<<<dim3(blocks, 4), ops::SCAN_T, 0, st>>>(col, stride, log, tmp, btot, claimed, n_inv);
    *claimed_dev = claimed;
    return cudaGetLastError();
}

cudaError_t launch_commit_on_layer(cudaStream_t st, const uint32_t* prev, const uint32_t* const* cols_dev, int n_cols, uint32_t n_nodes,
                                   uint32_t* out) {
    ops::commit_on_layer_kernel<<<(n_nodes + 127) / 128, 128, 0, st>>>(prev, cols_dev, n_cols, n_nodes, out);
    return cudaGetLastError();
}
