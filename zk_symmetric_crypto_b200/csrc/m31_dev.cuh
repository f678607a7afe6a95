// Device-only M31 arithmetic tuned for sm_100a issue slots (same field as m31.cuh; all results canonical in [0,p)).
//
// reduce uses the DPX fused add+min (VIADDMNMX): one ALU-pipe instruction instead of subtract+compare+select.
// mulw takes the multiplier pre-doubled (w2 = 2w): a*w2 = hi*2^32 + lo with lo even, so a*w = hi*2^31 + lo/2 == hi + (lo>>1)
// (mod p); ptxas fuses "hi + (lo>>1)" into one LEA.HI.  A butterfly is 7 instructions (FFT section).
#pragma once
#include "m31.cuh"

namespace m31d {
using m31::P;

// [0, 2p) -> [0, p)
__device__ __forceinline__ uint32_t redp(uint32_t x) { return __viaddmin_u32(x, 0u - P, x); }
__device__ __forceinline__ uint32_t addm(uint32_t a, uint32_t b) { return redp(a + b); }
// a, b in [0,p): a-b wraps to a huge value when negative, then (d + p) is the small one
__device__ __forceinline__ uint32_t subm(uint32_t a, uint32_t b) {
    uint32_t d = a - b;
    return __viaddmin_u32(d, P, d);
}
// a in [0, 2^31], w2 = 2*w with w in [0,p)  ->  a*w mod p
__device__ __forceinline__ uint32_t mulw(uint32_t a, uint32_t w2) {
    uint64_t v = (uint64_t)a * w2;
    return redp((uint32_t)(v >> 32) + (((uint32_t)v) >> 1));
}
__device__ __forceinline__ uint32_t mulm(uint32_t a, uint32_t b) { return mulw(a, b << 1); }
__device__ __forceinline__ uint32_t dbl(uint32_t a) { return redp(a + a); }

// general 64-bit value -> [0,p)
__device__ __forceinline__ uint32_t red64(uint64_t v) {
    uint64_t t = (v & P) + (v >> 31);                           // < 2^34
    uint32_t u = (uint32_t)(t & P) + (uint32_t)(t >> 31);       // < 2^31 + 8
    return redp(u);
}


// Accumulator of sum_k alpha_k * C_k with alpha_k in QM31 (4 coordinates) and C_k in M31.  Every coordinate of alpha_k is
// pre-split into 16-bit halves (two tables), so each product C * half < 2^48 and 2^16 products fit a 64-bit accumulator
// without intermediate reduction: a multiply-accumulate is 8 IMAD.WIDE on the FMA pipe and nothing on the ALU pipe, which is
// the pipe every other instruction of this kernel needs.  C may be any 32-bit representative (not necessarily < p).
struct AccSplit {
    uint64_t lo[4], hi[4];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int c = 0; c < 4; c++) lo[c] = hi[c] = 0;
    }
    __device__ __forceinline__ void mac(const uint4* __restrict__ tlo, const uint4* __restrict__ thi, int k, uint32_t C) {
        const uint4 l = __ldg(tlo + k), h = __ldg(thi + k);
        lo[0] += (uint64_t)C * l.x; lo[1] += (uint64_t)C * l.y; lo[2] += (uint64_t)C * l.z; lo[3] += (uint64_t)C * l.w;
        hi[0] += (uint64_t)C * h.x; hi[1] += (uint64_t)C * h.y; hi[2] += (uint64_t)C * h.z; hi[3] += (uint64_t)C * h.w;
    }
    __device__ __forceinline__ uint32_t result(int c) const { return addm(mulm(red64(hi[c]), 1u << 16), red64(lo[c])); }
};

// lazily reduced helpers: inputs canonical, outputs any representative < 2^32 unless noted
// b - b^2 + p  in (0, 2p):  b2 = 2b (unreduced)
__device__ __forceinline__ uint32_t bool_c(uint32_t b, uint32_t b2) {
    const uint64_t v = (uint64_t)b * b2;
    const uint32_t sq = redp((uint32_t)(v >> 32) + (((uint32_t)v) >> 1));
    return b + P - sq;
}

}  // namespace m31d
