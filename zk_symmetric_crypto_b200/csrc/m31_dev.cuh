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

}  // namespace m31d
