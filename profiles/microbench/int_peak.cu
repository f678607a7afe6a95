// Integer issue-rate microbenchmarks for sm_100a (B200): measured peaks for the "integer roof" the FFT / Blake2s / constraint
// kernels are judged against (MEASURED_PEAKS.json has no INT32 figure).  Every kernel runs `iters` iterations of
// an unrolled body of independent dependency chains (8 per thread) on 148 x 8 blocks x 256 threads; the result is
// thread-instructions per second = warp-instructions x 32 (counted from the SASS the body compiles to: one SASS instruction
// per listed op, verified with cuobjdump, see profiles/int_peak_sass_r02.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o profiles/microbench/int_peak profiles/microbench/int_peak.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>

constexpr uint32_t P = 0x7fffffffu;
constexpr int CH = 8;   // independent chains per thread
constexpr int RP = 16;  // unrolled repetitions per loop iteration

enum Op { IADD3, LOP3, SHF, IMAD, IMADWIDE, IMADHI, VIADDMIN, LEAHI, PAIR_IADD_IMAD, PAIR_MIN_IMAD, PAIR_MIN_WIDE, PAIR_MIN_LOP, PAIR_WIDE_LOP, PAIR_HI_LOP,
          BFLY_CUR, BFLY_ALU, BFLY_SHOUP, BFLY_MIX, BLAKE_G,
          DFMA, DADD, I2D, PAIR_DFMA_IMAD, PAIR_DFMA_LOP, MAC_WIDE, MAC_WIDE4, MAC_F64, MAC_F64_MAGIC, N_OPS };

__device__ __forceinline__ uint32_t redp(uint32_t x) { return __viaddmin_u32(x, 0u - P, x); }
__device__ __forceinline__ uint32_t mulw(uint32_t a, uint32_t w2) {
    uint64_t v = (uint64_t)a * w2;
    return redp((uint32_t)(v >> 32) + (((uint32_t)v) >> 1));
}
// current butterfly (kernels_fft2.cu): adds on the FMA pipe via run-time 1 / -1
__device__ __forceinline__ void bfly_cur(uint32_t& a, uint32_t& b, uint32_t w2, uint32_t one, uint32_t mone) {
    uint32_t t = mulw(b, w2);
    uint32_t s = a * one + t, d = t * mone + a;
    a = redp(s);
    b = __viaddmin_u32(d, P, d);
}
// same with the adds on the ALU pipe
__device__ __forceinline__ void bfly_alu(uint32_t& a, uint32_t& b, uint32_t w2) {
    uint32_t t = mulw(b, w2);
    uint32_t s = a + t, d = a - t;
    a = redp(s);
    b = __viaddmin_u32(d, P, d);
}
// Shoup-style multiply entirely on the FMA pipe: q = hi(b * w'), t = b*w - q*p in [0, 2p)
__device__ __forceinline__ void bfly_shoup(uint32_t& a, uint32_t& b, uint32_t w, uint32_t wp, uint32_t one, uint32_t mone) {
    uint32_t q = __umulhi(b, wp);
    uint32_t t = redp(b * w + q * (0u - P));
    uint32_t s = a * one + t, d = t * mone + a;
    a = redp(s);
    b = __viaddmin_u32(d, P, d);
}

template <int OP>
__global__ void __launch_bounds__(256) bench_kernel(uint32_t* out, uint32_t seed, int iters, uint32_t one, uint32_t mone) {
    uint32_t v[CH], u[CH];
    uint64_t acc[CH];
    double dacc[CH], dm = (double)(seed & 0xffffu), da = (double)((seed >> 7) & 0xffffu);
    // split-alpha tables as the constraint kernel uses them: 16-bit halves of 4 coordinates (integers / doubles)
    const uint4 tl = make_uint4(seed & 0xffffu, (seed >> 3) & 0xffffu, (seed >> 5) & 0xffffu, (seed >> 7) & 0xffffu);
    const uint4 th = make_uint4((seed >> 9) & 0xffffu, (seed >> 11) & 0xffffu, (seed >> 13) & 0xffffu, (seed >> 15) & 0xffffu);
    const double tdl[4] = {(double)tl.x, (double)tl.y, (double)tl.z, (double)tl.w}, tdh[4] = {(double)th.x, (double)th.y, (double)th.z, (double)th.w};
    uint64_t m8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double d8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < CH; i++) {
        v[i] = (seed * (threadIdx.x + 1) + i * 0x9e3779b9u) & P;
        u[i] = (seed * (blockIdx.x + 7) + threadIdx.x * 0x27d4eb2fu + i * 0x85ebca6bu) & P;
        acc[i] = v[i];
        dacc[i] = (double)v[i];
    }
    const uint32_t w2 = ((seed * 77u) & P) << 1, w = (seed * 77u) & P, wp = (w << 1) + (2 * (uint64_t)w >= P);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < RP; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (OP == IADD3) { v[i] = v[i] + u[i] + seed; u[i] = u[i] + v[i] + one; }  // two dependent 3-input adds
                else if (OP == LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(u[i]), "r"(seed)); }
                else if (OP == SHF) { asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(v[i]) : "r"(u[i])); }
                else if (OP == IMAD) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(u[i]), "r"(seed)); }
                else if (OP == IMADWIDE) { acc[i] = (uint64_t)(uint32_t)acc[i] * u[i] + acc[i]; }  // multiplicand = low word of the accumulator: not loop-invariant
                else if (OP == IMADHI) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(u[i])); }
                else if (OP == VIADDMIN) { v[i] = __viaddmin_u32(v[i], u[i], v[i]); }
                else if (OP == LEAHI) { uint32_t hi = v[i], lo = u[i]; v[i] = hi + (lo >> 1); u[i] = lo ^ hi; }  // LEA.HI + LOP3
                else if (OP == PAIR_IADD_IMAD) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(seed), "r"(one));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(seed), "r"(one));
                } else if (OP == PAIR_MIN_IMAD) {
                    v[i] = __viaddmin_u32(v[i], seed, v[i]);
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(seed), "r"(one));
                } else if (OP == PAIR_MIN_WIDE) {
                    v[i] = __viaddmin_u32(v[i], seed, v[i]);
                    acc[i] = (uint64_t)(uint32_t)acc[i] * seed + acc[i];
                } else if (OP == PAIR_MIN_LOP) {
                    v[i] = __viaddmin_u32(v[i], seed, v[i]);
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(seed), "r"(one));
                } else if (OP == PAIR_WIDE_LOP) {
                    acc[i] = (uint64_t)(uint32_t)acc[i] * seed + acc[i];
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(seed), "r"(one));
                } else if (OP == PAIR_HI_LOP) {
                    asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(v[i]) : "r"(seed));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(seed), "r"(one));
                } else if (OP == BFLY_CUR) { bfly_cur(v[i], u[i], w2, one, mone); }
                else if (OP == BFLY_ALU) { bfly_alu(v[i], u[i], w2); }
                else if (OP == BFLY_SHOUP) { bfly_shoup(v[i], u[i], w, wp, one, mone); }
                else if (OP == BFLY_MIX) { if (i & 1) bfly_shoup(v[i], u[i], w, wp, one, mone); else bfly_cur(v[i], u[i], w2, one, mone); }
                else if (OP == DFMA) { dacc[i] = fma(dacc[i], dm, da); }
                else if (OP == DADD) { dacc[i] = dacc[i] + dm; }
                else if (OP == I2D) { dacc[i] += (double)v[i]; v[i] ^= (uint32_t)__double2loint(dacc[i]); }  // I2F.F64.U32 + DADD + LOP3
                else if (OP == PAIR_DFMA_IMAD) {
                    dacc[i] = fma(dacc[i], dm, da);
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(seed), "r"(one));
                } else if (OP == PAIR_DFMA_LOP) {
                    dacc[i] = fma(dacc[i], dm, da);
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(seed), "r"(one));
                } else if (OP == MAC_WIDE) {   // the constraint kernel's multiply-accumulate today: 8 IMAD.WIDE per constraint value
                    const uint32_t C = v[i]; v[i] = v[i] * one + u[i];
                    m8[0] += (uint64_t)C * tl.x; m8[1] += (uint64_t)C * tl.y; m8[2] += (uint64_t)C * tl.z; m8[3] += (uint64_t)C * tl.w;
                    m8[4] += (uint64_t)C * th.x; m8[5] += (uint64_t)C * th.y; m8[6] += (uint64_t)C * th.z; m8[7] += (uint64_t)C * th.w;
                } else if (OP == MAC_WIDE4) {  // unsplit coefficients: 4 IMAD.WIDE + a fold of the accumulators every 4 values
                    const uint32_t C = v[i] & P; v[i] = v[i] * one + u[i];
                    m8[0] += (uint64_t)C * tl.x * 3; m8[1] += (uint64_t)C * tl.y; m8[2] += (uint64_t)C * tl.z; m8[3] += (uint64_t)C * tl.w;
                    if ((i & 3) == 3) {
#pragma unroll
                        for (int c = 0; c < 4; c++) m8[c] = (m8[c] & P) + (m8[c] >> 31);
                    }
                } else if (OP == MAC_F64) {    // the same in FP64: convert the value once, 8 DFMA
                    const double C = (double)v[i]; v[i] = v[i] * one + u[i];
#pragma unroll
                    for (int c = 0; c < 4; c++) { d8[c] = fma(C, tdl[c], d8[c]); d8[4 + c] = fma(C, tdh[c], d8[4 + c]); }
                } else if (OP == MAC_F64_MAGIC) {  // conversion by exponent trick (2^52 + C) - 2^52: one DADD instead of I2F
                    const double C = __hiloint2double(0x43300000, (int)v[i]) - 4503599627370496.0; v[i] = v[i] * one + u[i];
#pragma unroll
                    for (int c = 0; c < 4; c++) { d8[c] = fma(C, tdl[c], d8[c]); d8[4 + c] = fma(C, tdh[c], d8[4 + c]); }
                }
                else if (OP == BLAKE_G) {  // one Blake2s half-G: a += b + m; d = rotr(d ^ a, 16); c += d; b = rotr(b ^ c, 12)
                    uint32_t a = v[i], b = u[i], c = (uint32_t)acc[i], d = (uint32_t)(acc[i] >> 32);
                    a = a * one + b; a = a * one + seed; d = __funnelshift_r(d ^ a, d ^ a, 16);
                    c = c * one + d; b = __funnelshift_r(b ^ c, b ^ c, 12);
                    v[i] = a; u[i] = b; acc[i] = ((uint64_t)d << 32) | c;
                }
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < CH; i++)
        x ^= v[i] ^ u[i] ^ (uint32_t)acc[i] ^ (uint32_t)(acc[i] >> 32) ^ (uint32_t)__double2loint(dacc[i]) ^ (uint32_t)__double2hiint(dacc[i]) ^
             (uint32_t)m8[i] ^ (uint32_t)(m8[i] >> 32) ^ (uint32_t)__double2loint(d8[i]) ^ (uint32_t)__double2hiint(d8[i]);
    if (x == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

struct Spec { const char* name; int op; double sass_per_item; const char* what; };

template <int OP>
float run(uint32_t* d_out, int blocks, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench_kernel<OP><<<blocks, 256>>>(d_out, 12345u, iters / 8, 1u, 0xffffffffu);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        bench_kernel<OP><<<blocks, 256>>>(d_out, 12345u + rep, iters, 1u, 0xffffffffu);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const int blocks = pr.multiProcessorCount * 8, iters = 1024;
    uint32_t* d_out; cudaMalloc(&d_out, (size_t)blocks * 256 * 4);
    const double items = (double)blocks * 256 * iters * RP * CH;  // op-items executed per launch
    struct R { const char* name; float ms; double per_item; const char* what; };
    std::vector<R> rs;
#define RUN(OPN, per, what) rs.push_back({#OPN, run<OPN>(d_out, blocks, iters), per, what});
    RUN(IADD3, 2, "3-input adds (IADD3 on the ALU pipe; ptxas may move some to IMAD)")
    RUN(LOP3, 1, "LOP3 (ALU pipe)")
    RUN(SHF, 1, "SHF (ALU pipe)")
    RUN(IMAD, 1, "IMAD (FMA pipe)")
    RUN(IMADWIDE, 1, "IMAD.WIDE.U32 with 64-bit accumulate (FMA pipe)")
    RUN(IMADHI, 1, "IMAD.HI.U32 (FMA pipe)")
    RUN(VIADDMIN, 1, "VIADDMNMX.U32 (DPX add+min, ALU pipe)")
    RUN(LEAHI, 2, "LEA.HI + LOP3 (both ALU pipe)")
    RUN(PAIR_IADD_IMAD, 2, "LOP3 + IMAD issued together (ALU + FMA pipes)")
    RUN(PAIR_MIN_IMAD, 2, "VIADDMNMX + IMAD (ALU + FMA pipes)")
    RUN(PAIR_MIN_WIDE, 2, "VIADDMNMX + IMAD.WIDE (ALU + FMA pipes)")
    RUN(PAIR_MIN_LOP, 2, "VIADDMNMX + LOP3")
    RUN(PAIR_WIDE_LOP, 2, "IMAD.WIDE + LOP3 (FMA + ALU pipes)")
    RUN(PAIR_HI_LOP, 2, "IMAD.HI + LOP3 (FMA + ALU pipes)")
    RUN(BFLY_CUR, 1, "M31 butterfly as in kernels_fft2.cu: IMAD.WIDE, LEA.HI, 3 VIADDMNMX, 2 IMAD (4 ALU + 3 FMA)")
    RUN(BFLY_ALU, 1, "M31 butterfly, adds on the ALU pipe: IMAD.WIDE, LEA.HI, 3 VIADDMNMX, 2 IADD3 (6 ALU + 1 FMA)")
    RUN(BFLY_SHOUP, 1, "M31 butterfly, Shoup multiply on the FMA pipe: IMAD.HI, 2 IMAD, 3 VIADDMNMX, 2 IMAD (3 ALU + 5 FMA)")
    RUN(BFLY_MIX, 1, "alternating the two butterflies (3.5 ALU + 4 FMA)")
    RUN(BLAKE_G, 1, "Blake2s half-G: 3 IMAD adds (FMA) + 2 LOP3 + 2 SHF (ALU)")
    RUN(DFMA, 1, "DFMA (FP64 pipe)")
    RUN(DADD, 1, "DADD (FP64 pipe)")
    RUN(I2D, 3, "I2F.F64.U32 + DADD + LOP3")
    RUN(PAIR_DFMA_IMAD, 2, "DFMA + IMAD (FP64 + FMA pipes)")
    RUN(PAIR_DFMA_LOP, 2, "DFMA + LOP3 (FP64 + ALU pipes)")
    RUN(MAC_WIDE, 1, "constraint multiply-accumulate as in kernels_stream.cu: 8 IMAD.WIDE per value (items = values)")
    RUN(MAC_WIDE4, 1, "unsplit coefficients: 4 IMAD.WIDE per value + accumulator folds every 4 values")
    RUN(MAC_F64, 1, "FP64 form: I2F + 8 DFMA per value")
    RUN(MAC_F64_MAGIC, 1, "FP64 form with exponent-trick conversion: DADD + 8 DFMA per value")
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d, \"blocks\": %d, \"threads\": 256, \"results\": [\n", pr.name,
           pr.multiProcessorCount, clk, blocks);
    for (size_t i = 0; i < rs.size(); i++) {
        const double rate = items / (rs[i].ms * 1e-3);  // items per second
        printf("  {\"op\": \"%s\", \"ms\": %.4f, \"items_per_s\": %.4e, \"thread_instr_per_s\": %.4e, \"per_sm_per_clk_at_1965MHz\": %.1f, \"what\": \"%s\"}%s\n",
               rs[i].name, rs[i].ms, rate, rate * rs[i].per_item, rate * rs[i].per_item / pr.multiProcessorCount / 1.965e9, rs[i].what,
               i + 1 < rs.size() ? "," : "");
    }
    printf("]}\n");
    return 0;
}
