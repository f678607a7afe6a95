// Shared declarations for the CUDA side of the B200 stwo backend.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "m31.cuh"

// Where a column's trace-domain values come from.
enum { SRC_M31 = 0, SRC_BITS = 1, SRC_BYTES = 2 };
struct ColSrc {
    int kind;              // SRC_M31: base[col*stride + row];  SRC_BITS: bit (c&31) of base[(c>>5)*stride + row];
                           // SRC_BYTES: byte (c&3) of base[(c>>2)*stride + row];  c = col + first_col
    const uint32_t* base;
    size_t stride;         // words between consecutive columns (SRC_M31) or packed words (SRC_BITS/BYTES)
    uint32_t first_col;
};

struct FftTables {
    const uint32_t *X, *Y, *IX, *IY;  // flattened twiddle tables, see kernels_fft.cu
    int max_log;                      // largest canonic domain covered
};

// Equally sized, equally strided columns feeding Merkle leaves (see kernels_merkle.cu)
struct LeafGroup {
    const uint32_t* base;
    size_t stride;
    int ncols;
    int log_size;
};
#define MAX_LEAF_GROUPS 8
struct LeafGroups {
    LeafGroup g[MAX_LEAF_GROUPS];
    int n;
};

// One sample batch of the FRI quotient accumulation (see kernels_pcs.cu quotients_kernel)
struct QuotBatch {
    m31::CM31 prx, pry, pix, piy;  // sample point P = Pr + u*Pi
    m31::QM31 lin_a, lin_b;        // sum_j alpha_j a_j, sum_j alpha_j b_j
    m31::QM31 batch_coeff;         // multiplier applied to the running row accumulator before adding this batch
    const uint32_t* coefs;         // device [n_cols][4] = alpha_j * c_j
    const uint32_t* col_idx;       // device [n_cols] column indices (nullptr = identity)
    int n_cols;
};

// optional per-kernel profiling callback (begin=1 before a launch, begin=0 after it)
struct StageHook {
    void (*fn)(void* user, const char* name, int begin);
    void* user;
};

cudaError_t launch_fft(cudaStream_t st, const ColSrc& src, int ncols, int log_n, int ext, int mode, uint32_t* coef_out,
                       size_t coef_stride, uint32_t* eval_out, size_t eval_stride, const FftTables& tw, uint32_t* scratch,
                       size_t scratch_stride, const StageHook* hook = nullptr);
void fft_init_attrs();
