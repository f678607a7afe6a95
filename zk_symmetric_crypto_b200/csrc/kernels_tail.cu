// Tail kernels of the streaming ChaCha20 prover: the per-column QM31 loops of upstream `prove()` that do not touch the trace
// (closing constraint check at the sampled mask, adder-sum samples, FRI quotient line coefficients).  They are a few
// hundred microseconds of host work each in the reference (negligible next to its FFTs); here they dominate a product-size
// proof once the heavy passes run on the GPU, so they are single-block kernels over data that is already device-resident.
// Semantics: /root/reference/stwo/src/chacha/bitwise/constraints_stream.rs:20-131 (the AIR), upstream
// core/pcs/quotients.rs (column_line_coeffs, fri_answers' random-coefficient powers).
#include "common.cuh"

namespace tail {
using namespace m31;

__device__ __forceinline__ QM31 block_sum(QM31 v, QM31* sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = qadd(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    QM31 r = sh[0];
    __syncthreads();
    return r;
}

struct MaskQ {
    const uint4* m;
    __device__ QM31 operator[](int c) const { uint4 x = __ldg(m + c); return {{x.x, x.y, x.z, x.w}}; }
};
struct MaskB {
    const uint32_t* m;
    __device__ QM31 operator[](int c) const { return {{__ldg(m + c), 0, 0, 0}}; }
};

template <typename Mask>
__device__ QM31 rec_eval(const ConsRec& r, const Mask& v) {
    switch (r.type) {
        case CR_BOOL: { QM31 b = v[r.c0]; return qmul(b, qsub(qone(), b)); }
        case CR_ADD: {
            QM31 car = v[r.c1];
            QM31 t = qsub(qsub(qadd(v[r.c0], qadd(car, car)), v[r.c2]), v[r.c3]);
            return r.c4 >= 0 ? qsub(t, v[r.c4]) : t;
        }
        case CR_XOR: {
            QM31 a = v[r.c1], b = v[r.c2], ab = qmul(a, b);
            return qadd(qsub(qsub(v[r.c0], a), b), qadd(ab, ab));
        }
        default: {
            QM31 k = v[r.c0], p = v[r.c1], kp = qmul(k, p);
            return qsub(qsub(qadd(k, p), qadd(kp, kp)), v[r.c2]);
        }
    }
}

__global__ void __launch_bounds__(1024) mask_constraints_kernel(const ConsRec* __restrict__ table, int K, const uint32_t* __restrict__ mask,
                                                                int mask_is_base, const uint4* __restrict__ apr, uint32_t* __restrict__ out) {
    __shared__ QM31 sh[1024];
    QM31 acc = qzero();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const ConsRec r = table[k];
        const QM31 c = mask_is_base ? rec_eval(r, MaskB{mask}) : rec_eval(r, MaskQ{(const uint4*)mask});
        const uint4 a = __ldg(apr + k);
        acc = qadd(acc, qmul(c, QM31{{a.x, a.y, a.z, a.w}}));
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0)
        for (int c = 0; c < 4; c++) out[c] = acc.v[c];
}

// thread = (bit i, coordinate k); a thread only ever reads entries it wrote itself or entries of carry / independent words
__global__ void oods_fill_sums_kernel(uint32_t* __restrict__ s, const SumComb* __restrict__ combs, int n) {
    const int i = threadIdx.x >> 2, k = threadIdx.x & 3;
    for (int t = 0; t < n; t++) {
        const SumComb cb = combs[t];
        const uint32_t cv = s[((size_t)cb.c * 32 + i) * 4 + k];
        const uint32_t cin = i ? s[((size_t)cb.c * 32 + i - 1) * 4 + k] : 0;
        const uint32_t av = s[((size_t)cb.a * 32 + i) * 4 + k], bv = s[((size_t)cb.b * 32 + i) * 4 + k];
        s[((size_t)cb.res * 32 + i) * 4 + k] = sub(add(add(av, bv), cin), add(cv, cv));
    }
}

__global__ void __launch_bounds__(1024) quot_coefs_kernel(const uint4* __restrict__ sampled, int nc, const uint4* __restrict__ pw_rev, QM31 zy,
                                                          uint4* __restrict__ coefs, uint32_t* __restrict__ lin) {
    __shared__ QM31 sh[1024];
    const QM31 c = qsub(qconj(zy), zy);
    QM31 la = qzero(), lb = qzero();
    for (int j = threadIdx.x; j < nc; j += blockDim.x) {
        const uint4 sv = __ldg(sampled + j), pv = __ldg(pw_rev + (nc - 1 - j));
        const QM31 v{{sv.x, sv.y, sv.z, sv.w}}, alpha{{pv.x, pv.y, pv.z, pv.w}};
        const QM31 a = qsub(qconj(v), v);
        const QM31 b = qsub(qmul(v, c), qmul(a, zy));
        la = qadd(la, qmul(alpha, a));
        lb = qadd(lb, qmul(alpha, b));
        const QM31 ac = qmul(alpha, c);
        coefs[j] = make_uint4(ac.v[0], ac.v[1], ac.v[2], ac.v[3]);
    }
    la = block_sum(la, sh);
    lb = block_sum(lb, sh);
    if (threadIdx.x == 0)
        for (int k = 0; k < 4; k++) { lin[k] = la.v[k]; lin[4 + k] = lb.v[k]; }
}

// reverse order: a sum may feed a later sum.  kap_i = fc[res][i]: operands a and b gain kap_i, the carry word gains
// -2 kap_i + kap_(i+1) (bit i+1's carry-in).  thread = (bit i, coordinate k)
__global__ void quot_fold_sums_kernel(uint32_t* __restrict__ fc, const SumComb* __restrict__ combs, int n) {
    const int i = threadIdx.x >> 2, k = threadIdx.x & 3;
    for (int t = n - 1; t >= 0; t--) {
        const SumComb cb = combs[t];
        const uint32_t kap = fc[((size_t)cb.res * 32 + i) * 4 + k];
        const uint32_t kap_up = i < 31 ? fc[((size_t)cb.res * 32 + i + 1) * 4 + k] : 0;
        uint32_t* fa = fc + ((size_t)cb.a * 32 + i) * 4 + k;
        *fa = add(*fa, kap);
        uint32_t* fb = fc + ((size_t)cb.b * 32 + i) * 4 + k;
        *fb = add(*fb, kap);
        uint32_t* fy = fc + ((size_t)cb.c * 32 + i) * 4 + k;
        *fy = add(sub(*fy, add(kap, kap)), kap_up);
        __syncthreads();  // kap_up of a later iteration is an entry thread i+1 may have updated in this one
    }
}

}  // namespace tail

cudaError_t launch_mask_constraints(cudaStream_t st, const ConsRec* table, int K, const uint32_t* mask, int mask_is_base,
                                    const uint32_t* apr, uint32_t* out) {
    tail::mask_constraints_kernel<<<1, 1024, 0, st>>>(table, K, mask, mask_is_base, (const uint4*)apr, out);
    return cudaGetLastError();
}
cudaError_t launch_oods_fill_sums(cudaStream_t st, uint32_t* sampled, const SumComb* combs, int n_combs) {
    tail::oods_fill_sums_kernel<<<1, 128, 0, st>>>(sampled, combs, n_combs);
    return cudaGetLastError();
}
cudaError_t launch_quot_coefs(cudaStream_t st, const uint32_t* sampled, int nc, const uint32_t* pw_rev, m31::QM31 zy, uint32_t* coefs,
                              uint32_t* lin) {
    tail::quot_coefs_kernel<<<1, 1024, 0, st>>>((const uint4*)sampled, nc, (const uint4*)pw_rev, zy, (uint4*)coefs, lin);
    return cudaGetLastError();
}
cudaError_t launch_quot_fold_sums(cudaStream_t st, uint32_t* fc, const SumComb* combs, int n_combs) {
    tail::quot_fold_sums_kernel<<<1, 128, 0, st>>>(fc, combs, n_combs);
    return cudaGetLastError();
}
