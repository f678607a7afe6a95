// ChaCha20 stream AIR on the GPU: witness generation and constraint-quotient evaluation.
//
// Witness: follows /root/reference/stwo/src/chacha/bitwise/gen_stream.rs:83-221 (append_u32_bits, build_state, generate,
// quarter_round, add_u32, xor_rotl_u32) and :226-261 (zero default rows), with the product API's lane packing of
// /root/reference/stwo/src/wasm_api.rs:511-575 (counter+row, padding lanes pt=0 / ct=keystream).  Instead of 33,280
// one-bit M31 columns the kernel stores the 1,040 32-bit words the bits come from ("packed witness", 32x smaller);
// the FFT kernels expand bit columns on load (ColSrc kind SRC_BITS).  Column j = bit (j & 31) of word (j >> 5):
//   words 0..15 initial state | 80 QRs x [sum,carry,xor]x4 | 16 final adds x [sum,carry] | 16 plaintext | 16 ciphertext
//
// Constraints: follows constraints_stream.rs:20-70,74-82,85-101,104-131,134-152,179-189 driven by upstream
// `FrameworkComponent::evaluate_constraint_quotients_on_domain` (constraint-framework prover/component_prover.rs):
//   acc[row] += (sum_k alpha^(K-1-k) C_k(row)) * 1/Z_H(row).
// One thread block = 32 rows x 8 bit-groups; the 16 live state words (512 LDE values per row) stay in shared memory
// so every LDE column is read from HBM exactly once (plus the 512 initial-state columns re-read for the final adds).
#include "common.cuh"

namespace chacha {
using namespace m31;

constexpr int N_WORDS = 1040;
constexpr int N_COLS = 33280;
constexpr int N_CONSTRAINTS = 54784;

__device__ __forceinline__ uint32_t rotl(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

struct WitnessArgs {
    uint32_t key[8];
    uint32_t nonce[3];
    uint32_t counter;
    uint32_t num_blocks;     // rows with caller-supplied plaintext/ciphertext
    uint32_t n_active_rows;  // rows covered by provided vec-rows (multiple of 16); rows beyond are all-zero inputs
};

// thread per row.  pt/ct: [num_blocks][16] little-endian words (the caller's byte buffers).
__global__ void __launch_bounds__(128) witness_kernel(WitnessArgs a, const uint32_t* __restrict__ pt, const uint32_t* __restrict__ ct,
                                                      int log_size, uint32_t* __restrict__ W, size_t stride, int* __restrict__ invalid) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= (1u << log_size)) return;
    const bool active = row < a.n_active_rows;
    const bool real = row < a.num_blocks;
    uint32_t init[16];
    init[0] = 0x61707865u; init[1] = 0x3320646eu; init[2] = 0x79622d32u; init[3] = 0x6b206574u;
#pragma unroll
    for (int i = 0; i < 8; i++) init[4 + i] = active ? a.key[i] : 0u;
    init[12] = active ? a.counter + row : 0u;
#pragma unroll
    for (int i = 0; i < 3; i++) init[13 + i] = active ? a.nonce[i] : 0u;
    uint32_t* w = W + row;
    int wi = 0;
#define PUT(x) { w[(size_t)(wi++) * stride] = (x); }
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { v[i] = init[i]; PUT(init[i]); }
#define ADD(x, y) { uint32_t s_ = x + y; uint32_t c_ = (x & y) | ((x | y) & ~s_); PUT(s_); PUT(c_); x = s_; }
#define XR(x, y, r) { y = rotl(x ^ y, r); PUT(y); }
#define QR(A, B, C, D) ADD(v[A], v[B]) XR(v[A], v[D], 16) ADD(v[C], v[D]) XR(v[C], v[B], 12) \
                       ADD(v[A], v[B]) XR(v[A], v[D], 8)  ADD(v[C], v[D]) XR(v[C], v[B], 7)
    for (int r = 0; r < 10; r++) {
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#pragma unroll
    for (int i = 0; i < 16; i++) ADD(v[i], init[i])
    uint32_t p[16], c[16];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        p[i] = real ? pt[(size_t)row * 16 + i] : 0u;
        c[i] = real ? ct[(size_t)row * 16 + i] : (active ? v[i] : 0u);
        if ((v[i] ^ p[i]) != c[i]) ok = false;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) PUT(p[i])
#pragma unroll
    for (int i = 0; i < 16; i++) PUT(c[i])
    if (active && !ok) atomicOr(invalid, 1);
#undef PUT
#undef ADD
#undef XR
#undef QR
}

// ------------------------------------------------------------------------------------------------ constraints
struct Acc {
    uint64_t a[4];
    int pending;
    __device__ __forceinline__ void init() { a[0] = a[1] = a[2] = a[3] = 0; pending = 0; }
    __device__ __forceinline__ void fold() {
#pragma unroll
        for (int c = 0; c < 4; c++) a[c] = (a[c] & P) + (a[c] >> 31);
        pending = 0;
    }
    // acc += C * alpha^(K-1-k); C in [0,p)
    __device__ __forceinline__ void mac(const uint4* __restrict__ apr, int k, uint32_t C) {
        uint4 al = __ldg(apr + k);
        a[0] += (uint64_t)C * al.x; a[1] += (uint64_t)C * al.y; a[2] += (uint64_t)C * al.z; a[3] += (uint64_t)C * al.w;
        if (++pending == 4) fold();  // 4*(p-1)^2 + fold residual < 2^64
    }
};

__device__ __forceinline__ uint32_t boolc(uint32_t b) { return mul(b, sub(1, b)); }

constexpr int ROWS = 32;      // rows per block
constexpr int GROUPS = 8;     // bit groups per row (4 bits each)

// lde: column-major LDE values, column j at lde + j*stride.  apr[k] = alpha^(K-1-k) as 4 coordinates.
// den_inv[row >> trace_log] = 1/Z_H on that chunk of the evaluation domain.  out: 4 coordinate columns (out + c*out_stride).
__global__ void __launch_bounds__(ROWS * GROUPS) constraints_kernel(const uint32_t* __restrict__ lde, size_t stride, int eval_log,
                                                                     int trace_log, const uint4* __restrict__ apr,
                                                                     const uint32_t* __restrict__ den_inv,
                                                                     uint32_t* __restrict__ out, size_t out_stride, int accumulate) {
    extern __shared__ uint32_t st[];  // [16 words][32 bits][ROWS]
    const int r = threadIdx.x & (ROWS - 1);
    const int g = threadIdx.x / ROWS;
    const uint32_t row = blockIdx.x * ROWS + r;
    const uint32_t* col = lde + row;
#define L(c) __ldg(col + (size_t)(c) * stride)
#define ST(w, i) st[(((w) << 5) + (i)) * ROWS + r]
    Acc acc;
    acc.init();
    // initial state: 16 x next_u32
    for (int w = 0; w < 16; w++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int i = 4 * g + j;
            uint32_t b = L(w * 32 + i);
            acc.mac(apr, w * 32 + i, boolc(b));
            ST(w, i) = b;
        }
    }
    __syncthreads();

    auto add_op = [&](int a, int c0, int k0, auto getb) {
        // sum cols c0..c0+31, carry cols c0+32..c0+63; constraints: 32 bool(sum), then per bit [bool(carry), adder]
        uint32_t cin = (g == 0) ? 0u : L(c0 + 32 + 4 * g - 1);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int i = 4 * g + j;
            uint32_t s = L(c0 + i), cy = L(c0 + 32 + i);
            uint32_t av = ST(a, i), bv = getb(i);
            acc.mac(apr, k0 + i, boolc(s));
            acc.mac(apr, k0 + 32 + 2 * i, boolc(cy));
            // s + 2cy - a - b - cin
            uint32_t e = sub(sub(sub(add(s, add(cy, cy)), av), bv), cin);
            acc.mac(apr, k0 + 32 + 2 * i + 1, e);
            ST(a, i) = s;
            cin = cy;
        }
    };
    auto xor_op = [&](int a, int d, int rot, int c0, int k0) {
        __syncthreads();  // writes of v[a] by the preceding add are visible
        uint32_t res[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int i = 4 * g + j;
            int src = (i + 32 - rot) & 31;
            uint32_t rv = L(c0 + i);
            uint32_t av = ST(a, src), dv = ST(d, src);
            acc.mac(apr, k0 + i, boolc(rv));
            uint32_t ab = mul(av, dv);
            uint32_t e = add(sub(sub(rv, av), dv), add(ab, ab));
            acc.mac(apr, k0 + 32 + i, e);
            res[j] = rv;
        }
        __syncthreads();  // everyone has read the old v[d]
#pragma unroll
        for (int j = 0; j < 4; j++) ST(d, 4 * g + j) = res[j];
    };

    int c0 = 512, k0 = 512;
    for (int rnd = 0; rnd < 10; rnd++) {
#pragma unroll 1
        for (int q = 0; q < 8; q++) {
            int a = q & 3, b, c, d;
            if (q < 4) { b = 4 + a; c = 8 + a; d = 12 + a; }
            else { b = 4 + ((a + 1) & 3); c = 8 + ((a + 2) & 3); d = 12 + ((a + 3) & 3); }
            auto gb = [&](int i) { return ST(b, i); };
            auto gd = [&](int i) { return ST(d, i); };
            add_op(a, c0, k0, gb);            xor_op(a, d, 16, c0 + 64, k0 + 96);
            add_op(c, c0 + 96, k0 + 160, gd);  xor_op(c, b, 12, c0 + 160, k0 + 256);
            add_op(a, c0 + 192, k0 + 320, gb); xor_op(a, d, 8, c0 + 256, k0 + 416);
            add_op(c, c0 + 288, k0 + 480, gd); xor_op(c, b, 7, c0 + 352, k0 + 576);
            c0 += 384;
            k0 += 640;
        }
    }
    // final adds: keystream[i] = v[i] + initial[i]   (initial re-read from its LDE columns 0..511)
    for (int w = 0; w < 16; w++) {
        auto gi = [&](int i) { return L(w * 32 + i); };
        add_op(w, c0, k0, gi);
        c0 += 64;
        k0 += 96;
    }
    // plaintext / ciphertext booleans, then (ks xor pt) - ct
    const int c_ks = 512 + 80 * 384, c_pt = c_ks + 1024, c_ct = c_pt + 512;
    const int k_pt = k0, k_ct = k0 + 512, k_eq = k0 + 1024;
    for (int w = 0; w < 16; w++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int i = 4 * g + j;
            uint32_t p = L(c_pt + w * 32 + i), c = L(c_ct + w * 32 + i), ks = L(c_ks + w * 64 + i);
            acc.mac(apr, k_pt + w * 32 + i, boolc(p));
            acc.mac(apr, k_ct + w * 32 + i, boolc(c));
            uint32_t kp = mul(ks, p);
            uint32_t e = sub(sub(add(ks, p), add(kp, kp)), c);
            acc.mac(apr, k_eq + w * 32 + i, e);
        }
    }
    // reduce the 8 bit-groups of each row, multiply by 1/Z_H, store
    acc.fold();
    __syncthreads();
    uint32_t* red = st;  // [GROUPS][4][ROWS]
#pragma unroll
    for (int c = 0; c < 4; c++) red[(g * 4 + c) * ROWS + r] = reduce64_full(acc.a[c]);
    __syncthreads();
    if (g < 4) {
        const int c = g;
        uint64_t s = 0;
#pragma unroll
        for (int gg = 0; gg < GROUPS; gg++) s += red[(gg * 4 + c) * ROWS + r];
        uint32_t v = mul(reduce64_full(s), den_inv[row >> trace_log]);
        uint32_t* o = out + (size_t)c * out_stride + row;
        if (accumulate) v = add(v, *o);
        *o = v;
    }
#undef L
#undef ST
}

}  // namespace chacha

cudaError_t launch_chacha_witness(cudaStream_t st, const uint32_t key[8], const uint32_t nonce[3], uint32_t counter,
                                  uint32_t num_blocks, uint32_t n_active_rows, const uint32_t* pt, const uint32_t* ct,
                                  int log_size, uint32_t* W, size_t stride, int* invalid) {
    chacha::WitnessArgs a;
    for (int i = 0; i < 8; i++) a.key[i] = key[i];
    for (int i = 0; i < 3; i++) a.nonce[i] = nonce[i];
    a.counter = counter;
    a.num_blocks = num_blocks;
    a.n_active_rows = n_active_rows;
    uint32_t n = 1u << log_size;
    int threads = n < 128 ? 32 : 128;
    chacha::witness_kernel<<<(n + threads - 1) / threads, threads, 0, st>>>(a, pt, ct, log_size, W, stride, invalid);
    return cudaGetLastError();
}

void chacha_init_attrs() {
    cudaFuncSetAttribute(chacha::constraints_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}

cudaError_t launch_chacha_constraints(cudaStream_t st, const uint32_t* lde, size_t stride, int eval_log, int trace_log,
                                      const uint32_t* apr, const uint32_t* den_inv, uint32_t* out, size_t out_stride,
                                      int accumulate) {
    uint32_t rows = 1u << eval_log;
    chacha::constraints_kernel<<<rows / chacha::ROWS, chacha::ROWS * chacha::GROUPS, 64 * 1024, st>>>(
        lde, stride, eval_log, trace_log, (const uint4*)apr, den_inv, out, out_stride, accumulate);
    return cudaGetLastError();
}
