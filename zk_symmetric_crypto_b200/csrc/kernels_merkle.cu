// Blake2s lifted-Merkle kernels for sm_100a: one thread per leaf, integer pipe only.
//
// Replaces upstream stwo `MerkleProverLifted::commit` / `MerkleOpsLifted::{build_leaves, build_next_layer}` for
// SimdBackend (prover/vcs_lifted/prover.rs, backend/simd/blake2s_lifted.rs), reached from the reference through
// `TreeBuilder::commit` (/root/reference/stwo/src/chacha/bitwise/air_stream.rs:197,212).
//   leaf i  = Blake2s( LE u32 of col_0[lift_0(i)] || col_1[lift_1(i)] || ... )      (all columns of the tree, in order)
//   lift(i) for a column of log size k < L:  ((i >> (L-k+1)) << 1) | (i & 1)
//   node    = Blake2s( left || right )
// Columns arrive as groups of equally sized, equally strided columns; leaf state (h[8]) can be carried across calls
// so that a tree can be absorbed tile by tile while the LDE is streamed (column count per non-final call must be a
// multiple of 16 = one 64-byte Blake2s block).
#include <stdlib.h>
#include "common.cuh"
#include "blake2s.cuh"
#include "m31_dev.cuh"


namespace merk {

// state: h_state[w * n_leaves + leaf] (SoA) when carrying; bytes_before = bytes absorbed by earlier calls.
template <bool FMA>
__global__ void __launch_bounds__(256) leaves_kernel(LeafGroups groups, int lifting_log, uint32_t* __restrict__ h_state,
                                                     uint64_t bytes_before, int is_first, int is_final,
                                                     uint32_t* __restrict__ out, uint32_t one) {
    const uint32_t n_leaves = 1u << lifting_log;
    const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n_leaves) return;
    uint32_t h[8];
    if (is_first) {
        blake2s::init(h);
    } else {
#pragma unroll
        for (int w = 0; w < 8; w++) h[w] = h_state[(size_t)w * n_leaves + leaf];
    }
    uint64_t t = bytes_before;
    uint32_t m[16];
    int k = 0;  // words buffered in m (only non-zero on the slow path / at the very end)
    int total_cols = 0;
    for (int gi = 0; gi < groups.n; gi++) total_cols += groups.g[gi].ncols;
    int done_cols = 0;
    for (int gi = 0; gi < groups.n; gi++) {
        const LeafGroup g = groups.g[gi];
        uint32_t row = leaf;
        if (g.log_size < lifting_log) {
            int sh = lifting_log - g.log_size;
            row = ((leaf >> (sh + 1)) << 1) | (leaf & 1);
        }
        const uint32_t* p = g.base + row;
        int c = 0;
        if (g.cy != nullptr && k == 0) {
            // computed adder-sum word: 32 columns = two blocks; operands may have been written by this thread earlier in
            // this launch (plain loads, not the read-only path)
            const uint32_t* pb = g.b + row;
            const uint32_t* pc = g.cy + row;
            uint32_t* pr = g.res + row;
            uint32_t cin = 0;
            for (; c + 16 <= g.ncols; c += 16) {
                uint32_t bv[16], cv[16];
#pragma unroll
                for (int w = 0; w < 16; w++) {  // all loads first: the stores below may alias the operand tiles
                    const size_t o = (size_t)(c + w) * g.stride;
                    m[w] = p[o];
                    bv[w] = pb[o];
                    cv[w] = pc[o];
                }
#pragma unroll
                for (int w = 0; w < 16; w++) {
                    const uint32_t v = m31d::subm(m31d::addm(m31d::addm(m[w], bv[w]), cin), m31d::dbl(cv[w]));
                    pr[(size_t)(c + w) * g.stride] = v;
                    m[w] = v;
                    cin = cv[w];
                }
                t += 64;
                bool last = is_final && (done_cols + c + 16 == total_cols);
                if (FMA) blake2s::compress_fma(h, m, t, last, one); else blake2s::compress(h, m, t, last);
            }
        }
        // a group that starts in the middle of a 64-byte block (mixed-size trees: a few small columns hashed first) is
        // brought to the block boundary column by column, then takes the fast path
        for (; c < g.ncols && k != 0; c++) {
            uint32_t v = __ldg(p + (size_t)c * g.stride);
#pragma unroll
            for (int w = 0; w < 16; w++)
                if (w == k) m[w] = v;
            k++;
            if (k == 16) {
                t += 64;
                bool last = is_final && (done_cols + c + 1 == total_cols);
                blake2s::compress(h, m, t, last);
                k = 0;
            }
        }
        if (k == 0) {
            // fast path: whole 64-byte blocks straight from 16 coalesced column loads
            for (; c + 16 <= g.ncols; c += 16) {
#pragma unroll
                for (int w = 0; w < 16; w++) m[w] = __ldg(p + (size_t)(c + w) * g.stride);
                t += 64;
                bool last = is_final && (done_cols + c + 16 == total_cols);
                if (FMA) blake2s::compress_fma(h, m, t, last, one); else blake2s::compress(h, m, t, last);
            }
        }
        for (; c < g.ncols; c++) {
            uint32_t v = __ldg(p + (size_t)c * g.stride);
#pragma unroll
            for (int w = 0; w < 16; w++)
                if (w == k) m[w] = v;
            k++;
            if (k == 16) {
                t += 64;
                bool last = is_final && (done_cols + c + 1 == total_cols);
                blake2s::compress(h, m, t, last);
                k = 0;
            }
        }
        done_cols += g.ncols;
    }
    if (is_final && (k > 0 || total_cols == 0 && is_first)) {
#pragma unroll
        for (int w = 0; w < 16; w++)
            if (w >= k) m[w] = 0;
        t += 4 * k;
        blake2s::compress(h, m, t, true);
    }
    if (is_final) {
#pragma unroll
        for (int w = 0; w < 8; w++) out[(size_t)leaf * 8 + w] = h[w];
    } else {
#pragma unroll
        for (int w = 0; w < 8; w++) h_state[(size_t)w * n_leaves + leaf] = h[w];
    }
}

// Streaming prover form: every group is a whole number of 64-byte blocks (a multiple of 16 columns) of lifting-size columns,
// i.e. no lifting, no partially filled block between groups.  The hot loop has ONE compress call site: the general kernel
// above inlines the 10 unrolled rounds (~18 KB of code each) at five places, and the two that alternate inside a quarter-round
// group (fused adder sums / plain columns) thrash the instruction cache -- ncu showed `no_instruction` as its largest stall
// reason (5.1 stalled warps per issue, profiles/ncu_r02_leaves_kernel.csv).
template <bool FMA>
__global__ void __launch_bounds__(128) leaves_tiles_kernel(LeafGroups groups, int lifting_log, uint32_t* __restrict__ h_state,
                                                           uint64_t bytes_before, int is_first, int is_final,
                                                           uint32_t* __restrict__ out, uint32_t one) {
    const uint32_t n_leaves = 1u << lifting_log;
    const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n_leaves) return;
    uint32_t h[8];
    if (is_first) {
        blake2s::init(h);
    } else {
#pragma unroll
        for (int w = 0; w < 8; w++) h[w] = h_state[(size_t)w * n_leaves + leaf];
    }
    uint64_t t = bytes_before;
    int blocks_left = 0;
    for (int gi = 0; gi < groups.n; gi++) blocks_left += groups.g[gi].ncols >> 4;
    int gi = 0, c = 0;
    uint32_t cin = 0;
    // Addresses are formed as (uniform 64-bit tile base) + 4 * (32-bit word index): one IMAD.WIDE on the FMA pipe per load
    // instead of a 64-bit pointer increment (IADD3 + IADD3.X) on the ALU pipe, which is the pipe Blake2s saturates.
    // A tile has at most 2^26 words (32 columns of 2^21 rows), so the index fits 32 bits.
    const uint32_t *p = nullptr, *pb = nullptr, *pc = nullptr;
    uint32_t* pr = nullptr;
    uint32_t stride = 0;
    int ncols = 0;
#pragma unroll 1
    for (; blocks_left > 0; blocks_left--) {
        if (c == ncols) {  // next group
            const LeafGroup& g = groups.g[gi++];
            p = g.base;
            pb = g.b;
            pc = g.cy;
            pr = g.res;
            stride = (uint32_t)g.stride;
            ncols = g.ncols;
            c = 0;
            cin = 0;
        }
        uint32_t m[16];
        const uint32_t i0 = (uint32_t)c * stride + leaf;
        if (pc != nullptr) {
            // adder-sum word computed on the fly from its operand tiles (8 columns at a time: all loads of a half first);
            // plain loads: the operands may have been written by this thread earlier
#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint32_t av[8], bv[8], cv[8];
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    const uint32_t o = i0 + (uint32_t)(8 * half + w) * stride;
                    av[w] = p[o];
                    bv[w] = pb[o];
                    cv[w] = pc[o];
                }
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    const uint32_t v = m31d::subm(m31d::addm(m31d::addm(av[w], bv[w]), cin), m31d::dbl(cv[w]));
                    pr[i0 + (uint32_t)(8 * half + w) * stride] = v;
                    m[8 * half + w] = v;
                    cin = cv[w];
                }
            }
        } else {
#pragma unroll
            for (int w = 0; w < 16; w++) m[w] = __ldg(p + (i0 + (uint32_t)w * stride));
        }
        c += 16;
        t += 64;
        const bool last = is_final && blocks_left == 1;
        if (FMA) blake2s::compress_fma(h, m, t, last, one); else blake2s::compress(h, m, t, last);
    }
    if (is_final) {
#pragma unroll
        for (int w = 0; w < 8; w++) out[(size_t)leaf * 8 + w] = h[w];
    } else {
#pragma unroll
        for (int w = 0; w < 8; w++) h_state[(size_t)w * n_leaves + leaf] = h[w];
    }
}

// Product-size traces with the whole LDE materialised: one launch absorbs all words of the tree, tile(w) = arena + w * tile_words,
// [32][n_leaves].  A leaf's 2 n_words compressions are one dependent chain on a nearly empty GPU, so the next block's 16 column
// loads are issued before the current compression.  The additions stay IADD3 here: the FMA-pipe form (two IMAD per 3-input add)
// lengthens the chain and measured slower for a lone warp (2.6 vs 2.2 ms for the 2,080 compressions of a ChaCha leaf).
__global__ void __launch_bounds__(64) leaves_seq_kernel(const uint32_t* __restrict__ arena, size_t tile_words, uint32_t n_leaves, int n_words,
                                                        uint32_t* __restrict__ out, uint32_t one, int variant) {
    const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n_leaves) return;
    uint32_t h[8];
    blake2s::init(h);
    const uint32_t* __restrict__ p = arena + leaf;
    uint32_t nxt[16], m[16];
#pragma unroll
    for (int w = 0; w < 16; w++) nxt[w] = __ldg(p + (size_t)w * n_leaves);
    const int n_blocks = 2 * n_words;
    uint64_t t = 0;
#pragma unroll 1
    for (int b = 0; b < n_blocks; b++) {
#pragma unroll
        for (int w = 0; w < 16; w++) m[w] = nxt[w];
        if (b + 1 < n_blocks) {
            const uint32_t* __restrict__ q = p + (size_t)((b + 1) >> 1) * tile_words + (size_t)(((b + 1) & 1) * 16) * n_leaves;
#pragma unroll
            for (int w = 0; w < 16; w++) nxt[w] = __ldg(q + (size_t)w * n_leaves);
        }
        t += 64;
        if (variant) blake2s::compress_mix(h, m, t, b + 1 == n_blocks, one);
        else blake2s::compress(h, m, t, b + 1 == n_blocks);
    }
#pragma unroll
    for (int w = 0; w < 8; w++) out[(size_t)leaf * 8 + w] = h[w];
}

// parent[i] = Blake2s(child[2i] || child[2i+1]); hashes are 8 consecutive u32
__global__ void __launch_bounds__(256) nodes_kernel(const uint4* __restrict__ prev, uint32_t n_parents, uint4* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_parents) return;
    uint4 a = prev[4 * (size_t)i], b = prev[4 * (size_t)i + 1], c = prev[4 * (size_t)i + 2], d = prev[4 * (size_t)i + 3];
    uint32_t m[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
    uint32_t h[8];
    blake2s::init(h);
    blake2s::compress(h, m, 64, true);
    out[2 * (size_t)i] = make_uint4(h[0], h[1], h[2], h[3]);
    out[2 * (size_t)i + 1] = make_uint4(h[4], h[5], h[6], h[7]);
}

// all upper layers of a small tree in one launch: nodes = all layers concatenated, layer 0 (2^L hashes) already filled
__global__ void __launch_bounds__(256) nodes_all_kernel(uint32_t* nodes, int L) {
    size_t off = 0;
    for (int l = 0; l < L; l++) {
        const uint32_t n_parents = 1u << (L - l - 1);
        const uint4* __restrict__ prev = (const uint4*)(nodes + off * 8);
        uint4* __restrict__ out = (uint4*)(nodes + (off + ((size_t)1 << (L - l))) * 8);
        for (uint32_t i = threadIdx.x; i < n_parents; i += blockDim.x) {
            uint4 a = prev[4 * (size_t)i], b = prev[4 * (size_t)i + 1], c = prev[4 * (size_t)i + 2], d = prev[4 * (size_t)i + 3];
            uint32_t m[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
            uint32_t h[8];
            blake2s::init(h);
            blake2s::compress(h, m, 64, true);
            out[2 * (size_t)i] = make_uint4(h[0], h[1], h[2], h[3]);
            out[2 * (size_t)i + 1] = make_uint4(h[4], h[5], h[6], h[7]);
        }
        __syncthreads();
        off += (size_t)1 << (L - l);
    }
}

}  // namespace merk

// upper layers of a tree with 2^log_leaves <= 2^11 leaves in one launch (product-size proofs)
cudaError_t launch_merkle_tree_small(cudaStream_t st, uint32_t* nodes, int log_leaves) {
    if (log_leaves <= 0) return cudaSuccess;
    merk::nodes_all_kernel<<<1, 256, 0, st>>>(nodes, log_leaves);
    return cudaGetLastError();
}

cudaError_t launch_merkle_leaves(cudaStream_t st, const LeafGroups& groups, int lifting_log, uint32_t* h_state,
                                 uint64_t bytes_before, int is_first, int is_final, uint32_t* out) {
    uint32_t n = 1u << lifting_log;
    int threads = n >= 128 * 148 * 2 ? 128 : 64;
    if (n < 64) threads = 32;
    static const int variant = getenv("S2C_BLAKE_FMA") ? atoi(getenv("S2C_BLAKE_FMA")) : 1;
    static const bool general_only = getenv("S2C_LEAVES_GENERAL") != nullptr;  // A/B switch
    bool tiles = !general_only && groups.n > 0 && variant;
    for (int gi = 0; gi < groups.n; gi++)
        if (groups.g[gi].ncols % 16 != 0 || groups.g[gi].log_size != lifting_log) tiles = false;
    if (tiles) {
        merk::leaves_tiles_kernel<true><<<(n + threads - 1) / threads, threads, 0, st>>>(groups, lifting_log, h_state, bytes_before,
                                                                                         is_first, is_final, out, 1u);
        return cudaGetLastError();
    }
    if (variant)
        merk::leaves_kernel<true><<<(n + threads - 1) / threads, threads, 0, st>>>(groups, lifting_log, h_state, bytes_before,
                                                                                   is_first, is_final, out, 1u);
    else
        merk::leaves_kernel<false><<<(n + threads - 1) / threads, threads, 0, st>>>(groups, lifting_log, h_state, bytes_before,
                                                                                    is_first, is_final, out, 1u);
    return cudaGetLastError();
}

cudaError_t launch_merkle_leaves_seq(cudaStream_t st, const uint32_t* arena, size_t tile_words, int n_words, int lifting_log, uint32_t* out) {
    const uint32_t n = 1u << lifting_log;
    const int threads = n >= 64 * 148 ? 64 : 32;
    static const int variant = getenv("S2C_LEAVES_SEQ_MIX") ? atoi(getenv("S2C_LEAVES_SEQ_MIX")) : 0;  // A/B switch (1: c += d on the FMA pipe)
    merk::leaves_seq_kernel<<<(n + threads - 1) / threads, threads, 0, st>>>(arena, tile_words, n, n_words, out, 1u, variant);
    return cudaGetLastError();
}

cudaError_t launch_merkle_nodes(cudaStream_t st, const uint32_t* prev, uint32_t n_parents, uint32_t* out) {
    int threads = 128;
    merk::nodes_kernel<<<(n_parents + threads - 1) / threads, threads, 0, st>>>((const uint4*)prev, n_parents, (uint4*)out);
    return cudaGetLastError();
}
