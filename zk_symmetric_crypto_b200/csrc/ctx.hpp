// Backend context shared by the C ABI (cb_*) and the proof drivers (s2c_*).
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>
#include <stdexcept>
#include "common.cuh"
#include "host_util.hpp"
#include "comm.hpp"

struct CbError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define CB_CUDA(x)                                                                                          \
    do {                                                                                                    \
        cudaError_t e_ = (x);                                                                               \
        if (e_ != cudaSuccess)                                                                              \
            throw CbError(std::string(#x) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

struct StageTime {
    std::string name;
    float ms;
};

struct cb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    // producer stream of the streaming provers: LDE tiles of group g+1 are transformed here while `stream` consumes group g
    cudaStream_t stream2 = nullptr;
    cudaStream_t stream3 = nullptr;  // row-sharded mode: the last transform pass (peer stores) and the group barrier
    std::vector<cudaEvent_t> ev_pool;  // timing-disabled events for the producer/consumer hand-off
    cudaEvent_t event(size_t i);
    bool overlap = true;
    // row-sharded single-proof mode: NCCL communicator over the ranks that prove ONE trace together (world 1 = off)
    Comm comm;
    std::string err;
    // twiddles (device) for canonic domains up to tw.max_log
    FftTables tw{nullptr, nullptr, nullptr, nullptr, 0};
    uint32_t* tw_dev = nullptr;
    FftTables tw_shift{nullptr, nullptr, nullptr, nullptr, 0};  // host::make_twiddles(.., shifted = true)
    uint32_t* tw_shift_dev = nullptr;
    // per-stage device timing of the last proof (CUDA events on `stream`)
    bool profile = false;
    std::vector<StageTime> stages;
    std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> pending_events;
    // host wall-clock marks of the last profiled proof (stage name, seconds): where the host side of a small proof spends its time
    std::vector<std::pair<std::string, double>> host_marks;
    void host_mark(const char* name);
    int max_cached_tiles = -1;  // streaming prover: cap on LDE tiles kept between passes (-1 = as many as memory allows)
    uint64_t launches = 0;
    // counters of the last streaming proof: packed words transformed (x32 columns), tiles cached between passes, transient slots
    uint64_t fft_words = 0, fft_words_half = 0;  // packed words transformed by the last streaming proof (of which: on half of the domain)
    int cached_tiles = 0, transient_tiles = 0;  // kernels launched by this context (reported by bench.py as gpu_launches)

    // persistent work arena of the streaming provers (LDE tile slots): allocated once with cudaMalloc and kept between
    // proofs -- re-allocating ~170 GB from the stream-ordered pool per proof costs up to 0.3 s when the pool has fragmented
    void* arena = nullptr;
    size_t arena_bytes = 0;
    // pinned staging for reading device-resident inputs back before they are hashed on the host
    uint8_t* hash_stage = nullptr;
    size_t hash_stage_bytes = 0;
    uint64_t hash_wait_us = 0;  // host time the last proof waited for the public-input hashes after the commitment pass
    // Peer windows of the row-sharded mode: every rank's arena mapped into this process with CUDA IPC, so that the last
    // transform pass stores row shards straight into their owner's tile slots over NVLink (no staging, no all-to-all).
    std::vector<uint32_t*> peer_arena;  // [world]; own entry = arena.  Empty: not mapped
    int p2p_state = 0;                  // 0 not tried, 1 mapped, -1 unavailable on this box (the NCCL all-to-all path is used)
    bool last_p2p = false;              // the last sharded proof used the peer windows
    void close_peers();
    // collective over ctx->comm: (re)allocates the arena when `realloc` says so on this rank and (re)maps all arenas when any
    // rank re-allocated.  Returns false when peer mapping is unavailable (every rank gets the same answer).
    bool sync_peer_arenas(bool realloc, size_t bytes);
    void* small_jobs = nullptr;         // product-size ChaCha proofs: constraint job list of the whole AIR (device), valid for
    void* small_jobs_arena = nullptr;   // this arena base and tile pitch
    size_t small_jobs_tile_words = 0;
    int small_jobs_variant = -1;
    void* chacha_cidx[2] = {nullptr, nullptr};  // ChaCha alpha-table index list (consumption order), per AIR variant
    uint32_t* pin_buf = nullptr;        // pinned host staging for read-backs (pinned_words)
    size_t pin_words = 0;
    uint32_t* pinned_words(size_t words);
    void* chacha_consts[2] = {nullptr, nullptr};  // ChaCha AIR constraint table + adder-sum list, per AIR variant (chacha_dev)
    void* ensure_arena(size_t bytes);
    void release_arena();
    void ensure_twiddles(int max_log);
    void ensure_twiddles_shifted(int max_log);
    void* dmalloc(size_t bytes);
    void dfree(void* p);
    void sync() { CB_CUDA(cudaStreamSynchronize(stream)); }
    void stage_begin(const char* name);
    void stage_end();
    void collect_stages();
    static void hook_fn(void* user, const char* name, int begin) {
        cb_ctx* c = (cb_ctx*)user;
        if (begin) c->stage_begin(name); else c->stage_end();
    }
    StageHook hook() { return StageHook{&cb_ctx::hook_fn, this}; }
};

// RAII device buffer on the context's stream-ordered pool
template <typename T>
struct DBuf {
    cb_ctx* ctx = nullptr;
    T* p = nullptr;
    size_t n = 0;
    DBuf() {}
    DBuf(cb_ctx* c, size_t count) : ctx(c), n(count) { p = count ? (T*)c->dmalloc(count * sizeof(T)) : nullptr; }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : ctx(o.ctx), p(o.p), n(o.n) { o.p = nullptr; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) {
            release();
            ctx = o.ctx; p = o.p; n = o.n; o.p = nullptr;
        }
        return *this;
    }
    void release() {
        if (p) ctx->dfree(p);
        p = nullptr;
    }
    ~DBuf() { release(); }
};

// Device Merkle tree: all layers concatenated, layer 0 = leaves (2^log_leaves hashes of 8 words)
struct DevMerkle {
    DBuf<uint32_t> nodes;
    int log_leaves = 0;
    host::Hash32 root;
    size_t layer_offset(int layer) const {  // in hashes
        size_t off = 0;
        for (int l = 0; l < layer; l++) off += (size_t)1 << (log_leaves - l);
        return off;
    }
};

// kernel launchers (kernels_*.cu)
cudaError_t launch_merkle_leaves(cudaStream_t st, const LeafGroups& groups, int lifting_log, uint32_t* h_state, uint64_t bytes_before,
                                 int is_first, int is_final, uint32_t* out);
cudaError_t launch_merkle_nodes(cudaStream_t st, const uint32_t* prev, uint32_t n_parents, uint32_t* out);
cudaError_t launch_merkle_tree_small(cudaStream_t st, uint32_t* nodes, int log_leaves);
cudaError_t launch_chacha_witness(cudaStream_t st, const uint32_t key[8], const uint32_t nonce[3], uint32_t counter,
                                  uint32_t num_blocks, uint32_t n_active_rows, const uint32_t* pt, const uint32_t* ct, int log_size,
                                  uint32_t* W, size_t stride, int* invalid);
cudaError_t launch_chacha_constraints(cudaStream_t st, const uint32_t* lde, size_t stride, int eval_log, int trace_log,
                                      const uint32_t* apr, const uint32_t* den_inv, uint32_t* out, size_t out_stride, int accumulate);
void chacha_init_attrs();
cudaError_t launch_basis(cudaStream_t st, uint32_t* basis, size_t stride, int log_n, const m31::QM31* maps);
cudaError_t launch_oods_dot(cudaStream_t st, const uint32_t* coeffs, size_t stride, int n_cols, int log_n, const uint32_t* basis,
                            size_t b_stride, uint32_t* out);
cudaError_t launch_quotients(cudaStream_t st, const uint32_t* cols, size_t stride, int n_main, const uint32_t* extra,
                             size_t extra_stride, const void* batches_dev, int n_batches, int m, const FftTables& tw, uint32_t* out,
                             size_t out_stride);
cudaError_t launch_fold_circle(cudaStream_t st, const uint32_t* src, size_t s_stride, int m, m31::QM31 alpha, const FftTables& tw,
                               uint32_t* dst, size_t d_stride, int dst_is_zero);
cudaError_t launch_fold_line(cudaStream_t st, const uint32_t* src, size_t s_stride, int L, m31::QM31 alpha, const FftTables& tw,
                             uint32_t* dst, size_t d_stride);
cudaError_t launch_grind(cudaStream_t st, const uint32_t* prefixed_digest_dev, uint32_t pow_bits, uint64_t base, uint64_t count,
                         unsigned long long* best_dev);
cudaError_t launch_gather_rows(cudaStream_t st, const uint32_t* cols, size_t stride, int n_cols, const uint32_t* rows_dev, int n_rows,
                               uint32_t* out);
cudaError_t launch_gather_hashes(cudaStream_t st, const uint32_t* hashes, const uint32_t* idx_dev, int n, uint32_t* out);
cudaError_t launch_secure_powers_rev(cudaStream_t st, m31::QM31 alpha, int K, uint32_t* apr);


// shared prover building blocks (prover_common.cu)
DevMerkle build_merkle(cb_ctx* ctx, const LeafGroups& groups, int lifting_log, const char* leaf_stage = nullptr);
std::vector<host::Hash32> merkle_decommit(cb_ctx* ctx, const DevMerkle& t, const std::vector<uint32_t>& positions);
