// Context plumbing and prover building blocks shared by the AIR-specific drivers.
#include <chrono>
#include "ctx.hpp"

void cb_ctx::ensure_twiddles(int max_log) {
    if (tw.max_log >= max_log) return;
    if (max_log < 6) max_log = 6;
    host::TwiddleTables t = host::make_twiddles(max_log);
    if (tw_dev) { sync(); CB_CUDA(cudaFree(tw_dev)); tw_dev = nullptr; }
    size_t nx = t.X.size(), ny = t.Y.size();
    CB_CUDA(cudaMalloc(&tw_dev, (2 * nx + 2 * ny) * sizeof(uint32_t)));
    uint32_t* p = tw_dev;
    CB_CUDA(cudaMemcpyAsync(p, t.X.data(), nx * 4, cudaMemcpyHostToDevice, stream)); tw.X = p; p += nx;
    CB_CUDA(cudaMemcpyAsync(p, t.Y.data(), ny * 4, cudaMemcpyHostToDevice, stream)); tw.Y = p; p += ny;
    CB_CUDA(cudaMemcpyAsync(p, t.IX.data(), nx * 4, cudaMemcpyHostToDevice, stream)); tw.IX = p; p += nx;
    CB_CUDA(cudaMemcpyAsync(p, t.IY.data(), ny * 4, cudaMemcpyHostToDevice, stream)); tw.IY = p;
    tw.max_log = max_log;
    sync();
}

void cb_ctx::ensure_twiddles_shifted(int max_log) {
    if (tw_shift.max_log >= max_log) return;
    if (max_log < 6) max_log = 6;
    host::TwiddleTables t = host::make_twiddles(max_log, true);
    if (tw_shift_dev) { sync(); CB_CUDA(cudaFree(tw_shift_dev)); tw_shift_dev = nullptr; }
    size_t nx = t.X.size(), ny = t.Y.size();
    CB_CUDA(cudaMalloc(&tw_shift_dev, (2 * nx + 2 * ny) * sizeof(uint32_t)));
    uint32_t* p = tw_shift_dev;
    CB_CUDA(cudaMemcpyAsync(p, t.X.data(), nx * 4, cudaMemcpyHostToDevice, stream)); tw_shift.X = p; p += nx;
    CB_CUDA(cudaMemcpyAsync(p, t.Y.data(), ny * 4, cudaMemcpyHostToDevice, stream)); tw_shift.Y = p; p += ny;
    CB_CUDA(cudaMemcpyAsync(p, t.IX.data(), nx * 4, cudaMemcpyHostToDevice, stream)); tw_shift.IX = p; p += nx;
    CB_CUDA(cudaMemcpyAsync(p, t.IY.data(), ny * 4, cudaMemcpyHostToDevice, stream)); tw_shift.IY = p;
    tw_shift.max_log = max_log;
    sync();
}

cudaEvent_t cb_ctx::event(size_t i) {
    while (ev_pool.size() <= i) {
        cudaEvent_t e;
        CB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ev_pool.push_back(e);
    }
    return ev_pool[i];
}

void* cb_ctx::ensure_arena(size_t bytes) {
    if (arena && arena_bytes >= bytes) return arena;
    release_arena();
    // give cached pool memory back to the device before the big allocation
    cudaMemPool_t pool;
    CB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    sync();
    CB_CUDA(cudaMemPoolTrimTo(pool, 0));
    CB_CUDA(cudaMalloc(&arena, bytes));
    arena_bytes = bytes;
    return arena;
}
uint32_t* cb_ctx::pinned_words(size_t words) {
    if (pin_words < words) {
        if (pin_buf) cudaFreeHost(pin_buf);
        pin_buf = nullptr;
        pin_words = 0;
        CB_CUDA(cudaHostAlloc((void**)&pin_buf, words * 4, cudaHostAllocDefault));
        pin_words = words;
    }
    return pin_buf;
}

void cb_ctx::close_peers() {
    for (size_t r = 0; r < peer_arena.size(); r++)
        if ((int)r != comm.rank && peer_arena[r]) cudaIpcCloseMemHandle(peer_arena[r]);
    peer_arena.clear();
    if (p2p_state == 1) p2p_state = 0;
}

bool cb_ctx::sync_peer_arenas(bool realloc, size_t bytes) {
    static const bool disabled = getenv("S2C_NO_P2P") != nullptr;
    const int G = comm.world, R = comm.rank;
    const bool want = !disabled && p2p_state >= 0 && G <= MAX_PEERS;
    // does any rank need a new arena or a first mapping?
    const bool mine = realloc || (want && peer_arena.empty());
    const bool any = comm_min_int(comm, mine ? 0 : 1, stream) == 0;
    if (!any) return p2p_state == 1;
    // nobody may keep a mapping of memory that is about to be freed
    close_peers();
    comm_barrier(comm, stream);
    sync();
    if (realloc) {
        release_arena();
        ensure_arena(bytes);
    }
    if (!want) return false;
    // exchange the IPC handles (64 bytes each) through the communicator
    cudaIpcMemHandle_t mine_h;
    int ok = cudaIpcGetMemHandle(&mine_h, arena) == cudaSuccess ? 1 : 0;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    DBuf<uint32_t> d_send(this, 16), d_recv(this, (size_t)16 * G);
    CB_CUDA(cudaMemcpyAsync(d_send.p, &mine_h, 64, cudaMemcpyHostToDevice, stream));
    comm_allgather_u32(comm, d_send.p, d_recv.p, 16, stream);
    std::vector<cudaIpcMemHandle_t> all(G);
    CB_CUDA(cudaMemcpyAsync(all.data(), d_recv.p, (size_t)64 * G, cudaMemcpyDeviceToHost, stream));
    sync();
    peer_arena.assign(G, nullptr);
    peer_arena[R] = (uint32_t*)arena;
    for (int r = 0; r < G && ok; r++) {
        if (r == R) continue;
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        } else {
            peer_arena[r] = (uint32_t*)p;
        }
    }
    ok = comm_min_int(comm, ok, stream);
    if (!ok) {
        close_peers();
        p2p_state = -1;
        return false;
    }
    p2p_state = 1;
    return true;
}

void cb_ctx::release_arena() {
    if (arena) {
        sync();
        cudaFree(arena);
    }
    arena = nullptr;
    arena_bytes = 0;
}

void* cb_ctx::dmalloc(size_t bytes) {
    void* p = nullptr;
    CB_CUDA(cudaMallocAsync(&p, bytes, stream));
    return p;
}
void cb_ctx::dfree(void* p) {
    if (p) cudaFreeAsync(p, stream);
}

void cb_ctx::host_mark(const char* name) {
    if (!profile) return;
    host_marks.push_back({name, std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count()});
}
void cb_ctx::stage_begin(const char* name) {
    if (!profile) return;
    host_mark(name);
    cudaEvent_t a, b;
    CB_CUDA(cudaEventCreate(&a));
    CB_CUDA(cudaEventCreate(&b));
    CB_CUDA(cudaEventRecord(a, stream));
    pending_events.push_back({name, {a, b}});
}
void cb_ctx::stage_end() {
    if (!profile) return;
    CB_CUDA(cudaEventRecord(pending_events.back().second.second, stream));
}
void cb_ctx::collect_stages() {
    if (!profile) return;
    sync();
    host_mark("end");
    stages.clear();
    for (auto& e : pending_events) {
        float ms = 0;
        cudaEventElapsedTime(&ms, e.second.first, e.second.second);
        stages.push_back({e.first, ms});
        cudaEventDestroy(e.second.first);
        cudaEventDestroy(e.second.second);
    }
    pending_events.clear();
}

DevMerkle build_merkle(cb_ctx* ctx, const LeafGroups& groups, int lifting_log, const char* leaf_stage) {
    DevMerkle t;
    t.log_leaves = lifting_log;
    size_t n_hashes = ((size_t)2 << lifting_log) - 1;
    t.nodes = DBuf<uint32_t>(ctx, n_hashes * 8);
    if (leaf_stage) ctx->stage_begin(leaf_stage);
    CB_CUDA(launch_merkle_leaves(ctx->stream, groups, lifting_log, nullptr, 0, 1, 1, t.nodes.p));
    ctx->launches++;
    if (leaf_stage) ctx->stage_end();
    if (lifting_log <= 11) {
        CB_CUDA(launch_merkle_tree_small(ctx->stream, t.nodes.p, lifting_log));
        ctx->launches++;
    } else
    for (int l = 0; l < lifting_log; l++) {
        const uint32_t* prev = t.nodes.p + t.layer_offset(l) * 8;
        uint32_t* out = t.nodes.p + t.layer_offset(l + 1) * 8;
        CB_CUDA(launch_merkle_nodes(ctx->stream, prev, 1u << (lifting_log - l - 1), out));
        ctx->launches++;
    }
    CB_CUDA(cudaMemcpyAsync(t.root.b, t.nodes.p + t.layer_offset(lifting_log) * 8, 32, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    return t;
}

// MerkleProverLifted::decommit (prover/vcs_lifted/prover.rs): bottom-up, for every needed node whose sibling is not itself
// derivable from the queried set, the sibling hash, in position order.  (Order pinned against the reference proof.)
std::vector<host::Hash32> merkle_decommit(cb_ctx* ctx, const DevMerkle& t, const std::vector<uint32_t>& positions) {
    std::vector<uint32_t> idx;  // global hash indices to fetch
    std::vector<uint32_t> cur(positions.begin(), positions.end());
    std::sort(cur.begin(), cur.end());
    cur.erase(std::unique(cur.begin(), cur.end()), cur.end());
    for (int l = 0; l < t.log_leaves; l++) {
        size_t off = t.layer_offset(l);
        std::vector<uint32_t> nxt;
        for (size_t i = 0; i < cur.size(); i++) {
            uint32_t p = cur[i], sib = p ^ 1;
            bool have = (i + 1 < cur.size() && cur[i + 1] == sib) || (i > 0 && cur[i - 1] == sib);
            if (!have) idx.push_back((uint32_t)(off + sib));
            if (nxt.empty() || nxt.back() != (p >> 1)) nxt.push_back(p >> 1);
        }
        cur.swap(nxt);
    }
    std::vector<host::Hash32> out(idx.size());
    if (idx.empty()) return out;
    DBuf<uint32_t> d_idx(ctx, idx.size()), d_out(ctx, idx.size() * 8);
    CB_CUDA(cudaMemcpyAsync(d_idx.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(launch_gather_hashes(ctx->stream, t.nodes.p, d_idx.p, (int)idx.size(), d_out.p));
    ctx->launches++;
    CB_CUDA(cudaMemcpyAsync(out.data(), d_out.p, idx.size() * 32, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    return out;
}

// ------------------------------------------------------------------------------------------------ FRI
#include "prover.hpp"
using namespace m31;

QM31 coset_vanishing_q(int trace_log, const host::CirclePointQ& z) {
    host::Coset c = host::Coset::odds(trace_log);
    host::Pt t = host::index_to_point((0x80000000u - c.initial + (c.step >> 1)) & 0x7fffffffu);
    QM31 x = qsub(qmul_m(z.x, t.x), qmul_m(z.y, t.y));
    for (int i = 1; i < trace_log; i++) x = qsub(qmul_m(qmul(x, x), 2), qone());
    return x;
}

static DevMerkle merkle4(cb_ctx* ctx, const uint32_t* base, size_t stride, int log) {
    LeafGroups g{};
    g.n = 1;
    g.g[0] = {base, stride, 4, log};
    return build_merkle(ctx, g, log);
}

// FriProver::commit (prover/fri.rs): first (circle) layer, inner line layers, last-layer polynomial.
FriProverState fri_commit(cb_ctx* ctx, host::Channel& ch, const PcsConfig& cfg, DBuf<uint32_t>&& quot, int m) {
    FriProverState f;
    cudaStream_t st = ctx->stream;
    const size_t M = (size_t)1 << m;
    f.evals.push_back(std::move(quot));
    f.trees.push_back(merkle4(ctx, f.evals[0].p, M, m));
    f.logs.push_back(m);
    ch.mix_root(f.trees[0].root);
    QM31 alpha = ch.draw_secure_felt();
    int L = m - 1;
    DBuf<uint32_t> line(ctx, (size_t)4 << L);
    CB_CUDA(launch_fold_circle(st, f.evals[0].p, M, m, alpha, ctx->tw, line.p, (size_t)1 << L, 1));
    ctx->launches++;
    const int last_log = (int)(cfg.log_last_layer_degree_bound + cfg.log_blowup);
    while (L > last_log) {
        size_t n = (size_t)1 << L;
        f.evals.push_back(std::move(line));
        f.trees.push_back(merkle4(ctx, f.evals.back().p, n, L));
        f.logs.push_back(L);
        ch.mix_root(f.trees.back().root);
        alpha = ch.draw_secure_felt();
        line = DBuf<uint32_t>(ctx, (size_t)4 << (L - 1));
        CB_CUDA(launch_fold_line(st, f.evals.back().p, n, L, alpha, ctx->tw, line.p, n >> 1));
        ctx->launches++;
        L--;
    }
    // last layer: LineEvaluation::interpolate on the host (prover/line.rs), natural order in, ordered coefficients out
    const size_t n = (size_t)1 << L;
    std::vector<uint32_t> raw(4 * n);
    CB_CUDA(cudaMemcpyAsync(raw.data(), line.p, raw.size() * 4, cudaMemcpyDeviceToHost, st));
    ctx->sync();
    std::vector<QM31> v(n);
    for (size_t i = 0; i < n; i++) {
        size_t src = host::bit_reverse((uint32_t)i, L);
        v[i] = {{raw[src], raw[n + src], raw[2 * n + src], raw[3 * n + src]}};
    }
    host::Coset cs = host::Coset::half_odds(L);
    uint32_t init = cs.initial, step = cs.step;
    for (size_t size = n; size > 1; size >>= 1) {
        for (size_t blk = 0; blk < n; blk += size)
            for (size_t i = 0; i < size / 2; i++) {
                uint32_t x = host::index_to_point((init + step * (uint32_t)i) & 0x7fffffffu).x;
                QM31 a = v[blk + i], b = v[blk + size / 2 + i];
                v[blk + i] = qadd(a, b);
                v[blk + size / 2 + i] = qmul_m(qsub(a, b), inv(x));
            }
        init = (init * 2) & 0x7fffffffu;
        step = (step * 2) & 0x7fffffffu;
    }
    uint32_t ninv = inv((uint32_t)n);
    std::vector<QM31> ordered(n);
    for (size_t i = 0; i < n; i++) ordered[i] = qmul_m(v[host::bit_reverse((uint32_t)i, L)], ninv);
    size_t bound = (size_t)1 << cfg.log_last_layer_degree_bound;
    for (size_t i = bound; i < n; i++)
        if (!qeq(ordered[i], qzero())) throw CbError("invalid degree");
    f.last_poly.assign(ordered.begin(), ordered.begin() + bound);
    ch.mix_felts(f.last_poly.data(), f.last_poly.size());
    return f;
}

uint64_t grind(cb_ctx* ctx, const host::Channel& ch, uint32_t pow_bits) {
    host::Hash32 pd = ch.pow_prefixed_digest(pow_bits);
    DBuf<uint32_t> d_pd(ctx, 8);
    DBuf<unsigned long long> d_best(ctx, 1);
    unsigned long long best = ~0ull;
    CB_CUDA(cudaMemcpyAsync(d_pd.p, pd.b, 32, cudaMemcpyHostToDevice, ctx->stream));
    CB_CUDA(cudaMemcpyAsync(d_best.p, &best, 8, cudaMemcpyHostToDevice, ctx->stream));
    const uint64_t batch = 1ull << 20;
    for (uint64_t base = 0;; base += batch) {
        CB_CUDA(launch_grind(ctx->stream, d_pd.p, pow_bits, base, batch, d_best.p));
        ctx->launches++;
        CB_CUDA(cudaMemcpyAsync(&best, d_best.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->sync();
        if (best != ~0ull) return best;
    }
}

// FriProver::decommit_on_queries: per layer {fri_witness, decommitment, commitment}; then last_layer_poly{coeffs, log_size}
std::vector<uint8_t> fri_decommit(cb_ctx* ctx, FriProverState& f, const PcsConfig& cfg, const std::vector<uint32_t>& queries) {
    std::vector<uint8_t> out;
    std::vector<uint32_t> q = queries;
    for (size_t li = 0; li < f.trees.size(); li++) {
        if (li == 1) host::put_u64(out, f.trees.size() - 1);  // Vec<inner layers> length prefix
        const size_t n = (size_t)1 << f.logs[li];
        // compute_decommitment_positions_and_witness_evals (fold_step 1)
        std::vector<uint32_t> positions, wpos;
        for (size_t i = 0; i < q.size();) {
            uint32_t start = (q[i] >> 1) << 1;
            bool has0 = false, has1 = false;
            while (i < q.size() && ((q[i] >> 1) << 1) == start) {
                if (q[i] & 1) has1 = true; else has0 = true;
                i++;
            }
            positions.push_back(start);
            positions.push_back(start + 1);
            if (!has0) wpos.push_back(start);
            if (!has1) wpos.push_back(start + 1);
        }
        std::vector<uint32_t> wv(wpos.size() * 4);
        if (!wpos.empty()) {
            DBuf<uint32_t> d_rows(ctx, wpos.size()), d_out(ctx, wv.size());
            CB_CUDA(cudaMemcpyAsync(d_rows.p, wpos.data(), wpos.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
            CB_CUDA(launch_gather_rows(ctx->stream, f.evals[li].p, n, 4, d_rows.p, (int)wpos.size(), d_out.p));
            ctx->launches++;
            CB_CUDA(cudaMemcpyAsync(wv.data(), d_out.p, wv.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
            ctx->sync();
        }
        host::put_u64(out, wpos.size());
        for (size_t w = 0; w < wpos.size(); w++)
            for (int c = 0; c < 4; c++) host::put_u32(out, wv[(size_t)c * wpos.size() + w]);
        std::vector<host::Hash32> dec = merkle_decommit(ctx, f.trees[li], positions);
        host::put_u64(out, dec.size());
        for (auto& h : dec) host::put_bytes(out, h.b, 32);
        host::put_bytes(out, f.trees[li].root.b, 32);
        q = host::fold_positions(q, 1);
    }
    if (f.trees.size() == 1) host::put_u64(out, 0);
    host::put_u64(out, f.last_poly.size());
    for (auto& c : f.last_poly) host::put_qm31(out, c);
    host::put_u32(out, cfg.log_last_layer_degree_bound);
    return out;
}

// StarkProof::size_estimate() as the reference reports it in "proof_size_bytes" (wasm_api.rs:593): payload bytes of hashes,
// field elements and the nonce, plus size_of::<PcsConfig>() = 28 on the reference's wasm32 target (pinned vs. the reference:
// 933,332 for a one-block ChaCha proof).  Walks the bincode layout produced by the drivers.
size_t stark_proof_size_estimate(const uint8_t* p, size_t len, size_t stark_off) {
    size_t pos = stark_off + 25;  // config
    size_t total = 28;
    auto u64 = [&]() {
        uint64_t v = 0;
        if (pos + 8 > len) throw CbError("proof truncated");
        for (int i = 0; i < 8; i++) v |= (uint64_t)p[pos + i] << (8 * i);
        pos += 8;
        return v;
    };
    uint64_t n = u64();                       // commitments
    pos += 32 * n; total += 32 * n;
    uint64_t nt = u64();                      // sampled values
    for (uint64_t t = 0; t < nt; t++) {
        uint64_t nc = u64();
        for (uint64_t c = 0; c < nc; c++) { uint64_t k = u64(); pos += 16 * k; total += 16 * k; }
    }
    nt = u64();                               // decommitments
    for (uint64_t t = 0; t < nt; t++) { uint64_t k = u64(); pos += 32 * k; total += 32 * k; }
    nt = u64();                               // queried values
    for (uint64_t t = 0; t < nt; t++) {
        uint64_t nc = u64();
        for (uint64_t c = 0; c < nc; c++) { uint64_t k = u64(); pos += 4 * k; total += 4 * k; }
    }
    pos += 8; total += 8;                     // proof of work
    auto layer = [&]() {
        uint64_t k = u64(); pos += 16 * k; total += 16 * k;
        k = u64(); pos += 32 * k; total += 32 * k;
        pos += 32; total += 32;
    };
    layer();
    uint64_t ni = u64();
    for (uint64_t i = 0; i < ni; i++) layer();
    uint64_t k = u64(); pos += 16 * k; total += 16 * k;
    return total;
}
