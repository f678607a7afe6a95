// ChaCha20 stream proof driver: replays the reference's prove flow on the GPU backend.
//
// Mirrors /root/reference/stwo/src/wasm_api.rs:467-602 (generate_chacha20_proof: validation, log_size, lane packing) and
// /root/reference/stwo/src/chacha/bitwise/air_stream.rs:143-234 (prove_stream_with_inputs / prove_stream_internal:
// empty preprocessed tree, witness, statement mixing :66-123, trace commit, prove), followed by upstream
// stwo::prover::prove / CommitmentSchemeProver::prove_values / FriProver::{commit,decommit}.
// The Fiat-Shamir channel (tiny Blake2s calls) stays on the host; every heavy step is a kernel from kernels_*.cu.
// Output bytes = bincode(StreamProof{stmt, stark_proof}) exactly as the reference serialises it
// (air_stream.rs:30-131, wasm_api.rs:588): byte-identical to the reference, checked in tests/.
#include <array>
#include <chrono>
#include <thread>
#include "prover.hpp"

using namespace m31;
using host::Channel;
using host::Hash32;

namespace {

constexpr int N_WORDS = 1040;  // packed witness words of the stream AIR (the block AIR uses the first 1,008)
// The two AIRs served here.  [0] stream AIR (chacha/bitwise/{gen_stream,constraints_stream,air_stream}.rs): state, 80 quarter
// rounds, final additions, plaintext, ciphertext, keystream xor plaintext = ciphertext.  [1] block AIR
// (chacha/bitwise/{gen,constraints,air}.rs, `prove_bitwise`): the same trace without the last 1,024 columns and the same
// constraints without the plaintext / ciphertext booleans and the 512 xor equalities - a prefix of [0] in both orders.
struct AirDims {
    int words, cols, cons, indep;  // packed words, columns, constraints, words transformed (the rest are adder sums)
};
constexpr AirDims DIMS[2] = {{1040, 33280, 54784, 704}, {1008, 32256, 53248, 672}};

// ---- host evaluation of the AIR on QM31 mask values (prove()'s closing sanity check; same sequence as the kernel) ----
struct QAcc {
    const std::vector<QM31>& apr;  // apr[k] = alpha^(K-1-k)
    QM31 acc = qzero();
    int k = 0;
    void add(QM31 c) { acc = qadd(acc, qmul(c, apr[k++])); }
};

QM31 eval_constraints_at_mask(const std::vector<QM31>& v, const std::vector<QM31>& apr, bool block) {
    QAcc A{apr};
    int col = 0;
    const QM31 one = qone();
    auto boolc = [&](QM31 b) { return qmul(b, qsub(one, b)); };
    using U32 = std::array<QM31, 32>;
    auto next_u32 = [&]() {
        U32 r;
        for (int i = 0; i < 32; i++) { r[i] = v[col++]; A.add(boolc(r[i])); }
        return r;
    };
    auto add_u32 = [&](const U32& a, const U32& b) {
        U32 res = next_u32();
        U32 car;
        for (int i = 0; i < 32; i++) car[i] = v[col++];
        for (int i = 0; i < 32; i++) {
            QM31 cin = i == 0 ? qzero() : car[i - 1];
            A.add(boolc(car[i]));
            A.add(qsub(qsub(qsub(qadd(res[i], qadd(car[i], car[i])), a[i]), b[i]), cin));
        }
        return res;
    };
    auto xor_rotl = [&](const U32& a, const U32& b, int r) {
        U32 res = next_u32();
        for (int i = 0; i < 32; i++) {
            int s = (i + 32 - r) % 32;
            QM31 ab = qmul(a[s], b[s]);
            A.add(qadd(qsub(qsub(res[i], a[s]), b[s]), qadd(ab, ab)));
        }
        return res;
    };
    std::array<U32, 16> init, s;
    for (int i = 0; i < 16; i++) init[i] = next_u32();
    s = init;
    static const int QRS[8][4] = {{0, 4, 8, 12}, {1, 5, 9, 13}, {2, 6, 10, 14}, {3, 7, 11, 15},
                                  {0, 5, 10, 15}, {1, 6, 11, 12}, {2, 7, 8, 13}, {3, 4, 9, 14}};
    for (int rnd = 0; rnd < 10; rnd++)
        for (auto& q : QRS) {
            int a = q[0], b = q[1], c = q[2], d = q[3];
            s[a] = add_u32(s[a], s[b]); s[d] = xor_rotl(s[a], s[d], 16);
            s[c] = add_u32(s[c], s[d]); s[b] = xor_rotl(s[c], s[b], 12);
            s[a] = add_u32(s[a], s[b]); s[d] = xor_rotl(s[a], s[d], 8);
            s[c] = add_u32(s[c], s[d]); s[b] = xor_rotl(s[c], s[b], 7);
        }
    std::array<U32, 16> ks, pt, ct;
    for (int i = 0; i < 16; i++) ks[i] = add_u32(s[i], init[i]);
    if (block) return A.acc;
    for (int i = 0; i < 16; i++) pt[i] = next_u32();
    for (int i = 0; i < 16; i++) ct[i] = next_u32();
    for (int i = 0; i < 16; i++)
        for (int b = 0; b < 32; b++) {
            QM31 kp = qmul(ks[i][b], pt[i][b]);
            A.add(qsub(qsub(qadd(ks[i][b], pt[i][b]), qadd(kp, kp)), ct[i][b]));
        }
    return A.acc;
}

// The same traversal recorded as a table: one ConsRec per constraint, in constraint order (ConsRec: common.cuh).  The table
// is what both the device kernel (mask_constraints_kernel: prove()'s closing check and the row-N evaluation of the
// half-domain composition) and the host verifier evaluate; the walk above stays as the generator's cross-check.
std::vector<ConsRec> build_cons_recs(bool block) {
    const AirDims D = DIMS[block ? 1 : 0];
    std::vector<ConsRec> T;
    T.reserve(D.cons);
    int col = 0;
    using U32 = std::array<int, 32>;
    auto next_u32 = [&]() {
        U32 r;
        for (int i = 0; i < 32; i++) { r[i] = col++; T.push_back({CR_BOOL, r[i], -1, -1, -1, -1}); }
        return r;
    };
    auto add_u32 = [&](const U32& a, const U32& b) {
        U32 res = next_u32();
        U32 car;
        for (int i = 0; i < 32; i++) car[i] = col++;
        for (int i = 0; i < 32; i++) {
            T.push_back({CR_BOOL, car[i], -1, -1, -1, -1});
            T.push_back({CR_ADD, res[i], car[i], a[i], b[i], i == 0 ? -1 : car[i - 1]});
        }
        return res;
    };
    auto xor_rotl = [&](const U32& a, const U32& b, int r) {
        U32 res = next_u32();
        for (int i = 0; i < 32; i++) {
            int sft = (i + 32 - r) % 32;
            T.push_back({CR_XOR, res[i], a[sft], b[sft], -1, -1});
        }
        return res;
    };
    std::array<U32, 16> init, st;
    for (int i = 0; i < 16; i++) init[i] = next_u32();
    st = init;
    static const int QRS[8][4] = {{0, 4, 8, 12}, {1, 5, 9, 13}, {2, 6, 10, 14}, {3, 7, 11, 15},
                                  {0, 5, 10, 15}, {1, 6, 11, 12}, {2, 7, 8, 13}, {3, 4, 9, 14}};
    for (int rnd = 0; rnd < 10; rnd++)
        for (auto& q : QRS) {
            int a = q[0], b = q[1], c = q[2], d = q[3];
            st[a] = add_u32(st[a], st[b]); st[d] = xor_rotl(st[a], st[d], 16);
            st[c] = add_u32(st[c], st[d]); st[b] = xor_rotl(st[c], st[b], 12);
            st[a] = add_u32(st[a], st[b]); st[d] = xor_rotl(st[a], st[d], 8);
            st[c] = add_u32(st[c], st[d]); st[b] = xor_rotl(st[c], st[b], 7);
        }
    std::array<U32, 16> ks, pt, ct;
    for (int i = 0; i < 16; i++) ks[i] = add_u32(st[i], init[i]);
    if (!block) {
        for (int i = 0; i < 16; i++) pt[i] = next_u32();
        for (int i = 0; i < 16; i++) ct[i] = next_u32();
        for (int i = 0; i < 16; i++)
            for (int b = 0; b < 32; b++) T.push_back({CR_EQ, ks[i][b], pt[i][b], ct[i][b], -1, -1});
    }
    if ((int)T.size() != D.cons || col != D.cols) throw CbError("internal: constraint table size");
    return T;
}

const std::vector<ConsRec>& cons_recs(bool block) {
    auto make = [](bool blk) {
        const AirDims D = DIMS[blk ? 1 : 0];
        std::vector<ConsRec> t = build_cons_recs(blk);
        // one-time cross-check against the walk on a pseudo-random mask
        std::vector<QM31> mask(D.cols), apr(D.cons);
        uint64_t x = 0x9e3779b97f4a7c15ull;
        auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (uint32_t)(x % P); };
        for (auto& m : mask) m = {{rnd(), rnd(), rnd(), rnd()}};
        for (auto& a : apr) a = {{rnd(), rnd(), rnd(), rnd()}};
        QM31 acc = qzero();
        for (int k = 0; k < D.cons; k++) acc = qadd(acc, qmul(cons_rec_eval(t[k], mask.data()), apr[k]));
        if (!qeq(acc, eval_constraints_at_mask(mask, apr, blk))) throw CbError("internal: constraint table differs from the AIR walk");
        return t;
    };
    static const std::vector<ConsRec> T0 = make(false), T1 = make(true);
    return block ? T1 : T0;
}

QM31 eval_cons_table(const std::vector<QM31>& mask, const std::vector<QM31>& apr, bool block) {
    const std::vector<ConsRec>& T = cons_recs(block);
    QM31 acc = qzero();
    for (size_t k = 0; k < T.size(); k++) acc = qadd(acc, qmul(cons_rec_eval(T[k], mask.data()), apr[k]));
    return acc;
}

}  // namespace

// shared with the host verifier (verify.cu)
QM31 chacha_constraints_at_mask(const std::vector<QM31>& mask, const std::vector<QM31>& alpha_powers_rev, bool block_air) {
    return eval_cons_table(mask, alpha_powers_rev, block_air);
}

// ------------------------------------------------------------------------------------------------ streaming plan
// Packed witness word indices (kernels_chacha.cu): 0..15 initial state | 80 quarter rounds x [sum,carry,xor]x4 |
// 16 final adds x [sum,carry] | 16 plaintext | 16 ciphertext.  Sum words are "dependent" (combined from operand tiles,
// kernels_stream.cu fact 1); all other words are transformed from the packed witness.
namespace {

struct Comb { int res, a, b, c; };
struct CJ { int type, w0, w1, w2, kx, kb0, kb1, kb2, arg, w3, w4, wres, kbc; };
struct Group {
    std::vector<int> fft;         // independent words transformed in this group
    std::vector<Comb> comb;       // adder sum words, in dependency order
    std::vector<int> hash;        // words absorbed into the Merkle leaves, in column order (<= MAX_LEAF_GROUPS)
    std::vector<CJ> cons;         // constraints evaluated while the group's tiles are live (<= MAX_CONSTRAINT_JOBS)
    std::vector<int> free_after;  // tiles dead after this group
};

std::vector<Group> build_plan(bool block) {
    std::vector<Group> plan;
    int state[16];
    {
        Group g;
        for (int w = 0; w < 16; w++) {
            state[w] = w;
            g.fft.push_back(w);
            g.hash.push_back(w);
            g.cons.push_back({CJ_BOOL, w, -1, -1, -1, 32 * w, -1, -1, 1, -1, -1, -1, -1});
        }
        plan.push_back(g);
    }
    static const int ROT[4] = {16, 12, 8, 7};
    for (int q = 0; q < 80; q++) {
        const int qi = q & 7, a0 = qi & 3;
        int a = a0, b, c, d;
        if (qi < 4) { b = 4 + a0; c = 8 + a0; d = 12 + a0; }
        else { b = 4 + ((a0 + 1) & 3); c = 8 + ((a0 + 2) & 3); d = 12 + ((a0 + 3) & 3); }
        const int base = 16 + 12 * q, kb = 512 + 640 * q;
        const int S1 = base, C1 = base + 1, X1 = base + 2, S2 = base + 3, C2 = base + 4, X2 = base + 5, S3 = base + 6,
                  C3 = base + 7, X3 = base + 8, S4 = base + 9, C4 = base + 10, X4 = base + 11;
        Group g;
        g.fft = {C1, X1, C2, X2, C3, X3, C4, X4};
        g.comb = {{S1, state[a], state[b], C1}, {S2, state[c], X1, C2}, {S3, S1, X2, C3}, {S4, S2, X3, C4}};
        for (int w = base; w < base + 12; w++) g.hash.push_back(w);
        const int S[4] = {S1, S2, S3, S4}, C[4] = {C1, C2, C3, C4}, X[4] = {X1, X2, X3, X4};
        const int XD[4] = {state[d], state[b], X1, X2};  // second xor operand (the rotated word's previous value)
        for (int t = 0; t < 4; t++) {
            const int k = kb + 160 * t;
            // adder t: 32 sum booleans at k, then [carry boolean, adder identity] pairs from k+32 (identity skipped: it is
            // identically zero, see kernels_stream.cu); xor t: 32 result booleans at k+96, 32 xor constraints at k+128
            // one fused job per adder + xor-rotate pair; the sum tile S[t] is computed inside it
            const Comb& cb = g.comb[t];
            g.cons.push_back({CJ_ADDX, X[t], cb.a, XD[t], k + 128, k + 96, k, -1, ROT[t], cb.b, C[t], S[t], k + 32});
        }
        for (int w : {state[a], state[b], state[c], state[d]})
            if (w >= 16) g.free_after.push_back(w);
        for (int w : {S1, C1, X1, S2, C2, X2, C3, C4}) g.free_after.push_back(w);
        state[a] = S3; state[b] = X4; state[c] = S4; state[d] = X3;
        plan.push_back(g);
    }
    const int kf = 512 + 640 * 80, k_pt = kf + 96 * 16, k_ct = k_pt + 512, k_eq = k_ct + 512;
    for (int half = 0; half < 2; half++) {
        Group g;
        for (int i = 8 * half; i < 8 * half + 8; i++) {
            const int S = 976 + 2 * i, C = 977 + 2 * i;
            g.fft.push_back(C);
            g.comb.push_back({S, state[i], i, C});
            g.hash.push_back(S);
            g.hash.push_back(C);
            g.cons.push_back({CJ_ADDX, -1, state[i], -1, -1, -1, kf + 96 * i, -1, 0, i, C, S, kf + 96 * i + 32});
            if (state[i] >= 16) g.free_after.push_back(state[i]);
            g.free_after.push_back(C);
            g.free_after.push_back(i);  // initial-state tile i is no longer needed
            if (block) g.free_after.push_back(S);  // block AIR: the keystream word has no later consumer
        }
        plan.push_back(g);
    }
    if (block) return plan;
    {
        Group g;
        for (int i = 0; i < 16; i++) {
            g.fft.push_back(1008 + i);
            g.hash.push_back(1008 + i);
        }
        plan.push_back(g);
    }
    {
        Group g;
        for (int i = 0; i < 16; i++) {
            g.fft.push_back(1024 + i);
            g.hash.push_back(1024 + i);
            // ciphertext = keystream xor plaintext, plus the booleans of the three words involved
            g.cons.push_back({CJ_XORN, 1024 + i, 976 + 2 * i, 1008 + i, k_eq + 32 * i, k_ct + 32 * i, -1, k_pt + 32 * i, 0, -1, -1, -1, -1});
            g.free_after.push_back(1024 + i);
            g.free_after.push_back(1008 + i);
            g.free_after.push_back(976 + 2 * i);
        }
        plan.push_back(g);
    }
    return plan;
}

// Alpha-table entries of every group's jobs in the order the FP64 constraint kernel consumes them (constraints_tiles_kernel2):
// per job 32 steps x {1 (CJ_BOOL), 2 (adder only), 4 (adder + xor-rotate, xor jobs)} constraint indices; -1 = no constraint.
// job_off[g][j] = offset of job j of group g in the list.
struct ConsTable {
    std::vector<int> idx;
    std::vector<std::vector<int>> job_off;
};
ConsTable build_cons_table(const std::vector<Group>& plan) {
    ConsTable t;
    for (auto& g : plan) {
        t.job_off.emplace_back();
        for (auto& c : g.cons) {
            t.job_off.back().push_back((int)t.idx.size());
            for (int s = 0; s < 32; s++) {
                if (c.type == CJ_BOOL) {
                    t.idx.push_back(c.kb0 + s * c.arg);
                } else if (c.type == CJ_ADDX) {
                    t.idx.push_back(c.kbc + 2 * s);
                    t.idx.push_back(c.kb1 + s);
                    if (c.w0 >= 0) {
                        const int i = (s + c.arg) & 31;
                        t.idx.push_back(c.kx + i);
                        t.idx.push_back(c.kb0 + i);
                    }
                } else {  // CJ_XOR / CJ_XORN: step i reads operand bit (i - rot) mod 32
                    const int i = s, sb = (i + 32 - c.arg) & 31;
                    t.idx.push_back(c.kx + i);
                    t.idx.push_back(c.kb0 >= 0 ? c.kb0 + i : -1);
                    t.idx.push_back(c.kb1 >= 0 ? c.kb1 + sb : -1);
                    t.idx.push_back(c.kb2 >= 0 ? c.kb2 + sb : -1);
                }
            }
        }
    }
    return t;
}

// LDE tile slots: a cache of independent tiles that survives from the commitment pass to the constraint pass, plus
// transient slots recycled as words die.
struct Tiles {
    int n_cache = 0, n_trans = 0, cache_used = 0;
    size_t tile_words = 0;
    uint32_t* arena = nullptr;
    std::vector<int> slot_of, cache_slot, free_trans;
    std::vector<char> want;  // independent words that get a cache slot in pass 1 (chosen up front, cache_choice())
    // slots released while group g is processed become reusable `lag` groups later (lag 2 when tiles are produced on a second
    // stream: the producer of group g only waits for the consumer of group g-2)
    std::vector<std::vector<int>> pending;
    int lag = 0, group = 0;
    int in_use = 0, peak = 0;
    void init(int cache, int trans, size_t tw_, uint32_t* mem, const std::vector<char>& want_) {
        n_cache = cache; n_trans = trans; tile_words = tw_; arena = mem; cache_used = 0;
        want = want_;
        slot_of.assign(N_WORDS, -1);
        cache_slot.assign(N_WORDS, -1);
        free_trans.clear();
        for (int i = trans - 1; i >= 0; i--) free_trans.push_back(i);
        in_use = peak = 0;
        pending.clear();
        group = 0;
    }
    void begin_group(int g) {
        group = g;
        if ((int)pending.size() <= g) pending.resize(g + 1);
        if (g - lag >= 0)
            for (int s = 0; s <= g - lag; s++) {
                for (int x : pending[s]) { free_trans.push_back(x); in_use--; }
                pending[s].clear();
            }
    }
    void flush() {  // all work using the released slots has completed
        for (auto& v : pending) {
            for (int x : v) { free_trans.push_back(x); in_use--; }
            v.clear();
        }
    }
    uint32_t* ptr(int w) const {
        if (slot_of[w] < 0) throw CbError("internal: tile of word " + std::to_string(w) + " is not live");
        return arena + (size_t)slot_of[w] * tile_words;
    }
    // returns true when the tile must be (re)computed
    bool acquire(int w, bool indep, int pass) {
        if (indep && pass == 2 && cache_slot[w] >= 0) { slot_of[w] = cache_slot[w]; return false; }
        if (indep && pass == 1 && want[w] && cache_used < n_cache) { cache_slot[w] = cache_used++; slot_of[w] = cache_slot[w]; return true; }
        if (free_trans.empty()) throw CbError("internal: transient tile slots exhausted");
        slot_of[w] = n_cache + free_trans.back();
        free_trans.pop_back();
        if (++in_use > peak) peak = in_use;
        return true;
    }
    void release(int w) {
        if (slot_of[w] < 0) return;
        if (slot_of[w] >= n_cache) {
            if ((int)pending.size() <= group) pending.resize(group + 1);
            pending[group].push_back(slot_of[w] - n_cache);
        }
        slot_of[w] = -1;
    }
};

// Which independent words keep their LDE tile between the passes when only n_cache fit: the words of the last groups
// (final additions, plaintext, ciphertext) and of the first group come first, then the quarter-round groups in order.
// Those tail groups hold up to 48 tiles live at once; left uncached they alone would set the transient-slot peak
// (48 instead of 28 slots, i.e. 20 fewer cached tiles).
std::vector<char> cache_choice(const std::vector<Group>& plan, int n_cache) {
    std::vector<int> order;
    for (size_t g = plan.size() - 4; g < plan.size(); g++)
        for (int w : plan[g].fft) order.push_back(w);
    for (int w : plan[0].fft) order.push_back(w);
    for (size_t g = 1; g + 4 < plan.size(); g++)
        for (int w : plan[g].fft) order.push_back(w);
    std::vector<char> want(N_WORDS, 0);
    for (int i = 0; i < n_cache && i < (int)order.size(); i++) want[order[i]] = 1;
    return want;
}

// transient slots a pass needs when the words in `want` are cached (both passes have the same live sets)
int plan_peak_transient(const std::vector<Group>& plan, int lag, const std::vector<char>& want) {
    Tiles t;
    int n_want = 0;
    for (char c : want) n_want += c;
    t.init(n_want, N_WORDS, 0, nullptr, want);
    t.lag = lag;
    int gi = 0;
    for (auto& g : plan) {
        t.begin_group(gi++);
        for (int w : g.fft) t.acquire(w, true, 1);
        for (auto& c : g.comb) t.acquire(c.res, false, 1);
        for (int w : g.free_after) t.release(w);
    }
    return t.peak;
}

}  // namespace

namespace {
// constraint table + adder-sum list, uploaded once per context
struct ChaChaDev {
    const ConsRec* table;
    const SumComb* combs;
    int n_combs;
    const int* indep;  // the N_INDEP_WORDS transformed words, plan order
};
ChaChaDev chacha_dev(cb_ctx* ctx, const std::vector<Group>& plan, bool block) {
    void*& consts = ctx->chacha_consts[block ? 1 : 0];
    const AirDims D = DIMS[block ? 1 : 0];
    std::vector<SumComb> cl;
    for (auto& g : plan)
        for (auto& c : g.comb) cl.push_back({c.res, c.a, c.b, c.c});
    const size_t tb = (size_t)D.cons * sizeof(ConsRec), cb = cl.size() * sizeof(SumComb);
    if (!consts) {
        const std::vector<ConsRec>& T = cons_recs(block);
        std::vector<int> iw;
        for (auto& g : plan)
            for (int w : g.fft) iw.push_back(w);
        CB_CUDA(cudaMalloc(&consts, tb + cb + iw.size() * sizeof(int)));
        CB_CUDA(cudaMemcpy(consts, T.data(), tb, cudaMemcpyHostToDevice));
        CB_CUDA(cudaMemcpy((char*)consts + tb, cl.data(), cb, cudaMemcpyHostToDevice));
        CB_CUDA(cudaMemcpy((char*)consts + tb + cb, iw.data(), iw.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    return {(const ConsRec*)consts, (const SumComb*)((char*)consts + tb), (int)cl.size(), (const int*)((char*)consts + tb + cb)};
}
}  // namespace

// Proves ChaCha20 encryption of `len` bytes (multiple of 64).  On success fills proof bytes (bincode StreamProof).
// Returns "" on success, else the reference's error string.
std::string prove_chacha20(cb_ctx* ctx, const uint8_t key[32], const uint8_t nonce[12], uint32_t counter, const uint8_t* plaintext,
                           const uint8_t* ciphertext, size_t len, std::vector<uint8_t>& proof, ProveOptions opt) {
    const PcsConfig cfg;
    const bool block_air = opt.block_air;
    const int variant = block_air ? 1 : 0;
    const AirDims D = DIMS[variant];
    const uint32_t num_blocks = (uint32_t)(len / 64);
    int log_size = 4;
    while (((size_t)1 << log_size) < num_blocks) log_size++;
    if (block_air && (len != ((size_t)64 << log_size) || plaintext || ciphertext || opt.pt_dev))
        return "block AIR: the trace is generated from log_size alone";
    if (log_size > 24) return "log_size (" + std::to_string(log_size) + ") must be <= MAX_LOG_SIZE (24)";
    const int n = log_size, m = n + cfg.log_blowup;  // trace / LDE domain logs
    const size_t N = (size_t)1 << n, M = (size_t)1 << m;
    const uint32_t rows_needed = (num_blocks + 15) / 16;
    const uint32_t inv_n = 1u << (31 - n);
    // row-sharded mode (several ranks prove this one trace): each rank holds rows [rank*Mr, (rank+1)*Mr) of every LDE tile
    const Comm& cm = ctx->comm;
    const int G = cm.world, R = cm.rank;
    int logG = 0;
    while ((1 << logG) < G) logG++;
    if (G > 1 && block_air) return "sharded proving serves the stream AIR only";
    if (G > 1 && n < 16) return "sharded proving needs log_size >= 16 (got " + std::to_string(n) + ")";
    const int lr = m - logG;             // log2 of the rows per shard
    const size_t Mr = M >> logG;
    cudaStream_t st = ctx->stream;
    ctx->ensure_twiddles(m);
    ctx->pending_events.clear();
    ctx->host_marks.clear();
    ctx->host_mark("setup");
    StageHook hk = ctx->hook();
    const StageHook* hkp = ctx->profile ? &hk : nullptr;

    uint32_t key_w[8], nonce_w[3];
    for (int i = 0; i < 8; i++) key_w[i] = host::load_le32(key + 4 * i);
    for (int i = 0; i < 3; i++) nonce_w[i] = host::load_le32(nonce + 4 * i);

    // ChaChaPublicInputs::new (air_stream.rs:44-53) hashes the whole plaintext and ciphertext on the host; run both hashes on
    // their own threads so they overlap the GPU's commitment pass
    Hash32 pth, cth;
    struct Hashers {
        std::thread a, b;
        void join() { if (a.joinable()) a.join(); if (b.joinable()) b.join(); }
        ~Hashers() { join(); }
    } hashers;
    std::string hash_err;
    // (sharded mode: only rank 0 hashes and broadcasts the two digests - G processes x 2 hashing threads would fight for the
    // host cores, and at 8 ranks the commitment pass is no longer than one 64 MiB hash)
    if (!opt.pt_hash && !opt.empty_public_hashes && cm.rank == 0 && !block_air) {
        if (opt.pt_dev) {
            // inputs resident in HBM and no hashes supplied: read both buffers back into the context's pinned staging area
            // (2 x len bytes of D2H at PCIe speed, a few ms) and hash them on host threads like host-resident inputs
            if (ctx->hash_stage_bytes < 2 * len) {
                if (ctx->hash_stage) cudaFreeHost(ctx->hash_stage);
                ctx->hash_stage = nullptr;
                ctx->hash_stage_bytes = 0;
                CB_CUDA(cudaHostAlloc((void**)&ctx->hash_stage, 2 * len, cudaHostAllocDefault));
                ctx->hash_stage_bytes = 2 * len;
            }
            CB_CUDA(cudaMemcpyAsync(ctx->hash_stage, opt.pt_dev, len, cudaMemcpyDeviceToHost, st));
            CB_CUDA(cudaMemcpyAsync(ctx->hash_stage + len, opt.ct_dev, len, cudaMemcpyDeviceToHost, st));
            ctx->sync();
            const uint8_t* hp = ctx->hash_stage;
            hashers.a = std::thread([&pth, hp, len] { pth = host::blake2s_bytes(hp, len); });
            hashers.b = std::thread([&cth, hp, len] { cth = host::blake2s_bytes(hp + len, len); });
        } else if (len <= ((size_t)1 << 16)) {  // product-size inputs: cheaper than starting two threads
            pth = host::blake2s_bytes(plaintext, len);
            cth = host::blake2s_bytes(ciphertext, len);
        } else {
            hashers.a = std::thread([&] { pth = host::blake2s_bytes(plaintext, len); });
            hashers.b = std::thread([&] { cth = host::blake2s_bytes(ciphertext, len); });
        }
    }

    Channel ch;
    std::vector<Hash32> roots;
    // tree 0: empty preprocessed tree -> root = Blake2s("")
    roots.push_back(host::blake2s_bytes(nullptr, 0));
    ch.mix_root(roots[0]);

    // ---- witness (packed: 1,040 words per row)
    ctx->stage_begin("witness");
    DBuf<uint32_t> d_pt, d_ct, W(ctx, (size_t)N_WORDS * N);
    DBuf<int> d_invalid(ctx, 1);
    const uint32_t *pt_d = opt.pt_dev, *ct_d = opt.ct_dev;
    if (block_air) {  // no plaintext / ciphertext columns: the witness kernel's last 32 words are written (zeros) and never read
        d_pt = DBuf<uint32_t>(ctx, len / 4);
        CB_CUDA(cudaMemsetAsync(d_pt.p, 0, len, st));
        pt_d = ct_d = d_pt.p;
    } else if (!pt_d) {
        d_pt = DBuf<uint32_t>(ctx, len / 4);
        d_ct = DBuf<uint32_t>(ctx, len / 4);
        CB_CUDA(cudaMemcpyAsync(d_pt.p, plaintext, len, cudaMemcpyHostToDevice, st));
        CB_CUDA(cudaMemcpyAsync(d_ct.p, ciphertext, len, cudaMemcpyHostToDevice, st));
        pt_d = d_pt.p;
        ct_d = d_ct.p;
    }
    CB_CUDA(cudaMemsetAsync(d_invalid.p, 0, sizeof(int), st));
    CB_CUDA(launch_chacha_witness(st, key_w, nonce_w, counter, num_blocks, rows_needed * 16, pt_d, ct_d, n, W.p, N, d_invalid.p));
    ctx->launches++;
    ctx->stage_end();
    int invalid = 0;
    CB_CUDA(cudaMemcpyAsync(&invalid, d_invalid.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    ctx->sync();
    if (invalid && !block_air) return "Ciphertext does not match encryption - invalid witness";
    d_pt.release();
    d_ct.release();

    // ---- tile arena: as many independent tiles as fit stay cached between the two LDE passes
    static const std::vector<Group> plans[2] = {build_plan(false), build_plan(true)};
    const std::vector<Group>& plan = plans[variant];
    const ChaChaDev cdev = chacha_dev(ctx, plan, block_air);
    // tiles are transformed on a second stream one group ahead of their consumer (not while per-kernel profiling is on)
    const bool overlap = ctx->overlap && !ctx->profile && ctx->stream2 != nullptr && G == 1;
    // Row-sharded mode over peer windows (G > 1, CUDA IPC available): LDE rows are dealt to the ranks in 2G "virtual shards" of
    // Mv = M / 2G rows (rank r owns global rows [r Mv, (r+1) Mv) of the first half of the evaluation domain and the same range
    // of the second half), the owner of a column writes them straight into the other ranks' tile slots from the last transform
    // pass, and one tiny all-reduce per plan group orders producers and consumers.  A slot released in group g is written by
    // remote ranks from group g+2 on (they passed the barrier of group g+1, which this rank entered after consuming group g).
    const bool want_p2p = G > 1 && ctx->p2p_state >= 0 && getenv("S2C_NO_P2P") == nullptr;
    const int lag = (overlap || want_p2p) ? 2 : 0;
    static const int peak_none[2] = {plan_peak_transient(plans[0], 2, std::vector<char>(N_WORDS, 0)),
                                     plan_peak_transient(plans[1], 2, std::vector<char>(N_WORDS, 0))};  // nothing cached
    const int peak_trans_none = peak_none[variant];
    cudaStream_t sf = overlap ? ctx->stream2 : st;
    // placement knobs (KiB) for measuring how the power-of-two strides of the transform passes interact with the DRAM
    // address map: extra pitch between tile slots, offset of the FFT scratch behind the slots
    static const size_t tile_pad_words = getenv("S2C_TILE_PAD_KB") ? (size_t)atol(getenv("S2C_TILE_PAD_KB")) * 256 : 0;
    static const size_t scratch_off_words = getenv("S2C_SCRATCH_OFF_KB") ? (size_t)atol(getenv("S2C_SCRATCH_OFF_KB")) * 256 : 0;
    const size_t tile_words = 32 * Mr + tile_pad_words;
    // sharded mode: whole transformed tiles ([G][32][Mr]) wait here for the all-to-all; at most ceil(16/G) jobs per rank and group
    const int n_stage = G > 1 ? (16 + G - 1) / G : 0;
    const size_t scratch_words = fft_packed_scratch_words(SRC_BITS, G > 1 ? n_stage : MAX_FFT_JOBS, n);
    // (two sets in peer-window mode: the first two passes of group g+1 fill one while the last pass of group g drains the other)
    const size_t stage_words = (size_t)n_stage * 32 * M * (want_p2p ? 2 : 1);
    // Product-size traces (log_size <= 10): the whole LDE (1,040 tiles, <= 272 MB) is materialised, so each pass is a handful
    // of launches over all words instead of 85 plan groups - at these sizes a proof is bound by launch count and by the dependent
    // chains inside the per-row kernels, not by arithmetic (profiles/r02: log 4, 417 launches, 12 ms).
    const bool small_mode = G == 1 && n <= 10 && getenv("S2C_NO_SMALL") == nullptr;
    int n_cache = D.indep;
    if (!small_mode) {
        size_t free_b = 0, total_b = 0;
        CB_CUDA(cudaMemGetInfo(&free_b, &total_b));
        cudaMemPool_t pool;
        CB_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
        uint64_t reserved = 0, used = 0;
        CB_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved));
        CB_CUDA(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used));
        const size_t avail = free_b + (size_t)(reserved - used) + ctx->arena_bytes;  // the arena is re-used (or re-made)
        // everything else this proof allocates: scratch, leaf state + tree (24 M words), accumulators / composition /
        // quotient / FRI columns (~48 M words), plus slack for the allocator
        const size_t other = (scratch_words + stage_words + 96 * M) * 4 + ((size_t)3 << 30);
        const size_t tile_bytes = tile_words * 4;
        size_t can = avail > other ? (avail - other) / tile_bytes : 0;
        // hysteresis: free memory moves by a few MB between proofs; do not re-make a 170 GB arena to gain or lose a few tiles
        if (ctx->arena) {
            const size_t fixed = (scratch_off_words + scratch_words + stage_words) * 4;
            const size_t have = ctx->arena_bytes > fixed ? (ctx->arena_bytes - fixed) / tile_bytes : 0;
            if (have + 8 >= can && have <= can + 8) can = have;
        }
        if (can < (size_t)peak_trans_none) throw CbError("not enough device memory for the tile arena at log_size " + std::to_string(n));
        const int cap = opt.max_cached_tiles >= 0 ? opt.max_cached_tiles : ctx->max_cached_tiles;
        if (cap >= 0 && n_cache > cap) n_cache = cap;
        // the largest cache whose tiles plus the transient slots the plan then needs fit
        // (n_cache + peak(n_cache) never decreases with n_cache: binary search; the answer is remembered per thread)
        static thread_local size_t memo_can = 0;
        static thread_local int memo_in = -1, memo_lag = -1, memo_out = 0, memo_variant = -1;
        if (memo_can == can && memo_in == n_cache && memo_lag == lag && memo_variant == variant) {
            n_cache = memo_out;
        } else {
            auto fits = [&](int c) { return (size_t)(c + plan_peak_transient(plan, lag, cache_choice(plan, c))) <= can; };
            int lo = 0, hi = n_cache;  // fits(0) holds (checked above)
            while (lo < hi) {
                const int mid = (lo + hi + 1) / 2;
                if (fits(mid)) lo = mid; else hi = mid - 1;
            }
            memo_can = can; memo_in = n_cache; memo_lag = lag; memo_out = lo; memo_variant = variant;
            n_cache = lo;
        }
        n_cache = comm_min_int(cm, n_cache, st);  // every rank must take the same caching decisions
    }
    const std::vector<char> want = small_mode ? std::vector<char>(N_WORDS, 1) : cache_choice(plan, n_cache);
    const int peak_trans = small_mode ? D.words - D.indep : plan_peak_transient(plan, lag, want);
    // tile slots + FFT scratch live in the context's persistent arena
    const size_t arena_words = small_mode ? (size_t)N_WORDS * tile_words
                                          : (size_t)(n_cache + peak_trans) * tile_words + scratch_off_words + scratch_words + stage_words;
    uint32_t* arena_p;
    const bool arena_ok = ctx->arena && ctx->arena_bytes >= arena_words * 4 && ctx->arena_bytes <= arena_words * 4 + ((size_t)8 << 30);
    bool p2p = false;
    if (G > 1) {
        p2p = ctx->sync_peer_arenas(!arena_ok, arena_words * 4) && want_p2p;
        arena_p = (uint32_t*)ctx->arena;
    } else if (arena_ok) {
        arena_p = (uint32_t*)ctx->arena;
    } else {
        ctx->close_peers();
        ctx->release_arena();
        arena_p = (uint32_t*)ctx->ensure_arena(arena_words * 4);
    }
    ctx->last_p2p = p2p;
    const int lv = lr - 1;        // virtual shards (peer-window mode)
    const size_t Mv = Mr >> 1;
    uint32_t* scratch_p = arena_p + (size_t)(n_cache + peak_trans) * tile_words + scratch_off_words;
    uint32_t* stage_p = scratch_p + scratch_words;
    Tiles tiles;
    tiles.init(n_cache, peak_trans, tile_words, arena_p, want);
    tiles.lag = lag;
    if (small_mode)  // tile of word w = slot w, all live for the whole proof
        for (int w = 0; w < N_WORDS; w++) tiles.slot_of[w] = tiles.cache_slot[w] = w;
    ctx->fft_words = 0;
    ctx->fft_words_half = 0;
    ctx->cached_tiles = n_cache;
    ctx->transient_tiles = peak_trans;

    // single-GPU mode evaluates the constraints on storage rows [0, N] only (see "half-domain evaluation" below): tiles
    // recomputed for the constraint pass need only their first half
    const bool half_mode = (G == 1 && !small_mode) || p2p;
    // peer-window mode, three streams: passes A/B of the transforms on `stream2`, the last pass (throttled by NVLink) and the
    // group barrier on `stream3`, the consumers on `stream` - so that the compute of group g+1 and the hashing of group g-1 fill
    // the SMs while the stores of group g cross the links.  Per group g: stream2: wait stage set free, passes A/B, record;
    // stream3: wait A/B, last pass -> peers, record set free, wait consumed(g-1), barrier, record ready(g); stream: wait ready(g).
    const bool ov2 = p2p && ctx->stream2 != nullptr && ctx->stream3 != nullptr && !ctx->profile && getenv("S2C_P2P_1STREAM") == nullptr;
    auto run_pass = [&](int pass, auto&& consume) {
        const size_t NG = plan.size();
        tiles.flush();
        cudaStream_t sp = ov2 ? ctx->stream2 : st;   // passes A/B
        cudaStream_t sc = ov2 ? ctx->stream3 : st;   // last pass + barrier
        int fft_seq = 0;                             // groups with local transform jobs so far (stage set = fft_seq & 1)
        if (ov2) {  // the producer streams start after everything enqueued so far
            CB_CUDA(cudaEventRecord(ctx->event(2 * NG), st));
            CB_CUDA(cudaStreamWaitEvent(sp, ctx->event(2 * NG), 0));
            CB_CUDA(cudaStreamWaitEvent(sc, ctx->event(2 * NG), 0));
        }
        for (size_t gi = 0; gi < NG; gi++) {
            const Group& g = plan[gi];
            tiles.begin_group((int)gi);
            std::vector<const uint32_t*> src;
            std::vector<uint32_t*> out;
            std::vector<int> jobs_w;
            for (int w : g.fft)
                if (tiles.acquire(w, true, pass)) jobs_w.push_back(w);
            if (p2p) {
                // column-sharded transform fused with the exchange: job j belongs to rank j mod G, whose last transform pass
                // stores each 4096-row chunk into the slot of the rank owning those rows
                std::vector<unsigned long long> offs;
                int k = 0;
                for (size_t j = 0; j < jobs_w.size(); j++)
                    if ((int)(j % G) == R) {
                        src.push_back(W.p + (size_t)jobs_w[j] * N);
                        out.push_back(stage_p + ((size_t)(fft_seq & 1) * n_stage + (k++)) * 32 * M);
                        offs.push_back((unsigned long long)(tiles.ptr(jobs_w[j]) - arena_p));
                    }
                if (!src.empty()) {
                    PeerDst pd{};
                    const int set = fft_seq & 1;
                    if (ov2) {
                        pd.last_stream = sc;
                        pd.ab_done = ctx->event(2 * NG + 1 + set);
                        if (fft_seq >= 2) CB_CUDA(cudaStreamWaitEvent(sp, ctx->event(2 * NG + 3 + set), 0));  // the set's last reader
                    }
                    for (int r = 0; r < G; r++) pd.base[r] = ctx->peer_arena[r];
                    pd.logG = logG;
                    pd.lv = lv;
                    pd.off = offs.data();
                    int nl = 0;
                    CB_CUDA(launch_fft_packed(sp, SRC_BITS, src.data(), out.data(), (int)src.size(), n, ctx->tw, scratch_p, hkp, &nl, 0,
                                              pass == 2, &pd));
                    ctx->launches += nl;
                    ctx->fft_words += src.size();
                    if (pass == 2) ctx->fft_words_half += src.size();
                    if (ov2) CB_CUDA(cudaEventRecord(ctx->event(2 * NG + 3 + set), sc));
                    fft_seq++;
                }
                if (pass == 1 || n_cache < D.indep) {
                    ctx->stage_begin("group_barrier");
                    if (ov2 && gi >= 1) CB_CUDA(cudaStreamWaitEvent(sc, ctx->event(NG + gi - 1), 0));
                    comm_barrier(ctx->comm, sc);
                    ctx->stage_end();
                    if (ov2) {
                        CB_CUDA(cudaEventRecord(ctx->event(gi), sc));
                        CB_CUDA(cudaStreamWaitEvent(st, ctx->event(gi), 0));
                    }
                }
                src.clear();
                out.clear();
            } else if (G > 1 && !jobs_w.empty()) {
                // column-sharded transform: job j of the group belongs to rank j mod G, which transforms the whole columns
                // into a staging tile laid out [G][32][Mr]; then the grouped send/recv all-to-all hands every rank its row
                // shard of every tile of the group
                int k = 0;
                for (size_t j = 0; j < jobs_w.size(); j++)
                    if ((int)(j % G) == R) {
                        src.push_back(W.p + (size_t)jobs_w[j] * N);
                        out.push_back(stage_p + (size_t)(k++) * 32 * M);
                    }
                if (!src.empty()) {
                    int nl = 0;
                    CB_CUDA(launch_fft_packed(st, SRC_BITS, src.data(), out.data(), (int)src.size(), n, ctx->tw, scratch_p, hkp, &nl, lr));
                    ctx->launches += nl;
                    ctx->fft_words += src.size();
                }
                ctx->stage_begin("all_to_all");
                comm_group_start();
                k = 0;
                for (size_t j = 0; j < jobs_w.size(); j++) {
                    const int owner = (int)(j % G);
                    uint32_t* slot = tiles.ptr(jobs_w[j]);
                    if (owner == R) {
                        const uint32_t* stg = stage_p + (size_t)(k++) * 32 * M;
                        for (int r = 0; r < G; r++)
                            if (r != R) comm_send_u32(cm, stg + (size_t)r * (32 * Mr), 32 * Mr, r, st);
                        CB_CUDA(cudaMemcpyAsync(slot, stg + (size_t)R * (32 * Mr), (size_t)32 * Mr * 4, cudaMemcpyDeviceToDevice, st));
                    } else {
                        comm_recv_u32(cm, slot, 32 * Mr, owner, st);
                    }
                }
                comm_group_end();
                ctx->stage_end();
                src.clear();
                out.clear();
            } else {
                for (int w : jobs_w) {
                    src.push_back(W.p + (size_t)w * N);
                    out.push_back(tiles.ptr(w));
                }
            }
            if (!src.empty()) {
                // producer: may overwrite slots released two groups ago -> wait for that group's consumer
                if (overlap && gi >= 2) CB_CUDA(cudaStreamWaitEvent(sf, ctx->event(NG + gi - 2), 0));
                int nl = 0;
                CB_CUDA(launch_fft_packed(sf, SRC_BITS, src.data(), out.data(), (int)src.size(), n, ctx->tw, scratch_p, hkp, &nl, 0,
                                          pass == 2 && half_mode));
                ctx->launches += nl;
                ctx->fft_words += src.size();
                if (pass == 2 && half_mode) ctx->fft_words_half += src.size();
                if (overlap) {
                    CB_CUDA(cudaEventRecord(ctx->event(gi), sf));
                    CB_CUDA(cudaStreamWaitEvent(st, ctx->event(gi), 0));
                }
            }
            for (auto& c : g.comb) tiles.acquire(c.res, false, pass);
            consume(gi, g);
            if (overlap || ov2) CB_CUDA(cudaEventRecord(ctx->event(NG + gi), st));
            for (int w : g.free_after) tiles.release(w);
        }
        for (int w = 0; w < N_WORDS; w++) tiles.release(w);
        if (overlap) CB_CUDA(cudaStreamWaitEvent(sf, ctx->event(2 * NG - 1), 0));  // next pass's producer starts after this pass
        if (ov2) {
            CB_CUDA(cudaStreamWaitEvent(sp, ctx->event(2 * NG - 1), 0));
            CB_CUDA(cudaStreamWaitEvent(sc, ctx->event(2 * NG - 1), 0));
        }
    };

    // ---- tree 1 (pass 1): LDE tiles in column order -> Blake2s leaf states -> Merkle tree.  Sharded mode: every rank builds
    //      the subtree over its Mr leaves; rank 0 collects the layers, adds the top log2(G) layers and broadcasts the root.
    DBuf<uint32_t> d_rowN(ctx, half_mode ? (size_t)D.cols : 1);
    DevMerkle tree1;
    tree1.log_leaves = m;
    if (R == 0) tree1.nodes = DBuf<uint32_t>(ctx, (((size_t)2 << m) - 1) * 8);
    {
        DevMerkle local;
        local.log_leaves = lr;
        DBuf<uint32_t> local_nodes;
        if (G > 1) local_nodes = DBuf<uint32_t>(ctx, (((size_t)2 << lr) - 1) * 8);
        uint32_t* ln = G > 1 ? local_nodes.p : tree1.nodes.p;  // unsharded: the local subtree IS the tree
        DBuf<uint32_t> hstate(ctx, 8 * Mr);
        uint64_t bytes_before = 0;
        if (small_mode) {
            ctx->stage_begin("fft_small");
            CB_CUDA(launch_fft_packed_list(st, cdev.indep, D.indep, W.p, N, arena_p, tile_words, n, ctx->tw));
            CB_CUDA(launch_sum_tiles(st, arena_p, tile_words, M, cdev.combs, cdev.n_combs));
            ctx->stage_end();
            ctx->stage_begin("trace_merkle_leaves");
            CB_CUDA(launch_merkle_leaves_seq(st, arena_p, tile_words, D.words, m, ln));
            ctx->stage_end();
            ctx->launches += 3;
            ctx->fft_words = D.indep;
        } else
        run_pass(1, [&](size_t gi, const Group& g) {
            LeafGroups lg{};
            lg.n = (int)g.hash.size();
            for (int i = 0; i < lg.n; i++) {
                lg.g[i] = {tiles.ptr(g.hash[i]), Mr, 32, lr, nullptr, nullptr, nullptr};
                for (auto& c : g.comb)
                    if (c.res == g.hash[i]) lg.g[i] = {tiles.ptr(c.a), Mr, 32, lr, tiles.ptr(c.b), tiles.ptr(c.c), tiles.ptr(c.res)};
            }
            ctx->stage_begin("trace_merkle_leaves");
            CB_CUDA(launch_merkle_leaves(st, lg, lr, hstate.p, bytes_before, gi == 0, gi + 1 == plan.size(), ln));
            ctx->stage_end();
            ctx->launches++;
            bytes_before += 128ull * g.hash.size();
            if (half_mode && R == 0) {  // every column's value at storage row N (the one row of the second half the composition needs)
                TileRowJobs tj{};
                for (int w : g.hash) { tj.word[tj.n] = w; tj.tile[tj.n] = tiles.ptr(w); tj.n++; }
                CB_CUDA(launch_gather_tile_row(st, tj, Mr, G > 1 ? Mv : N, d_rowN.p));  // global row N = local row Mv of rank 0
                ctx->launches++;
            }
        });
        ctx->stage_begin("merkle_nodes");
        // local subtrees: one over Mr leaves, or (virtual shards) two over Mv leaves each, whose roots are the two nodes of layer lv
        const int local_top = p2p ? lv : lr;
        if (G == 1 && lr <= 11) {
            CB_CUDA(launch_merkle_tree_small(st, ln, lr));
            ctx->launches++;
        } else
        for (int l = 0; l < local_top; l++) {
            CB_CUDA(launch_merkle_nodes(st, ln + local.layer_offset(l) * 8, 1u << (lr - l - 1), ln + local.layer_offset(l + 1) * 8));
            ctx->launches++;
        }
        if (G > 1) {
            comm_group_start();
            for (int l = 0; l <= local_top; l++) {
                // a rank's layer l = `parts` runs of cnt words; run i of rank r sits at position (i G + r) of the global layer
                const int parts = p2p ? 2 : 1;
                const size_t cnt = ((size_t)(p2p ? Mv : Mr) >> l) * 8;
                for (int i = 0; i < parts; i++) {
                    const uint32_t* mine = ln + local.layer_offset(l) * 8 + (size_t)i * cnt;
                    uint32_t* glob = R == 0 ? tree1.nodes.p + tree1.layer_offset(l) * 8 + (size_t)i * G * cnt : nullptr;
                    if (R == 0) {
                        CB_CUDA(cudaMemcpyAsync(glob, mine, cnt * 4, cudaMemcpyDeviceToDevice, st));
                        for (int r = 1; r < G; r++) comm_recv_u32(cm, glob + (size_t)r * cnt, cnt, r, st);
                    } else {
                        comm_send_u32(cm, mine, cnt, 0, st);
                    }
                }
            }
            comm_group_end();
            DBuf<uint32_t> d_root(ctx, 8);
            if (R == 0) {
                for (int l = local_top; l < m; l++) {
                    CB_CUDA(launch_merkle_nodes(st, tree1.nodes.p + tree1.layer_offset(l) * 8, 1u << (m - l - 1),
                                                tree1.nodes.p + tree1.layer_offset(l + 1) * 8));
                    ctx->launches++;
                }
                CB_CUDA(cudaMemcpyAsync(d_root.p, tree1.nodes.p + tree1.layer_offset(m) * 8, 32, cudaMemcpyDeviceToDevice, st));
            }
            comm_group_start();
            if (R == 0) for (int r = 1; r < G; r++) comm_send_u32(cm, d_root.p, 8, r, st);
            else comm_recv_u32(cm, d_root.p, 8, 0, st);
            comm_group_end();
            CB_CUDA(cudaMemcpyAsync(tree1.root.b, d_root.p, 32, cudaMemcpyDeviceToHost, st));
            ctx->sync();
        } else {
            CB_CUDA(cudaMemcpyAsync(tree1.root.b, tree1.nodes.p + tree1.layer_offset(m) * 8, 32, cudaMemcpyDeviceToHost, st));
        }
        ctx->stage_end();
    }
    // ---- statement (the two public-input hashes were computed on host threads while the GPU ran pass 1)
    std::vector<uint8_t> stmt;
    host::put_u32(stmt, (uint32_t)log_size);
    if (!block_air) {
    host::put_bytes(stmt, opt.stmt_nonce ? opt.stmt_nonce : nonce, 12);
    host::put_u32(stmt, counter);
    {
        const auto t0 = std::chrono::steady_clock::now();
        hashers.join();
        ctx->hash_wait_us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    }
    if (!hash_err.empty()) throw CbError("public-input hashing: " + hash_err);
    if (opt.pt_hash) {
        memcpy(pth.b, opt.pt_hash, 32);
        memcpy(cth.b, opt.ct_hash, 32);
    } else if (opt.empty_public_hashes) {
        pth = host::blake2s_bytes(nullptr, 0);
        cth = pth;
    }
    if (G > 1) {
        DBuf<uint32_t> d_h(ctx, 16);
        uint8_t hb[64];
        memcpy(hb, pth.b, 32);
        memcpy(hb + 32, cth.b, 32);
        if (R == 0) CB_CUDA(cudaMemcpyAsync(d_h.p, hb, 64, cudaMemcpyHostToDevice, st));
        comm_bcast_u32(cm, d_h.p, 16, 0, st);
        CB_CUDA(cudaMemcpyAsync(hb, d_h.p, 64, cudaMemcpyDeviceToHost, st));
        ctx->sync();
        memcpy(pth.b, hb, 32);
        memcpy(cth.b, hb + 32, 32);
    }
    host::put_bytes(stmt, pth.b, 32);
    host::put_bytes(stmt, cth.b, 32);
    }
    ch.mix_u64((uint64_t)log_size);  // BitwiseStatement::mix_into (bitwise/air.rs:45-47) stops here
    if (!block_air) {
        for (int i = 0; i < 3; i++) ch.mix_u64(host::load_le32(&stmt[4 + 4 * i]));
        ch.mix_u64(counter);
        for (int i = 0; i < 16; i++) ch.mix_u64(host::load_le32(&stmt[20 + 4 * i]));
    }

    ctx->sync();
    roots.push_back(tree1.root);
    ch.mix_root(tree1.root);

    // ---- composition polynomial (pass 2): constraint quotients accumulated tile by tile
    QM31 random_coeff = ch.draw_secure_felt();
    DBuf<uint32_t> apr(ctx, (size_t)D.cons * 4), d_den(ctx, (size_t)1 << cfg.log_blowup), acc(ctx, R == 0 ? 4 * M : 4);
    DBuf<uint32_t> acc_local;
    if (G > 1) acc_local = DBuf<uint32_t>(ctx, 4 * Mr);
    uint32_t* accp = G > 1 ? acc_local.p : acc.p;  // this rank's rows of the 4 accumulator columns
    static const ConsTable ctabs[2] = {build_cons_table(plans[0]), build_cons_table(plans[1])};
    const ConsTable& ctab = ctabs[variant];
    static const bool cons_v1 = getenv("S2C_CONS_V1") != nullptr;  // A/B switch: the integer (IMAD.WIDE) accumulation
    DBuf<uint32_t> apr_lo(ctx, cons_v1 ? (size_t)D.cons * 4 : 4), apr_hi(ctx, cons_v1 ? (size_t)D.cons * 4 : 4);
    DBuf<double> gtab(ctx, ctab.idx.size() * 8);
    if (!cons_v1 && !ctx->chacha_cidx[variant]) {  // consumption-order index list of the alpha table: static, uploaded once per context
        CB_CUDA(cudaMalloc(&ctx->chacha_cidx[variant], ctab.idx.size() * sizeof(int)));
        CB_CUDA(cudaMemcpy(ctx->chacha_cidx[variant], ctab.idx.data(), ctab.idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    CB_CUDA(launch_secure_powers_rev(st, random_coeff, D.cons, apr.p));
    // the constraint sum at storage row N (half-domain evaluation below): one block over the AIR's constraint table
    DBuf<uint32_t> d_qrow(ctx, 8);
    if (half_mode && R == 0) {
        CB_CUDA(launch_mask_constraints(st, cdev.table, D.cons, d_rowN.p, 1, apr.p, d_qrow.p));
        ctx->launches++;
    }
    if (cons_v1) {
        CB_CUDA(launch_split16(st, apr.p, D.cons, apr_lo.p, apr_hi.p));
    } else {
        CB_CUDA(launch_cons_table(st, apr.p, (const int*)ctx->chacha_cidx[variant], (int)ctab.idx.size(), gtab.p));
    }
    ctx->launches += 2;
    std::vector<uint32_t> den((size_t)1 << cfg.log_blowup);  // 1 / Z_H on the two halves of the (bit-reversed) evaluation domain
    for (uint32_t i = 0; i < den.size(); i++) {
        uint32_t row = i << n;
        host::Pt p = host::index_to_point(host::canonic_index_at(m, host::bit_reverse(row, m)));
        den[i] = inv(host::coset_vanishing_m31(n, p));
    }
    CB_CUDA(cudaMemcpyAsync(d_den.p, den.data(), den.size() * 4, cudaMemcpyHostToDevice, st));
    ctx->sync();
    // Half-domain evaluation.  The constraints have degree 2, so the quotient q = C / Z_H lies in the (N+1)-dimensional space
    // {p_left + c * Z_H}: N coefficients of the log-n circle basis plus ONE constant c in front of Z_H = pi^(n-1)(x) (the
    // reference's "right half" of the composition polynomial is that constant).  Storage rows [0, N) of the evaluation domain
    // form a circle domain of log size n on which Z_H is constant (= v): q restricted to it is p_left + v c, interpolated with
    // the shifted twiddle tower (host::make_twiddles(.., true)); one more row (row N, where Z_H = -v) separates c.
    // So only N + 1 of the 2N rows are evaluated; coefficients - and therefore the proof bytes - are unchanged.
    // (Row-sharded mode keeps the full-domain evaluation: the first-half rows live on half of the ranks only.)
    const size_t cons_rows = half_mode ? (G > 1 ? Mv : N) : 0;  // virtual shards: local rows [0, Mv) are first-half rows
    if (small_mode) {
        // the job list of the whole AIR (tile pointers into the arena), kept on the device while the arena does not move
        int n_jobs = 0;
        for (auto& g : plan) n_jobs += (int)g.cons.size();
        if (ctx->small_jobs == nullptr || ctx->small_jobs_arena != (void*)arena_p || ctx->small_jobs_tile_words != tile_words ||
            ctx->small_jobs_variant != variant) {
            std::vector<ConstraintJob> jl;
            for (size_t gi = 0; gi < plan.size(); gi++)
                for (size_t k = 0; k < plan[gi].cons.size(); k++) {
                    const CJ& c = plan[gi].cons[k];
                    auto tp = [&](int w) -> uint32_t* { return w >= 0 ? arena_p + (size_t)w * tile_words : nullptr; };
                    jl.push_back({tp(c.w0), tp(c.w1), tp(c.w2), tp(c.w3), tp(c.w4), tp(c.wres), ctab.job_off[gi][k], c.kb0, c.kb1, c.kb2,
                                  c.kbc, c.arg, c.type});
                }
            if (!ctx->small_jobs) CB_CUDA(cudaMalloc(&ctx->small_jobs, 512 * sizeof(ConstraintJob)));
            CB_CUDA(cudaMemcpyAsync(ctx->small_jobs, jl.data(), jl.size() * sizeof(ConstraintJob), cudaMemcpyHostToDevice, st));
            ctx->sync();
            ctx->small_jobs_variant = variant;
            ctx->small_jobs_arena = arena_p;
            ctx->small_jobs_tile_words = tile_words;
        }
        DBuf<uint32_t> partial(ctx, (size_t)n_jobs * 4 * M);
        ctx->stage_begin("constraints");
        CB_CUDA(launch_constraints_jobs(st, (const ConstraintJob*)ctx->small_jobs, n_jobs, M, gtab.p, partial.p, accp, M));
        ctx->stage_end();
        ctx->launches += 2;
    } else
    run_pass(2, [&](size_t gi, const Group& g) {
        ConstraintJobs cj{};
        for (auto& c : g.cons)
            cj.j[cj.n++] = {c.w0 >= 0 ? tiles.ptr(c.w0) : nullptr, c.w1 >= 0 ? tiles.ptr(c.w1) : nullptr,
                            c.w2 >= 0 ? tiles.ptr(c.w2) : nullptr, c.w3 >= 0 ? tiles.ptr(c.w3) : nullptr,
                            c.w4 >= 0 ? tiles.ptr(c.w4) : nullptr, c.wres >= 0 ? tiles.ptr(c.wres) : nullptr,
                            c.kx, c.kb0, c.kb1, c.kb2, c.kbc, c.arg, c.type};
        if (cj.n == 0) return;
        ctx->stage_begin("constraints");
        if (cons_v1) {
            CB_CUDA(launch_constraints_tiles(st, cj, Mr, apr_lo.p, apr_hi.p, accp, gi == 0, cons_rows));
        } else {
            for (int k = 0; k < cj.n; k++) cj.j[k].kx = ctab.job_off[gi][k];
            CB_CUDA(launch_constraints_tiles2(st, cj, Mr, gtab.p, accp, gi == 0, cons_rows));
        }
        ctx->stage_end();
        ctx->launches++;
    });
    CB_CUDA(launch_scale_rows(st, accp, Mr, n, d_den.p, p2p ? (size_t)R * Mv : (size_t)R * Mr));
    ctx->launches++;
    QM31 q_rowN = qzero();
    if (G > 1) {
        // the row shards of the accumulator go to rank 0, which finishes the proof alone (4-8 columns from here on)
        comm_group_start();
        const size_t rows_each = p2p ? Mv : Mr;  // virtual shards: only the first-half rows were evaluated
        for (int c = 0; c < 4; c++) {
            if (R == 0) {
                CB_CUDA(cudaMemcpyAsync(acc.p + (size_t)c * M, accp + (size_t)c * Mr, rows_each * 4, cudaMemcpyDeviceToDevice, st));
                for (int r = 1; r < G; r++) comm_recv_u32(cm, acc.p + (size_t)c * M + (size_t)r * rows_each, rows_each, r, st);
            } else {
                comm_send_u32(cm, accp + (size_t)c * Mr, rows_each, 0, st);
            }
        }
        comm_group_end();
        ctx->sync();
    }
    // From here on 4-12 columns remain.  Rank 0 ("lead") owns them and the transcript; the other ranks of a sharded proof
    // stay for the three passes that still touch all 33,280 trace columns - out-of-domain samples and the FRI numerator
    // (both passes over the packed witness, split by witness word) and the queried values (read from the row-sharded tiles) -
    // whose disjoint partial results are merged with a sum all-reduce.
    const bool lead = R == 0;

    ctx->stage_begin("composition_commit");
    // interpolate the 4 coordinate columns (log m), split into halves, evaluate each half (log n) on the LDE domain
    DBuf<uint32_t> comp_coef(ctx, lead ? 4 * M : 4), comp_lde(ctx, lead ? 8 * M : 8);
    DevMerkle tree2;
    if (lead) {
        DBuf<uint32_t> scratch4(ctx, 4 * M);
        ColSrc src{SRC_M31, acc.p, M, 0};
        if (half_mode) {
            ctx->ensure_twiddles_shifted(n);
            CB_CUDA(cudaMemsetAsync(comp_coef.p, 0, 4 * M * 4, st));
            CB_CUDA(launch_fft(st, src, 4, n, 0, 1 | 2, comp_coef.p, M, nullptr, 0, ctx->tw_shift, scratch4.p, N));  // E = p_left + v c
            // E at the point of storage row N, and q there
            const host::Pt ps = host::index_to_point(host::canonic_index_at(m, host::bit_reverse((uint32_t)N, m)));
            std::vector<QM31> maps(n);
            maps[0] = qfrom(ps.y);
            uint32_t x = ps.x;
            for (int j = 1; j < n; j++) { maps[j] = qfrom(x); x = sub(mul(2, mul(x, x)), 1); }
            DBuf<uint32_t> bas(ctx, 4 * N), d_e(ctx, 16);
            CB_CUDA(launch_basis(st, bas.p, N, n, maps.data()));
            CB_CUDA(launch_oods_dot(st, comp_coef.p, M, 4, n, bas.p, N, d_e.p));
            uint32_t e16[16], qn[4], c0[4];
            CB_CUDA(cudaMemcpyAsync(e16, d_e.p, sizeof e16, cudaMemcpyDeviceToHost, st));
            CB_CUDA(cudaMemcpyAsync(q_rowN.v, d_qrow.p, 16, cudaMemcpyDeviceToHost, st));
            for (int c = 0; c < 4; c++) CB_CUDA(cudaMemcpyAsync(&c0[c], comp_coef.p + (size_t)c * M, 4, cudaMemcpyDeviceToHost, st));
            ctx->sync();
            for (int c = 0; c < 4; c++) qn[c] = mul(q_rowN.v[c], den[1]);
            const uint32_t v = inv(den[0]), inv2v = inv(add(v, v));
            for (int c = 0; c < 4; c++) {
                const uint32_t cc = mul(sub(e16[4 * c], qn[c]), inv2v);  // E(p*) - q(p*) = 2 v c
                const uint32_t left0 = sub(c0[c], mul(v, cc));
                CB_CUDA(cudaMemcpyAsync(comp_coef.p + (size_t)c * M, &left0, 4, cudaMemcpyHostToDevice, st));
                CB_CUDA(cudaMemcpyAsync(comp_coef.p + (size_t)c * M + N, &cc, 4, cudaMemcpyHostToDevice, st));
            }
            ctx->sync();
            ctx->launches += (n <= 10 ? 1 : n) + 4;
        } else {
            CB_CUDA(launch_fft(st, src, 4, m, 0, 1 | 2, comp_coef.p, M, nullptr, 0, ctx->tw, scratch4.p, M));
        }
        for (int half = 0; half < 2; half++) {
            ColSrc cs{SRC_M31, comp_coef.p + half * N, M, 0};
            CB_CUDA(launch_fft(st, cs, 4, n, cfg.log_blowup, 4, nullptr, 0, comp_lde.p + (size_t)half * 4 * M, M, ctx->tw, nullptr, 0));
        }
        ctx->launches += 6;
        LeafGroups g2{};
        g2.n = 1;
        g2.g[0] = {comp_lde.p, M, 8, m};
        tree2 = build_merkle(ctx, g2, m);
    }
    if (G > 1) {
        DBuf<uint32_t> d_r(ctx, 8);
        if (lead) CB_CUDA(cudaMemcpyAsync(d_r.p, tree2.root.b, 32, cudaMemcpyHostToDevice, st));
        comm_bcast_u32(cm, d_r.p, 8, 0, st);
        CB_CUDA(cudaMemcpyAsync(tree2.root.b, d_r.p, 32, cudaMemcpyDeviceToHost, st));
        ctx->sync();
    }
    ctx->stage_end();
    roots.push_back(tree2.root);
    ch.mix_root(tree2.root);

    // ---- OODS sampling: f_j(z) = 2^-n <bits_j, (FFT with inverse twiddles)(basis(z))>  (kernels_stream.cu fact 2)
    host::CirclePointQ z = host::get_random_point(ch);
    ctx->stage_begin("oods");
    std::vector<QM31> sampled((size_t)D.cols + 8);
    const FftTables tw_t{ctx->tw.IX, ctx->tw.IY, ctx->tw.X, ctx->tw.Y, ctx->tw.max_log};  // transposed inverse transform
    DBuf<uint32_t> basis(ctx, 4 * N), wt(ctx, 4 * N);
    // the 336 adder-sum words are skipped in every pass over the packed witness: s_i = a_i + b_i + c_(i-1) - 2 c_i is an
    // identity of polynomials (kernels_stream.cu fact 1), so sampled values follow from the operands' and the sum columns'
    // quotient coefficients are folded into the operands' coefficients
    std::vector<int> indep_words;
    for (auto& g : plan)
        for (int w : g.fft) indep_words.push_back(w);
    if (G > 1) {  // this rank's share of the witness words
        std::vector<int> mine;
        for (size_t i = 0; i < indep_words.size(); i++)
            if ((int)(i % G) == R) mine.push_back(indep_words[i]);
        indep_words.swap(mine);
    }
    DBuf<int> d_indep_own(ctx, G > 1 ? indep_words.size() : 1);
    if (G > 1) CB_CUDA(cudaMemcpyAsync(d_indep_own.p, indep_words.data(), indep_words.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    struct { const int* p; } d_indep{G > 1 ? d_indep_own.p : cdev.indep};
    DBuf<uint32_t> d_sampled(ctx, ((size_t)D.cols + 8) * 4), d_close(ctx, 4);
    {
        std::vector<QM31> maps(n);
        maps[0] = z.y;
        QM31 x = z.x;
        for (int j = 1; j < n; j++) { maps[j] = x; x = qsub(qmul_m(qmul(x, x), 2), qone()); }
        CB_CUDA(launch_basis(st, basis.p, N, n, maps.data()));
        ColSrc bs{SRC_M31, basis.p, N, 0};
        CB_CUDA(launch_fft(st, bs, 4, n, 0, 4, nullptr, 0, wt.p, N, tw_t, nullptr, 0));
        if (G > 1) CB_CUDA(cudaMemsetAsync(d_sampled.p, 0, ((size_t)D.cols + 8) * 16, st));
        CB_CUDA(launch_bitcol_dot(st, W.p, N, (int)indep_words.size(), wt.p, inv_n, d_sampled.p, d_indep.p));
        if (lead)
            for (int half = 0; half < 2; half++)
                CB_CUDA(launch_oods_dot(st, comp_coef.p + half * N, M, 4, n, basis.p, N, d_sampled.p + ((size_t)D.cols + 4 * half) * 4));
        if (G > 1) comm_allreduce_sum_u32(cm, d_sampled.p, ((size_t)D.cols + 8) * 4, st);
        CB_CUDA(launch_oods_fill_sums(st, d_sampled.p, cdev.combs, cdev.n_combs));  // the adder-sum words' samples, from their operands'
        ctx->launches += (n <= 10 ? 1 : n) + 7;
        uint32_t* const pin = ctx->pinned_words(sampled.size() * 4);  // read-backs go through the context's pinned buffer
        CB_CUDA(cudaMemcpyAsync(pin, d_sampled.p, sampled.size() * 16, cudaMemcpyDeviceToHost, st));
        // prove()'s closing check (numerator): the AIR on the sampled mask, read back at the end
        if (lead) {
            CB_CUDA(launch_mask_constraints(st, cdev.table, D.cons, d_sampled.p, 0, apr.p, d_close.p));
            ctx->launches++;
        }
        ctx->sync();
        memcpy(sampled.data(), pin, sampled.size() * 16);
    }
    ctx->stage_end();
    ch.mix_felts(sampled.data(), sampled.size());

    // ---- FRI quotients: numerator of the 33,280 trace columns = extension of one row-wise combination of the packed
    //      witness (kernels_stream.cu fact 3); the 8 composition columns are read from their LDE
    QM31 rc = ch.draw_secure_felt();
    ctx->stage_begin("quotients");
    DBuf<uint32_t> quot(ctx, lead ? 4 * M : 4);
    {
        // line coefficients of all 33,288 sampled columns, the powers of the random coefficient and the folding of the sum
        // columns' coefficients into their operands': device kernels over the sampled values (kernels_tail.cu)
        const int nc = (int)sampled.size();
        DBuf<uint32_t> d_pw(ctx, (size_t)nc * 4), d_call(ctx, (size_t)nc * 4), d_lin(ctx, 8);
        CB_CUDA(launch_secure_powers_rev(st, rc, nc, d_pw.p));
        CB_CUDA(launch_quot_coefs(st, d_sampled.p, nc, d_pw.p, z.y, d_call.p, d_lin.p));
        DBuf<uint32_t> g(ctx, 4 * N), g_lde(ctx, lead ? 4 * M : 4), d_bc(ctx, 12 * 4);
        const uint32_t ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        CB_CUDA(cudaMemcpyAsync(d_bc.p, ident, sizeof ident, cudaMemcpyHostToDevice, st));
        CB_CUDA(cudaMemcpyAsync(d_bc.p + 16, d_call.p + (size_t)D.cols * 4, 8 * 16, cudaMemcpyDeviceToDevice, st));  // composition columns
        CB_CUDA(launch_quot_fold_sums(st, d_call.p, cdev.combs, cdev.n_combs));
        uint32_t* const d_coefs_p = d_call.p;
        ctx->launches += 3;
        CB_CUDA(launch_bitrow_comb(st, W.p, N, (int)indep_words.size(), d_coefs_p, g.p, d_indep.p));
        if (G > 1) {  // the ranks' partial combinations (disjoint word sets) are added on the lead
            DBuf<uint32_t> parts(ctx, lead ? (size_t)G * 4 * N : 4);
            comm_group_start();
            if (lead) {
                CB_CUDA(cudaMemcpyAsync(parts.p, g.p, 4 * N * 4, cudaMemcpyDeviceToDevice, st));
                for (int r = 1; r < G; r++) comm_recv_u32(cm, parts.p + (size_t)r * 4 * N, 4 * N, r, st);
            } else {
                comm_send_u32(cm, g.p, 4 * N, 0, st);
            }
            comm_group_end();
            if (lead) CB_CUDA(launch_bitrow_reduce(st, parts.p, N, G, g.p));
            ctx->launches++;
        }
        if (lead) {
        ColSrc gs{SRC_M31, g.p, N, 0};
        CB_CUDA(launch_fft(st, gs, 4, n, cfg.log_blowup, 1 | 4, nullptr, 0, g_lde.p, M, ctx->tw, g.p, N));
        QuotBatch qb{};
        qb.prx = {z.x.v[0], z.x.v[1]}; qb.pix = {z.x.v[2], z.x.v[3]};
        qb.pry = {z.y.v[0], z.y.v[1]}; qb.piy = {z.y.v[2], z.y.v[3]};
        qb.lin_a = qzero(); qb.lin_b = qzero(); qb.batch_coeff = qzero();  // lin_a / lin_b: filled on the device below
        qb.coefs = d_bc.p; qb.col_idx = nullptr; qb.n_cols = 12;
        DBuf<QuotBatch> d_qb(ctx, 1);
        CB_CUDA(cudaMemcpyAsync(d_qb.p, &qb, sizeof(qb), cudaMemcpyHostToDevice, st));
        static_assert(offsetof(QuotBatch, lin_b) == offsetof(QuotBatch, lin_a) + 16, "lin_a, lin_b contiguous");
        CB_CUDA(cudaMemcpyAsync((char*)d_qb.p + offsetof(QuotBatch, lin_a), d_lin.p, 32, cudaMemcpyDeviceToDevice, st));
        CB_CUDA(launch_quotients(st, g_lde.p, M, 4, comp_lde.p, M, d_qb.p, 1, m, ctx->tw, quot.p, M));
        ctx->launches += 5;
        }
        ctx->sync();
    }
    ctx->stage_end();

    FriProverState fri;
    uint64_t pow_nonce = 0;
    std::vector<uint32_t> queries;
    if (lead) {
        // ---- FRI commit
        ctx->stage_begin("fri_commit");
        fri = fri_commit(ctx, ch, cfg, std::move(quot), m);
        ctx->stage_end();

        // ---- proof of work, queries
        ctx->stage_begin("grind");
        pow_nonce = grind(ctx, ch, cfg.pow_bits);
        ctx->stage_end();
        ch.mix_u64(pow_nonce);
        queries = host::queries_generate(ch, m, cfg.n_queries);
    }
    if (G > 1) {  // the query positions go to every rank
        std::vector<uint32_t> qb(cfg.n_queries + 1, 0);
        if (lead) {
            qb[0] = (uint32_t)queries.size();
            for (size_t i = 0; i < queries.size(); i++) qb[1 + i] = queries[i];
        }
        DBuf<uint32_t> d_qs(ctx, qb.size());
        if (lead) CB_CUDA(cudaMemcpyAsync(d_qs.p, qb.data(), qb.size() * 4, cudaMemcpyHostToDevice, st));
        comm_bcast_u32(cm, d_qs.p, qb.size(), 0, st);
        CB_CUDA(cudaMemcpyAsync(qb.data(), d_qs.p, qb.size() * 4, cudaMemcpyDeviceToHost, st));
        ctx->sync();
        if (!lead) queries.assign(qb.begin() + 1, qb.begin() + 1 + qb[0]);
    }

    // ---- decommit: queried LDE values of the trace columns.  Tiles that stayed cached since the commitment pass are read
    //      directly; adder-sum words follow from their operands (kernels_stream.cu fact 1, which holds row by row on the
    //      extended domain); only the independent words without a cached tile are evaluated from the packed witness like the
    //      OODS samples (all independent words when the tiles are row-sharded over several ranks).
    ctx->stage_begin("decommit");
    std::vector<uint8_t> fri_bytes;
    std::vector<Hash32> dec1, dec2;
    if (lead) {
        fri_bytes = fri_decommit(ctx, fri, cfg, queries);
        dec1 = merkle_decommit(ctx, tree1, queries);
        dec2 = merkle_decommit(ctx, tree2, queries);
    }
    const int nq = (int)queries.size();
    // (rank, local row) holding global LDE row q of the row-sharded tiles
    auto locate = [&](uint32_t q, int& owner, uint32_t& local) {
        if (G == 1) { owner = 0; local = q; }
        else if (p2p) { const uint32_t v = q >> lv; owner = (int)(v & (uint32_t)(G - 1)); local = ((v >> logG) << lv) | (q & (uint32_t)(Mv - 1)); }
        else { owner = (int)(q >> lr); local = q & (uint32_t)(Mr - 1); }
    };
    std::vector<uint32_t> qv1((size_t)D.cols * nq), qv2((size_t)8 * nq);
    {
        std::vector<int> slot(N_WORDS, -1);
        std::vector<char> indep(N_WORDS, 0);
        for (auto& g : plan)
            for (int w : g.fft) indep[w] = 1;
        std::vector<int> need;  // independent words whose tile is not cached (all of them when the tiles are row-sharded)
        for (int w = 0; w < N_WORDS; w++) {
            if (!indep[w]) continue;
            if (tiles.cache_slot[w] >= 0) slot[w] = tiles.cache_slot[w];
            else need.push_back(w);
        }
        DBuf<int> d_slot(ctx, N_WORDS), d_need(ctx, need.size() + 1);
        CB_CUDA(cudaMemcpyAsync(d_slot.p, slot.data(), N_WORDS * sizeof(int), cudaMemcpyHostToDevice, st));
        CB_CUDA(cudaMemcpyAsync(d_need.p, need.data(), need.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        DBuf<uint32_t> d_rows(ctx, nq), d_q2(ctx, qv2.size()), d_q1(ctx, (size_t)D.cols * 4);
        std::vector<uint32_t> q4((size_t)D.cols * 4);
        for (int q0 = 0; q0 < nq; q0 += 4) {
            const int nqc = nq - q0 < 4 ? nq - q0 : 4;
            if (G > 1) CB_CUDA(cudaMemsetAsync(d_q1.p, 0, (size_t)D.cols * 16, st));
            if (!need.empty() && lead) {
                uint32_t init[4] = {0, 0, 0, 0};
                std::vector<std::array<uint32_t, 4>> maps(n);
                for (int c = 0; c < nqc; c++) {
                    host::Pt p = host::index_to_point(host::canonic_index_at(m, host::bit_reverse(queries[q0 + c], m)));
                    init[c] = 1;
                    maps[0][c] = p.y;
                    uint32_t x = p.x;
                    for (int j = 1; j < n; j++) { maps[j][c] = x; x = sub(mul(2, mul(x, x)), 1); }
                }
                for (int c = nqc; c < 4; c++)
                    for (int j = 0; j < n; j++) maps[j][c] = 0;
                CB_CUDA(launch_basis4(st, basis.p, N, n, init, (const uint32_t(*)[4])maps.data()));
                ColSrc bs{SRC_M31, basis.p, N, 0};
                CB_CUDA(launch_fft(st, bs, 4, n, 0, 4, nullptr, 0, wt.p, N, tw_t, nullptr, 0));
                CB_CUDA(launch_bitcol_dot(st, W.p, N, (int)need.size(), wt.p, inv_n, d_q1.p, d_need.p));
                ctx->launches += (n <= 10 ? 1 : n) + 3;
            }
            if (need.size() < (size_t)D.indep) {
                uint32_t rows4[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};  // 0xffffffff: row held by another rank
                for (int c = 0; c < nqc; c++) {
                    int owner;
                    uint32_t local;
                    locate(queries[q0 + c], owner, local);
                    if (owner == R) rows4[c] = local;
                }
                CB_CUDA(launch_gather_cached(st, tiles.arena, tiles.tile_words, Mr, d_slot.p, N_WORDS, rows4, nqc, d_q1.p));
                ctx->launches++;
            }
            if (G > 1) comm_allreduce_sum_u32(cm, d_q1.p, (size_t)D.cols * 4, st);
            if (!lead) continue;
            uint32_t* const pin = ctx->pinned_words(q4.size());
            CB_CUDA(cudaMemcpyAsync(pin, d_q1.p, q4.size() * 4, cudaMemcpyDeviceToHost, st));
            ctx->sync();
            memcpy(q4.data(), pin, q4.size() * 4);
            for (auto& g : plan)
                for (auto& cb : g.comb)
                    for (int c = 0; c < nqc; c++) {
                        uint32_t cin = 0;
                        for (int i = 0; i < 32; i++) {
                            const uint32_t av = q4[((size_t)cb.a * 32 + i) * 4 + c], bv = q4[((size_t)cb.b * 32 + i) * 4 + c],
                                           cv = q4[((size_t)cb.c * 32 + i) * 4 + c];
                            q4[((size_t)cb.res * 32 + i) * 4 + c] = sub(add(add(av, bv), cin), add(cv, cv));
                            cin = cv;
                        }
                    }
            for (int j = 0; j < D.cols; j++)
                for (int c = 0; c < nqc; c++) qv1[(size_t)j * nq + q0 + c] = q4[(size_t)j * 4 + c];
        }
        if (!lead) {
            ctx->sync();
            ctx->stage_end();
            proof.clear();
            ctx->collect_stages();
            return "";
        }
        CB_CUDA(cudaMemcpyAsync(d_rows.p, queries.data(), nq * 4, cudaMemcpyHostToDevice, st));
        CB_CUDA(launch_gather_rows(st, comp_lde.p, M, 8, d_rows.p, nq, d_q2.p));
        ctx->launches++;
        CB_CUDA(cudaMemcpyAsync(qv2.data(), d_q2.p, qv2.size() * 4, cudaMemcpyDeviceToHost, st));
        ctx->sync();
    }
    ctx->stage_end();

    // ---- prove()'s closing check: composition OODS value == constraints on the sampled mask / Z_H(z)
    {
        QM31 num;
        CB_CUDA(cudaMemcpyAsync(num.v, d_close.p, 16, cudaMemcpyDeviceToHost, st));
        ctx->sync();
        QM31 zh = coset_vanishing_q(n, z);
        QM31 expect = qmul(num, qinv(zh));
        const QM31 units[4] = {{{1, 0, 0, 0}}, {{0, 1, 0, 0}}, {{0, 0, 1, 0}}, {{0, 0, 0, 1}}};
        QM31 left = qzero(), right = qzero();
        for (int k = 0; k < 4; k++) {
            left = qadd(left, qmul(sampled[D.cols + k], units[k]));
            right = qadd(right, qmul(sampled[D.cols + 4 + k], units[k]));
        }
        QM31 pix = z.x;
        for (int i = 0; i < n - 1; i++) pix = qsub(qmul_m(qmul(pix, pix), 2), qone());
        if (!qeq(qadd(left, qmul(pix, right)), expect)) return "Proof generation failed: ConstraintsNotSatisfied";
    }

    // ---- serialise: StreamProof{stmt, StarkProof(CommitmentSchemeProof{config, commitments, sampled_values, decommitments,
    //                 queried_values, proof_of_work, fri_proof})}
    proof.clear();
    proof.reserve(stmt.size() + 64 + sampled.size() * 24 + (size_t)D.cols * (8 + 4 * nq) + 4096);
    host::put_bytes(proof, stmt.data(), stmt.size());
    cfg.serialize(proof);
    host::put_u64(proof, roots.size());
    for (auto& r : roots) host::put_bytes(proof, r.b, 32);
    host::put_u64(proof, 3);
    host::put_u64(proof, 0);
    auto bulk = [&](size_t bytes) {  // appends `bytes` and returns where to write them
        const size_t o = proof.size();
        proof.resize(o + bytes);
        return proof.data() + o;
    };
    host::put_u64(proof, D.cols);
    {
        uint8_t* w = bulk((size_t)D.cols * 24);
        const uint64_t one = 1;
        for (int j = 0; j < D.cols; j++, w += 24) { memcpy(w, &one, 8); memcpy(w + 8, sampled[j].v, 16); }
    }
    host::put_u64(proof, 8);
    for (int j = 0; j < 8; j++) { host::put_u64(proof, 1); host::put_qm31(proof, sampled[D.cols + j]); }
    host::put_u64(proof, 3);
    host::put_u64(proof, 0);
    host::put_u64(proof, dec1.size());
    for (auto& h : dec1) host::put_bytes(proof, h.b, 32);
    host::put_u64(proof, dec2.size());
    for (auto& h : dec2) host::put_bytes(proof, h.b, 32);
    host::put_u64(proof, 3);
    host::put_u64(proof, 0);
    host::put_u64(proof, D.cols);
    {
        const size_t each = 8 + 4 * (size_t)nq;
        uint8_t* w = bulk((size_t)D.cols * each);
        const uint64_t cnt = (uint64_t)nq;
        for (int j = 0; j < D.cols; j++, w += each) { memcpy(w, &cnt, 8); memcpy(w + 8, &qv1[(size_t)j * nq], 4 * (size_t)nq); }
    }
    host::put_u64(proof, 8);
    for (int j = 0; j < 8; j++) { host::put_u64(proof, nq); host::put_bytes(proof, &qv2[(size_t)j * nq], 4 * nq); }
    host::put_u64(proof, pow_nonce);
    host::put_bytes(proof, fri_bytes.data(), fri_bytes.size());
    ctx->collect_stages();
    return "";
}
