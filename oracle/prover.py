"""CPU restatement (numpy) of upstream stwo's `prove` flow as the reference drives it, plus the reference's proof
containers.  TEST INFRASTRUCTURE ONLY.

Restates (un-vendored stwo rev f117d487, module paths as embedded in the reference's WASM build):
  prover/mod.rs `prove`, prover/pcs/mod.rs `CommitmentSchemeProver::{tree_builder,commit,prove_values}`,
  prover/pcs/quotient_ops.rs + core/pcs/quotients.rs, prover/fri.rs `FriProver::{commit,decommit}`,
  core/fri.rs, core/queries.rs, prover/air/accumulation.rs, constraint-framework prover/component_prover.rs.
Reference call sites: /root/reference/stwo/src/chacha/bitwise/air_stream.rs:160-234 (prove_stream_internal),
containers air_stream.rs:30-131, serialisation wasm_api.rs:588-601 (bincode 1.3 defaults + base64).
Pinned byte-for-byte against oracle/_ref (tests/test_oracle.py, tests/test_oracle_aes.py).
"""
import struct
import numpy as np

from stwo_core import (P, U64, QM31, Blake2sChannel, MerkleTree, canonic_domain, Coset, circle_ifft, circle_fft,
                       eval_at_point, get_random_point, bit_reverse_indices, bit_reverse_index, blake2s,
                       m_add, m_sub, m_mul, m_neg, m_inv, q_mul, q_add, q_sub, q_inv, q_mul_m31, q_from_m31,
                       index_to_point, subgroup_gen, double_x, cm_mul)


class PcsConfig:
    """PcsConfig::default() at the pinned rev as serialised by the reference: pow_bits=10,
    FriConfig{log_blowup_factor=1, log_last_layer_degree_bound=0, n_queries=3, fold_step=1}, trailing Option = None."""
    pow_bits = 10
    log_blowup_factor = 1
    log_last_layer_degree_bound = 0
    n_queries = 3
    fold_step = 1

    def serialize(self):
        return struct.pack("<IIIQI", self.pow_bits, self.log_blowup_factor, self.log_last_layer_degree_bound,
                           self.n_queries, self.fold_step) + b"\x00"


# ------------------------------------------------------------------------------------------------ commitment scheme
class CommitmentTree:
    """CommitmentTreeProver: polys (coefficients), their evaluations on CanonicCoset(log+blowup), lifted Merkle tree."""

    def __init__(self, coeffs_list, log_blowup):
        self.coeffs = coeffs_list                              # list of 1-D arrays (sizes may differ)
        self.evals = [circle_fft(c, (len(c).bit_length() - 1) + log_blowup) for c in coeffs_list]
        self.tree = MerkleTree(self.evals)

    @staticmethod
    def from_matrix(coeffs, log_blowup):
        """Fast path for equal-size columns given as [C, N]."""
        t = CommitmentTree.__new__(CommitmentTree)
        ev = circle_fft(coeffs, (coeffs.shape[1].bit_length() - 1) + log_blowup)
        t.coeffs = [coeffs[j] for j in range(coeffs.shape[0])]
        t.evals = [ev[j] for j in range(ev.shape[0])]
        t.tree = MerkleTree(t.evals)
        return t


class CommitmentSchemeProver:
    def __init__(self, config):
        self.config = config
        self.trees = []

    def commit_evals(self, evals_matrix_or_list, channel):
        """TreeBuilder::extend_evals + commit: interpolate, LDE, Merkle, mix_root."""
        if isinstance(evals_matrix_or_list, np.ndarray):
            tree = CommitmentTree.from_matrix(circle_ifft(evals_matrix_or_list), self.config.log_blowup_factor)
        else:
            tree = CommitmentTree([circle_ifft(e) for e in evals_matrix_or_list], self.config.log_blowup_factor)
        self.trees.append(tree)
        channel.mix_root(tree.tree.root())
        return tree

    def commit_polys(self, coeffs_list, channel):
        tree = CommitmentTree(coeffs_list, self.config.log_blowup_factor)
        self.trees.append(tree)
        channel.mix_root(tree.tree.root())
        return tree


# ------------------------------------------------------------------------------------------------ composition
def secure_powers(alpha, n):
    """generate_secure_powers: [1, a, a^2, ...] as array [n,4]."""
    out = np.empty((n, 4), dtype=U64)
    cur = QM31(1)
    for i in range(n):
        out[i] = cur.v
        cur = cur * alpha
    return out


def coset_vanishing_on_domain(trace_log, eval_log):
    """core/constraints.rs coset_vanishing(CanonicCoset(trace_log).coset, p) for every p of
    CanonicCoset(eval_log).circle_domain() in storage (bit-reversed) order."""
    dom = canonic_domain(eval_log)
    xs, ys = dom.points_bitrev()
    coset = Coset.odds(trace_log)
    # p - initial + step/2
    t = index_to_point((-coset.initial_index + (coset.step_size >> 1)) & ((1 << 31) - 1))
    tx, ty = U64(t[0]), U64(t[1])
    x = m_sub(m_mul(xs, tx), m_mul(ys, ty))
    for _ in range(1, trace_log):
        x = m_sub(m_mul(U64(2), m_mul(x, x)), U64(1))
    return x


def finalize_composition(acc, eval_log):
    """DomainEvaluationAccumulator::finalize for a single log size + the split into two half-degree polys:
    returns 8 coefficient vectors of length 2^(eval_log-1): [left c0..c3, right c0..c3]."""
    coeffs = circle_ifft(np.ascontiguousarray(acc.T))           # [4, 2N]
    half = coeffs.shape[1] // 2
    return [coeffs[c, :half].copy() for c in range(4)] + [coeffs[c, half:].copy() for c in range(4)]


# ------------------------------------------------------------------------------------------------ FRI quotients
def fri_quotients(columns, sample_batches, random_coeff, domain_log):
    """core/pcs/quotients.rs accumulate_row_quotients over the whole domain (lifted-protocol form of the pinned rev).
    columns: list of 1-D eval arrays of size 2^domain_log; sample_batches: list of
    (point(x,y QM31), [(col_idx, value QM31, alpha_power QM31)...]) -- every (column, sample) pair carries its own power of
    the random coefficient (build_samples_with_randomness_and_periodicity) and the batches are simply summed:
        q(p) = sum_batches [ sum_j alpha_j (c f_j(p) - a_j p.y - b_j) ] / den_batch(p).
    Pinned against the reference binary: the NumeratorData / PointSampleWithRandomness tables of an AES-CTR proof were read
    out of its linear memory (oracle/dev notes in DESIGN.md).  `random_coeff` is unused (kept for the call signature).
    Returns [R,4]."""
    dom = canonic_domain(domain_log)
    xs, ys = dom.points_bitrev()
    R = len(xs)
    row_acc = np.zeros((R, 4), dtype=U64)
    for (px, py), cav in sample_batches:
        num = np.zeros((R, 4), dtype=U64)
        lin_a = QM31(0)
        lin_b = QM31(0)
        c_coefs = np.empty((len(cav), 4), dtype=U64)
        c = py.conj() - py
        for k, (ci, val, alpha) in enumerate(cav):
            a = val.conj() - val
            b = val * c - a * py
            lin_a = lin_a + alpha * a
            lin_b = lin_b + alpha * b
            c_coefs[k] = (alpha * c).v
        # numerator = sum_k (alpha c)_k * f_k(row) - (sum alpha a) * y - sum alpha b
        colmat = np.stack([_lift(columns[ci], domain_log) for ci, _, _ in cav], axis=0)          # [k, R]
        for co in range(4):
            tot = np.zeros(R, dtype=U64)
            ck = c_coefs[:, co]
            for s in range(0, len(cav), 4096):
                tot = (tot + ((colmat[s:s + 4096] * ck[s:s + 4096, None]) % U64(P)).sum(axis=0)) % U64(P)
            num[:, co] = tot
        lin = q_add(q_mul_m31(np.broadcast_to(lin_a.arr(), (R, 4)), ys), lin_b.arr())
        num = q_sub(num, lin)
        # denominator (CM31): (Prx - x) * Piy - (Pry - y) * Pix
        prx, pix = (px.v[0], px.v[1]), (px.v[2], px.v[3])
        pry, piy = (py.v[0], py.v[1]), (py.v[2], py.v[3])
        d0a, d1a = cm_mul(m_sub(U64(prx[0]), xs), np.full(R, prx[1], dtype=U64), U64(piy[0]), U64(piy[1]))
        d0b, d1b = cm_mul(m_sub(U64(pry[0]), ys), np.full(R, pry[1], dtype=U64), U64(pix[0]), U64(pix[1]))
        d0, d1 = m_sub(d0a, d0b), m_sub(d1a, d1b)
        nrm = m_inv(m_add(m_mul(d0, d0), m_mul(d1, d1)))
        i0, i1 = m_mul(d0, nrm), m_mul(m_neg(d1), nrm)
        # numerator.mul_cm31(den_inv)
        n0, n1 = cm_mul(num[:, 0], num[:, 1], i0, i1)
        n2, n3 = cm_mul(num[:, 2], num[:, 3], i0, i1)
        row_acc = q_add(row_acc, np.stack([n0, n1, n2, n3], axis=-1))
    return row_acc


def _lift(col, lifting_log):
    """Column of a smaller domain seen on the lifting domain (vcs_lifted index map)."""
    lg = len(col).bit_length() - 1
    if lg == lifting_log:
        return col
    from stwo_core import lifted_index
    idx = np.arange(1 << lifting_log)
    sh = lifting_log - lg
    return np.asarray(col)[((idx >> (sh + 1)) << 1) + (idx & 1)]


def build_sample_batches(sample_points_per_tree, sampled, random_coeff, col_lifts=None):
    """build_samples_with_randomness_and_periodicity + ColumnSampleBatch::new_vec: walk the columns of all trees in order;
    every sample gets the next power of the random coefficient (alpha^0 first); a column with more than one sample (a mask
    with a non-zero offset) first gets an extra "periodicity" copy of its offset-0 sample (the last of its samples), which
    takes a power of its own; samples are then grouped by point."""
    batches = {}
    alpha = QM31(1)
    ci = 0
    for ti, (pts, tv) in enumerate(zip(sample_points_per_tree, sampled)):
        for cj, (plist, vals) in enumerate(zip(pts, tv)):
            entries = list(zip(plist, vals))
            if len(entries) > 1:
                # periodicity sample: the lifted column g(pi^k(p)) has period h_k (the point of order 2^k), so its value at
                # z + h_k must equal its value at z; for k = 0 this is a second copy of the offset-0 sample
                k_lift = col_lifts[ti][cj] if col_lifts else 0
                (zx, zy), zval = entries[-1]
                if k_lift > 0:
                    from stwo_core import index_to_point
                    hx, hy = index_to_point(1 << (31 - k_lift))
                    zx, zy = zx * hx - zy * hy, zx * hy + zy * hx
                entries = [((zx, zy), zval)] + entries
            for (pt, val) in entries:
                key = (tuple(pt[0].v), tuple(pt[1].v))
                batches.setdefault(key, (pt, []))[1].append((ci, val, alpha))
                alpha = alpha * random_coeff
            ci += 1
    return list(batches.values())


# ------------------------------------------------------------------------------------------------ FRI
def _ibutterfly_q(v0, v1, itw):
    return q_add(v0, v1), q_mul_m31(q_sub(v0, v1), itw)


def fold_circle_into_line(dst, src, alpha, domain_log):
    """cpu/fri.rs fold_circle_into_line: src [2M,4] on canonic circle domain, dst [M,4]."""
    dom = canonic_domain(domain_log)
    xs, ys = dom.points_bitrev()
    f0, f1 = _ibutterfly_q(src[0::2], src[1::2], m_inv(ys[0::2]))
    fp = q_add(q_mul(f1, alpha.arr()), f0)
    return q_add(q_mul(dst, (alpha * alpha).arr()), fp)


def line_domain_xs_bitrev(coset):
    xs, _ = coset.points()
    return xs[bit_reverse_indices(coset.log_size)]


def fold_line(ev, alpha, coset):
    """cpu/fri.rs fold_line: ev [M,4] on LineDomain(coset) in bit-reversed order -> [M/2,4] on coset.double()."""
    xs = line_domain_xs_bitrev(coset)
    f0, f1 = _ibutterfly_q(ev[0::2], ev[1::2], m_inv(xs[0::2]))
    return q_add(f0, q_mul(f1, alpha.arr()))


def line_interpolate(ev, coset):
    """LineEvaluation::interpolate (prover/line.rs): bit-reversed evals -> coefficients, then ordered."""
    n = ev.shape[0]
    vals = ev[bit_reverse_indices(coset.log_size)].copy()        # natural order
    c = coset
    size = n
    while size > 1:
        xs, _ = c.points()
        itw = m_inv(xs[:size // 2])
        v = vals.reshape(n // size, size, 4)
        l, r = v[:, :size // 2].copy(), v[:, size // 2:].copy()
        v[:, :size // 2] = q_add(l, r)
        v[:, size // 2:] = q_mul_m31(q_sub(l, r), itw[None, :])
        c = c.double()
        size //= 2
    vals = q_mul_m31(vals, np.full(n, pow(n, P - 2, P), dtype=U64))
    return vals[bit_reverse_indices(coset.log_size)]              # into_ordered_coefficients


def queries_generate(channel, log_domain_size, n_queries):
    """core/queries.rs Queries::generate."""
    qs = set()
    cnt = 0
    mask = (1 << log_domain_size) - 1
    while True:
        for w in channel.draw_u32s():
            qs.add(w & mask)
            cnt += 1
            if cnt == n_queries:
                return sorted(qs)


def fold_positions(pos, n_folds):
    return sorted(set(p >> n_folds for p in pos))


def decommit_positions_and_witness(column, query_positions, fold_step):
    """prover/fri.rs compute_decommitment_positions_and_witness_evals."""
    positions = []
    witness = []
    qs = set(query_positions)
    for start in sorted(set((q >> fold_step) << fold_step for q in query_positions)):
        for pos in range(start, start + (1 << fold_step)):
            positions.append(pos)
            if pos not in qs:
                witness.append(column[pos])
    return positions, witness


class FriLayer:
    def __init__(self, evaluation):
        self.evaluation = evaluation                                   # [M,4]
        self.tree = MerkleTree([evaluation[:, c].copy() for c in range(4)])

    def decommit(self, queries, fold_step=1):
        positions, witness = decommit_positions_and_witness(self.evaluation, queries, fold_step)
        return witness, self.tree.decommit(positions), self.tree.root()


def fri_commit(channel, config, quotient_eval, domain_log):
    first = FriLayer(quotient_eval)
    channel.mix_root(first.tree.root())
    circle_alpha = channel.draw_secure_felt()
    line_log = domain_log - 1
    coset = Coset.half_odds(line_log)
    layer_eval = np.zeros((1 << line_log, 4), dtype=U64)
    layer_eval = fold_circle_into_line(layer_eval, quotient_eval, circle_alpha, domain_log)
    inner = []
    last_size = 1 << (config.log_last_layer_degree_bound + config.log_blowup_factor)
    while layer_eval.shape[0] > last_size:
        layer = FriLayer(layer_eval)
        channel.mix_root(layer.tree.root())
        alpha = channel.draw_secure_felt()
        layer_eval = fold_line(layer_eval, alpha, coset)
        coset = coset.double()
        inner.append(layer)
    coeffs = line_interpolate(layer_eval, coset)
    bound = 1 << config.log_last_layer_degree_bound
    import os
    if not os.environ.get("ORACLE_DEV_SKIP_DEGREE_CHECK"):
        assert not coeffs[bound:].any(), "invalid degree"
    last_poly = coeffs[:bound]
    channel.mix_felts(last_poly)
    return first, inner, last_poly


# ------------------------------------------------------------------------------------------------ serialisation
def ser_hashes(hs):
    return struct.pack("<Q", len(hs)) + b"".join(hs)


def ser_qm31_vec(arr):
    arr = np.asarray(arr, dtype=U64).reshape(-1, 4)
    return struct.pack("<Q", arr.shape[0]) + arr.astype("<u4").tobytes()


def ser_m31_vec(arr):
    arr = np.asarray(arr, dtype=U64).reshape(-1)
    return struct.pack("<Q", arr.shape[0]) + arr.astype("<u4").tobytes()


def prove_values(scheme, sample_points_per_tree, channel, lifting_log):
    """CommitmentSchemeProver::prove_values.  sample_points_per_tree[t][col] = list of (x,y) QM31 points.
    Returns the serialised CommitmentSchemeProof (bincode) and a dict of intermediates."""
    cfg = scheme.config
    # sampled values
    sampled = []
    for tree, pts in zip(scheme.trees, sample_points_per_tree):
        tv = [[None] * len(plist) for plist in pts]
        # batch columns that share (size, point): one vectorised eval_at_point per group
        groups = {}
        for j, (coeffs, plist) in enumerate(zip(tree.coeffs, pts)):
            # points are given on the lifting domain; a column whose LDE is 2^k times smaller is the lift g(pi^k(p)) of its
            # polynomial g, so g is evaluated at the k-fold doubled point
            k_lift = lifting_log - ((len(coeffs).bit_length() - 1) + cfg.log_blowup_factor)
            for k, (px, py) in enumerate(plist):
                for _ in range(k_lift):
                    px, py = px * px * 2 - 1, px * py * 2
                groups.setdefault((len(coeffs), px.v, py.v), []).append((j, k, px, py))
        for (_, _, _), members in groups.items():
            px, py = members[0][2], members[0][3]
            mat = np.stack([tree.coeffs[j] for (j, _, _, _) in members], axis=0)
            res = eval_at_point(mat, px, py)
            for (j, k, _, _), v in zip(members, res):
                tv[j][k] = QM31(int(v[0]), int(v[1]), int(v[2]), int(v[3]))
        sampled.append(tv)
    flat = [v for tv in sampled for cv in tv for v in cv]
    channel.mix_felts(flat)
    random_coeff = channel.draw_secure_felt()
    columns = [e for tree in scheme.trees for e in tree.evals]
    col_lifts = [[lifting_log - ((len(c).bit_length() - 1) + cfg.log_blowup_factor) for c in tree.coeffs] for tree in scheme.trees]
    sample_batches = build_sample_batches(sample_points_per_tree, sampled, random_coeff, col_lifts)
    quot = fri_quotients(columns, sample_batches, random_coeff, lifting_log)
    first, inner, last_poly = fri_commit(channel, cfg, quot, lifting_log)
    nonce = channel.grind(cfg.pow_bits)
    channel.mix_u64(nonce)
    queries = queries_generate(channel, lifting_log, cfg.n_queries)
    # FRI decommit
    fw, fd, fc = first.decommit(queries)
    fri_bytes = ser_qm31_vec(fw) if fw else struct.pack("<Q", 0)
    fri_bytes += ser_hashes(fd) + fc
    lq = fold_positions(queries, 1)
    inner_bytes = struct.pack("<Q", len(inner))
    for layer in inner:
        w, dcm, com = layer.decommit(lq)
        inner_bytes += (ser_qm31_vec(w) if w else struct.pack("<Q", 0)) + ser_hashes(dcm) + com
        lq = fold_positions(lq, 1)
    fri_bytes += inner_bytes + ser_qm31_vec(last_poly) + struct.pack("<I", cfg.log_last_layer_degree_bound)
    # tree decommitments
    decomm = struct.pack("<Q", len(scheme.trees))
    qvals = struct.pack("<Q", len(scheme.trees))
    for tree in scheme.trees:
        if tree.evals:
            from stwo_core import lifted_index as _li
            tlog = tree.tree.lifting_log            # a tree whose largest column is smaller than the lifting domain
            decomm += ser_hashes(tree.tree.decommit(sorted(set(_li(q, lifting_log, tlog) for q in queries))))
        else:
            decomm += ser_hashes([])
        qvals += struct.pack("<Q", len(tree.evals))
        for e in tree.evals:
            lg = len(e).bit_length() - 1
            from stwo_core import lifted_index
            qvals += ser_m31_vec([e[lifted_index(q, lifting_log, lg)] for q in queries])
    out = cfg.serialize()
    out += ser_hashes([t.tree.root() for t in scheme.trees])
    sv = struct.pack("<Q", len(sampled))
    for tv in sampled:
        sv += struct.pack("<Q", len(tv))
        for cv in tv:
            sv += struct.pack("<Q", len(cv)) + b"".join(struct.pack("<4I", *v.v) for v in cv)
    out += sv + decomm + qvals + struct.pack("<Q", nonce) + fri_bytes
    return out, dict(sampled=sampled, random_coeff=random_coeff, quot=quot, nonce=nonce, queries=queries)
