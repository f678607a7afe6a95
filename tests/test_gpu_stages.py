"""GPU parity tests (-m gpu) for the rest of the backend-trait boundary (include/s2c_b200.h, csrc/cb_api_ext.cu): ColumnOps /
FieldOps helpers, barycentric evaluation, coset twiddles, legacy commit_on_layer, multi-batch quotient accumulation and every
AES-CTR AIR stage (trace, LogUp interaction trace with the device-side finalize_last, both components' constraint quotients,
lift_and_accumulate) against the CPU oracle (oracle/stwo_core.py, oracle/prover.py, oracle/aes_air.py), bit-exact."""
import ctypes
import hashlib

import numpy as np
import pytest

import aes_air as aa
import aes_api
import prover as op
import stwo_core as sc
from make_golden_aes import aes_case_inputs

pytestmark = pytest.mark.gpu
U32P = ctypes.POINTER(ctypes.c_uint32)
P = sc.P


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def hp(a):
    return a.ctypes.data_as(U32P)


def q4(q):
    return (ctypes.c_uint32 * 4)(*[int(x) for x in (q.v if hasattr(q, "v") else q)])


def ptr_at(base, words):
    return ctypes.c_void_p(base.value + 4 * int(words))


@pytest.mark.parametrize("log_n", [0, 1, 5, 12, 17])
def test_bit_reverse_and_column_access(backend, log_n):
    be = backend
    rng = np.random.default_rng(log_n)
    v = rng.integers(0, P, size=1 << log_n, dtype=np.uint64).astype(np.uint32)
    d = be.upload(v)
    be._ck(be.L.cb_bit_reverse(be.ctx, d, log_n))
    assert np.array_equal(be.download(d, v.shape), v[sc.bit_reverse_indices(log_n)])
    idx = (1 << log_n) - 1
    got = ctypes.c_uint32()
    be._ck(be.L.cb_col_at(be.ctx, d, ctypes.c_size_t(idx), ctypes.byref(got)))
    assert got.value == v[sc.bit_reverse_indices(log_n)][idx]
    be._ck(be.L.cb_col_set(be.ctx, d, ctypes.c_size_t(idx), ctypes.c_uint32(12345)))
    assert be.download(d, v.shape)[idx] == 12345
    assert be.L.cb_col_set(be.ctx, d, ctypes.c_size_t(0), ctypes.c_uint32(P)) != 0   # not a canonical M31 word
    be.free(d)


def test_batch_inverse(backend):
    be = backend
    rng = np.random.default_rng(7)
    n = 5000
    v = rng.integers(1, P, size=n, dtype=np.uint64)
    d = be.upload(v)
    o = be.malloc(n * 4)
    be._ck(be.L.cb_batch_inverse_m31(be.ctx, d, o, ctypes.c_size_t(n)))
    assert np.array_equal(be.download(o, (n,)).astype(np.uint64), sc.m_inv(v))
    q = rng.integers(0, P, size=(4, n), dtype=np.uint64)
    dq = be.upload(q)
    oq = be.malloc(4 * n * 4)
    be._ck(be.L.cb_batch_inverse_qm31(be.ctx, dq, ctypes.c_size_t(n), oq, ctypes.c_size_t(n), ctypes.c_size_t(n)))
    assert np.array_equal(be.download(oq, (4, n)).astype(np.uint64).T, aa.q_batch_inv(q.T.copy()))
    for p in (d, o, dq, oq):
        be.free(p)


@pytest.mark.parametrize("log_n", [4, 9, 13])
def test_extend_and_barycentric_eval(backend, log_n):
    be = backend
    rng = np.random.default_rng(40 + log_n)
    n, ncols = 1 << log_n, 6
    ev = rng.integers(0, P, size=(ncols, n), dtype=np.uint64)
    coef = sc.circle_ifft(ev)
    d_c = be.upload(coef)
    d_x = be.malloc(ncols * 4 * n * 4)
    be._ck(be.L.cb_extend(be.ctx, d_c, ctypes.c_size_t(n), ncols, log_n, 2, d_x, ctypes.c_size_t(4 * n)))
    ext = be.download(d_x, (ncols, 4 * n))
    assert np.array_equal(ext[:, :n], coef) and not ext[:, n:].any()
    z = sc.get_random_point(sc.Blake2sChannel())
    pt8 = (ctypes.c_uint32 * 8)(*(z[0].v + z[1].v))
    d_w = be.malloc(4 * n * 4)
    be._ck(be.L.cb_barycentric_weights(be.ctx, log_n, pt8, d_w))
    d_e = be.upload(ev)
    got = np.empty((ncols, 4), dtype=np.uint32)
    be._ck(be.L.cb_barycentric_eval_at_point(be.ctx, d_e, ctypes.c_size_t(n), ncols, log_n, d_w, hp(got)))
    assert np.array_equal(got, sc.eval_at_point(coef, z[0], z[1]))
    for p in (d_c, d_x, d_w, d_e):
        be.free(p)


@pytest.mark.parametrize("log_n", [3, 8, 14])
def test_precompute_twiddles_for_a_coset(backend, log_n):
    """The half coset of the canonic domain of log size log_n + 1: layers 1.. of the oracle's twiddle tower."""
    be = backend
    tws, itws = sc._layer_twiddles(log_n + 1)
    coset = sc.canonic_domain(log_n + 1).half_coset
    want = np.concatenate([np.asarray(t, dtype=np.uint64) for t in tws[1:]] + [np.array([1], dtype=np.uint64)])
    iwant = np.concatenate([np.asarray(t, dtype=np.uint64) for t in itws[1:]] + [np.array([1], dtype=np.uint64)])
    n = 1 << log_n
    assert len(want) == n
    d_t, d_i = be.malloc(n * 4), be.malloc(n * 4)
    be._ck(be.L.cb_precompute_twiddles_coset(be.ctx, ctypes.c_uint32(coset.initial_index), log_n, d_t, d_i))
    assert np.array_equal(be.download(d_t, (n,)), want)
    assert np.array_equal(be.download(d_i, (n,)), iwant)
    be.free(d_t)
    be.free(d_i)


@pytest.mark.parametrize("with_prev,ncols", [(False, 3), (True, 0), (True, 21), (False, 16), (True, 48)])
def test_commit_on_layer(backend, with_prev, ncols):
    be = backend
    rng = np.random.default_rng(ncols + 100 * with_prev)
    log_n = 6
    n = 1 << log_n
    cols = rng.integers(0, P, size=(max(ncols, 1), n), dtype=np.uint64).astype(np.uint32)
    prev = rng.integers(0, 1 << 32, size=(2 * n, 8), dtype=np.uint64).astype(np.uint32)
    d_cols = be.upload(cols)
    d_prev = be.upload(prev)
    ptrs = (ctypes.c_void_p * max(ncols, 1))(*[d_cols.value + 4 * n * j for j in range(max(ncols, 1))])
    d_out = be.malloc(n * 32)
    be._ck(be.L.cb_commit_on_layer(be.ctx, log_n, d_prev if with_prev else None, ptrs, ncols, d_out))
    got = be.download(d_out, (n, 8))
    for i in (0, 1, n // 2, n - 1):
        msg = (prev[2 * i].tobytes() + prev[2 * i + 1].tobytes() if with_prev else b"") + cols[:ncols, i].tobytes()
        assert got[i].tobytes() == hashlib.blake2s(msg).digest(), i
    for p in (d_cols, d_prev, d_out):
        be.free(p)


def test_accumulate_quotients_batches_with_lifted_columns(backend):
    """Two sample points, columns of two sizes, every (column, sample) pair with its own power of the random coefficient."""
    be = backend
    rng = np.random.default_rng(5)
    m = 9
    big = rng.integers(0, P, size=(5, 1 << m), dtype=np.uint64)
    small = rng.integers(0, P, size=(2, 1 << (m - 2)), dtype=np.uint64)
    columns = [big[j] for j in range(5)] + [small[j] for j in range(2)]
    ch = sc.Blake2sChannel()
    z1 = sc.get_random_point(ch)
    ch.mix_u64(3)
    z2 = sc.get_random_point(ch)
    rc = sc.QM31(3, 1, 4, 1)
    entries = [(z1, [0, 1, 2, 5, 6, 4]), (z2, [3, 4, 6])]
    batches, k = [], 0
    for pt, cis in entries:
        cav = []
        for ci in cis:
            val = sc.QM31(*[int(x) for x in rng.integers(0, P, size=4)])
            cav.append((ci, val, rc ** k))
            k += 1
        batches.append(((pt[0], pt[1]), cav))
    want = op.fri_quotients(columns, batches, rc, m)
    d_big, d_small = be.upload(big), be.upload(small)
    ptrs = (ctypes.c_void_p * 7)(*([d_big.value + 4 * (1 << m) * j for j in range(5)] + [d_small.value + 4 * (1 << (m - 2)) * j for j in range(2)]))
    logs = (ctypes.c_int * 7)(*([m] * 5 + [m - 2] * 2))
    pts = u32([w for (px, py), _ in batches for w in (px.v + py.v)])
    offs = (ctypes.c_int * 3)(0, 6, 9)
    ecol = (ctypes.c_int * 9)(*[ci for _, cav in batches for ci, _, _ in cav])
    evals = u32([w for _, cav in batches for _, v, _ in cav for w in v.v])
    ealpha = u32([w for _, cav in batches for _, _, a in cav for w in a.v])
    d_out = be.malloc(4 * (1 << m) * 4)
    be._ck(be.L.cb_accumulate_quotients_batches(be.ctx, ptrs, logs, 7, m, 2, hp(pts), offs, ecol, hp(evals), hp(ealpha), d_out,
                                                ctypes.c_size_t(1 << m)))
    assert np.array_equal(be.download(d_out, (4, 1 << m)).astype(np.uint64), want.T)
    for p in (d_big, d_small, d_out):
        be.free(p)


def test_accumulate_and_lift_and_accumulate(backend):
    be = backend
    rng = np.random.default_rng(11)
    big = rng.integers(0, P, size=(4, 1 << 11), dtype=np.uint64)
    small = rng.integers(0, P, size=(4, 1 << 9), dtype=np.uint64)
    d_b, d_s, d_b2 = be.upload(big), be.upload(small), be.upload(big)
    be._ck(be.L.cb_lift_and_accumulate(be.ctx, d_b, ctypes.c_size_t(1 << 11), 11, d_s, 9))
    want = np.stack([(big[c] + op._lift(small[c], 11)) % P for c in range(4)])
    assert np.array_equal(be.download(d_b, (4, 1 << 11)).astype(np.uint64), want)
    be._ck(be.L.cb_accumulate(be.ctx, d_b2, d_b, ctypes.c_size_t(4 << 11)))
    assert np.array_equal(be.download(d_b2, (4, 1 << 11)).astype(np.uint64), (big + want) % P)
    for p in (d_b, d_s, d_b2):
        be.free(p)


@pytest.mark.parametrize("log_n", [0, 3, 8, 11, 13])
def test_logup_finalize_last(backend, log_n):
    """Device-side LogupTraceGenerator::finalize_last (prefix sum in coset order) against the oracle's."""
    be = backend
    rng = np.random.default_rng(70 + log_n)
    n = 1 << log_n
    col = rng.integers(0, P, size=(n, 4), dtype=np.uint64)
    want_cols, want_sum = aa.logup_finalize_last([col.copy()], log_n)
    d = be.upload(col.T.copy())
    claimed = (ctypes.c_uint32 * 4)()
    be._ck(be.L.cb_logup_finalize_last(be.ctx, d, ctypes.c_size_t(n), log_n, claimed))
    assert list(claimed) == [int(x) for x in want_sum.v]
    assert np.array_equal(be.download(d, (4, n)).astype(np.uint64), np.stack(want_cols))
    be.free(d)


@pytest.mark.parametrize("key_len,nb,seed", [(16, 5, 1), (32, 40, 2), (16, 300, 3)])
def test_aes_ctr_stages(backend, key_len, nb, seed):
    """cb_gen_trace_aes_ctr -> cb_gen_logup_interaction_aes_ctr -> LDE -> cb_eval_constraints_aes_ctr / _sbox_table ->
    cb_lift_and_accumulate, each stage compared with oracle/aes_air.py (log 8: one size; 300 blocks = log 9: lifted table)."""
    be = backend
    key, nonce, counter, pt, ct = aes_case_inputs(key_len, nb, seed)
    log, nonce_rows, counters, PT, CT = aes_api.build_aes_inputs(key, nonce, counter, pt, ct)
    trace, lookups, mults, valid = aa.generate_ctr_trace(log, key, nonce_rows, counters, PT, CT)
    assert valid
    n, m = 1 << log, 2 << log
    nc, nk, nl = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert be.L.cb_aes_ctr_layout(key_len, ctypes.byref(nc), ctypes.byref(nk), ctypes.byref(nl), None, None) == 0
    C, K, NL = nc.value, nk.value, nl.value
    assert (C, K, NL) == (aa.n_cols(key_len), aa.n_constraints(key_len), aa.n_lookups(key_len)) and trace.shape[0] == C
    lk_in, lk_out = (ctypes.c_int * NL)(), (ctypes.c_int * NL)()
    be.L.cb_aes_ctr_layout(key_len, None, None, None, lk_in, lk_out)
    for k in (0, 1, NL - 1):   # lookup k reads the (input, output) trace columns the oracle recorded
        assert np.array_equal(trace[lk_in[k]], lookups[k, 0]) and np.array_equal(trace[lk_out[k]], lookups[k, 1])
    # ---- trace + multiplicities
    d_t = be.malloc(C * n * 4)
    gm = (ctypes.c_uint32 * 256)()
    ok = ctypes.c_int()
    be._ck(be.L.cb_gen_trace_aes_ctr(be.ctx, key_len, key, nonce, ctypes.c_uint32(counter), pt, ct, ctypes.c_uint32(nb), log, d_t,
                                     ctypes.c_size_t(n), gm, ctypes.byref(ok)))
    assert ok.value == 1
    assert np.array_equal(be.download(d_t, (C, n)).astype(np.uint64), trace)
    assert [int(x) for x in gm] == [int(x) for x in mults]
    bad = bytearray(ct); bad[5] ^= 4
    be._ck(be.L.cb_gen_trace_aes_ctr(be.ctx, key_len, key, nonce, ctypes.c_uint32(counter), pt, bytes(bad), ctypes.c_uint32(nb), log, d_t,
                                     ctypes.c_size_t(n), None, ctypes.byref(ok)))
    assert ok.value == 0
    be._ck(be.L.cb_gen_trace_aes_ctr(be.ctx, key_len, key, nonce, ctypes.c_uint32(counter), pt, ct, ctypes.c_uint32(nb), log, d_t,
                                     ctypes.c_size_t(n), None, ctypes.byref(ok)))
    # ---- interaction trace (device-side finalize_last) and claimed sum
    elems = aa.SboxElements(sc.QM31(11, 22, 33, 44), sc.QM31(5, 6, 7, 8))
    icols, csum = aa.ctr_interaction_trace(log, lookups, elems)
    tcols, tsum = aa.table_interaction_trace(mults, elems)
    NI = len(icols)
    assert NI == 4 * (NL // 2)
    d_i = be.malloc(NI * n * 4)
    got_sum = (ctypes.c_uint32 * 4)()
    be._ck(be.L.cb_gen_logup_interaction_aes_ctr(be.ctx, key_len, d_t, ctypes.c_size_t(n), log, q4(elems.z), q4(elems.alpha), d_i,
                                                 ctypes.c_size_t(n), got_sum))
    assert list(got_sum) == [int(x) for x in csum.v]
    assert np.array_equal(be.download(d_i, (NI, n)).astype(np.uint64), np.stack(icols))
    # ---- constraint quotients of both components on their evaluation domains
    main_lde = sc.circle_fft(sc.circle_ifft(trace), log + 1)
    inter_lde = sc.circle_fft(sc.circle_ifft(np.stack(icols)), log + 1)
    tab = aa.sbox_table_columns()
    pre_lde = sc.circle_fft(sc.circle_ifft(np.stack([tab[0], tab[1]]).astype(np.uint64)), 9)
    mult_lde = sc.circle_fft(sc.circle_ifft(np.asarray(mults, dtype=np.uint64)[None, :]), 9)[0]
    tinter_lde = sc.circle_fft(sc.circle_ifft(np.stack(tcols)), 9)
    rc = sc.QM31(1234567, 7654321, 1111, 2222)
    apr = op.secure_powers(rc, K + 1)[::-1].copy()
    prev = aa.prev_row_index(log, log + 1)
    acc = aa.evaluate_ctr_constraints(main_lde, inter_lde, inter_lde[NI - 4:][:, prev], elems, csum, log, apr[:K], key_len)
    acc = sc.q_mul_m31(acc, sc.m_inv(op.coset_vanishing_on_domain(log, log + 1)))
    prev8 = aa.prev_row_index(8, 9)
    acc_t = aa.evaluate_table_constraint(pre_lde, mult_lde, tinter_lde, tinter_lde[:, prev8], elems, tsum, apr[K])
    acc_t = sc.q_mul_m31(acc_t, sc.m_inv(op.coset_vanishing_on_domain(8, 9)))
    d_ml, d_il = be.upload(main_lde), be.upload(inter_lde)
    d_apr = be.upload(apr)
    d_acc = be.malloc(4 * m * 4)
    be._ck(be.L.cb_eval_constraints_aes_ctr(be.ctx, key_len, d_ml, ctypes.c_size_t(m), d_il, ctypes.c_size_t(m), log, d_apr, q4(elems.z),
                                            q4(elems.alpha), q4(csum), d_acc, ctypes.c_size_t(m)))
    assert np.array_equal(be.download(d_acc, (4, m)).astype(np.uint64), acc.T)
    d_pre, d_mu, d_ti = be.upload(pre_lde), be.upload(mult_lde), be.upload(tinter_lde)
    d_acct = be.malloc(4 * 512 * 4)
    be._ck(be.L.cb_eval_constraints_sbox_table(be.ctx, d_pre, ptr_at(d_pre, 512), d_mu, d_ti, ctypes.c_size_t(512), q4(elems.z),
                                               q4(elems.alpha), q4(tsum), q4(apr[K]), d_acct))
    assert np.array_equal(be.download(d_acct, (4, 512)).astype(np.uint64), acc_t.T)
    # ---- DomainEvaluationAccumulator::finalize: the table component's accumulation lifted onto the larger domain
    be._ck(be.L.cb_lift_and_accumulate(be.ctx, d_acc, ctypes.c_size_t(m), log + 1, d_acct, 9))
    want = np.stack([(acc[:, c] + op._lift(acc_t[:, c], log + 1)) % P for c in range(4)])
    assert np.array_equal(be.download(d_acc, (4, m)).astype(np.uint64), want)
    for p in (d_t, d_i, d_ml, d_il, d_apr, d_acc, d_pre, d_mu, d_ti, d_acct):
        be.free(p)
