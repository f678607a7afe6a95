"""GPU parity tests (-m gpu): every CUDA stage through the C ABI against the CPU oracle on the same inputs (bit-exact:
all arithmetic is M31/QM31 integer), then whole proofs against the oracle, the golden fixtures and (when oracle/_ref
travelled to the box) the reference's own verifier."""
import ctypes
import hashlib
import json
import os
import struct

import numpy as np
import pytest

import api as oracle_api
import chacha_air as ca
import prover as op
import stwo_core as sc
import ref_wasm
from make_golden import case_inputs

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "chacha20_golden.json")))["cases"]
U32P = ctypes.POINTER(ctypes.c_uint32)


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def hp(a):
    return a.ctypes.data_as(U32P)


def q4(q):
    return (ctypes.c_uint32 * 4)(*q.v)


@pytest.mark.parametrize("log_n", [4, 6, 10, 12, 13, 14, 16])
def test_interpolate_evaluate(backend, log_n):
    be = backend
    rng = np.random.default_rng(log_n)
    ncols = 5 if log_n >= 14 else 37
    v = rng.integers(0, sc.P, size=(ncols, 1 << log_n), dtype=np.uint64)
    coef = sc.circle_ifft(v)
    lde = sc.circle_fft(coef, log_n + 1)
    n, m = 1 << log_n, 2 << log_n
    d_v = be.upload(v)
    be._ck(be.L.cb_interpolate_columns(be.ctx, d_v, ctypes.c_size_t(n), ncols, log_n))
    assert np.array_equal(be.download(d_v, (ncols, n)), coef)
    d_e = be.malloc(ncols * m * 4)
    be._ck(be.L.cb_evaluate_polynomials(be.ctx, d_v, ctypes.c_size_t(n), ncols, log_n, 1, d_e, ctypes.c_size_t(m)))
    assert np.array_equal(be.download(d_e, (ncols, m)), lde)
    # evaluate with no extension returns the original values
    d_b = be.malloc(ncols * n * 4)
    be._ck(be.L.cb_evaluate_polynomials(be.ctx, d_v, ctypes.c_size_t(n), ncols, log_n, 0, d_b, ctypes.c_size_t(n)))
    assert np.array_equal(be.download(d_b, (ncols, n)), v)
    # fused commit path from M31 values
    d_v2 = be.upload(v)
    d_c2 = be.malloc(ncols * n * 4)
    d_e2 = be.malloc(ncols * m * 4)
    be._ck(be.L.cb_commit_lde(be.ctx, 0, d_v2, ctypes.c_size_t(n), 0, ncols, log_n, 1, d_c2, ctypes.c_size_t(n), d_e2, ctypes.c_size_t(m)))
    assert np.array_equal(be.download(d_c2, (ncols, n)), coef)
    assert np.array_equal(be.download(d_e2, (ncols, m)), lde)
    for p in (d_v, d_e, d_b, d_v2, d_c2, d_e2):
        be.free(p)


@pytest.mark.parametrize("log_n,kind", [(4, 1), (5, 1), (8, 1), (11, 1), (12, 1), (13, 1), (14, 1), (15, 1), (16, 1), (17, 1), (18, 1), (19, 1), (20, 1),
                                        (22, 1), (6, 2), (12, 2), (15, 2), (5, 0), (12, 0), (14, 0), (17, 0)])
def test_lde_packed(backend, log_n, kind):
    """Packed-witness fused interpolate+extend (the transform of the streaming prover) against the oracle's circle
    iFFT/FFT, for every kernel schedule: whole-column (<=12), three-pass (>=13), padded strided tiles (>=22)."""
    be = backend
    rng = np.random.default_rng(100 + log_n)
    n, m = 1 << log_n, 2 << log_n
    n_words = 3 if log_n <= 16 else 1
    cpj = 32 if kind == 1 else 4
    if kind == 0:   # a job = 4 plain M31 columns
        words = rng.integers(0, sc.P, size=(n_words * 4, n), dtype=np.uint64).astype(np.uint32)
    else:
        words = rng.integers(0, 1 << 32, size=(n_words, n), dtype=np.uint64).astype(np.uint32)
    d_w = be.upload(words)
    d_t = be.malloc(n_words * cpj * m * 4)
    be._ck(be.L.cb_lde_packed(be.ctx, kind, d_w, n_words, log_n, d_t))
    tiles = be.download(d_t, (n_words, cpj, m))
    check_cols = range(cpj) if log_n <= 14 else ([0, 13, 31] if kind == 1 else [0, 3])
    for w in range(n_words):
        wv = words[w].astype(np.uint64) if kind else None
        if kind == 0:
            cols = np.stack([words[4 * w + c].astype(np.uint64) for c in check_cols])
        elif kind == 1:
            cols = np.stack([(wv >> np.uint64(c)) & np.uint64(1) for c in check_cols])
        else:
            cols = np.stack([(wv >> np.uint64(8 * c)) & np.uint64(0xFF) for c in check_cols])
        lde = sc.circle_fft(sc.circle_ifft(cols), log_n + 1)
        for i, c in enumerate(check_cols):
            assert np.array_equal(tiles[w, c], lde[i]), (w, c)
    be.free(d_w)
    be.free(d_t)


def test_lde_packed_generic_kernels(backend):
    """The runtime-schedule three-pass kernels (used above log 20) on a size the specialised kernels normally serve."""
    be = backend
    be.L.cb_debug_force_generic_fft(1)
    try:
        test_lde_packed(be, 14, 1)
        test_lde_packed(be, 17, 1)
    finally:
        be.L.cb_debug_force_generic_fft(0)


def _witness(be, nb, seed):
    key, nonce, counter, pt, ct = case_inputs(nb, seed)
    log, K, NO, C, PT, CT, mrows = oracle_api.build_chacha_inputs(key, nonce, counter, pt, ct)
    trace, valid = ca.generate_stream_trace(log, K, NO, C, PT, CT, mrows)
    n = 1 << log
    d_w = be.malloc(1040 * n * 4)
    ok = ctypes.c_int()
    be._ck(be.L.cb_gen_trace_chacha_stream(be.ctx, key, nonce, ctypes.c_uint32(counter), pt, ct, ctypes.c_uint32(nb), log, d_w,
                                           ctypes.c_size_t(n), ctypes.byref(ok)))
    return (key, nonce, counter, pt, ct), log, trace, valid, d_w, ok.value


@pytest.mark.parametrize("nb,seed", [(1, None), (16, 1), (40, 3), (100, 9)])
def test_chacha_witness(backend, nb, seed):
    be = backend
    _, log, trace, valid, d_w, ok = _witness(be, nb, seed)
    n = 1 << log
    words = be.download(d_w, (1040, n)).astype(np.uint64)
    bits = ((words[:, None, :] >> np.arange(32, dtype=np.uint64)[None, :, None]) & 1).reshape(1040 * 32, n)
    assert np.array_equal(bits, trace)
    assert bool(ok) == valid
    be.free(d_w)


def test_chacha_witness_flags_bad_ciphertext(backend):
    be = backend
    key, nonce, counter, pt, ct = case_inputs(3, 5)
    bad = bytearray(ct); bad[70] ^= 0x10
    ok = ctypes.c_int()
    d_w = be.malloc(1040 * 16 * 4)
    be._ck(be.L.cb_gen_trace_chacha_stream(be.ctx, key, nonce, ctypes.c_uint32(counter), pt, bytes(bad), ctypes.c_uint32(3), 4, d_w,
                                           ctypes.c_size_t(16), ctypes.byref(ok)))
    assert ok.value == 0
    be.free(d_w)


@pytest.mark.parametrize("nb,seed", [(16, 1), (64, 4)])
def test_chacha_commit_constraints_pipeline(backend, nb, seed):
    """packed witness -> LDE (bit expansion fused in the FFT load) -> Merkle root -> constraint quotients."""
    be = backend
    _, log, trace, valid, d_w, ok = _witness(be, nb, seed)
    n, m = 1 << log, 2 << log
    C = ca.N_COLS
    coef = sc.circle_ifft(trace)
    lde = sc.circle_fft(coef, log + 1)
    d_c = be.malloc(C * n * 4)
    d_l = be.malloc(C * m * 4)
    be._ck(be.L.cb_commit_lde(be.ctx, 1, d_w, ctypes.c_size_t(n), 0, C, log, 1, d_c, ctypes.c_size_t(n), d_l, ctypes.c_size_t(m)))
    assert np.array_equal(be.download(d_c, (C, n)), coef)
    assert np.array_equal(be.download(d_l, (C, m)), lde)
    # Merkle: leaves + all layers
    tree = sc.MerkleTree([lde[j] for j in range(C)])
    d_h = be.malloc(m * 32)
    bases = (ctypes.c_void_p * 1)(d_l.value)
    strides = (ctypes.c_size_t * 1)(m)
    ncols = (ctypes.c_int * 1)(C)
    logs = (ctypes.c_int * 1)(log + 1)
    be._ck(be.L.cb_merkle_build_leaves(be.ctx, bases, strides, ncols, logs, 1, log + 1, d_h))
    leaves = be.download(d_h, (m, 8))
    assert leaves.tobytes() == b"".join(tree.layers[0])
    # streaming absorb in 3 tiles gives the same leaves
    d_state = be.malloc(m * 32)
    d_h2 = be.malloc(m * 32)
    tiles = [(0, 1024), (1024, 16384), (16384, C)]
    for i, (c0, c1) in enumerate(tiles):
        ptr = ctypes.c_void_p(d_l.value + c0 * m * 4)
        be._ck(be.L.cb_merkle_leaves_absorb(be.ctx, ptr, ctypes.c_size_t(m), c1 - c0, log + 1, log + 1, d_state,
                                            ctypes.c_uint64(c0 * 4), int(i == 0), int(i == len(tiles) - 1), d_h2))
    assert np.array_equal(be.download(d_h2, (m, 8)), leaves)
    cur, n_cur = d_h, m
    for layer in tree.layers[1:]:
        d_n = be.malloc((n_cur // 2) * 32)
        be._ck(be.L.cb_merkle_next_layer(be.ctx, cur, ctypes.c_uint32(n_cur // 2), d_n))
        assert be.download(d_n, (n_cur // 2, 8)).tobytes() == b"".join(layer)
        cur, n_cur = d_n, n_cur // 2
    # constraints
    alpha = sc.QM31(123456789, 987654321, 55555, 2147483000)
    apr = op.secure_powers(alpha, ca.N_CONSTRAINTS)[::-1].copy()
    d_apr = be.malloc(ca.N_CONSTRAINTS * 16)
    be._ck(be.L.cb_generate_secure_powers_rev(be.ctx, q4(alpha), ca.N_CONSTRAINTS, d_apr))
    assert np.array_equal(be.download(d_apr, (ca.N_CONSTRAINTS, 4)), apr)
    acc = ca.evaluate_constraints(lde, apr)
    acc = sc.q_mul_m31(acc, sc.m_inv(op.coset_vanishing_on_domain(log, log + 1)))
    d_acc = be.malloc(4 * m * 4)
    be._ck(be.L.cb_eval_constraints_chacha_stream(be.ctx, d_l, ctypes.c_size_t(m), log + 1, log, d_apr, d_acc, ctypes.c_size_t(m), 0))
    assert np.array_equal(be.download(d_acc, (4, m)), acc.T)
    # eval_at_point + quotients + folds
    z = sc.get_random_point(sc.Blake2sChannel())
    pt8 = (ctypes.c_uint32 * 8)(*(z[0].v + z[1].v))
    ncheck = 640
    got = np.empty((ncheck, 4), dtype=np.uint32)
    be._ck(be.L.cb_eval_at_point(be.ctx, d_c, ctypes.c_size_t(n), ncheck, log, pt8, hp(got)))
    want = sc.eval_at_point(coef[:ncheck], z[0], z[1])
    assert np.array_equal(got, want)
    rc = sc.QM31(5, 6, 7, 8)
    pw = op.secure_powers(rc, ncheck)  # every (column, sample) pair has its own power of the random coefficient, alpha^0 first
    batches = [((z[0], z[1]), [(j, sc.QM31(*[int(x) for x in want[j]]), sc.QM31(*[int(x) for x in pw[j]])) for j in range(ncheck)])]
    quot = op.fri_quotients([lde[j] for j in range(ncheck)], batches, rc, log + 1)
    d_q = be.malloc(4 * m * 4)
    be._ck(be.L.cb_accumulate_quotients(be.ctx, d_l, ctypes.c_size_t(m), ncheck, log + 1, hp(u32(want)), pt8, q4(rc), d_q, ctypes.c_size_t(m)))
    assert np.array_equal(be.download(d_q, (4, m)), quot.T)
    a1 = sc.QM31(11, 22, 33, 44)
    line = op.fold_circle_into_line(np.zeros((m // 2, 4), dtype=np.uint64), quot, a1, log + 1)
    d_line = be.malloc(4 * (m // 2) * 4)
    be._ck(be.L.cb_fold_circle_into_line(be.ctx, d_q, ctypes.c_size_t(m), log + 1, q4(a1), d_line, ctypes.c_size_t(m // 2), 1))
    assert np.array_equal(be.download(d_line, (4, m // 2)), line.T)
    line2 = op.fold_line(line, a1, sc.Coset.half_odds(log))
    d_line2 = be.malloc(4 * (m // 4) * 4)
    be._ck(be.L.cb_fold_line(be.ctx, d_line, ctypes.c_size_t(m // 2), log, q4(a1), d_line2, ctypes.c_size_t(m // 4)))
    assert np.array_equal(be.download(d_line2, (4, m // 4)), line2.T)
    rows = u32([3, 7, m - 1])
    outv = np.empty((ncheck, 3), dtype=np.uint32)
    be._ck(be.L.cb_gather_rows(be.ctx, d_l, ctypes.c_size_t(m), ncheck, hp(rows), 3, hp(outv)))
    assert np.array_equal(outv, lde[:ncheck][:, rows])


def test_grind_returns_lowest_nonce(backend):
    be = backend
    ch = sc.Blake2sChannel()
    ch.mix_u64(42)
    want = ch.grind(10)
    pd = ch.pow_prefixed_digest(10)
    nonce = ctypes.c_uint64()
    be._ck(be.L.cb_grind_blake2s(be.ctx, pd, 10, ctypes.byref(nonce)))
    assert nonce.value == want


@pytest.mark.parametrize("case", [c for c in GOLDEN], ids=lambda c: c["name"])
def test_proof_bytes_match_golden(backend, case):
    key, nonce, counter, pt, ct = case_inputs(case["n_blocks"], case["seed"])
    res = backend.generate_chacha20_proof(key, nonce, counter, pt, ct)
    if "error" in case:
        assert res == {"error": case["error"]}
        return
    assert res["success"] is True and res["blocks"] == case["blocks"] and res["algorithm"] == "chacha20"
    assert res["proof_size_bytes"] == case["proof_size_bytes"]
    assert hashlib.sha256(res["proof"].encode()).hexdigest() == case["b64_sha256"]


def test_proof_matches_oracle_bytes(backend):
    key, nonce, counter, pt, ct = case_inputs(5, 21)
    want = oracle_api.generate_chacha20_proof(key, nonce, counter, pt, ct)["proof_bytes"]
    got = backend.prove_chacha20_raw(key, nonce, counter, pt, ct)
    assert got == want


@pytest.mark.parametrize("log_size", [4, 6])
def test_prove_stream_testdata_generator_matches_oracle(backend, log_size):
    """s2c_prove_chacha20_stream_testdata = the reference's own `prove_stream` generator (air_stream.rs:237-289: key 00..1f, witness
    nonce words [0, 0x4a, 0], counters r+1, plaintext word 16r+w, statement over an all-zero nonce and empty-string hashes),
    against the oracle's prover on the same inputs; the host verifier accepts it for exactly those public inputs."""
    import struct
    import zk_symmetric_crypto_b200 as z
    n = 1 << log_size
    key = bytes(range(32))
    wnonce = struct.pack("<3I", 0, 0x4A, 0)
    pt = np.arange(16 * n, dtype=np.uint32).reshape(n, 16)
    ks = np.frombuffer(ca.chacha20_keystream_bytes(key, wnonce, 1, n), dtype="<u4").reshape(n, 16)
    ct = pt ^ ks
    K = np.tile(np.frombuffer(key, dtype="<u4").astype(np.uint32), (n, 1))      # per-row arrays (the reference splats them)
    NO = np.tile(np.array([0, 0x4A, 0], dtype=np.uint32), (n, 1))
    C = (1 + np.arange(n)).astype(np.uint32)
    pub = oracle_api.chacha_public_inputs(bytes(12), 1, b"", b"")
    want = oracle_api.prove_stream_internal(log_size, K, NO, C, pt, ct, pub)
    got = backend.prove_chacha20_stream_testdata(log_size)
    assert got == want
    assert z.verify_chacha20_raw(got, bytes(12), 1, b"", b"") == (True, None)
    assert z.verify_chacha20_raw(got, wnonce, 1, b"", b"")[0] is False


def test_proof_independent_of_tile_cache(backend):
    """The streaming prover recomputes LDE tiles that do not fit the cache; the proof must not depend on the cache size."""
    key, nonce, counter, pt, ct = case_inputs(16, 33)
    want = backend.prove_chacha20_raw(key, nonce, counter, pt, ct)
    for cap in (0, 100):
        backend._ck(backend.L.cb_set_max_cached_tiles(backend.ctx, cap))
        try:
            assert backend.prove_chacha20_raw(key, nonce, counter, pt, ct) == want
        finally:
            backend._ck(backend.L.cb_set_max_cached_tiles(backend.ctx, -1))


def test_error_behaviour(backend):
    import zk_symmetric_crypto_b200 as z
    zero = bytes(64)
    assert backend.generate_chacha20_proof(bytes(31), bytes(12), 0, zero, zero) == {"error": "Key must be 32 bytes, got 31"}
    assert backend.generate_chacha20_proof(bytes(32), bytes(12), 0, zero, zero) == \
        {"error": "Ciphertext does not match encryption - invalid witness"}
    with pytest.raises(z.BackendError):
        backend.prove_chacha20_raw(bytes(32), bytes(12), 0, zero, zero)


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not on this box")
@pytest.mark.parametrize("nb", [500, 1024, 4096])  # ceil(nb/16)*16 must fill the trace (reference quirk, gen_stream.rs:240-250)
def test_reference_verifier_accepts_gpu_proofs(backend, nb):
    """Sizes beyond the golden set: the reference's own verifier (wasm_api.rs:609) is the acceptance test, and the
    reference prover must produce the same bytes."""
    key, nonce, counter, pt, ct = case_inputs(nb, 100 + nb)
    res = backend.generate_chacha20_proof(key, nonce, counter, pt, ct)
    assert res.get("success") is True, res
    assert ref_wasm.verify_chacha20_proof(res["proof"], nonce, counter, pt, ct) == {"algorithm": "chacha20", "valid": True}
    bad = bytearray(pt); bad[0] ^= 1
    assert ref_wasm.verify_chacha20_proof(res["proof"], nonce, counter, bytes(bad), ct)["valid"] is False
    if nb in (500, 4096):  # 4096 rows and up take the byte-table / row-sliced kernels of the witness-wide passes (~25 s of CPU)
        assert ref_wasm.generate_chacha20_proof(key, nonce, counter, pt, ct)["proof"] == res["proof"]


# ------------------------------------------------------------------------------ large sizes: every code path of the headline config
_GOLD_DOC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "chacha20_golden.json")))
LARGE_GOLDEN = _GOLD_DOC.get("large_cases", [])


def _with_cache_cap(be, cap, fn):
    be._ck(be.L.cb_set_max_cached_tiles(be.ctx, cap))
    try:
        return fn()
    finally:
        be._ck(be.L.cb_set_max_cached_tiles(be.ctx, -1))


@pytest.mark.parametrize("case", LARGE_GOLDEN, ids=lambda c: c["name"])
def test_large_proofs_match_reference_prover_for_every_cache_size(backend, case):
    """log 12 / log 13 proofs byte-identical to the REFERENCE PROVER's (fixtures from tests/golden/make_golden.py --large), with
    the full tile cache and with caps 0 / 100 / 300: at log 13 that runs the three-pass FFT kernels inside the prover, the
    partial cache with half-tile recomputation in the constraint pass, `gather_cached_kernel` and the uncached-word query path
    -- every path a log 20 proof takes -- and pins them to the reference, not to the path itself."""
    key, nonce, counter, pt, ct = case_inputs(case["n_blocks"], case["seed"])
    for cap in (-1, 0, 100, 300):
        proof = _with_cache_cap(backend, cap, lambda: backend.prove_chacha20_raw(key, nonce, counter, pt, ct))
        assert len(proof) == case["proof_len"], cap
        assert hashlib.sha256(proof).hexdigest() == case["proof_sha256"], "cache cap %d" % cap
        if cap >= 0:
            assert backend.counters()["cached_tiles"] <= cap


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not on this box")
def test_log13_proof_equals_live_reference_prover(backend):
    """The same comparison against the reference prover run live on this box (8,192 blocks: the largest trace its wasm32 build
    holds; ~80 s of CPU), on inputs no fixture was made from."""
    key, nonce, counter, pt, ct = case_inputs(8192, 4242)
    want = ref_wasm.generate_chacha20_proof(key, nonce, counter, pt, ct)
    assert want.get("success") is True, want
    for cap in (-1, 200):
        res = _with_cache_cap(backend, cap, lambda: backend.generate_chacha20_proof(key, nonce, counter, pt, ct))
        assert res == want, "cache cap %d" % cap


@pytest.mark.parametrize("log_n", [14, 16])
def test_partial_cache_paths_equal_full_cache_at_log_14_16(backend, log_n):
    """cb_set_max_cached_tiles in {0, 100, 300} at log 14 / 16 (three-pass FFT with 4 / 16 strided layers): the proof equals the
    full-cache proof, and the reference's verifier accepts it."""
    import bench
    key, nonce, counter, pt, ct = bench.synth_inputs(log_n, 3)
    ptb, ctb = pt.tobytes(), ct.tobytes()
    want = backend.prove_chacha20_raw(key, nonce, counter, ptb, ctb)
    assert backend.counters()["cached_tiles"] == 704
    for cap in (0, 100, 300):
        got = _with_cache_cap(backend, cap, lambda: backend.prove_chacha20_raw(key, nonce, counter, ptb, ctb))
        assert got == want, "cache cap %d" % cap
    import base64
    import zk_symmetric_crypto_b200 as z
    assert z.verify_chacha20_raw(want, nonce, counter, ptb, ctb) == (True, None)
    if ref_wasm.available():
        b64 = base64.b64encode(want).decode()
        assert ref_wasm.verify_chacha20_proof(b64, nonce, counter, ptb, ctb) == {"algorithm": "chacha20", "valid": True}


def test_headline_log20_proof_is_accepted_by_the_reference_verifier(backend):
    """BASELINE configs[1] itself: the log_n_rows = 20 proof bench.py times (same synthetic inputs) is accepted by the
    reference's own verifier (wasm_api.rs:609 -> air_stream.rs:343-421) and rejected for a plaintext with one flipped bit;
    the partial tile cache (642 of 704 tiles fit one B200) is in use, and a smaller cache gives the same bytes."""
    import base64
    import bench
    import zk_symmetric_crypto_b200 as z
    key, nonce, counter, pt, ct = bench.synth_inputs(20, 0)
    ptb, ctb = pt.tobytes(), ct.tobytes()
    be = z.Backend(0)   # its own context: the ~170 GB tile arena is released again when it closes
    try:
        proof = be.prove_chacha20_raw(key, nonce, counter, ptb, ctb)
        cnt = be.counters()
        assert 0 < cnt["cached_tiles"] <= 704
        again = _with_cache_cap(be, 500, lambda: be.prove_chacha20_raw(key, nonce, counter, ptb, ctb))
    finally:
        be.close()
    assert again == proof, "log 20 proof depends on the tile-cache size"
    bad = bytearray(ptb)
    bad[12345] ^= 0x20
    assert z.verify_chacha20_raw(proof, nonce, counter, ptb, ctb) == (True, None)
    assert z.verify_chacha20_raw(proof, nonce, counter, bytes(bad), ctb) == (False, "OodsNotMatching")
    if ref_wasm.available():
        b64 = base64.b64encode(proof).decode()
        assert ref_wasm.verify_chacha20_proof(b64, nonce, counter, ptb, ctb) == {"algorithm": "chacha20", "valid": True}
        assert ref_wasm.verify_chacha20_proof(b64, nonce, counter, bytes(bad), ctb) == {"error": "OodsNotMatching", "valid": False}


# ---------------------------------------------------------------------------------------------------- AES-CTR
import aes_api as oracle_aes
from make_golden_aes import aes_case_inputs
AES_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "aes_ctr_golden.json")))["cases"]


@pytest.mark.parametrize("case", AES_GOLDEN, ids=lambda c: c["name"])
def test_aes_proof_bytes_match_golden(backend, case):
    """AES-128/256-CTR proofs byte-identical to the reference's (fixtures generated by the reference binary): log 8 (all
    columns one size), log 9 / 10 (lifted S-box table columns), invalid witness."""
    key, nonce, counter, pt, ct = aes_case_inputs(case["key_len"], case["n_blocks"], case["seed"], case["corrupt"])
    fn = backend.generate_aes128_ctr_proof if case["key_len"] == 16 else backend.generate_aes256_ctr_proof
    res = fn(key, nonce, counter, pt, ct)
    if "error" in case:
        assert res == {"error": case["error"]}
        return
    assert res["success"] is True and res["blocks"] == case["blocks"] and res["algorithm"] == case["algorithm"]
    assert res["proof_size_bytes"] == case["proof_size_bytes"]
    assert hashlib.sha256(res["proof"].encode()).hexdigest() == case["b64_sha256"]


def test_aes_proof_matches_oracle_bytes(backend):
    key, nonce, counter, pt, ct = aes_case_inputs(16, 7, 77)
    want = oracle_aes.generate_aes128_ctr_proof(key, nonce, counter, pt, ct)["proof_bytes"]
    assert backend.prove_aes_ctr_raw(key, nonce, counter, pt, ct) == want


def test_aes_error_behaviour(backend):
    z = bytes(16)
    assert backend.generate_aes128_ctr_proof(bytes(15), bytes(12), 0, z, z) == {"error": "Key must be 16 bytes, got 15"}
    assert backend.generate_aes256_ctr_proof(bytes(16), bytes(12), 0, z, z) == {"error": "Key must be 32 bytes, got 16"}
    assert backend.generate_aes128_ctr_proof(bytes(16), bytes(12), 0, z, z) == \
        {"error": "Ciphertext does not match encryption - invalid witness"}


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not on this box")
def test_reference_verifier_accepts_gpu_aes_proofs(backend):
    """A size beyond the golden set (log 12, 4,096 blocks): the reference's verifier is the acceptance test."""
    key, nonce, counter, pt, ct = aes_case_inputs(32, 4096, 99)
    res = backend.generate_aes256_ctr_proof(key, nonce, counter, pt, ct)
    assert res.get("success") is True, res
    assert ref_wasm.verify_aes_ctr_proof(res["proof"], nonce, counter, pt, ct) == {"algorithm": "aes256-ctr", "valid": True}
    bad = bytearray(ct); bad[3] ^= 1
    assert ref_wasm.verify_aes_ctr_proof(res["proof"], nonce, counter, pt, bytes(bad))["valid"] is False


def test_pool_of_contexts_gives_identical_proofs(backend):
    """Throughput mode: several contexts on one GPU driven by host threads; every proof equals the single-context proof."""
    from zk_symmetric_crypto_b200.pool import ProverPool
    c = case_inputs(2, 0)
    a = aes_case_inputs(16, 5, None)
    want_c = backend.generate_chacha20_proof(*c)
    want_a = backend.generate_aes128_ctr_proof(*a)
    pool = ProverPool(0, 4)
    try:
        out = pool.prove_many([("chacha20",) + tuple(c), ("aes-128-ctr",) + tuple(a)] * 4)
    finally:
        pool.close()
    assert out[0::2] == [want_c] * 4 and out[1::2] == [want_a] * 4


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not on this box")
@pytest.mark.parametrize("key_len,log_n", [(16, 18), (32, 18), (16, 19)])
def test_reference_verifier_accepts_large_aes_proofs(backend, key_len, log_n):
    """BASELINE configs[2] towards its stated sizes: the stored-LDE AES driver at log 18 (both key sizes) and log 19 (AES-128,
    ~157 GB): the reference's own verifier accepts the proof and rejects it against a flipped ciphertext byte."""
    import base64
    import bench
    import zk_symmetric_crypto_b200 as z
    key, nonce, counter, pt, ct = bench.synth_aes_inputs(key_len, log_n, 0)
    proof = backend.prove_aes_ctr_raw(key, nonce, counter, pt, ct)
    b64 = base64.b64encode(proof).decode()
    name = "aes128-ctr" if key_len == 16 else "aes256-ctr"
    assert z.verify_aes_ctr_proof(b64, nonce, counter, pt, ct) == {"algorithm": name, "valid": True}
    assert ref_wasm.verify_aes_ctr_proof(b64, nonce, counter, pt, ct) == {"algorithm": name, "valid": True}
    bad = bytearray(ct)
    bad[len(bad) // 2] ^= 1
    assert ref_wasm.verify_aes_ctr_proof(b64, nonce, counter, pt, bytes(bad))["valid"] is False


BLOCK_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "chacha20_block_golden.json")))["cases"]


@pytest.mark.parametrize("case", BLOCK_GOLDEN, ids=lambda c: "log%d" % c["log_size"])
def test_block_air_proof_bytes_match_fixture(backend, case):
    """ChaCha20 block AIR (reference: bitwise/air.rs prove_bitwise): GPU proof bytes == the restatement's fixture (product-size path)."""
    import zk_symmetric_crypto_b200 as z
    proof = backend.prove_chacha20_block(case["log_size"])
    assert len(proof) == case["proof_bytes"] and hashlib.sha256(proof).hexdigest() == case["sha256"]
    assert z.verify_chacha20_block(proof) == ""


@pytest.mark.parametrize("log_size,cap", [(11, -1), (13, -1), (13, 40), (16, -1)])
def test_block_air_streaming_path(backend, log_size, cap):
    """Above log 10 the block AIR goes through the streaming pipeline (tile groups, half-domain constraint pass, partial cache):
    the host verifier accepts, a flipped byte is rejected, and the proof does not depend on the cache size."""
    import zk_symmetric_crypto_b200 as z
    backend.set_max_cached_tiles(cap)
    try:
        proof = backend.prove_chacha20_block(log_size)
    finally:
        backend.set_max_cached_tiles(-1)
    assert z.verify_chacha20_block(proof) == ""
    bad = bytearray(proof)
    bad[len(bad) // 3] ^= 4
    assert z.verify_chacha20_block(bytes(bad)) != ""
    if cap >= 0:
        assert proof == backend.prove_chacha20_block(log_size)


AES_BLOCK_GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "aes128_block_golden.json")))["cases"]


@pytest.mark.parametrize("case", AES_BLOCK_GOLDEN, ids=lambda c: "log%d" % c["log_size"])
def test_aes_block_air_proof_bytes_match_fixture(backend, case):
    """AES-128 block AIR (reference: aes/lookup/air.rs prove_aes_lookup): GPU proof bytes == the restatement's fixture."""
    import zk_symmetric_crypto_b200 as z
    proof = backend.prove_aes128_block(case["log_size"])
    assert len(proof) == case["proof_bytes"] and hashlib.sha256(proof).hexdigest() == case["sha256"]
    assert z.verify_aes128_block(proof) == ""


def test_aes_block_air_larger_trace_verifies(backend):
    import zk_symmetric_crypto_b200 as z
    proof = backend.prove_aes128_block(14)
    assert z.verify_aes128_block(proof) == ""
    bad = bytearray(proof)
    bad[len(bad) // 3] ^= 2
    assert z.verify_aes128_block(bytes(bad)) != ""
