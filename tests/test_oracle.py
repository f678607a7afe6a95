"""CPU tests (no GPU): the oracle against the reference's own KATs, the golden fixtures generated from the reference,
and (when oracle/_ref is built) the reference itself, byte for byte."""
import base64
import hashlib
import json
import os
import struct

import numpy as np
import pytest

import api as oracle_api
import chacha_air as ca
import stwo_core as sc
import ref_wasm
from make_golden import case_inputs

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "chacha20_golden.json")))["cases"]


def test_rfc7539_block_kat():
    # /root/reference/stwo/src/chacha/block.rs:116-139
    key = struct.unpack("<8I", bytes(range(32)))
    nonce = struct.unpack("<3I", bytes([0, 0, 0, 9, 0, 0, 0, 0x4A, 0, 0, 0, 0]))
    out = ca.chacha20_block_words(key, 1, nonce)
    assert out[:4] == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3]
    assert out[-1] == 0x4E3C50A2


def test_quarter_round_kat():
    # /root/reference/stwo/src/chacha/quarter_round.rs:135-140 (RFC 7539 2.1.1)
    a, b, c, d = 0x11111111, 0x01020304, 0x9B8D6F43, 0x01234567
    M = 0xFFFFFFFF
    rotl = lambda x, r: ((x << r) | (x >> (32 - r))) & M
    a = (a + b) & M; d = rotl(d ^ a, 16); c = (c + d) & M; b = rotl(b ^ c, 12)
    a = (a + b) & M; d = rotl(d ^ a, 8); c = (c + d) & M; b = rotl(b ^ c, 7)
    assert (a, b, c, d) == (0xEA2A92F4, 0xCB1CF8CE, 0x4581472E, 0x5881C4BB)


def test_field_identities():
    rng = np.random.default_rng(0)
    a = rng.integers(0, sc.P, size=(64, 4), dtype=np.uint64)
    b = rng.integers(1, sc.P, size=(64, 4), dtype=np.uint64)
    one = np.zeros((64, 4), dtype=np.uint64); one[:, 0] = 1
    assert np.array_equal(sc.q_mul(b, sc.q_inv(b)), one)
    assert np.array_equal(sc.q_mul(a, b), sc.q_mul(b, a))
    x = sc.QM31(*[int(v) for v in a[0]]); y = sc.QM31(*[int(v) for v in b[0]])
    assert tuple(int(v) for v in sc.q_mul(a[:1], b[:1])[0]) == (x * y).v
    assert (y * y.inv()).v == (1, 0, 0, 0)


def test_circle_generator_and_domain():
    g = sc.GEN
    assert (g[0] * g[0] + g[1] * g[1]) % sc.P == 1
    p = g
    for _ in range(30):
        p = sc.pt_double(p)
    assert p == (sc.P - 1, 0) or p == ((-1) % sc.P, 0)       # order 2^31
    d = sc.canonic_domain(5)
    xs, ys = d.points()
    assert np.array_equal(xs[16:], xs[:16]) and np.array_equal(ys[16:], sc.m_neg(ys[:16]))


@pytest.mark.parametrize("log_n", [1, 2, 4, 7, 10])
def test_fft_roundtrip_and_eval(log_n):
    rng = np.random.default_rng(log_n)
    v = rng.integers(0, sc.P, size=(3, 1 << log_n), dtype=np.uint64)
    c = sc.circle_ifft(v)
    assert np.array_equal(sc.circle_fft(c), v)
    # the LDE restricted through eval_at_point: value at a domain point equals the polynomial evaluated there
    lde = sc.circle_fft(c, log_n + 1)
    xs, ys = sc.canonic_domain(log_n + 1).points_bitrev()
    for row in (0, 1, (1 << log_n) + 3 if log_n > 1 else 1):
        got = sc.eval_at_point(c, sc.QM31(int(xs[row])), sc.QM31(int(ys[row])))
        assert np.array_equal(got[:, 0], lde[:, row]) and not got[:, 1:].any()


def test_merkle_empty_and_lifting():
    assert sc.MerkleTree([]).root().hex() == "69217a3079908094e11121d042354a7c1f55b6482ca1a51e1b250dfd1ed0eef9"
    assert [sc.lifted_index(i, 4, 2) for i in range(16)] == [0, 1, 0, 1, 0, 1, 0, 1, 2, 3, 2, 3, 2, 3, 2, 3]


def test_trace_satisfies_constraints_on_trace_domain():
    # semantics of assert_constraints_on_polys (/root/reference/stwo/src/chacha/bitwise/mod.rs:116-216)
    key, nonce, counter, pt, ct = case_inputs(16, 7)
    log, K, NO, C, PT, CT, m = oracle_api.build_chacha_inputs(key, nonce, counter, pt, ct)
    trace, valid = ca.generate_stream_trace(log, K, NO, C, PT, CT, m)
    assert valid and trace.shape == (ca.N_COLS, 16)
    apr = np.ones((ca.N_CONSTRAINTS, 4), dtype=np.uint64)
    assert not ca.evaluate_constraints(trace, apr).any()
    bad = bytearray(ct); bad[5] ^= 1
    log, K, NO, C, PT, CT, m = oracle_api.build_chacha_inputs(key, nonce, counter, pt, bytes(bad))
    trace, valid = ca.generate_stream_trace(log, K, NO, C, PT, CT, m)
    assert not valid and ca.evaluate_constraints(trace, apr).any()


@pytest.mark.parametrize("case", [c for c in GOLDEN if c["n_blocks"] <= 17], ids=lambda c: c["name"])
def test_oracle_matches_golden(case):
    key, nonce, counter, pt, ct = case_inputs(case["n_blocks"], case["seed"])
    res = oracle_api.generate_chacha20_proof(key, nonce, counter, pt, ct)
    assert len(res["proof_bytes"]) == case["proof_len"]
    assert hashlib.sha256(res["proof_bytes"]).hexdigest() == case["proof_sha256"]
    assert hashlib.sha256(res["proof"].encode()).hexdigest() == case["b64_sha256"]


def test_oracle_constraints_not_satisfied_quirk():
    case = [c for c in GOLDEN if "error" in c][0]
    key, nonce, counter, pt, ct = case_inputs(case["n_blocks"], case["seed"])
    assert oracle_api.generate_chacha20_proof(key, nonce, counter, pt, ct) == {"error": case["error"]}


def test_oracle_input_validation_messages():
    z = bytes(64)
    assert oracle_api.generate_chacha20_proof(bytes(31), bytes(12), 0, z, z)["error"] == "Key must be 32 bytes, got 31"
    assert oracle_api.generate_chacha20_proof(bytes(32), bytes(11), 0, z, z)["error"] == "Nonce must be 12 bytes, got 11"
    assert "non-empty multiple of 64" in oracle_api.generate_chacha20_proof(bytes(32), bytes(12), 0, b"", b"")["error"]
    assert "same length" in oracle_api.generate_chacha20_proof(bytes(32), bytes(12), 0, z, z + z)["error"]
    assert "Counter overflow" in oracle_api.generate_chacha20_proof(bytes(32), bytes(12), 0xFFFFFFFF, z + z, z + z)["error"]
    assert oracle_api.generate_chacha20_proof(bytes(32), bytes(12), 0, z, z)["error"].startswith("Ciphertext does not match")


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_reference_live():
    key, nonce, counter, pt, ct = case_inputs(3, 11)
    ref = ref_wasm.generate_chacha20_proof(key, nonce, counter, pt, ct)
    mine = oracle_api.generate_chacha20_proof(key, nonce, counter, pt, ct)
    assert mine["proof"] == ref["proof"]
    assert ref_wasm.verify_chacha20_proof(mine["proof"], nonce, counter, pt, ct) == {"algorithm": "chacha20", "valid": True}
    assert ref_wasm.verify_chacha20_proof(mine["proof"], nonce, counter + 1, pt, ct)["valid"] is False
    assert ref_wasm.get_circuits_info()["chacha20"] == {"cols": ca.N_COLS, "constraints": ca.N_CONSTRAINTS,
                                                        "block_bytes": 64, "key_bytes": 32}


def test_half_domain_facts_behind_the_constraint_pass():
    """The CUDA prover evaluates the constraints on storage rows [0, N) only.  Pinned here on the oracle:
    (i) those rows are the circle domain with half coset g_(n+2) + <g_(n-1)>, in its own bit-reversed order;
    (ii) the trace coset's vanishing polynomial is constant on them and takes the opposite value on rows [N, 2N);
    (iii) the composition polynomial of a real trace is p_left + c * Z_H: its right half is one constant per coordinate."""
    import prover as op
    for n in (4, 5, 7):
        N = 1 << n
        xs, ys = sc.canonic_domain(n + 1).points_bitrev()
        sub = sc.CircleDomain(sc.Coset(sc.subgroup_gen(n + 2), n - 1))   # initial g_(n+2), step g_(n-1)
        sx, sy = sub.points_bitrev()
        assert np.array_equal(xs[:N], sx) and np.array_equal(ys[:N], sy)
        z = op.coset_vanishing_on_domain(n, n + 1)
        assert len(set(int(v) for v in z[:N])) == 1 and len(set(int(v) for v in z[N:])) == 1
        assert (int(z[0]) + int(z[N])) % sc.P == 0
    key, nonce, counter, pt, ct = case_inputs(3, 7)
    log, K, NO, C, PT, CT, mrows = oracle_api.build_chacha_inputs(key, nonce, counter, pt, ct)
    trace, valid = ca.generate_stream_trace(log, K, NO, C, PT, CT, mrows)
    assert valid
    lde = sc.circle_fft(sc.circle_ifft(trace), log + 1)
    alpha = sc.QM31(123456789, 987654321, 55555, 2147483000)
    apr = op.secure_powers(alpha, ca.N_CONSTRAINTS)[::-1].copy()
    acc = ca.evaluate_constraints(lde, apr)
    acc = sc.q_mul_m31(acc, sc.m_inv(op.coset_vanishing_on_domain(log, log + 1)))
    halves = op.finalize_composition(acc, log + 1)        # [left c0..c3, right c0..c3], 2^log coefficients each
    for right in halves[4:]:
        assert not np.any(right[1:])                      # only the constant term survives
    assert any(int(r[0]) for r in halves[4:])


def _plan_combs():
    """(sum word, operand a, operand b, carry word) of every 32-bit adder of the AIR, in trace order -- the same bookkeeping
    as build_plan() in csrc/prove_chacha.cu (words of 32 columns: 16 initial, 80 x [S C X]x4, 16 x [S C], 16 pt, 16 ct)."""
    combs, state = [], list(range(16))
    for q in range(80):
        qi, a0 = q & 7, q & 3
        a = a0
        if qi < 4:
            b, c, d = 4 + a0, 8 + a0, 12 + a0
        else:
            b, c, d = 4 + ((a0 + 1) & 3), 8 + ((a0 + 2) & 3), 12 + ((a0 + 3) & 3)
        base = 16 + 12 * q
        S1, C1, X1, S2, C2, X2, S3, C3, X3, S4, C4, X4 = range(base, base + 12)
        combs += [(S1, state[a], state[b], C1), (S2, state[c], X1, C2), (S3, S1, X2, C3), (S4, S2, X3, C4)]
        state[a], state[b], state[c], state[d] = S3, X4, S4, X3
    for i in range(16):
        combs.append((976 + 2 * i, state[i], i, 977 + 2 * i))
    return combs


def test_adder_sum_words_follow_from_their_operands():
    """Facts behind the CUDA prover's skipping of the 336 adder-sum words, on the oracle's trace:
    s_i = a_i + b_i + c_(i-1) - 2 c_i (i) on the extended domain, (ii) at an out-of-domain QM31 point, and (iii) a random
    combination of all columns equals the combination with the sum columns' coefficients folded into their operands'."""
    key, nonce, counter, pt, ct = case_inputs(5, 11)
    log, K, NO, C, PT, CT, mrows = oracle_api.build_chacha_inputs(key, nonce, counter, pt, ct)
    trace, valid = ca.generate_stream_trace(log, K, NO, C, PT, CT, mrows)
    coef = sc.circle_ifft(trace)
    lde = sc.circle_fft(coef, log + 1).astype(np.int64)
    combs = _plan_combs()
    assert len(combs) == 336
    P = int(sc.P)
    z = sc.get_random_point(sc.Blake2sChannel())
    rng = np.random.default_rng(3)
    kappa = rng.integers(0, P, size=ca.N_COLS, dtype=np.int64)
    folded = kappa.copy()
    for (S, A, B, Cy) in reversed(combs):
        for i in range(32):
            k = folded[S * 32 + i]
            folded[A * 32 + i] = (folded[A * 32 + i] + k) % P
            folded[B * 32 + i] = (folded[B * 32 + i] + k) % P
            folded[Cy * 32 + i] = (folded[Cy * 32 + i] - 2 * k) % P
            if i > 0:
                folded[Cy * 32 + i - 1] = (folded[Cy * 32 + i - 1] + k) % P
            folded[S * 32 + i] = 0
    full = np.zeros(trace.shape[1], dtype=object)
    part = np.zeros(trace.shape[1], dtype=object)
    for j in range(ca.N_COLS):
        col = trace[j].astype(object)
        full = full + int(kappa[j]) * col
        if folded[j]:
            part = part + int(folded[j]) * col
    assert all(int(x) % P == int(y) % P for x, y in zip(full, part))
    for (S, A, B, Cy) in combs[:8] + combs[-4:]:
        cols = lambda w: lde[32 * w:32 * w + 32]
        cprev = np.vstack([np.zeros((1, lde.shape[1]), dtype=np.int64), cols(Cy)[:-1]])
        assert np.array_equal(cols(S) % P, (cols(A) + cols(B) + cprev - 2 * cols(Cy)) % P)
        ev = lambda w: [sc.QM31(*[int(v) for v in r]) for r in sc.eval_at_point(coef[32 * w:32 * w + 32], z[0], z[1])]
        s, a, b, c = ev(S), ev(A), ev(B), ev(Cy)
        for i in range(32):
            want = a[i] + b[i] - c[i] - c[i]
            if i > 0:
                want = want + c[i - 1]
            assert s[i].v == want.v


def test_block_air_oracle_proof_and_host_verifier():
    """ChaCha20 block AIR (bitwise/air.rs prove_bitwise / verify_bitwise): the restatement's proof matches its fixture, the C++ host
    verifier (an independent implementation of the same AIR as a constraint table) accepts it and rejects mutations."""
    import zk_symmetric_crypto_b200 as z
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "chacha20_block_golden.json")))["cases"][0]
    proof = oracle_api.prove_bitwise(gold["log_size"])
    assert len(proof) == gold["proof_bytes"] and hashlib.sha256(proof).hexdigest() == gold["sha256"]
    assert z.verify_chacha20_block(proof) == ""
    for pos, want in ((0, None), (300, None), (len(proof) // 2, None), (len(proof) - 40, None)):
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert z.verify_chacha20_block(bytes(bad)) != ""
    assert z.verify_chacha20_block(proof[:-1]).startswith("Invalid proof format")
