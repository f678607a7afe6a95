"""Host-side verifier (csrc/verify.cu) against the reference's own verifier.

CPU tests (no GPU needed: verification is host code inside libs2c_b200.so): proofs come from the reference prover
(oracle/_ref, the reference's shipped binary); every accept / reject decision and every error rendering of
`verify_chacha20_proof` / `verify_aes_ctr_proof` must equal the reference's on the same -- valid, wrongly-bound and
tampered -- inputs (wasm_api.rs:609-648, :904-946; air_stream.rs:284-421; air_ctr.rs:619-714).
GPU tests: proofs from the CUDA prover verify, `prove_*_encrypt` answers like the reference (wasm_api.rs:61-188, 210-463).
"""
import base64
import struct

import pytest

import ref_wasm
import zk_symmetric_crypto_b200 as z
from make_golden import case_inputs
from make_golden_aes import aes_case_inputs

needs_ref = pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not built (needs /root/reference)")


class Layout:
    """Byte offsets inside a bincode StarkProof (tests only: where to tamper)."""

    def __init__(self, raw, stark_off):
        self.raw = raw
        p = stark_off
        self.config = p
        p += 25
        self.commitments = p + 8
        n = self.u64(p)
        p += 8 + 32 * n
        self.sampled = {}
        nt = self.u64(p)
        p += 8
        for t in range(nt):
            nc = self.u64(p)
            p += 8
            for c in range(nc):
                k = self.u64(p)
                self.sampled[(t, c)] = (p + 8, k)
                p += 8 + 16 * k
        self.decommit = []
        nt = self.u64(p)
        p += 8
        for t in range(nt):
            k = self.u64(p)
            self.decommit.append((p + 8, k))
            p += 8 + 32 * k
        self.queried = {}
        nt = self.u64(p)
        p += 8
        for t in range(nt):
            nc = self.u64(p)
            p += 8
            for c in range(nc):
                k = self.u64(p)
                self.queried[(t, c)] = (p + 8, k)
                p += 8 + 4 * k
        self.pow = p
        p += 8
        k = self.u64(p)
        self.fri_first_witness = (p + 8, k)
        p += 8 + 16 * k
        k = self.u64(p)
        self.fri_first_decommit = (p + 8, k)
        p += 8 + 32 * k
        self.fri_first_commitment = p
        p += 32
        self.n_inner = self.u64(p)
        p += 8
        self.inner = []
        for i in range(self.n_inner):
            k = self.u64(p)
            w = (p + 8, k)
            p += 8 + 16 * k
            k = self.u64(p)
            d = (p + 8, k)
            p += 8 + 32 * k
            self.inner.append((w, d, p))
            p += 32
        k = self.u64(p)
        self.last_poly = (p + 8, k)

    def u64(self, p):
        return struct.unpack_from("<Q", self.raw, p)[0]


def flip(raw, off, bit=1):
    r = bytearray(raw)
    r[off] ^= bit
    return bytes(r)


def put(raw, off, val):
    r = bytearray(raw)
    r[off] = val
    return bytes(r)


def set_lifting(raw, config_off, v):
    """PcsConfig.lifting_log_size (the Option<u32> closing the config) := Some(v)."""
    off = config_off + 24
    assert raw[off] == 0
    return raw[:off] + b"\x01" + struct.pack("<I", v) + raw[off + 1:]


def set_last_poly(raw, n_coef, log_size):
    """Rewrites the trailing LinePoly {coeffs: Vec<QM31>, log_size: u32} (one coefficient in every proof of the default config)."""
    assert struct.unpack_from("<Q", raw, len(raw) - 28)[0] == 1
    return raw[:-28] + struct.pack("<Q", n_coef) + raw[-20:-4] * n_coef + struct.pack("<I", log_size)


def resize_vec(raw, data_off, old_n, new_n, elem):
    """Rewrites a Vec's length prefix, dropping or zero-padding elements."""
    r = bytes(raw)
    body = r[data_off:data_off + elem * min(old_n, new_n)] + bytes(elem * max(0, new_n - old_n))
    return r[:data_off - 8] + struct.pack("<Q", new_n) + body + r[data_off + elem * old_n:]


def both(mine_fn, ref_fn, raw_or_b64, nonce, counter, pt, ct):
    b64 = raw_or_b64 if isinstance(raw_or_b64, str) else base64.b64encode(raw_or_b64).decode()
    return mine_fn(b64, nonce, counter, pt, ct), ref_fn(b64, nonce, counter, pt, ct)


# ---------------------------------------------------------------------------------------------------- ChaCha20
@pytest.fixture(scope="module")
def chacha_ref_proof():
    key, nonce, counter, pt, ct = case_inputs(2, 0)
    res = ref_wasm.generate_chacha20_proof(key, nonce, counter, pt, ct)
    return base64.b64decode(res["proof"]), nonce, counter, pt, ct


def chacha_mutations(raw):
    L = Layout(raw, 84)
    yield "pow_bits below minimum", put(raw, L.config, 5)
    yield "blow-up below minimum", put(raw, L.config + 4, 0)
    yield "n_queries below minimum", put(raw, L.config + 12, 2)
    yield "last layer bound changed", put(raw, L.config + 8, 1)
    yield "fold step changed", put(raw, L.config + 20, 2)
    yield "pow_bits raised", put(raw, L.config, 11)
    yield "log_size changed", put(raw, 0, raw[0] + 1)
    for t in range(3):
        yield "root %d" % t, flip(raw, L.commitments + 32 * t + 5)
    yield "trace sample", flip(raw, L.sampled[(1, 100)][0] + 4)
    yield "composition sample", flip(raw, L.sampled[(2, 3)][0])
    yield "trace decommitment hash", flip(raw, L.decommit[1][0] + 40)
    yield "composition decommitment hash", flip(raw, L.decommit[2][0] + 3)
    yield "trace decommitment short", resize_vec(raw, L.decommit[1][0], L.decommit[1][1], L.decommit[1][1] - 1, 32)
    yield "trace decommitment long", resize_vec(raw, L.decommit[1][0], L.decommit[1][1], L.decommit[1][1] + 1, 32)
    yield "trace queried value", flip(raw, L.queried[(1, 7)][0])
    yield "trace queried value, last query", flip(raw, L.queried[(1, 33279)][0] + 8)
    yield "composition queried value", flip(raw, L.queried[(2, 0)][0] + 4)
    yield "too few queried values", resize_vec(raw, L.queried[(1, 9)][0], 3, 2, 4)
    yield "too many queried values", resize_vec(raw, L.queried[(1, 9)][0], 3, 4, 4)
    yield "proof of work nonce", flip(raw, L.pow)
    yield "fri first-layer witness", flip(raw, L.fri_first_witness[0])
    yield "fri first-layer witness short", resize_vec(raw, L.fri_first_witness[0], L.fri_first_witness[1], L.fri_first_witness[1] - 1, 16)
    yield "fri first-layer decommitment", flip(raw, L.fri_first_decommit[0] + 1)
    yield "fri first-layer commitment", flip(raw, L.fri_first_commitment)
    for i, (w, d, c) in enumerate(L.inner):
        if w[1]:
            yield "fri inner %d witness" % i, flip(raw, w[0] + 2)
        if d[1]:
            yield "fri inner %d decommitment" % i, flip(raw, d[0] + 2)
        yield "fri inner %d commitment" % i, flip(raw, c + 2)
    yield "fri last-layer coefficient", flip(raw, L.last_poly[0])
    log = raw[0]
    yield "lifting_log_size = the lifting log (empty preprocessed tree has no path of that height)", set_lifting(raw, L.config, log + 1)
    yield "lifting_log_size above the lifting log", set_lifting(raw, L.config, log + 2)
    yield "lifting_log_size far above", set_lifting(raw, L.config, log + 4)
    for nc, lg in ((1, 1), (1, 3), (2, 0), (2, 1), (0, 0), (0, 1), (1, 31), (1, 32), (1, 40), (4, 2)):
        yield "last layer poly with %d coefficients, log_size %d" % (nc, lg), set_last_poly(raw, nc, lg)
    yield "truncated", raw[:-10]
    yield "truncated statement", raw[:50]
    yield "trailing bytes", raw + b"abc"
    yield "length prefix beyond the 32-bit usize of the reference build", put(raw, L.queried[(1, 0)][0] - 4, 1)


@needs_ref
def test_chacha_verdicts_match_reference(chacha_ref_proof):
    raw, nonce, counter, pt, ct = chacha_ref_proof
    mine, ref = both(z.verify_chacha20_proof, ref_wasm.verify_chacha20_proof, raw, nonce, counter, pt, ct)
    assert mine == ref == {"algorithm": "chacha20", "valid": True}
    # verifier-supplied public inputs (ChaChaPublicInputs::verify, air_stream.rs:56-64)
    for args in ((nonce, counter + 1, pt, ct), (bytes(12), counter, pt, ct), (nonce, counter, bytes(len(pt)), ct),
                 (nonce, counter, pt, ct[:-1]), (nonce, counter, b"", b"")):
        mine, ref = both(z.verify_chacha20_proof, ref_wasm.verify_chacha20_proof, raw, *args)
        assert mine == ref == {"error": "OodsNotMatching", "valid": False}
    seen = set()
    for name, mutated in chacha_mutations(raw):
        mine, ref = both(z.verify_chacha20_proof, ref_wasm.verify_chacha20_proof, mutated, nonce, counter, pt, ct)
        assert mine == ref, name
        seen.add(mine.get("error", "valid"))
    # the mutations reach every stage of the verifier
    for expected in ("OodsNotMatching", "ProofOfWork", "Merkle(RootMismatch)", "Merkle(WitnessTooShort)", "Merkle(WitnessTooLong)",
                     "Fri(InvalidNumFriLayers)", "Fri(FirstLayerCommitmentInvalid { error: RootMismatch })", "Fri(LastLayerDegreeInvalid)",
                     "Invalid proof format: io error: unexpected end of file", "valid"):
        assert expected in seen, (expected, sorted(seen))


@needs_ref
@pytest.mark.parametrize("nb,seed", [(17, 2), (64, 4)])
def test_chacha_larger_reference_proofs(nb, seed):
    """log_size 5 and 6 (more FRI layers, deeper trees): accept, and reject like the reference on a sample of mutations."""
    key, nonce, counter, pt, ct = case_inputs(nb, seed)
    raw = base64.b64decode(ref_wasm.generate_chacha20_proof(key, nonce, counter, pt, ct)["proof"])
    assert z.verify_chacha20_raw(raw, nonce, counter, pt, ct) == (True, None)
    L = Layout(raw, 84)
    muts = [flip(raw, L.sampled[(1, 31000)][0] + 9), flip(raw, L.queried[(1, 12345)][0] + 4), flip(raw, L.decommit[1][0] + 70),
            flip(raw, L.inner[-1][2]), flip(raw, L.inner[1][0][0]) if L.inner[1][0][1] else flip(raw, L.inner[1][1][0]),
            flip(raw, L.last_poly[0] + 5), resize_vec(raw, L.fri_first_decommit[0], L.fri_first_decommit[1], L.fri_first_decommit[1] - 1, 32)]
    for i, mutated in enumerate(muts):
        mine, ref = both(z.verify_chacha20_proof, ref_wasm.verify_chacha20_proof, mutated, nonce, counter, pt, ct)
        assert mine == ref and mine.get("valid") is not True, i


@needs_ref
def test_verify_input_validation_matches_reference(chacha_ref_proof):
    raw, nonce, counter, pt, ct = chacha_ref_proof
    b64 = base64.b64encode(raw).decode()
    for bad in ("!!!!", b64[:-3], "QUJDR", "QUI", "QUJ=", "", "QUJD"):
        mine, ref = both(z.verify_chacha20_proof, ref_wasm.verify_chacha20_proof, bad, nonce, counter, pt, ct)
        assert mine == ref, bad[:16]
        mine, ref = both(z.verify_aes_ctr_proof, ref_wasm.verify_aes_ctr_proof, bad, nonce, counter, pt, ct)
        assert mine == ref, bad[:16]
    mine, ref = both(z.verify_chacha20_proof, ref_wasm.verify_chacha20_proof, b64, nonce[:11], counter, pt, ct)
    assert mine == ref == {"error": "Nonce must be 12 bytes, got 11"}
    assert z.verify_chacha20_proof("A" * (8 * 1024 * 1024 + 1), nonce, counter, pt, ct) == {"error": "Proof payload too large"}
    # a ChaCha proof handed to the AES verifier (and vice versa) is a format or binding error, never an accept
    mine, ref = both(z.verify_aes_ctr_proof, ref_wasm.verify_aes_ctr_proof, b64, nonce, counter, pt, ct)
    assert mine == ref and mine.get("valid") is not True


def test_raw_verify_rejects_garbage_without_reference():
    ok, err = z.verify_chacha20_raw(b"", bytes(12), 0, b"", b"")
    assert not ok and err == "Invalid proof format: io error: unexpected end of file"
    ok, err = z.verify_aes_ctr_raw(bytes(4) + struct.pack("<I", 7) + bytes(200), bytes(12), 0, b"", b"")
    assert not ok and err == "Invalid proof format: invalid value: integer `7`, expected variant index 0 <= i < 2"
    ok, err = z.verify_chacha20_raw(bytes(4096), bytes(12), 0, b"", b"")
    assert not ok and "pow_bits (0) below minimum (10)" in err


# ---------------------------------------------------------------------------------------------------- AES-CTR
@pytest.fixture(scope="module")
def aes_ref_proofs():
    out = []
    for key_len, n_blocks, seed in ((16, 5, None), (32, 3, 2), (16, 300, 3)):
        key, nonce, counter, pt, ct = aes_case_inputs(key_len, n_blocks, seed)
        fn = ref_wasm.generate_aes128_ctr_proof if key_len == 16 else ref_wasm.generate_aes256_ctr_proof
        res = fn(key, nonce, counter, pt, ct)
        out.append((key_len, base64.b64decode(res["proof"]), nonce, counter, pt, ct))
    return out


def aes_mutations(raw, key_len):
    L = Layout(raw, 136)
    n_main = 24480 if key_len == 16 else 34784
    n_inter = 320 if key_len == 16 else 448
    yield "log_size changed", put(raw, 0, raw[0] + 1)
    yield "ctr claimed sum", flip(raw, 88)
    yield "table claimed sum", flip(raw, 104)
    yield "interaction width beyond the cap", put(raw, 122, 2)
    yield "pow_bits below minimum", put(raw, L.config, 5)
    for t in range(4):
        yield "root %d" % t, flip(raw, L.commitments + 32 * t + 5)
    for t, c in ((0, 0), (0, 1), (1, 5), (1, n_main), (2, 0), (2, n_inter - 4), (2, n_inter), (2, n_inter + 3), (3, 0), (3, 7)):
        yield "sample %d/%d" % (t, c), flip(raw, L.sampled[(t, c)][0])
    yield "previous-row sample of the LogUp column", flip(raw, L.sampled[(2, n_inter - 1)][0])
    yield "current-row sample of the LogUp column", flip(raw, L.sampled[(2, n_inter - 1)][0] + 16)
    yield "previous-row sample of the table LogUp column", flip(raw, L.sampled[(2, n_inter + 2)][0] + 3)
    for t in range(4):
        yield "decommitment %d" % t, flip(raw, L.decommit[t][0] + 3)
    for t, c in ((0, 1), (1, 3), (1, n_main), (2, 5), (2, n_inter + 2), (3, 2)):
        yield "queried %d/%d" % (t, c), flip(raw, L.queried[(t, c)][0])
        yield "queried %d/%d, last query" % (t, c), flip(raw, L.queried[(t, c)][0] + 8)
    yield "proof of work nonce", flip(raw, L.pow)
    yield "fri first-layer witness", flip(raw, L.fri_first_witness[0])
    yield "fri last-layer coefficient", flip(raw, L.last_poly[0])
    log = raw[0]
    yield "lifting_log_size = the lifting log", set_lifting(raw, L.config, log + 1)   # valid at log 8, WitnessTooShort above
    yield "lifting_log_size above the lifting log", set_lifting(raw, L.config, log + 2)
    for nc, lg in ((1, 1), (2, 0), (0, 1), (1, 32)):
        yield "last layer poly with %d coefficients, log_size %d" % (nc, lg), set_last_poly(raw, nc, lg)
    yield "truncated", raw[:-10]
    yield "truncated statement", raw[:100]
    yield "key size variant out of range", put(raw, 4, 2)


@needs_ref
def test_aes_verdicts_match_reference(aes_ref_proofs):
    for key_len, raw, nonce, counter, pt, ct in aes_ref_proofs:
        alg = "aes128-ctr" if key_len == 16 else "aes256-ctr"
        mine, ref = both(z.verify_aes_ctr_proof, ref_wasm.verify_aes_ctr_proof, raw, nonce, counter, pt, ct)
        assert mine == ref == {"algorithm": alg, "valid": True}
        for args in ((nonce, counter + 1, pt, ct), (nonce, counter, pt, bytes(len(ct)))):
            mine, ref = both(z.verify_aes_ctr_proof, ref_wasm.verify_aes_ctr_proof, raw, *args)
            assert mine == ref == {"error": "OodsNotMatching", "valid": False}
        seen = set()
        for name, mutated in aes_mutations(raw, key_len):
            mine, ref = both(z.verify_aes_ctr_proof, ref_wasm.verify_aes_ctr_proof, mutated, nonce, counter, pt, ct)
            assert mine == ref, (key_len, name)
            seen.add(mine.get("error"))
        for expected in ("OodsNotMatching", "ProofOfWork", "Merkle(RootMismatch)", "Fri(LastLayerDegreeInvalid)"):
            assert expected in seen, (expected, sorted(seen, key=str))


@needs_ref
def test_aes_statement_the_reference_cannot_parse_is_rejected(aes_ref_proofs):
    """The reference panics (wasm trap) on a statement whose key size / interaction widths contradict the proof's shape;
    this verifier answers with a structured rejection instead."""
    key_len, raw, nonce, counter, pt, ct = aes_ref_proofs[0]
    for off, val in ((4, 1), (120, raw[120] ^ 1), (128, 5)):
        ok, err = z.verify_aes_ctr_raw(put(raw, off, val), nonce, counter, pt, ct)
        assert not ok and err.startswith("InvalidStructure(")


@needs_ref
def test_lifting_log_size_the_reference_panics_on_is_rejected(chacha_ref_proof, aes_ref_proofs):
    """lifting_log_size below the largest committed column (or beyond the circle group) is a panic in the reference (wasm trap);
    here it is a structured rejection, never an accept."""
    raw, nonce, counter, pt, ct = chacha_ref_proof
    for v in (0, raw[0], 40):
        bad = set_lifting(raw, 84, v)
        with pytest.raises(RuntimeError):
            ref_wasm.verify_chacha20_proof(base64.b64encode(bad).decode(), nonce, counter, pt, ct)
        ok, err = z.verify_chacha20_raw(bad, nonce, counter, pt, ct)
        assert not ok and err.startswith("InvalidStructure(")
    key_len, raw, nonce, counter, pt, ct = aes_ref_proofs[2]
    ok, err = z.verify_aes_ctr_raw(set_lifting(raw, 136, raw[0]), nonce, counter, pt, ct)
    assert not ok and err.startswith("InvalidStructure(")


@needs_ref
def test_operator_verify(chacha_ref_proof):
    """js/src/stwo/operator.ts:135-180 groth16Verify: bool result, bytes or base64 proof, 'in' = ciphertext, 'out' = plaintext."""
    raw, nonce, counter, pt, ct = chacha_ref_proof
    op = z.make_stwo_zk_operator("chacha20")
    sig = {"noncesAndCounters": [{"nonce": nonce, "counter": counter}], "in": ct, "out": pt}
    assert op.groth16_verify(sig, raw) is True
    assert op.groth16_verify(sig, base64.b64encode(raw).decode()) is True
    assert op.groth16_verify(dict(sig, out=bytes(len(pt))), raw) is False
    assert op.groth16_verify(dict(sig, noncesAndCounters=[]), raw) is False
    warnings = []

    class Log:
        def warn(self, m):
            warnings.append(m)
    assert op.groth16_verify(sig, flip(raw, 300), Log()) is False and "OodsNotMatching" in warnings[0]
    op.release()


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("nb", [1, 33 - 1, 1 << 10, 1 << 13])
def test_gpu_chacha_proofs_verify(backend, nb):
    key, nonce, counter, pt, ct = case_inputs(nb, 11)
    proof = backend.prove_chacha20_raw(key, nonce, counter, pt, ct)
    assert z.verify_chacha20_raw(proof, nonce, counter, pt, ct) == (True, None)
    assert z.verify_chacha20_raw(proof, nonce, counter + 1, pt, ct) == (False, "OodsNotMatching")
    L = Layout(proof, 84)
    assert z.verify_chacha20_raw(flip(proof, L.queried[(1, 5)][0]), nonce, counter, pt, ct) == (False, "Merkle(RootMismatch)")
    assert z.verify_chacha20_raw(flip(proof, L.sampled[(1, 5)][0]), nonce, counter, pt, ct) == (False, "OodsNotMatching")
    assert z.verify_chacha20_raw(flip(proof, L.pow), nonce, counter, pt, ct) == (False, "ProofOfWork")


@pytest.mark.gpu
@pytest.mark.parametrize("key_len,nb", [(16, 5), (32, 300), (16, 1 << 12), (32, 1 << 11)])
def test_gpu_aes_proofs_verify(backend, key_len, nb):
    key, nonce, counter, pt, ct = aes_case_inputs(key_len, nb, 21)
    proof = backend.prove_aes_ctr_raw(key, nonce, counter, pt, ct)
    assert z.verify_aes_ctr_raw(proof, nonce, counter, pt, ct) == (True, None)
    assert z.verify_aes_ctr_raw(proof, nonce, counter, pt, bytes(len(ct))) == (False, "OodsNotMatching")
    L = Layout(proof, 136)
    assert z.verify_aes_ctr_raw(flip(proof, L.queried[(2, 1)][0]), nonce, counter, pt, ct) == (False, "Merkle(RootMismatch)")
    assert z.verify_aes_ctr_raw(flip(proof, 104), nonce, counter, pt, ct) == (False, "OodsNotMatching")


@pytest.mark.gpu
def test_prove_encrypt_matches_reference(backend):
    key, nonce, counter, pt, ct = case_inputs(2, 0)
    res = backend.prove_chacha20_encrypt(key, nonce, counter, pt, ct)
    assert res == {"algorithm": "chacha20", "blocks": 2, "success": True}
    bad = bytearray(ct)
    bad[3] ^= 1
    cases = [(key, nonce, counter, pt, bytes(bad)), (key[:31], nonce, counter, pt, ct), (key, nonce, counter, pt[:63], ct[:63]),
             (key, nonce, 0xFFFFFFFF, pt, ct)]
    for args in cases:
        mine = backend.prove_chacha20_encrypt(*args)
        assert "error" in mine
        if ref_wasm.available():
            assert mine == ref_wasm.prove_chacha20_encrypt(*args)
    if ref_wasm.available():
        assert res == ref_wasm.prove_chacha20_encrypt(key, nonce, counter, pt, ct)
    for key_len, fn, alg in ((16, backend.prove_aes128_ctr_encrypt, "aes128-ctr"), (32, backend.prove_aes256_ctr_encrypt, "aes256-ctr")):
        key, nonce, counter, pt, ct = aes_case_inputs(key_len, 5, 1)
        assert fn(key, nonce, counter, pt, ct) == {"algorithm": alg, "blocks": 5, "success": True}
        bad = bytearray(ct)
        bad[0] ^= 1
        mine = fn(key, nonce, counter, pt, bytes(bad))
        assert mine == {"error": "Ciphertext does not match encryption - invalid witness"}
        if ref_wasm.available():
            rfn = ref_wasm.prove_aes128_ctr_encrypt if key_len == 16 else ref_wasm.prove_aes256_ctr_encrypt
            assert mine == rfn(key, nonce, counter, pt, bytes(bad))


@pytest.mark.gpu
def test_operator_round_trip_on_gpu():
    """js/src/tests/lib.test.ts:27-158 in miniature: generateWitness -> groth16Prove -> groth16Verify; zeroed plaintext fails."""
    for alg, (key, nonce, counter, pt, ct) in (("chacha20", case_inputs(2, 5)), ("aes-128-ctr", aes_case_inputs(16, 5, 5)),
                                              ("aes-256-ctr", aes_case_inputs(32, 5, 6))):
        op = z.make_stwo_zk_operator(alg)
        w = op.generate_witness({"key": key, "nonce": nonce, "counter": counter, "in": ct, "out": pt})
        proof = op.groth16_prove(w)["proof"]
        sig = {"noncesAndCounters": [{"nonce": nonce, "counter": counter}], "in": ct, "out": pt}
        assert op.groth16_verify(sig, proof) is True
        assert op.groth16_verify(dict(sig, out=bytes(len(pt))), proof) is False
        op.release()


@needs_ref
def test_fuzzed_proofs_get_the_reference_answer(chacha_ref_proof, aes_ref_proofs):
    """Seeded random damage (bit flips, truncation, forged length fields, non-canonical field words, deleted / inserted
    bytes): never an accept, never a crash, and the same JSON as the reference verifier wherever the reference itself does
    not panic."""
    import random
    rnd = random.Random(20260117)

    def mutate(raw):
        r = bytearray(raw)
        kind = rnd.randrange(6)
        if kind == 0:
            for _ in range(rnd.randrange(1, 4)):
                r[rnd.randrange(len(r))] ^= 1 << rnd.randrange(8)
        elif kind == 1:
            r = r[:rnd.randrange(len(r))]
        elif kind == 2:
            p = rnd.randrange(80, 400)
            r[p:p + 8] = struct.pack("<Q", rnd.choice([0, 1, 2, 3, 5, 1 << 20, 1 << 40, (1 << 64) - 1]))
        elif kind == 3:
            p = rnd.randrange(len(r))
            r[p:p + 4] = struct.pack("<I", rnd.choice([0x7FFFFFFF, 0x80000000, 0xFFFFFFFF]))
        elif kind == 4:
            p = rnd.randrange(len(r))
            del r[p:p + rnd.randrange(1, 64)]
        else:
            p = rnd.randrange(len(r))
            r[p:p] = bytes(rnd.randrange(1, 64))
        return bytes(r)

    raw, nonce, counter, pt, ct = chacha_ref_proof
    _, araw, anonce, acounter, apt, act = aes_ref_proofs[0]
    for R, args, mine_fn, ref_fn in ((raw, (nonce, counter, pt, ct), z.verify_chacha20_proof, ref_wasm.verify_chacha20_proof),
                                     (araw, (anonce, acounter, apt, act), z.verify_aes_ctr_proof, ref_wasm.verify_aes_ctr_proof)):
        for i in range(60):
            m = mutate(R)
            if m == R:
                continue
            b64 = base64.b64encode(m).decode()
            mine = mine_fn(b64, *args)
            assert mine.get("valid") is not True, i
            try:
                ref = ref_fn(b64, *args)
            except RuntimeError:        # the reference trapped (panic inside the wasm module)
                continue
            assert mine == ref, (i, mine, ref)
