"""CPU tests (no GPU): the AES-CTR oracle against the reference's KATs and the golden fixtures generated from the reference
binary (tests/golden/make_golden_aes.py), byte for byte."""
import hashlib
import json
import os

import numpy as np
import pytest

import aes_air as aa
import aes_api
import ref_wasm
from make_golden_aes import aes_case_inputs

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "aes_ctr_golden.json")))["cases"]


def test_fips197_kats():
    # /root/reference/stwo/src/aes/mod.rs:431-470
    pt = list(bytes.fromhex("00112233445566778899aabbccddeeff"))
    assert bytes(aa.encrypt_block(aa.expand_key(bytes(range(16))), pt)).hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"
    assert bytes(aa.encrypt_block(aa.expand_key(bytes(range(32))), pt)).hex() == "8ea2b7ca516745bfeafc49904b496089"
    assert aa.SBOX[0] == 0x63 and aa.SBOX[0x53] == 0xED and len(set(int(x) for x in aa.SBOX)) == 256


def test_counts_match_reference_circuit_info():
    # get_circuits_info() of the reference (wasm_api.rs:993-1008)
    assert (aa.n_cols(16), aa.n_constraints(16)) == (24480, 34464)
    assert (aa.n_cols(32), aa.n_constraints(32)) == (34784, 49024)


def test_trace_satisfies_constraints_on_trace_domain():
    """assert_constraints_on_polys semantics: every base-field constraint vanishes on the generated trace; the LogUp
    constraints vanish with the generated interaction trace."""
    key, nonce, counter, pt, ct = aes_case_inputs(16, 20, 9)
    log_size, nonce_rows, counters, PT, CT = aes_api.build_aes_inputs(key, nonce, counter, pt, ct)
    trace, lookups, mults, valid = aa.generate_ctr_trace(log_size, key, nonce_rows, counters, PT, CT)
    assert valid and trace.shape == (24480, 256) and int(mults.sum()) == 160 * 256
    from stwo_core import Blake2sChannel
    el = aa.SboxElements.draw(Blake2sChannel())
    icols, csum = aa.ctr_interaction_trace(log_size, lookups, el)
    tcols, tsum = aa.table_interaction_trace(mults, el)
    assert (csum + tsum).v == (0, 0, 0, 0)
    inter = np.stack(icols, axis=0)
    order = aa.coset_order_to_storage(log_size)
    prev = np.empty(256, dtype=np.int64)
    prev[order] = order[np.arange(256) - 1]          # previous row in coset order
    apr = np.ones((aa.n_constraints(16), 4), dtype=np.uint64)
    acc = aa.evaluate_ctr_constraints(trace, inter, inter[-4:][:, prev], el, csum, log_size, apr, 16)
    assert not acc.any()


@pytest.mark.parametrize("case", [c for c in GOLDEN if c["n_blocks"] <= 5], ids=lambda c: c["name"])
def test_oracle_matches_golden(case):
    key, nonce, counter, pt, ct = aes_case_inputs(case["key_len"], case["n_blocks"], case["seed"], case["corrupt"])
    fn = aes_api.generate_aes128_ctr_proof if case["key_len"] == 16 else aes_api.generate_aes256_ctr_proof
    res = fn(key, nonce, counter, pt, ct)
    if "error" in case:
        assert res == {"error": case["error"]}
        return
    assert len(res["proof_bytes"]) == case["proof_len"]
    assert hashlib.sha256(res["proof_bytes"]).hexdigest() == case["proof_sha256"]
    assert hashlib.sha256(res["proof"].encode()).hexdigest() == case["b64_sha256"]


def test_oracle_matches_golden_mixed_sizes():
    """log_size 9: the log-8 S-box columns are lifted (size-sorted Merkle leaves, lifted accumulation, doubled sample
    points, periodicity samples)."""
    case = [c for c in GOLDEN if c["name"] == "aes128_300blocks_log9"][0]
    key, nonce, counter, pt, ct = aes_case_inputs(case["key_len"], case["n_blocks"], case["seed"])
    res = aes_api.generate_aes128_ctr_proof(key, nonce, counter, pt, ct)
    assert hashlib.sha256(res["proof_bytes"]).hexdigest() == case["proof_sha256"]


def test_input_validation_messages():
    z = bytes(16)
    assert aes_api.generate_aes128_ctr_proof(bytes(15), bytes(12), 0, z, z)["error"] == "Key must be 16 bytes, got 15"
    assert aes_api.generate_aes256_ctr_proof(bytes(16), bytes(12), 0, z, z)["error"] == "Key must be 32 bytes, got 16"
    assert "multiple of 16" in aes_api.generate_aes128_ctr_proof(bytes(16), bytes(12), 0, b"", b"")["error"]


@pytest.mark.skipif(not ref_wasm.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_verifier_accepts_oracle_proof():
    key, nonce, counter, pt, ct = aes_case_inputs(32, 2, 12)
    mine = aes_api.generate_aes256_ctr_proof(key, nonce, counter, pt, ct)
    assert ref_wasm.verify_aes_ctr_proof(mine["proof"], nonce, counter, pt, ct) == {"algorithm": "aes256-ctr", "valid": True}


def test_aes_block_air_oracle_proof_and_host_verifier():
    """AES-128 block AIR (aes/lookup/air.rs prove_aes_lookup / verify_aes_lookup): the restatement's proof matches its fixture, the C++
    host verifier accepts it and rejects mutations."""
    import hashlib
    import json
    import os
    import zk_symmetric_crypto_b200 as z
    import aes_api
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "aes128_block_golden.json")))["cases"][0]
    proof = aes_api.prove_aes_lookup(gold["log_size"])
    assert len(proof) == gold["proof_bytes"] and hashlib.sha256(proof).hexdigest() == gold["sha256"]
    assert z.verify_aes128_block(proof) == ""
    for pos in (0, 10, 60, len(proof) // 2, len(proof) - 40):
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert z.verify_aes128_block(bytes(bad)) != ""
    assert z.verify_aes128_block(proof[:-1]).startswith("Invalid proof format")
