"""Fixtures for the block AIRs: ChaCha20 (reference: stwo/src/chacha/bitwise/air.rs prove_bitwise) and AES-128
(stwo/src/aes/lookup/air.rs prove_aes_lookup).  The reference's product
binary does not export this prover, so unlike chacha20_golden.json these hashes come from the numpy restatement (oracle/api.py
prove_bitwise), whose shared machinery (commitments, prove_values, FRI, channel) is pinned to the reference binary by the stream
AIR fixtures: parity for this variant is pinned to the restatement, not to the reference itself.
Run from the repo root: python tests/golden/make_golden_block.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import aes_api
import api

if __name__ == "__main__":
    aes_cases = []
    for log_size in (8, 9):
        p = aes_api.prove_aes_lookup(log_size)
        aes_cases.append({"log_size": log_size, "proof_bytes": len(p), "sha256": hashlib.sha256(p).hexdigest()})
    json.dump({"generator": "oracle/aes_api.py prove_aes_lookup (numpy restatement; not exported by the reference binary)", "cases": aes_cases},
              open(os.path.join(ROOT, "tests", "golden", "aes128_block_golden.json"), "w"), indent=1)
    print(aes_cases)
    cases = []
    for log_size in (4, 5, 6):
        p = api.prove_bitwise(log_size)
        cases.append({"log_size": log_size, "proof_bytes": len(p), "sha256": hashlib.sha256(p).hexdigest()})
    json.dump({"generator": "oracle/api.py prove_bitwise (numpy restatement; not exported by the reference binary)", "cases": cases},
              open(os.path.join(ROOT, "tests", "golden", "chacha20_block_golden.json"), "w"), indent=1)
    print(cases)
