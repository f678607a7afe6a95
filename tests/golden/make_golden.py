"""Generates tests/golden/chacha20_golden.json from the REFERENCE ITSELF (oracle/_ref/libs2c_ref.so = the reference's
shipped WASM prover compiled natively; needs /root/reference, so it only runs in the build container).

Each case stores the deterministic inputs (numpy default_rng seed), and for the reference's proof: byte length, sha256,
the three commitment roots, the proof-of-work nonce, and sha256 of the base64 string.  The parity tests regenerate the
inputs from the seed and compare the oracle's / the CUDA backend's proof against these digests.
"""
import base64
import hashlib
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import numpy as np
import ref_wasm
import chacha_air as ca

CASES = [  # (name, n_blocks, seed)
    ("rfc7539_1block", 1, None),
    ("rand_2blocks", 2, 0),
    ("rand_16blocks", 16, 1),
    ("rand_17blocks_log5", 17, 2),
    ("rand_64blocks_log6", 64, 4),
    ("rand_128blocks_log7", 128, 5),
    ("rand_40blocks_constraints_fail", 40, 3),
]


# Sizes only the GPU parity tests use (`python make_golden.py --large` regenerates them; ~2.5 minutes of reference proving):
# log 12 (whole-column FFT kernel, byte-table kernels) and log 13 = 8,192 blocks, the largest trace the reference's wasm32
# build can hold (3.06 GB of its 4 GB address space; 16,384 blocks trap with out-of-memory).  Log 13 is the smallest size that
# runs the three-pass FFT kernels, the partial tile cache with half-tile recomputation and the cached-tile query path.
LARGE_CASES = [
    ("rand_4096blocks_log12", 4096, 100 + 4096),
    ("rand_8192blocks_log13", 8192, 100 + 8192),
]


def case_inputs(n_blocks, seed):
    if seed is None:  # RFC 7539 2.3.2 key/nonce/counter (chacha/block.rs:116-139)
        key = bytes(range(32)); nonce = bytes([0, 0, 0, 9, 0, 0, 0, 0x4A, 0, 0, 0, 0]); counter = 1
        pt = bytes((i * 7) & 0xFF for i in range(64 * n_blocks))
    else:
        rng = np.random.default_rng(seed)
        key = rng.bytes(32); nonce = rng.bytes(12); counter = int(rng.integers(0, 2 ** 31)); pt = rng.bytes(64 * n_blocks)
    ks = ca.chacha20_keystream_bytes(key, nonce, counter, n_blocks)
    ct = bytes(a ^ b for a, b in zip(pt, ks))
    return key, nonce, counter, pt, ct


def main():
    path = os.path.join(HERE, "chacha20_golden.json")
    large = "--large" in sys.argv
    prev = json.load(open(path)) if os.path.exists(path) else {}
    out = []
    for name, nb, seed in (LARGE_CASES if large else CASES):
        key, nonce, counter, pt, ct = case_inputs(nb, seed)
        res = ref_wasm.generate_chacha20_proof(key, nonce, counter, pt, ct)
        entry = {"name": name, "n_blocks": nb, "seed": seed}
        if "error" in res:
            entry["error"] = res["error"]
        else:
            pb = base64.b64decode(res["proof"])
            entry.update(proof_len=len(pb), proof_sha256=hashlib.sha256(pb).hexdigest(),
                         b64_sha256=hashlib.sha256(res["proof"].encode()).hexdigest(),
                         proof_size_bytes=res["proof_size_bytes"], blocks=res["blocks"],
                         roots=[pb[117 + 32 * i:149 + 32 * i].hex() for i in range(3)],
                         log_size=struct.unpack_from("<I", pb, 0)[0])
            v = ref_wasm.verify_chacha20_proof(res["proof"], nonce, counter, pt, ct)
            assert v.get("valid") is True, v
        out.append(entry)
        print(entry)
    doc = {"generator": "tests/golden/make_golden.py", "reference": "resources/stwo/s2circuits_bg.wasm via oracle/_ref",
           "cases": prev.get("cases", []), "large_cases": prev.get("large_cases", [])}
    doc["large_cases" if large else "cases"] = out
    json.dump(doc, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
