"""Generates tests/golden/aes_ctr_golden.json from the REFERENCE ITSELF (oracle/_ref/libs2c_ref.so), like make_golden.py.

Cases: AES-128/256-CTR proofs at log_size 8 (all columns one size), 9 and 10 (log-8 S-box table columns lifted into the
larger trees: mixed-size Merkle leaves, lifted composition accumulation, periodicity samples), plus an invalid witness."""
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
import aes_air as aa

CASES = [  # (name, key_len, n_blocks, seed, corrupt)
    ("aes128_fips_key_5blocks", 16, 5, None, False),
    ("aes128_1block", 16, 1, 1, False),
    ("aes256_3blocks", 32, 3, 2, False),
    ("aes128_300blocks_log9", 16, 300, 3, False),
    ("aes256_257blocks_log9", 32, 257, 4, False),
    ("aes128_600blocks_log10", 16, 600, 5, False),
    ("aes128_bad_ciphertext", 16, 2, 6, True),
]


def aes_case_inputs(key_len, n_blocks, seed, corrupt=False):
    if seed is None:
        key = bytes(range(key_len)); nonce = bytes(range(100, 112)); counter = 7
        pt = bytes((i * 7 + 3) & 0xFF for i in range(16 * n_blocks))
    else:
        rng = np.random.default_rng(1000 + seed)
        key = rng.bytes(key_len); nonce = rng.bytes(12); counter = int(rng.integers(0, 2 ** 31)); pt = rng.bytes(16 * n_blocks)
    ct = aa.ctr_encrypt(key, nonce, counter, pt)
    if corrupt:
        ct = bytes([ct[0] ^ 1]) + ct[1:]
    return key, nonce, counter, pt, ct


def main():
    out = []
    for name, kl, nb, seed, corrupt in CASES:
        key, nonce, counter, pt, ct = aes_case_inputs(kl, nb, seed, corrupt)
        fn = ref_wasm.generate_aes128_ctr_proof if kl == 16 else ref_wasm.generate_aes256_ctr_proof
        res = fn(key, nonce, counter, pt, ct)
        entry = {"name": name, "key_len": kl, "n_blocks": nb, "seed": seed, "corrupt": corrupt}
        if "error" in res:
            entry["error"] = res["error"]
        else:
            pb = base64.b64decode(res["proof"])
            entry.update(proof_len=len(pb), proof_sha256=hashlib.sha256(pb).hexdigest(),
                         b64_sha256=hashlib.sha256(res["proof"].encode()).hexdigest(),
                         proof_size_bytes=res["proof_size_bytes"], blocks=res["blocks"], algorithm=res["algorithm"],
                         roots=[pb[169 + 32 * i:201 + 32 * i].hex() for i in range(4)],
                         log_size=struct.unpack_from("<I", pb, 0)[0])
            v = ref_wasm.verify_aes_ctr_proof(res["proof"], nonce, counter, pt, ct)
            assert v.get("valid") is True, v
        out.append(entry)
        print(entry)
    json.dump({"generator": "tests/golden/make_golden_aes.py", "reference": "resources/stwo/s2circuits_bg.wasm via oracle/_ref",
               "cases": out}, open(os.path.join(HERE, "aes_ctr_golden.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
