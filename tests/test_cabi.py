"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/s2c_b200.h declares, host-only entry points
work, and compute entry points fail loudly without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

import zk_symmetric_crypto_b200 as z

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "s2c_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:cb|s2c)_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = z.lib()
    declared = header_symbols()
    assert len(declared) >= 30
    for s in declared:
        assert hasattr(L, s), "libs2c_b200.so does not export %s" % s
    assert sorted(z.EXPORTED_SYMBOLS) == declared


def test_host_only_entry_points():
    assert z.get_circuits_info()["chacha20"] == {"block_bytes": 64, "cols": 33280, "constraints": 54784, "key_bytes": 32}
    ks = z.debug_chacha20_keystream(bytes(range(32)), bytes([0, 0, 0, 9, 0, 0, 0, 0x4A, 0, 0, 0, 0]), 1)
    assert ks["keystream_hex"].startswith("10f1e7e4d13b5915500fdd1fa32071c4")   # RFC 7539 2.3.2


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(z.BackendError):
        z.Backend(0)
    with pytest.raises(z.BackendError):
        z.generate_chacha20_proof(bytes(32), bytes(12), 0, bytes(64), bytes(64))


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "zk_symmetric_crypto_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import ref_wasm" not in txt and "stwo_core" not in txt and "oracle/" not in txt.replace("oracle/trace_blake.py", "").replace("oracle/prover.py", ""), f


def test_raw_entry_points_validate_lengths_before_reaching_c():
    """s2c_*_raw take key[32] / nonce[12] and one length for both buffers: the Python binding rejects anything that would make the
    C side read past a Python object, with the reference's messages (wasm_api.rs:475-493, 660-678)."""
    ck = z.Backend._check_raw
    ck(bytes(32), (32,), bytes(12), 0, 64, 64, 64)
    for args, msg in (((bytes(31), (32,), bytes(12), 0, 64, 64, 64), "Key must be 32 bytes, got 31"),
                      ((bytes(24), (16, 32), bytes(12), 0, 16, 16, 16), "Key must be 16 or 32 bytes, got 24"),
                      ((bytes(32), (32,), bytes(11), 0, 64, 64, 64), "Nonce must be 12 bytes, got 11"),
                      ((bytes(32), (32,), bytes(12), 0, 0, 0, 64), "Plaintext must be non-empty multiple of 64 bytes, got 0"),
                      ((bytes(32), (32,), bytes(12), 0, 100, 100, 64), "Plaintext must be non-empty multiple of 64 bytes, got 100"),
                      ((bytes(32), (32,), bytes(12), 0, 128, 64, 64), "Ciphertext must be same length as plaintext, got 64 vs 128"),
                      ((bytes(32), (32,), bytes(12), 0xFFFFFFFF, 128, 128, 64), "Counter overflow: counter 4294967295 + 2 blocks would exceed u32::MAX")):
        with pytest.raises(z.BackendError, match=msg.replace("+", r"\+")):
            ck(*args)
    ck(bytes(32), (32,), bytes(12), 0xFFFFFFFF, 64, 64, 64)   # a single block at the last counter value is fine


def test_host_blake2s_matches_hashlib():
    """The library's host Blake2s (scalar tail + SIMD blocks with run-time dispatch) against hashlib on every length class:
    empty, sub-block, block boundaries, the SIMD threshold (4 full non-final blocks) and a long message."""
    import ctypes
    import hashlib
    import random
    L = z.lib()
    rng = random.Random(7)
    for n in [0, 1, 31, 63, 64, 65, 127, 128, 255, 256, 257, 319, 320, 321, 1000, 4096, 65536 + 7, 532608 + 32]:
        data = bytes(rng.getrandbits(8) for _ in range(n))
        out = (ctypes.c_uint8 * 32)()
        assert L.s2c_debug_blake2s(data, ctypes.c_size_t(n), out) == 0
        assert bytes(out) == hashlib.blake2s(data, digest_size=32).digest(), n
