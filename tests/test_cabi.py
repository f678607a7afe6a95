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
