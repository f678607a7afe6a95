"""ctypes front-end for oracle/_ref/libs2c_ref.so -- the reference's own prover/verifier.

TEST INFRASTRUCTURE ONLY.  libs2c_ref.so is the reference's shipped WASM build
(`/root/reference/resources/stwo/s2circuits_bg.wasm`; Rust source `/root/reference/stwo/src/wasm_api.rs`)
translated to C by oracle/wasm2c.py and compiled by oracle/Makefile.  This module plays the role of
the wasm-bindgen glue `/root/reference/js/src/stwo/s2circuits.cjs:1-330`: copy byte arrays into the
module's linear memory with `__wbindgen_malloc`, call the export, read the `(ptr,len)` UTF-8 result,
release it with `__wbindgen_free`.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline/reference arm may import this.
"""
import ctypes
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libs2c_ref.so")
_lib = None


def available():
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libs2c_ref.so missing: run `make -C oracle ref` where /root/reference exists")
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.w2c_memory.restype = ctypes.c_void_p
        _lib.w2c_last_trap.restype = ctypes.c_char_p
        _lib.w2c_memory_bytes.restype = ctypes.c_uint64
    return _lib


def _call(name, *args):
    L = lib()
    out = (ctypes.c_uint64 * 4)()
    fn = getattr(L, "w2c_" + name)
    rc = fn(*[ctypes.c_uint32(a) for a in args], out)
    if rc != 0:
        raise RuntimeError("reference trapped in %s: %s" % (name, L.w2c_last_trap().decode(errors="replace")))
    return out


def _mem():
    return lib().w2c_memory()


def _pass_bytes(b):
    n = len(b)
    ptr = _call("__wbindgen_malloc", n, 1)[0] if n else 1
    if n:
        ctypes.memmove(_mem() + ptr, bytes(b), n)
    return int(ptr), n


def _take_string(out):
    ptr, n = int(out[0]), int(out[1])
    s = ctypes.string_at(_mem() + ptr, n)
    _call("__wbindgen_free", ptr, n, 1)
    return s.decode()


def _prove_like(name, key, nonce, counter, plaintext, ciphertext):
    p0, l0 = _pass_bytes(key)
    p1, l1 = _pass_bytes(nonce)
    p2, l2 = _pass_bytes(plaintext)
    p3, l3 = _pass_bytes(ciphertext)
    return json.loads(_take_string(_call(name, p0, l0, p1, l1, counter & 0xFFFFFFFF, p2, l2, p3, l3)))


def generate_chacha20_proof(key, nonce, counter, plaintext, ciphertext):
    """wasm_api.rs:467 -- returns the parsed JSON ({"success","blocks","algorithm","proof",...} or {"error"})."""
    return _prove_like("generate_chacha20_proof", key, nonce, counter, plaintext, ciphertext)


def generate_aes128_ctr_proof(key, nonce, counter, plaintext, ciphertext):
    return _prove_like("generate_aes128_ctr_proof", key, nonce, counter, plaintext, ciphertext)


def generate_aes256_ctr_proof(key, nonce, counter, plaintext, ciphertext):
    return _prove_like("generate_aes256_ctr_proof", key, nonce, counter, plaintext, ciphertext)


def prove_chacha20_encrypt(key, nonce, counter, plaintext, ciphertext):
    return _prove_like("prove_chacha20_encrypt", key, nonce, counter, plaintext, ciphertext)


def prove_aes128_ctr_encrypt(key, nonce, counter, plaintext, ciphertext):
    return _prove_like("prove_aes128_ctr_encrypt", key, nonce, counter, plaintext, ciphertext)


def prove_aes256_ctr_encrypt(key, nonce, counter, plaintext, ciphertext):
    return _prove_like("prove_aes256_ctr_encrypt", key, nonce, counter, plaintext, ciphertext)


def verify_chacha20_proof(proof_b64, nonce, counter, plaintext, ciphertext):
    """wasm_api.rs:609 -- (proof string, nonce, counter, plaintext, ciphertext) -> {"valid": bool, ...}."""
    return _prove_like("verify_chacha20_proof", proof_b64.encode(), nonce, counter, plaintext, ciphertext)


def verify_aes_ctr_proof(proof_b64, nonce, counter, plaintext, ciphertext):
    return _prove_like("verify_aes_ctr_proof", proof_b64.encode(), nonce, counter, plaintext, ciphertext)


def debug_chacha20_keystream(key, nonce, counter):
    p0, l0 = _pass_bytes(key)
    p1, l1 = _pass_bytes(nonce)
    return json.loads(_take_string(_call("debug_chacha20_keystream", p0, l0, p1, l1, counter)))


def get_circuits_info():
    return json.loads(_take_string(_call("get_circuits_info")))


def memory_bytes():
    return int(lib().w2c_memory_bytes())
