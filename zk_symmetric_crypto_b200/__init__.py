"""B200-native stwo prover backend for reclaimprotocol/zk-symmetric-crypto (ChaCha20 / AES-CTR AIRs).

The compute path is hand-written CUDA for sm_100a in `csrc/`, exposed through the C ABI `include/s2c_b200.h`
(libs2c_b200.so).  This package is the thin host-side mirror of the reference's operator interface
(/root/reference/js/src/stwo/operator.ts:87-191, /root/reference/stwo/src/wasm_api.rs): same function names,
argument meaning and JSON results.  There is NO CPU fallback: importing works without a GPU, any proving call
raises `BackendError` if libs2c_b200.so or a CUDA device is missing (verification is host work in the same library).
"""
from .backend import (BackendError, Backend, lib, lib_path, generate_chacha20_proof, prove_chacha20_raw,
                      generate_aes128_ctr_proof, generate_aes256_ctr_proof,
                      verify_chacha20_proof, verify_aes_ctr_proof, verify_chacha20_raw, verify_aes_ctr_raw, verify_chacha20_block, verify_aes128_block,
                      prove_chacha20_encrypt, prove_aes128_ctr_encrypt, prove_aes256_ctr_encrypt,
                      debug_chacha20_keystream, get_circuits_info, EXPORTED_SYMBOLS)
from .operator import make_stwo_zk_operator

__all__ = ["BackendError", "Backend", "lib", "lib_path", "generate_chacha20_proof", "prove_chacha20_raw", "generate_aes128_ctr_proof", "generate_aes256_ctr_proof",
           "verify_chacha20_proof", "verify_aes_ctr_proof", "verify_chacha20_raw", "verify_aes_ctr_raw", "verify_chacha20_block", "verify_aes128_block",
           "prove_chacha20_encrypt", "prove_aes128_ctr_encrypt", "prove_aes256_ctr_encrypt",
           "debug_chacha20_keystream", "get_circuits_info", "make_stwo_zk_operator", "EXPORTED_SYMBOLS"]
