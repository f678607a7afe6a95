"""ctypes binding of libs2c_b200.so (C ABI: include/s2c_b200.h)."""
import ctypes
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u32p = ctypes.POINTER(ctypes.c_uint32)

# every symbol include/s2c_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "cb_init", "cb_destroy", "cb_last_error", "cb_set_stream", "cb_sync", "cb_launch_count",
    "cb_malloc", "cb_free", "cb_h2d", "cb_d2h", "cb_memset_zero",
    "cb_precompute_twiddles", "cb_interpolate_columns", "cb_evaluate_polynomials", "cb_commit_lde", "cb_lde_packed",
    "cb_comm_unique_id", "cb_comm_init", "cb_comm_destroy", "cb_set_max_cached_tiles", "cb_debug_force_generic_fft", "cb_eval_at_point",
    "cb_merkle_build_leaves", "cb_merkle_leaves_absorb", "cb_merkle_next_layer",
    "cb_generate_secure_powers_rev", "cb_eval_constraints_chacha_stream",
    "cb_accumulate_quotients", "cb_fold_circle_into_line", "cb_fold_line", "cb_grind_blake2s", "cb_gather_rows",
    "cb_gen_trace_chacha_stream",
    "s2c_generate_chacha20_proof", "s2c_generate_aes128_ctr_proof", "s2c_generate_aes256_ctr_proof", "s2c_prove_aes_ctr_raw",
    "s2c_prove_chacha20_raw", "s2c_prove_chacha20_dev", "cb_set_profile", "cb_stage_times", "cb_host_times", "cb_counters",
    "s2c_verify_chacha20_proof", "s2c_verify_aes_ctr_proof", "s2c_verify_chacha20_raw", "s2c_verify_aes_ctr_raw",
    "s2c_prove_chacha20_encrypt", "s2c_prove_aes128_ctr_encrypt", "s2c_prove_aes256_ctr_encrypt",
    "s2c_debug_chacha20_keystream", "s2c_debug_blake2s", "s2c_get_circuits_info", "s2c_free",
    "s2c_prove_chacha20_stream_testdata", "s2c_prove_chacha20_block", "s2c_verify_chacha20_block",
    "s2c_prove_aes128_block", "s2c_verify_aes128_block",
    "cb_bit_reverse", "cb_col_at", "cb_col_set", "cb_batch_inverse_m31", "cb_batch_inverse_qm31", "cb_extend",
    "cb_barycentric_weights", "cb_barycentric_eval_at_point", "cb_precompute_twiddles_coset", "cb_commit_on_layer",
    "cb_accumulate", "cb_lift_and_accumulate", "cb_accumulate_quotients_batches", "cb_aes_ctr_layout", "cb_gen_trace_aes_ctr",
    "cb_gen_logup_interaction_aes_ctr", "cb_logup_finalize_last", "cb_eval_constraints_aes_ctr", "cb_eval_constraints_sbox_table",
]


class BackendError(RuntimeError):
    pass


def lib_path():
    """S2C_B200_LIB overrides the in-tree library (used to compare kernel build variants; same switch as the koffi stub in
    INTEGRATION.md)."""
    return os.environ.get("S2C_B200_LIB") or os.path.join(_HERE, "libs2c_b200.so")


def lib():
    """Load the CUDA backend; fails loudly when it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise BackendError("%s missing: run `make` (or __graft_entry__.build()) -- there is no CPU fallback" % p)
        L = ctypes.CDLL(p)
        L.cb_last_error.restype = ctypes.c_char_p
        L.cb_stage_times.restype = ctypes.c_char_p
        L.cb_counters.restype = ctypes.c_char_p
        L.cb_host_times.restype = ctypes.c_char_p
        L.cb_launch_count.restype = ctypes.c_uint64
        L.s2c_free.argtypes = [ctypes.c_void_p]
        _LIB = L
    return _LIB


def _bytes(b):
    b = bytes(b)
    return (ctypes.c_uint8 * max(len(b), 1)).from_buffer_copy(b if b else b"\0"), len(b)


def _take_json(L, out, n):
    s = ctypes.string_at(out.value, n.value).decode()
    L.s2c_free(out)
    return json.loads(s)


def comm_unique_id():
    """128-byte NCCL unique id (create on one rank, broadcast to the others)."""
    buf = (ctypes.c_uint8 * 128)()
    if lib().cb_comm_unique_id(buf) != 0:
        raise BackendError("NCCL is not available (libnccl.so.2 could not be loaded)")
    return bytes(buf)


class Backend:
    """One backend context = one (GPU, stream).  Thin object wrapper over the cb_* entry points."""

    def __init__(self, device=0):
        self.L = lib()
        self.ctx = ctypes.c_void_p()
        rc = self.L.cb_init(int(device), ctypes.byref(self.ctx))
        if rc != 0:
            raise BackendError("cb_init(device=%d) failed with status %d: no usable CUDA device; no CPU fallback" % (device, rc))
        self.device = device

    def close(self):
        if self.ctx:
            self.L.cb_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise BackendError(self.L.cb_last_error(self.ctx).decode(errors="replace"))

    # ---- memory
    def malloc(self, nbytes):
        p = ctypes.c_void_p()
        self._ck(self.L.cb_malloc(self.ctx, ctypes.c_size_t(nbytes), ctypes.byref(p)))
        return p

    def free(self, p):
        self._ck(self.L.cb_free(self.ctx, p))

    def upload(self, arr):
        """numpy uint32 array -> device pointer."""
        import numpy as np
        a = np.ascontiguousarray(arr, dtype=np.uint32)
        p = self.malloc(a.nbytes)
        self._ck(self.L.cb_h2d(self.ctx, p, a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.nbytes)))
        return p

    def download(self, p, shape):
        import numpy as np
        a = np.empty(shape, dtype=np.uint32)
        self._ck(self.L.cb_d2h(self.ctx, a.ctypes.data_as(ctypes.c_void_p), p, ctypes.c_size_t(a.nbytes)))
        return a

    def sync(self):
        self._ck(self.L.cb_sync(self.ctx))

    def launch_count(self):
        return int(self.L.cb_launch_count(self.ctx))

    def set_profile(self, on=True):
        self.L.cb_set_profile(self.ctx, int(on))

    def stage_times(self):
        s = self.L.cb_stage_times(self.ctx).decode()
        out = {}
        for item in s.split(";"):
            if "=" in item:
                k, v = item.split("=")
                out[k] = out.get(k, 0.0) + float(v)
        return out

    def host_times(self):
        """Host wall-clock ms per stage of the last profiled proof (includes the synchronisations inside the stage)."""
        out = {}
        for item in self.L.cb_host_times(self.ctx).decode().split(";"):
            if "=" in item:
                k, v = item.split("=")
                out[k] = out.get(k, 0.0) + float(v)
        return out

    def comm_init(self, rank, world, unique_id):
        """Join the communicator of the ranks that prove one trace together (unique_id: 128 bytes from comm_unique_id())."""
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.L.cb_comm_init(self.ctx, int(rank), int(world), buf))

    def comm_destroy(self):
        self._ck(self.L.cb_comm_destroy(self.ctx))

    def counters(self):
        s = self.L.cb_counters(self.ctx).decode()
        return {k: int(v) for k, v in (item.split("=") for item in s.split(";") if "=" in item)}

    # ---- product level
    @staticmethod
    def _check_raw(key, key_lens, nonce, counter, pt_len, ct_len, block):
        """The *_raw C entry points take fixed-size key / nonce arrays and ONE length for both buffers: validate here, with
        the reference's messages (wasm_api.rs:475-493, 660-678), so that the C side never reads past a Python object."""
        if len(key) not in key_lens:
            raise BackendError("Key must be %s bytes, got %d" % (" or ".join(str(k) for k in key_lens), len(key)))
        if len(nonce) != 12:
            raise BackendError("Nonce must be 12 bytes, got %d" % len(nonce))
        if pt_len == 0 or pt_len % block:
            raise BackendError("Plaintext must be non-empty multiple of %d bytes, got %d" % (block, pt_len))
        if ct_len is not None and ct_len != pt_len:
            raise BackendError("Ciphertext must be same length as plaintext, got %d vs %d" % (ct_len, pt_len))
        nblk = pt_len // block
        if nblk > 1 and (counter & 0xFFFFFFFF) + nblk - 1 > 0xFFFFFFFF:
            raise BackendError("Counter overflow: counter %d + %d blocks would exceed u32::MAX" % (counter & 0xFFFFFFFF, nblk))

    def prove_chacha20_raw(self, key, nonce, counter, plaintext, ciphertext):
        pb = bytes(plaintext)
        cbuf = bytes(ciphertext)
        self._check_raw(bytes(key), (32,), bytes(nonce), counter, len(pb), len(cbuf), 64)
        kb, _ = _bytes(key)
        nb, _ = _bytes(nonce)
        out = ctypes.POINTER(ctypes.c_uint8)()
        n = ctypes.c_size_t()
        rc = self.L.s2c_prove_chacha20_raw(self.ctx, kb, nb, ctypes.c_uint32(counter & 0xFFFFFFFF), pb, cbuf,
                                           ctypes.c_size_t(len(pb)), ctypes.byref(out), ctypes.byref(n))
        self._ck(rc)
        proof = ctypes.string_at(out, n.value)
        self.L.s2c_free(out)
        return proof

    def prove_chacha20_stream_testdata(self, log_size):
        """The reference's `prove_stream` test-data generator (air_stream.rs:237-289) at `log_size`; returns the proof bytes."""
        out = ctypes.POINTER(ctypes.c_uint8)()
        n = ctypes.c_size_t()
        self._ck(self.L.s2c_prove_chacha20_stream_testdata(self.ctx, int(log_size), ctypes.byref(out), ctypes.byref(n)))
        proof = ctypes.string_at(out, n.value)
        self.L.s2c_free(out)
        return proof

    def prove_aes128_block(self, log_size):
        """The reference's AES-128 block-AIR prover `prove_aes_lookup` (aes/lookup/air.rs:139-260) at `log_size`; proof bytes."""
        out = ctypes.POINTER(ctypes.c_uint8)()
        n = ctypes.c_size_t()
        self._ck(self.L.s2c_prove_aes128_block(self.ctx, int(log_size), ctypes.byref(out), ctypes.byref(n)))
        proof = ctypes.string_at(out, n.value)
        self.L.s2c_free(out)
        return proof

    def set_max_cached_tiles(self, n_tiles):
        """Cap on LDE tiles the streaming prover keeps between its two passes (-1: as many as memory allows)."""
        self._ck(self.L.cb_set_max_cached_tiles(self.ctx, int(n_tiles)))

    def prove_chacha20_block(self, log_size):
        """The reference's block-AIR prover `prove_bitwise` (chacha/bitwise/air.rs:53-137) at `log_size`; returns the proof bytes."""
        out = ctypes.POINTER(ctypes.c_uint8)()
        n = ctypes.c_size_t()
        self._ck(self.L.s2c_prove_chacha20_block(self.ctx, int(log_size), ctypes.byref(out), ctypes.byref(n)))
        proof = ctypes.string_at(out, n.value)
        self.L.s2c_free(out)
        return proof

    def set_stream(self, cuda_stream_ptr):
        """Run this context's work on an existing CUDA stream (e.g. torch.cuda.Stream().cuda_stream)."""
        self._ck(self.L.cb_set_stream(self.ctx, ctypes.c_void_p(cuda_stream_ptr)))

    def prove_chacha20_ptr(self, key, nonce, counter, pt_ptr, ct_ptr, nbytes, on_device=False, pt_hash=None, ct_hash=None):
        """Raw-pointer form: host pointers (e.g. pinned memory) or, with on_device=True, device pointers plus the two
        public-input hashes.  Returns the proof bytes.  Both buffers must hold `nbytes` bytes (the caller owns them)."""
        self._check_raw(bytes(key), (32,), bytes(nonce), counter, nbytes, None, 64)
        if on_device and (pt_hash is None) != (ct_hash is None):
            raise BackendError("give both public-input hashes or neither")
        if on_device and pt_hash is not None and (len(bytes(pt_hash)) != 32 or len(bytes(ct_hash)) != 32):
            raise BackendError("public-input hashes are 32 bytes each")
        kb, _ = _bytes(key)
        nb, _ = _bytes(nonce)
        out = ctypes.POINTER(ctypes.c_uint8)()
        n = ctypes.c_size_t()
        if on_device:
            rc = self.L.s2c_prove_chacha20_dev(self.ctx, kb, nb, ctypes.c_uint32(counter & 0xFFFFFFFF), ctypes.c_void_p(pt_ptr),
                                               ctypes.c_void_p(ct_ptr), ctypes.c_size_t(nbytes),
                                               bytes(pt_hash) if pt_hash is not None else None,
                                               bytes(ct_hash) if ct_hash is not None else None, ctypes.byref(out), ctypes.byref(n))
        else:
            rc = self.L.s2c_prove_chacha20_raw(self.ctx, kb, nb, ctypes.c_uint32(counter & 0xFFFFFFFF), ctypes.c_void_p(pt_ptr),
                                               ctypes.c_void_p(ct_ptr), ctypes.c_size_t(nbytes), ctypes.byref(out), ctypes.byref(n))
        self._ck(rc)
        proof = ctypes.string_at(out, n.value)
        self.L.s2c_free(out)
        return proof

    def generate_chacha20_proof(self, key, nonce, counter, plaintext, ciphertext):
        key, nonce, pb, cbuf = bytes(key), bytes(nonce), bytes(plaintext), bytes(ciphertext)
        out = ctypes.c_void_p()
        n = ctypes.c_size_t()
        rc = self.L.s2c_generate_chacha20_proof(self.ctx, key, ctypes.c_size_t(len(key)), nonce, ctypes.c_size_t(len(nonce)),
                                                ctypes.c_uint32(counter & 0xFFFFFFFF), pb, ctypes.c_size_t(len(pb)), cbuf,
                                                ctypes.c_size_t(len(cbuf)), ctypes.byref(out), ctypes.byref(n))
        if not out:
            raise BackendError("s2c_generate_chacha20_proof failed with status %d" % rc)
        return _take_json(self.L, out, n)


    def _generate_aes(self, fn, key, nonce, counter, plaintext, ciphertext):
        key, nonce, pb, cbuf = bytes(key), bytes(nonce), bytes(plaintext), bytes(ciphertext)
        out = ctypes.c_void_p()
        n = ctypes.c_size_t()
        rc = fn(self.ctx, key, ctypes.c_size_t(len(key)), nonce, ctypes.c_size_t(len(nonce)), ctypes.c_uint32(counter & 0xFFFFFFFF),
                pb, ctypes.c_size_t(len(pb)), cbuf, ctypes.c_size_t(len(cbuf)), ctypes.byref(out), ctypes.byref(n))
        if not out:
            raise BackendError("AES-CTR proof call failed with status %d" % rc)
        return _take_json(self.L, out, n)

    def generate_aes128_ctr_proof(self, key, nonce, counter, plaintext, ciphertext):
        """wasm_api.rs:652 generate_aes128_ctr_proof -> dict (same keys as the reference's JSON)."""
        return self._generate_aes(self.L.s2c_generate_aes128_ctr_proof, key, nonce, counter, plaintext, ciphertext)

    def generate_aes256_ctr_proof(self, key, nonce, counter, plaintext, ciphertext):
        """wasm_api.rs:776 generate_aes256_ctr_proof."""
        return self._generate_aes(self.L.s2c_generate_aes256_ctr_proof, key, nonce, counter, plaintext, ciphertext)

    def prove_chacha20_encrypt(self, key, nonce, counter, plaintext, ciphertext):
        """wasm_api.rs:61 prove_chacha20_encrypt: prove on the GPU, verify on the host -> {"success", "blocks", "algorithm"}."""
        return self._generate_aes(self.L.s2c_prove_chacha20_encrypt, key, nonce, counter, plaintext, ciphertext)

    def prove_aes128_ctr_encrypt(self, key, nonce, counter, plaintext, ciphertext):
        """wasm_api.rs:210 prove_aes128_ctr_encrypt."""
        return self._generate_aes(self.L.s2c_prove_aes128_ctr_encrypt, key, nonce, counter, plaintext, ciphertext)

    def prove_aes256_ctr_encrypt(self, key, nonce, counter, plaintext, ciphertext):
        """wasm_api.rs:343 prove_aes256_ctr_encrypt."""
        return self._generate_aes(self.L.s2c_prove_aes256_ctr_encrypt, key, nonce, counter, plaintext, ciphertext)

    def prove_aes_ctr_raw(self, key, nonce, counter, plaintext, ciphertext):
        key, nonce, pb, cbuf = bytes(key), bytes(nonce), bytes(plaintext), bytes(ciphertext)
        self._check_raw(key, (16, 32), nonce, counter, len(pb), len(cbuf), 16)
        out = ctypes.POINTER(ctypes.c_uint8)()
        n = ctypes.c_size_t()
        self._ck(self.L.s2c_prove_aes_ctr_raw(self.ctx, len(key), key, nonce, ctypes.c_uint32(counter & 0xFFFFFFFF), pb, cbuf,
                                              ctypes.c_size_t(len(pb)), ctypes.byref(out), ctypes.byref(n)))
        proof = ctypes.string_at(out, n.value)
        self.L.s2c_free(out)
        return proof


_DEFAULT = None


def _default():
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Backend(int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("S2C_USE_LOCAL_RANK") else 0)
    return _DEFAULT


def generate_chacha20_proof(key, nonce, counter, plaintext, ciphertext):
    """wasm_api.rs:467 generate_chacha20_proof -> dict (same keys as the reference's JSON)."""
    return _default().generate_chacha20_proof(key, nonce, counter, plaintext, ciphertext)


def prove_chacha20_raw(key, nonce, counter, plaintext, ciphertext):
    return _default().prove_chacha20_raw(key, nonce, counter, plaintext, ciphertext)


def generate_aes128_ctr_proof(key, nonce, counter, plaintext, ciphertext):
    return _default().generate_aes128_ctr_proof(key, nonce, counter, plaintext, ciphertext)


def generate_aes256_ctr_proof(key, nonce, counter, plaintext, ciphertext):
    return _default().generate_aes256_ctr_proof(key, nonce, counter, plaintext, ciphertext)


def _verify_json(name, proof_b64, nonce, counter, plaintext, ciphertext):
    L = lib()
    pb64 = proof_b64.encode() if isinstance(proof_b64, str) else bytes(proof_b64)
    nonce, pb, cbuf = bytes(nonce), bytes(plaintext), bytes(ciphertext)
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    rc = getattr(L, name)(pb64, ctypes.c_size_t(len(pb64)), nonce, ctypes.c_size_t(len(nonce)), ctypes.c_uint32(counter & 0xFFFFFFFF),
                          pb, ctypes.c_size_t(len(pb)), cbuf, ctypes.c_size_t(len(cbuf)), ctypes.byref(out), ctypes.byref(n))
    if not out:
        raise BackendError("%s failed with status %d" % (name, rc))
    return _take_json(L, out, n)


def verify_chacha20_block(proof):
    """verify_bitwise (chacha/bitwise/air.rs:139-171) on proof bytes: "" when the proof verifies, else the error rendering."""
    L = lib()
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    pb = bytes(proof)
    L.s2c_verify_chacha20_block(pb, ctypes.c_size_t(len(pb)), ctypes.byref(out), ctypes.byref(n))
    msg = ctypes.string_at(out, n.value).decode() if out else ""
    if out:
        L.s2c_free(out)
    return msg


def verify_aes128_block(proof):
    """verify_aes_lookup (aes/lookup/air.rs:262-305) on proof bytes: "" when the proof verifies, else the error rendering."""
    L = lib()
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    pb = bytes(proof)
    L.s2c_verify_aes128_block(pb, ctypes.c_size_t(len(pb)), ctypes.byref(out), ctypes.byref(n))
    msg = ctypes.string_at(out, n.value).decode() if out else ""
    if out:
        L.s2c_free(out)
    return msg


def verify_chacha20_proof(proof_b64, nonce, counter, plaintext, ciphertext):
    """wasm_api.rs:609 verify_chacha20_proof -> {"valid": True, "algorithm": ...} | {"valid": False, "error": ...} | {"error": ...}.
    Host-side (csrc/verify.cu); needs no GPU."""
    return _verify_json("s2c_verify_chacha20_proof", proof_b64, nonce, counter, plaintext, ciphertext)


def verify_aes_ctr_proof(proof_b64, nonce, counter, plaintext, ciphertext):
    """wasm_api.rs:904 verify_aes_ctr_proof (AES-128 and AES-256)."""
    return _verify_json("s2c_verify_aes_ctr_proof", proof_b64, nonce, counter, plaintext, ciphertext)


def _verify_raw(name, proof, nonce, counter, plaintext, ciphertext):
    L = lib()
    proof, nonce, pb, cbuf = bytes(proof), bytes(nonce), bytes(plaintext), bytes(ciphertext)
    if len(nonce) != 12:
        raise BackendError("Nonce must be 12 bytes, got %d" % len(nonce))
    err = ctypes.c_void_p()
    rc = getattr(L, name)(proof, ctypes.c_size_t(len(proof)), nonce, ctypes.c_uint32(counter & 0xFFFFFFFF), pb, ctypes.c_size_t(len(pb)),
                          cbuf, ctypes.c_size_t(len(cbuf)), ctypes.byref(err))
    msg = None
    if err:
        msg = ctypes.string_at(err.value).decode()
        L.s2c_free(err)
    return rc == 0, msg


def verify_chacha20_raw(proof, nonce, counter, plaintext, ciphertext):
    """bincode StreamProof bytes -> (valid, error rendering or None)."""
    return _verify_raw("s2c_verify_chacha20_raw", proof, nonce, counter, plaintext, ciphertext)


def verify_aes_ctr_raw(proof, nonce, counter, plaintext, ciphertext):
    return _verify_raw("s2c_verify_aes_ctr_raw", proof, nonce, counter, plaintext, ciphertext)


def prove_chacha20_encrypt(key, nonce, counter, plaintext, ciphertext):
    return _default().prove_chacha20_encrypt(key, nonce, counter, plaintext, ciphertext)


def prove_aes128_ctr_encrypt(key, nonce, counter, plaintext, ciphertext):
    return _default().prove_aes128_ctr_encrypt(key, nonce, counter, plaintext, ciphertext)


def prove_aes256_ctr_encrypt(key, nonce, counter, plaintext, ciphertext):
    return _default().prove_aes256_ctr_encrypt(key, nonce, counter, plaintext, ciphertext)


def debug_chacha20_keystream(key, nonce, counter):
    L = lib()
    key, nonce = bytes(key), bytes(nonce)
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    L.s2c_debug_chacha20_keystream(key, ctypes.c_size_t(len(key)), nonce, ctypes.c_size_t(len(nonce)),
                                   ctypes.c_uint32(counter), ctypes.byref(out), ctypes.byref(n))
    return _take_json(L, out, n)


def get_circuits_info():
    L = lib()
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    L.s2c_get_circuits_info(ctypes.byref(out), ctypes.byref(n))
    return _take_json(L, out, n)
