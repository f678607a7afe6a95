"""CPU restatement (numpy) of the reference's ChaCha20 *stream* AIR: witness generation and constraint evaluation.

TEST INFRASTRUCTURE ONLY.

Follows (file:line relative to /root/reference/stwo/src):
  * trace generator     chacha/bitwise/gen_stream.rs:83-221 (append_u32_bits, build_state, generate, quarter_round,
                        add_u32, xor_rotl_u32) and :226-261 (generate_stream_trace: default all-zero rows)
  * constraint sequence chacha/bitwise/constraints_stream.rs:20-70 (eval), :74-82 (next_u32), :85-101 (quarter_round),
                        :104-131 (add_u32), :134-152 (xor_rotl_u32), :179-189 (xor_u32_no_trace)
  * native cipher       chacha/block.rs:95, chacha/quarter_round.rs:19 (KAT: RFC 7539 2.3.2 at block.rs:116-139)
Column/constraint counts (33,280 / 54,784) are confirmed by the reference's get_circuits_info().
"""
import numpy as np

from stwo_core import P, U64, m_add, m_sub, m_mul, m_neg, QM31, q_mul_m31

N_COLS = 33280
N_CONSTRAINTS = 54784
CONSTANTS = (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)
QR_SCHEDULE = ((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15),
               (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14))
_M32 = np.uint64(0xFFFFFFFF)


# ---------------------------------------------------------------- native cipher (chacha/block.rs, quarter_round.rs)
def _rotl(x, r):
    return ((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & _M32


def chacha20_block_words(key_words, counter, nonce_words):
    """Keystream words for one block (python ints).  chacha/block.rs:95 chacha20_block_from_key."""
    init = list(CONSTANTS) + list(key_words) + [counter & 0xFFFFFFFF] + list(nonce_words)
    v = list(init)

    def rotl(x, r):
        return ((x << r) | (x >> (32 - r))) & 0xFFFFFFFF

    def qr(a, b, c, d):
        v[a] = (v[a] + v[b]) & 0xFFFFFFFF; v[d] = rotl(v[d] ^ v[a], 16)
        v[c] = (v[c] + v[d]) & 0xFFFFFFFF; v[b] = rotl(v[b] ^ v[c], 12)
        v[a] = (v[a] + v[b]) & 0xFFFFFFFF; v[d] = rotl(v[d] ^ v[a], 8)
        v[c] = (v[c] + v[d]) & 0xFFFFFFFF; v[b] = rotl(v[b] ^ v[c], 7)

    for _ in range(10):
        for q in QR_SCHEDULE:
            qr(*q)
    return [(v[i] + init[i]) & 0xFFFFFFFF for i in range(16)]


def chacha20_keystream_bytes(key, nonce, counter, n_blocks):
    import struct
    kw = struct.unpack("<8I", key)
    nw = struct.unpack("<3I", nonce)
    out = b""
    for b in range(n_blocks):
        out += struct.pack("<16I", *chacha20_block_words(kw, counter + b, nw))
    return out


# ---------------------------------------------------------------- witness (gen_stream.rs)
def generate_stream_trace(log_size, key, nonce, counters, plaintext, ciphertext, n_input_rows=None):
    """n_input_rows: rows covered by caller-provided inputs; only those enter the validity flag
    (gen_stream.rs:240-250 ignores the result for default rows).  All arguments are per-row uint arrays: key[N,8], nonce[N,3], counters[N], plaintext[N,16], ciphertext[N,16]
    (the reference splats key/nonce over the 16 lanes of a vec-row and zero-fills rows beyond the inputs,
    gen_stream.rs:237-250 -- callers build those arrays).  Returns (trace[N_COLS, N] uint64 bits, valid)."""
    n = 1 << log_size
    key = np.asarray(key, dtype=U64).reshape(n, 8)
    nonce = np.asarray(nonce, dtype=U64).reshape(n, 3)
    counters = np.asarray(counters, dtype=U64).reshape(n)
    pt = np.asarray(plaintext, dtype=U64).reshape(n, 16)
    ct = np.asarray(ciphertext, dtype=U64).reshape(n, 16)
    trace = np.empty((N_COLS, n), dtype=U64)
    col = [0]
    shifts = np.arange(32, dtype=U64)[:, None]

    def append_u32_bits(val):
        trace[col[0]:col[0] + 32] = (val[None, :] >> shifts) & U64(1)
        col[0] += 32

    def add_u32(a, b):
        res = (a + b) & _M32
        append_u32_bits(res)
        carry = np.zeros(n, dtype=U64)
        for i in range(32):
            s = ((a >> U64(i)) & U64(1)) + ((b >> U64(i)) & U64(1)) + carry
            carry = s >> U64(1)
            trace[col[0]] = carry
            col[0] += 1
        return res

    def xor_rotl_u32(a, b, r):
        res = _rotl(a ^ b, r)
        append_u32_bits(res)
        return res

    init = [np.full(n, c, dtype=U64) for c in CONSTANTS] + [key[:, i] for i in range(8)] + [counters] + \
           [nonce[:, i] for i in range(3)]
    for s in init:
        append_u32_bits(s)
    v = list(init)
    for _ in range(10):
        for (a, b, c, d) in QR_SCHEDULE:
            v[a] = add_u32(v[a], v[b]); v[d] = xor_rotl_u32(v[a], v[d], 16)
            v[c] = add_u32(v[c], v[d]); v[b] = xor_rotl_u32(v[c], v[b], 12)
            v[a] = add_u32(v[a], v[b]); v[d] = xor_rotl_u32(v[a], v[d], 8)
            v[c] = add_u32(v[c], v[d]); v[b] = xor_rotl_u32(v[c], v[b], 7)
    ks = [add_u32(v[i], init[i]) for i in range(16)]
    for i in range(16):
        append_u32_bits(pt[:, i])
    valid = True
    for i in range(16):
        append_u32_bits(ct[:, i])
        m = n if n_input_rows is None else n_input_rows
        if np.any((ks[i][:m] ^ pt[:m, i]) != ct[:m, i]):
            valid = False
    assert col[0] == N_COLS
    return trace, valid


# ---------------------------------------------------------------- constraints (constraints_stream.rs)
class _Acc:
    """sum_k alpha^(K-1-k) * C_k(row): add_constraint semantics of the framework's domain evaluator."""

    def __init__(self, n_rows, alpha_pows_rev, ext):
        self.acc = np.zeros((n_rows, 4), dtype=U64)
        self.k = 0
        self.apr = alpha_pows_rev           # [K,4]: entry k = alpha^(K-1-k)
        self.ext = ext

    def emit(self, cmat):
        m = cmat.shape[0]
        co = self.apr[self.k:self.k + m]                       # [m,4]
        if self.ext:                                           # cmat [m,R,4] QM31 values
            from stwo_core import q_mul
            prod = q_mul(cmat, co[:, None, :])
        else:                                                  # cmat [m,R] M31 values
            prod = (cmat[:, :, None] * co[:, None, :]) % U64(P)
        self.acc = (self.acc + prod.sum(axis=0)) % U64(P)
        self.k += m


N_COLS_BLOCK = 32256          # block AIR (chacha/bitwise/{gen,constraints,air}.rs): no plaintext / ciphertext columns
N_CONSTRAINTS_BLOCK = 53248   # ... and no plaintext / ciphertext booleans, no xor equalities


def generate_block_trace(log_size, key_words, nonce_words):
    """chacha/bitwise/gen.rs generate_trace on prove_bitwise's inputs (air.rs:66-90): initial state = build_state(key, 0, nonce)
    with word 12 = the row index.  The columns are the first 32,256 columns of the stream trace."""
    n = 1 << log_size
    key = np.tile(np.asarray(key_words, dtype=U64), (n, 1))
    nonce = np.tile(np.asarray(nonce_words, dtype=U64), (n, 1))
    zeros = np.zeros((n, 16), dtype=U64)
    trace, _ = generate_stream_trace(log_size, key, nonce, np.arange(n, dtype=U64), zeros, zeros)
    return trace[:N_COLS_BLOCK]


def evaluate_constraints(lde, alpha_pows_rev, block=False):
    """lde: [N_COLS, R] uint64 M31 column values on the evaluation domain (any row order), or [N_COLS, R, 4]
    QM31 values (the verifier / prove()'s sanity check evaluate the same constraints on the OODS mask values).
    Returns acc[R,4] = sum_k alpha_pows_rev[k] * C_k(row), before multiplication by the vanishing inverse."""
    ext = lde.ndim == 3
    R = lde.shape[1]
    A = _Acc(R, alpha_pows_rev, ext)
    col = [0]
    if ext:
        from stwo_core import q_mul, q_add, q_sub
        mul, add, sub = q_mul, q_add, q_sub
        one = np.array([1, 0, 0, 0], dtype=U64)
        two = np.array([2, 0, 0, 0], dtype=U64)
        zrow = np.zeros((1, R, 4), dtype=U64)
        cshape = (64, R, 4)
    else:
        mul, add, sub = m_mul, m_add, m_sub
        one, two = U64(1), U64(2)
        zrow = np.zeros((1, R), dtype=U64)
        cshape = (64, R)

    def boolean(b):
        return mul(b, sub(one, b))

    def next_u32():
        b = lde[col[0]:col[0] + 32]
        col[0] += 32
        A.emit(boolean(b))
        return b

    def add_u32(a, b):
        res = next_u32()
        car = lde[col[0]:col[0] + 32]
        col[0] += 32
        cin = np.concatenate([zrow, car[:-1]], axis=0)
        cm = np.empty(cshape, dtype=U64)
        cm[0::2] = boolean(car)
        # result + 2*carry - a - b - carry_in
        cm[1::2] = sub(sub(sub(add(res, mul(two, car)), a), b), cin)
        A.emit(cm)
        return res

    def xor_rotl_u32(a, b, r):
        res = next_u32()
        src = (np.arange(32) + 32 - r) % 32
        sa, sb = a[src], b[src]
        A.emit(add(sub(sub(res, sa), sb), mul(two, mul(sa, sb))))
        return res

    init = [next_u32() for _ in range(16)]
    v = list(init)
    for _ in range(10):
        for (a, b, c, d) in QR_SCHEDULE:
            v[a] = add_u32(v[a], v[b]); v[d] = xor_rotl_u32(v[a], v[d], 16)
            v[c] = add_u32(v[c], v[d]); v[b] = xor_rotl_u32(v[c], v[b], 12)
            v[a] = add_u32(v[a], v[b]); v[d] = xor_rotl_u32(v[a], v[d], 8)
            v[c] = add_u32(v[c], v[d]); v[b] = xor_rotl_u32(v[c], v[b], 7)
    ks = [add_u32(v[i], init[i]) for i in range(16)]
    if block:   # ChaChabitwiseEvalAtRow::eval (bitwise/constraints.rs:31-58) ends with the final additions
        assert col[0] == N_COLS_BLOCK and A.k == N_CONSTRAINTS_BLOCK
        return A.acc
    pt = [next_u32() for _ in range(16)]
    ct = [next_u32() for _ in range(16)]
    for i in range(16):
        comp = sub(add(ks[i], pt[i]), mul(two, mul(ks[i], pt[i])))
        A.emit(sub(comp, ct[i]))
    assert col[0] == N_COLS and A.k == N_CONSTRAINTS
    return A.acc
