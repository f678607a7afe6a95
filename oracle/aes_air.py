"""CPU restatement (numpy) of the reference's AES-128/256-CTR AIR: cipher, witness, LogUp interaction trace, constraints.

TEST INFRASTRUCTURE ONLY.

Follows (file:line relative to /root/reference/stwo/src):
  * native cipher        aes/mod.rs:10-30 (S-box), :213-270 (key expansion), :371-409 (aes{128,256}_ctr_block);
                         KATs FIPS-197 aes/mod.rs:431-470
  * trace generator      aes/lookup/gen_ctr.rs:70-145 (append_byte/append_bits/xor_byte_trace/xtime_trace/sbox_trace),
                         :152-195 (mix_columns_trace), :198-310 (process_ctr_block), :386-439 (default padding rows)
  * constraint sequence  aes/lookup/ctr.rs:26-45 (next_byte, sbox), :73-146 (xor_byte), :150-233 (xtime), :236-281
                         (gf_mul3, mix_columns), :293-364 (aes_block, ctr_block)
  * S-box table          aes/sbox_table.rs:35-48 (preprocessed columns), :52-76 (multiplicities), :94-120 (table component)
  * interaction traces   aes/lookup/gen_ctr.rs:640-683, aes/lookup/gen.rs:438-478; upstream LogupTraceGenerator
                         (constraint-framework/src/prover/logup.rs) and finalize_logup_in_pairs (src/lib.rs, logup.rs)
Column / constraint counts (24,480 / 34,464 and 34,784 / 49,024) are confirmed by the reference's get_circuits_info().
"""
import struct

import numpy as np

from stwo_core import P, U64, QM31, m_add, m_sub, m_mul, q_add, q_sub, q_mul, q_mul_m31, q_inv, q_from_m31

# ---------------------------------------------------------------- native cipher (aes/mod.rs)
SBOX = np.zeros(256, dtype=np.uint8)


def _init_sbox():
    p = q = 1
    while True:
        p = p ^ ((p << 1) & 0xFF) ^ (0x1B if p & 0x80 else 0)
        q ^= q << 1
        q ^= q << 2
        q ^= q << 4
        q &= 0xFF
        if q & 0x80:
            q ^= 0x09
        x = q ^ ((q << 1 | q >> 7) & 0xFF) ^ ((q << 2 | q >> 6) & 0xFF) ^ ((q << 3 | q >> 5) & 0xFF) ^ ((q << 4 | q >> 4) & 0xFF)
        SBOX[p] = (x ^ 0x63) & 0xFF
        if p == 1:
            break
    SBOX[0] = 0x63


_init_sbox()


def _xt(a):
    return ((a << 1) & 0xFF) ^ (0x1B if a & 0x80 else 0)


def expand_key(key):
    """aes/mod.rs:213 expand_key_128 / :240 expand_key_256 -> list of 11 / 15 round keys (16 ints each)."""
    nk = len(key) // 4
    nr = nk + 6
    w = [list(key[4 * i:4 * i + 4]) for i in range(nk)]
    rc = 1
    for i in range(nk, 4 * (nr + 1)):
        t = list(w[i - 1])
        if i % nk == 0:
            t = t[1:] + t[:1]
            t = [int(SBOX[b]) for b in t]
            t[0] ^= rc
            rc = _xt(rc)
        elif nk > 6 and i % nk == 4:
            t = [int(SBOX[b]) for b in t]
        w.append([a ^ b for a, b in zip(w[i - nk], t)])
    return [sum(w[4 * r:4 * r + 4], []) for r in range(nr + 1)]


SHIFT_ROWS = (0, 5, 10, 15, 4, 9, 14, 3, 8, 13, 2, 7, 12, 1, 6, 11)


def encrypt_block(rk, blk):
    s = [a ^ b for a, b in zip(blk, rk[0])]
    nr = len(rk) - 1
    for r in range(1, nr + 1):
        s = [int(SBOX[b]) for b in s]
        s = [s[i] for i in SHIFT_ROWS]
        if r < nr:
            o = []
            for c in range(4):
                a = s[4 * c:4 * c + 4]
                o += [_xt(a[0]) ^ _xt(a[1]) ^ a[1] ^ a[2] ^ a[3], a[0] ^ _xt(a[1]) ^ _xt(a[2]) ^ a[2] ^ a[3],
                      a[0] ^ a[1] ^ _xt(a[2]) ^ _xt(a[3]) ^ a[3], _xt(a[0]) ^ a[0] ^ a[1] ^ a[2] ^ _xt(a[3])]
            s = o
        s = [a ^ b for a, b in zip(s, rk[r])]
    return s


def ctr_keystream_block(key, nonce, counter):
    """aes/mod.rs:371 aes128_ctr_block with zero plaintext: AES(key, nonce || counter_be)."""
    return bytes(encrypt_block(expand_key(key), list(nonce) + list(struct.pack(">I", counter & 0xFFFFFFFF))))


def ctr_encrypt(key, nonce, counter, data):
    rk = expand_key(key)
    out = bytearray()
    for i in range(len(data) // 16):
        ks = encrypt_block(rk, list(nonce) + list(struct.pack(">I", (counter + i) & 0xFFFFFFFF)))
        out += bytes(a ^ b for a, b in zip(ks, data[16 * i:16 * i + 16]))
    return bytes(out)


def n_cols(key_len, block=False):
    """block: the block AIR (aes/lookup/{gen,constraints,air}.rs) - no plaintext / ciphertext columns, no final xor."""
    r = 10 if key_len == 16 else 14
    return (12 + 4 + 16 * (r + 1) + (0 if block else 32)) + 400 + (r - 1) * (16 + 2144 + 400) + 16 + 400 + (0 if block else 400)


def n_lookups(key_len):
    return 16 * (10 if key_len == 16 else 14)


def n_constraints(key_len, block=False):
    r = 10 if key_len == 16 else 14
    return 560 + (r - 1) * (3072 + 560) + 560 + (0 if block else 560 + 16) + n_lookups(key_len) // 2


# ---------------------------------------------------------------- witness (gen_ctr.rs)
def generate_ctr_trace(log_size, key, nonce_rows, counters, plaintext, ciphertext, block=False):
    """Per-row inputs: nonce_rows[N,12], counters[N], plaintext[N,16], ciphertext[N,16] (uint8/uint32 arrays; callers build
    the padding rows).  Returns (trace [C, N] uint64, lookups [L, 2, N] uint64, mults[256], valid)."""
    n = 1 << log_size
    rk = expand_key(key)
    nr = len(rk) - 1
    nonce_rows = np.asarray(nonce_rows, dtype=np.uint8).reshape(n, 12)
    counters = np.asarray(counters, dtype=np.uint64).reshape(n)
    pt = np.asarray(plaintext, dtype=np.uint8).reshape(n, 16)
    ct = np.asarray(ciphertext, dtype=np.uint8).reshape(n, 16)
    cols = []
    lookups = []
    mults = np.zeros(256, dtype=np.uint64)

    def byte(v):
        cols.append(v.astype(U64))

    def bits(v):
        for i in range(8):
            cols.append(((v >> np.uint8(i)) & np.uint8(1)).astype(U64))

    def xor_byte(a, b):
        r = a ^ b
        bits(a); bits(b); bits(r); byte(r)
        return r

    def xtime(a):
        r = ((a << np.uint8(1)) ^ ((a >> np.uint8(7)) * np.uint8(0x1B))).astype(np.uint8)
        bits(a); bits(r); byte(r)
        return r

    def sbox(a):
        o = SBOX[a]
        np.add.at(mults, a.astype(np.int64), 1)
        lookups.append((a.astype(U64), o.astype(U64)))
        byte(o)
        return o

    def mul3(a):
        d = xtime(a)
        return xor_byte(d, a)

    def mix_columns(s):
        out = [None] * 16
        for c in range(4):
            i = 4 * c
            s0, s1, s2, s3 = s[i], s[i + 1], s[i + 2], s[i + 3]
            t0 = xtime(s0); t1 = mul3(s1); t2 = xor_byte(t0, t1); t3 = xor_byte(t2, s2); out[i] = xor_byte(t3, s3)
            t0 = xtime(s1); t1 = mul3(s2); t2 = xor_byte(s0, t0); t3 = xor_byte(t2, t1); out[i + 1] = xor_byte(t3, s3)
            t0 = xtime(s2); t1 = mul3(s3); t2 = xor_byte(s0, s1); t3 = xor_byte(t2, t0); out[i + 2] = xor_byte(t3, t1)
            t0 = mul3(s0); t1 = xtime(s3); t2 = xor_byte(t0, s1); t3 = xor_byte(t2, s2); out[i + 3] = xor_byte(t3, t1)
        return out

    rkv = [[np.full(n, b, dtype=np.uint8) for b in r] for r in rk]
    if block:   # gen.rs generate_trace: the input block (here `plaintext`), the round keys, then the cipher
        blk = [pt[:, i] for i in range(16)]
        for b in blk:
            byte(b)
        for r in rkv:
            for b in r:
                byte(b)
    else:
        for i in range(12):
            byte(nonce_rows[:, i])
        cbytes = [((counters >> np.uint64(8 * (3 - i))) & np.uint64(0xFF)).astype(np.uint8) for i in range(4)]
        for c in cbytes:
            byte(c)
        for r in rkv:
            for b in r:
                byte(b)
        for i in range(16):
            byte(pt[:, i])
        for i in range(16):
            byte(ct[:, i])
        blk = [nonce_rows[:, i] for i in range(12)] + cbytes
    state = [xor_byte(blk[i], rkv[0][i]) for i in range(16)]
    for rnd in range(1, nr):
        state = [sbox(state[i]) for i in range(16)]
        state = [state[i] for i in SHIFT_ROWS]
        state = mix_columns(state)
        state = [xor_byte(state[i], rkv[rnd][i]) for i in range(16)]
    state = [sbox(state[i]) for i in range(16)]
    state = [state[i] for i in SHIFT_ROWS]
    ks = [xor_byte(state[i], rkv[nr][i]) for i in range(16)]
    valid = True
    if not block:
        comp = [xor_byte(ks[i], pt[:, i]) for i in range(16)]
        valid = all(np.array_equal(comp[i], ct[:, i]) for i in range(16))
    trace = np.stack(cols, axis=0)
    assert trace.shape[0] == n_cols(len(key), block), trace.shape
    lk = np.stack([np.stack(l, axis=0) for l in lookups], axis=0)
    return trace, lk, mults, valid


def sbox_table_columns():
    """aes/sbox_table.rs:35-48 generate_sbox_trace: input 0..255, output SBOX[input] (log size 8)."""
    return np.stack([np.arange(256, dtype=U64), SBOX.astype(U64)], axis=0)


# ---------------------------------------------------------------- LogUp (upstream constraint-framework logup.rs)
class SboxElements:
    """relation!(SboxElements, 2): draw z, alpha from the channel; combine([a, b]) = a + alpha*b - z."""

    def __init__(self, z, alpha):
        self.z, self.alpha = z, alpha

    @staticmethod
    def draw(channel):
        z, alpha = channel.draw_secure_felts(2)
        return SboxElements(z, alpha)

    def combine_cols(self, a, b):
        """a, b: uint64 arrays [N] of M31 values -> [N,4]."""
        al = np.array(self.alpha.v, dtype=U64)[None, :]
        out = q_mul_m31(np.broadcast_to(al, (len(a), 4)).copy(), np.asarray(b, dtype=U64))
        out[:, 0] = m_add(out[:, 0], np.asarray(a, dtype=U64))
        return q_sub(out, np.broadcast_to(np.array(self.z.v, dtype=U64)[None, :], out.shape))


def q_batch_inv(x):
    """Element-wise QM31 inverse of [N,4]."""
    return q_inv(x)


def coset_order_to_storage(log_size):
    """storage index (bit-reversed circle-domain order) of the i-th point of CanonicCoset(log).coset in natural order
    (core/utils.rs coset_index_to_circle_domain_index + bit_reverse_index)."""
    from stwo_core import bit_reverse_indices
    n = 1 << log_size
    i = np.arange(n)
    cd = np.where(i % 2 == 0, i // 2, n - 1 - i // 2)      # coset index -> circle-domain index
    br = bit_reverse_indices(log_size)                      # br[k] = bit_reverse(k)
    return br[cd]


def logup_finalize_last(cols, log_size):
    """LogupTraceGenerator::finalize_last: claimed_sum = sum of the last cumulative column; the last column becomes the
    inclusive prefix sum (in coset order) of (value - claimed_sum/N).  cols: list of [N,4] arrays (cumulative columns).
    Returns (list of 4*len(cols) base columns, claimed_sum QM31)."""
    n = 1 << log_size
    last = cols[-1]
    tot = [int(np.sum(last[:, c].astype(object)) % P) for c in range(4)]
    claimed = QM31(*tot)
    ninv = pow(n, P - 2, P)
    shift = np.array([(t * ninv) % P for t in tot], dtype=U64)
    shifted = q_sub(last, np.broadcast_to(shift[None, :], last.shape))
    order = coset_order_to_storage(log_size)
    pref = np.empty_like(shifted)
    for c in range(4):
        v = shifted[order, c].astype(object)
        pref[order, c] = (np.cumsum(v) % P).astype(U64)
    cols = cols[:-1] + [pref]
    out = []
    for col in cols:
        for c in range(4):
            out.append(col[:, c].copy())
    return out, claimed


def ctr_interaction_trace(log_size, lookups, elems):
    """gen_ctr.rs:640-683: lookups in pairs, fraction (p0+p1)/(p0*p1), cumulative over pairs."""
    n = 1 << log_size
    cols = []
    prev = np.zeros((n, 4), dtype=U64)
    L = lookups.shape[0]
    for k in range(0, L - 1, 2):
        p0 = elems.combine_cols(lookups[k, 0], lookups[k, 1])
        p1 = elems.combine_cols(lookups[k + 1, 0], lookups[k + 1, 1])
        num = q_add(p0, p1)
        den = q_mul(p0, p1)
        cur = q_add(prev, q_mul(num, q_batch_inv(den)))
        cols.append(cur)
        prev = cur
    assert L % 2 == 0
    return logup_finalize_last(cols, log_size)


def table_interaction_trace(mults, elems):
    """aes/lookup/gen.rs:438-478: one column, fraction -mult / combine(i, SBOX[i]) over the 256 table rows."""
    t = sbox_table_columns()
    p = elems.combine_cols(t[0], t[1])
    num = np.zeros((256, 4), dtype=U64)
    num[:, 0] = (P - (np.asarray(mults, dtype=U64) % U64(P))) % U64(P)
    cur = q_mul(num, q_batch_inv(p))
    return logup_finalize_last([cur], 8)


# ---------------------------------------------------------------- constraints (aes/lookup/ctr.rs, sbox_table.rs)
class _Acc:
    """sum_k alpha^(K-1-k) * C_k(row) -- add_constraint of the framework's domain / point evaluators."""

    def __init__(self, n_rows, alpha_pows_rev, k0=0):
        self.acc = np.zeros((n_rows, 4), dtype=U64)
        self.k = k0
        self.apr = alpha_pows_rev

    def emit_base(self, cmat):                       # [m,R] M31 constraint values
        m = cmat.shape[0]
        co = self.apr[self.k:self.k + m]
        prod = (cmat[:, :, None] * co[:, None, :]) % U64(P)
        self.acc = (self.acc + prod.sum(axis=0)) % U64(P)
        self.k += m

    def emit_ext(self, cmat):                        # [m,R,4] QM31 constraint values
        m = cmat.shape[0]
        co = self.apr[self.k:self.k + m]
        prod = q_mul(cmat, np.broadcast_to(co[:, None, :], cmat.shape))
        self.acc = (self.acc + prod.sum(axis=0)) % U64(P)
        self.k += m


def _field_ops(ext):
    if ext:
        one = np.array([1, 0, 0, 0], dtype=U64)
        return q_mul, q_add, q_sub, one
    return m_mul, m_add, m_sub, U64(1)


def _to_ext(v, ext):
    return v if ext else q_from_m31(v)


def _combine(elems, a, b):
    """SboxElements::combine on QM31 arrays [R,4]: a + alpha*b - z."""
    al = np.broadcast_to(np.array(elems.alpha.v, dtype=U64), a.shape)
    z = np.broadcast_to(np.array(elems.z.v, dtype=U64), a.shape)
    return q_sub(q_add(a, q_mul(al, b)), z)


def _from_partial(cols4):
    """SecureField::from_partial_evals: sum_c cols4[c] * unit_c, cols4 [4,R,4] QM31."""
    units = [np.array(u, dtype=U64) for u in ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1))]
    out = np.zeros(cols4.shape[1:], dtype=U64)
    for c in range(4):
        out = q_add(out, q_mul(cols4[c], np.broadcast_to(units[c], cols4[c].shape)))
    return out


def evaluate_ctr_constraints(main, inter, inter_prev_last, elems, claimed_sum, log_size, alpha_pows_rev, key_len, block_air=False):
    """AESCtrEvalAtRow::ctr_block (ctr.rs:320-364) + finalize_logup_in_pairs.
    main [C,R] (M31) or [C,R,4] (QM31 mask values); inter [4*L/2, R(,4)] interaction coordinate columns at offset 0;
    inter_prev_last [4, R(,4)] = last interaction QM31 column at offset -1.  alpha_pows_rev[k] multiplies constraint k of
    this component.  Returns acc[R,4]."""
    ext = main.ndim == 3
    mul, add, sub, one = _field_ops(ext)
    R = main.shape[1]
    A = _Acc(R, alpha_pows_rev)
    col = [0]
    nr = 10 if key_len == 16 else 14
    pw = [U64(1 << i) for i in range(8)]
    rel = []

    def nxt(k=1):
        v = main[col[0]:col[0] + k]
        col[0] += k
        return v

    def boolean(b):
        return mul(b, sub(one, b))

    def scale(v, c):                                  # v * small constant
        return (v * c) % U64(P)

    def recompose(bits):
        s = scale(bits[0], pw[0])
        for i in range(1, 8):
            s = add(s, scale(bits[i], pw[i]))
        return s

    def bits8():
        b = nxt(8)
        A.emit_base(boolean(b)) if not ext else A.emit_ext(boolean(b))
        return b

    def emit1(v):
        (A.emit_ext if ext else A.emit_base)(v[None])

    def xor_byte(a, b):
        ab, bb, cb = bits8(), bits8(), bits8()
        emit1(sub(a, recompose(ab)))
        emit1(sub(b, recompose(bb)))
        x = add(sub(sub(cb, ab), bb), scale(mul(ab, bb), U64(2)))
        (A.emit_ext if ext else A.emit_base)(x)
        r = nxt()[0]
        emit1(sub(r, recompose(cb)))
        return r

    def xtime(a):
        ab = bits8()
        emit1(sub(a, recompose(ab)))
        rb = bits8()
        hb = ab[7]

        def x2(i, j):
            return add(sub(sub(rb[i], ab[j]), hb), scale(mul(ab[j], hb), U64(2)))
        cons = [sub(rb[0], hb), x2(1, 0), sub(rb[2], ab[1]), x2(3, 2), x2(4, 3), sub(rb[5], ab[4]), sub(rb[6], ab[5]),
                sub(rb[7], ab[6])]
        (A.emit_ext if ext else A.emit_base)(np.stack(cons, axis=0))
        r = nxt()[0]
        emit1(sub(r, recompose(rb)))
        return r

    def sbox(a):
        o = nxt()[0]
        rel.append((a, o))
        return o

    def mul3(a):
        return xor_byte(xtime(a), a)

    def mix_columns(s):
        out = [None] * 16
        for c in range(4):
            i = 4 * c
            s0, s1, s2, s3 = s[i], s[i + 1], s[i + 2], s[i + 3]
            t0 = xtime(s0); t1 = mul3(s1); t2 = xor_byte(t0, t1); t3 = xor_byte(t2, s2); out[i] = xor_byte(t3, s3)
            t0 = xtime(s1); t1 = mul3(s2); t2 = xor_byte(s0, t0); t3 = xor_byte(t2, t1); out[i + 1] = xor_byte(t3, s3)
            t0 = xtime(s2); t1 = mul3(s3); t2 = xor_byte(s0, s1); t3 = xor_byte(t2, t0); out[i + 2] = xor_byte(t3, t1)
            t0 = mul3(s0); t1 = xtime(s3); t2 = xor_byte(t0, s1); t3 = xor_byte(t2, s2); out[i + 3] = xor_byte(t3, t1)
        return out

    block = list(nxt(16))
    rks = [list(nxt(16)) for _ in range(nr + 1)]
    if not block_air:
        pt = list(nxt(16))
        ct = list(nxt(16))
    state = [xor_byte(block[i], rks[0][i]) for i in range(16)]
    for rnd in range(1, nr):
        state = [sbox(state[i]) for i in range(16)]
        state = [state[i] for i in SHIFT_ROWS]
        state = mix_columns(state)
        state = [xor_byte(state[i], rks[rnd][i]) for i in range(16)]
    state = [sbox(state[i]) for i in range(16)]
    state = [state[i] for i in SHIFT_ROWS]
    ks = [xor_byte(state[i], rks[nr][i]) for i in range(16)]
    if not block_air:
        comp = [xor_byte(ks[i], pt[i]) for i in range(16)]
        for i in range(16):
            emit1(sub(comp[i], ct[i]))
    assert col[0] == n_cols(key_len, block_air), (col[0], n_cols(key_len, block_air))
    # finalize_logup_in_pairs
    nb = len(rel) // 2
    prev_col = np.zeros((R, 4), dtype=U64)
    ninv = pow(1 << log_size, P - 2, P)
    shift = np.broadcast_to(np.array([(c * ninv) % P for c in claimed_sum.v], dtype=U64), (R, 4))
    for k in range(nb):
        p0 = _combine(elems, _to_ext(rel[2 * k][0], ext), _to_ext(rel[2 * k][1], ext))
        p1 = _combine(elems, _to_ext(rel[2 * k + 1][0], ext), _to_ext(rel[2 * k + 1][1], ext))
        num, den = q_add(p0, p1), q_mul(p0, p1)
        cur = _from_partial(np.stack([_to_ext(inter[4 * k + c], ext) for c in range(4)], axis=0))
        if k < nb - 1:
            diff = q_sub(cur, prev_col)
        else:
            prv = _from_partial(np.stack([_to_ext(inter_prev_last[c], ext) for c in range(4)], axis=0))
            diff = q_add(q_sub(q_sub(cur, prv), prev_col), shift)
        prev_col = cur
        A.emit_ext(q_sub(q_mul(diff, den), num)[None])
    assert A.k == n_constraints(key_len, block_air), (A.k, n_constraints(key_len, block_air))
    return A.acc


def evaluate_table_constraint(pre, mult, inter, inter_prev, elems, claimed_sum, alpha_pow):
    """SboxTableEval::evaluate (sbox_table.rs:103-120): one LogUp entry with multiplicity -mult, one constraint.
    pre [2,R(,4)], mult [R(,4)], inter / inter_prev [4,R(,4)]; alpha_pow = the (single) power for this constraint [4]."""
    ext = pre.ndim == 3
    R = pre.shape[1]
    p = _combine(elems, _to_ext(pre[0], ext), _to_ext(pre[1], ext))
    num = q_sub(np.zeros((R, 4), dtype=U64), _to_ext(mult, ext))
    cur = _from_partial(np.stack([_to_ext(inter[c], ext) for c in range(4)], axis=0))
    prv = _from_partial(np.stack([_to_ext(inter_prev[c], ext) for c in range(4)], axis=0))
    ninv = pow(256, P - 2, P)
    shift = np.broadcast_to(np.array([(c * ninv) % P for c in claimed_sum.v], dtype=U64), (R, 4))
    diff = q_add(q_sub(cur, prv), shift)
    c = q_sub(q_mul(diff, p), num)
    return q_mul(c, np.broadcast_to(np.asarray(alpha_pow, dtype=U64), c.shape))


def prev_row_index(log_size, eval_log):
    """core/utils.rs offset_bit_reversed_circle_domain_index(i, log_size, eval_log, -1) for every storage index i."""
    from stwo_core import bit_reverse_indices
    m = 1 << eval_log
    br = bit_reverse_indices(eval_log)
    half = m >> 1
    step = -(1 << (eval_log - log_size - 1))
    nat = br[np.arange(m)]
    prev = np.where(nat < half, (nat + step) % half, ((nat - step) % half) + half)
    return br[prev]
