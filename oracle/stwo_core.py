"""CPU restatement (numpy) of the upstream-stwo arithmetic the reference prover drives.

TEST INFRASTRUCTURE ONLY -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

The algorithms live in the un-vendored dependency `stwo` / `stwo-constraint-framework` 2.1.0, git rev
f117d487b39cd4441c2dbbdd7127b186f9c23564 (/root/reference/stwo/Cargo.toml:16-17).  Their source is not under
/root/reference, so every function below restates the *published* algorithm of that crate (module path given in
each docstring, as embedded in the reference's WASM build) and is PINNED against the reference itself:
oracle/_ref/libs2c_ref.so (the reference's shipped WASM, compiled natively) must produce byte-identical proofs
(tests/test_oracle.py, tests/test_oracle_aes.py, tests/golden/).  Call sites in the reference: /root/reference/stwo/src/
chacha/bitwise/air_stream.rs:185-231 and aes/lookup/air_ctr.rs:328-414.
"""
import hashlib
import struct
import numpy as np

P = (1 << 31) - 1
U64 = np.uint64
_P64 = np.uint64(P)


# ------------------------------------------------------------------------------------------------
# M31 / CM31 / QM31   (stwo core/fields/{m31,cm31,qm31}.rs)
# numpy arrays are uint64 holding canonical values in [0, P); QM31 = trailing axis of 4 (a + bi) + (c + di)u
# ------------------------------------------------------------------------------------------------
def m_add(a, b):
    return (a + b) % _P64


def m_sub(a, b):
    return (a + _P64 - b) % _P64


def m_mul(a, b):
    return (a * b) % _P64


def m_neg(a):
    return (_P64 - a) % _P64


def m_pow(a, e):
    """a ** e elementwise (a: uint64 array or int)."""
    if isinstance(a, (int, np.integer)):
        return pow(int(a), e, P)
    r = np.ones_like(a)
    base = a.copy()
    while e:
        if e & 1:
            r = m_mul(r, base)
        base = m_mul(base, base)
        e >>= 1
    return r


def m_inv(a):
    return m_pow(a, P - 2)


def m_batch_inv(a):
    return m_pow(a, P - 2)


# --- scalar (python int tuple) extension-field arithmetic, used for transcript-level values ---
def c_mul(x, y):
    return ((x[0] * y[0] - x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P)


def c_add(x, y):
    return ((x[0] + y[0]) % P, (x[1] + y[1]) % P)


def c_sub(x, y):
    return ((x[0] - y[0]) % P, (x[1] - y[1]) % P)


def c_inv(x):
    n = pow((x[0] * x[0] + x[1] * x[1]) % P, P - 2, P)
    return (x[0] * n % P, (-x[1]) * n % P)


class QM31:
    """Scalar QM31 = CM31[u]/(u^2 - (2+i)).  core/fields/qm31.rs."""
    __slots__ = ("v",)

    def __init__(self, a=0, b=0, c=0, d=0):
        self.v = (a % P, b % P, c % P, d % P)

    @staticmethod
    def from_m31(a):
        return QM31(int(a), 0, 0, 0)

    def __add__(self, o):
        o = _q(o)
        return QM31(*[(x + y) for x, y in zip(self.v, o.v)])

    __radd__ = __add__

    def __sub__(self, o):
        o = _q(o)
        return QM31(*[(x - y) for x, y in zip(self.v, o.v)])

    def __rsub__(self, o):
        return _q(o) - self

    def __neg__(self):
        return QM31(*[-x for x in self.v])

    def __mul__(self, o):
        o = _q(o)
        a0, a1 = self.v[:2], self.v[2:]
        b0, b1 = o.v[:2], o.v[2:]
        t = c_mul(a1, b1)
        rt = c_mul((2, 1), t)
        lo = c_add(c_mul(a0, b0), rt)
        hi = c_add(c_mul(a0, b1), c_mul(a1, b0))
        return QM31(lo[0], lo[1], hi[0], hi[1])

    __rmul__ = __mul__

    def inv(self):
        a, b = self.v[:2], self.v[2:]
        b2 = c_mul(b, b)
        # denom = a^2 - (2+i) b^2   (core/fields/qm31.rs inverse)
        denom = c_sub(c_mul(a, a), c_mul((2, 1), b2))
        di = c_inv(denom)
        lo = c_mul(a, di)
        hi = c_mul(((-b[0]) % P, (-b[1]) % P), di)
        return QM31(lo[0], lo[1], hi[0], hi[1])

    def conj(self):
        """complex_conjugate: a + bu -> a - bu."""
        return QM31(self.v[0], self.v[1], -self.v[2], -self.v[3])

    def __pow__(self, e):
        r = QM31(1)
        b = self
        while e:
            if e & 1:
                r = r * b
            b = b * b
            e >>= 1
        return r

    def __eq__(self, o):
        return self.v == _q(o).v

    def __hash__(self):
        return hash(self.v)

    def __repr__(self):
        return "QM31%s" % (self.v,)

    def is_zero(self):
        return self.v == (0, 0, 0, 0)

    def arr(self):
        return np.array(self.v, dtype=U64)


def _q(o):
    return o if isinstance(o, QM31) else QM31(int(o), 0, 0, 0)


# --- vectorised QM31 (arrays [...,4]) ---
def q_from_m31(a):
    z = np.zeros(a.shape + (4,), dtype=U64)
    z[..., 0] = a
    return z


def cm_mul(a0, a1, b0, b1):
    return m_sub(m_mul(a0, b0), m_mul(a1, b1)), m_add(m_mul(a0, b1), m_mul(a1, b0))


def q_add(x, y):
    return (x + y) % _P64


def q_sub(x, y):
    return (x + _P64 - y) % _P64


def q_mul(x, y):
    x = np.asarray(x, dtype=U64)
    y = np.asarray(y, dtype=U64)
    a0, a1, a2, a3 = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    b0, b1, b2, b3 = y[..., 0], y[..., 1], y[..., 2], y[..., 3]
    t0, t1 = cm_mul(a2, a3, b2, b3)                       # a1*b1 in CM31
    r0, r1 = m_sub(m_add(t0, t0), t1), m_add(m_add(t1, t1), t0)   # (2+i)*t
    l0, l1 = cm_mul(a0, a1, b0, b1)
    h0a, h1a = cm_mul(a0, a1, b2, b3)
    h0b, h1b = cm_mul(a2, a3, b0, b1)
    return np.stack([m_add(l0, r0), m_add(l1, r1), m_add(h0a, h0b), m_add(h1a, h1b)], axis=-1)


def q_mul_m31(x, m):
    return (x * np.asarray(m, dtype=U64)[..., None]) % _P64


def q_inv(x):
    """Elementwise QM31 inverse (arrays [...,4])."""
    a0, a1, b0, b1 = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    s0, s1 = cm_mul(b0, b1, b0, b1)
    r0, r1 = m_sub(m_add(s0, s0), s1), m_add(m_add(s1, s1), s0)
    q0, q1 = cm_mul(a0, a1, a0, a1)
    d0, d1 = m_sub(q0, r0), m_sub(q1, r1)
    n = m_inv(m_add(m_mul(d0, d0), m_mul(d1, d1)))
    i0, i1 = m_mul(d0, n), m_mul(m_neg(d1), n)
    l0, l1 = cm_mul(a0, a1, i0, i1)
    h0, h1 = cm_mul(m_neg(b0), m_neg(b1), i0, i1)
    return np.stack([l0, l1, h0, h1], axis=-1)


# ------------------------------------------------------------------------------------------------
# Circle group, cosets, domains   (stwo core/circle.rs, core/poly/circle/{canonic,domain}.rs)
# ------------------------------------------------------------------------------------------------
GEN = (2, 1268011823)      # M31_CIRCLE_GEN, order 2^31
LOG_ORDER = 31


def pt_add(p, q):
    return ((p[0] * q[0] - p[1] * q[1]) % P, (p[0] * q[1] + p[1] * q[0]) % P)


def pt_double(p):
    return pt_add(p, p)


_GEN_POW2 = [GEN]
for _ in range(31):
    _GEN_POW2.append(pt_double(_GEN_POW2[-1]))


def index_to_point(idx):
    """CirclePointIndex::to_point: GEN * idx (idx mod 2^31)."""
    idx &= (1 << 31) - 1
    r = (1, 0)
    k = 0
    while idx:
        if idx & 1:
            r = pt_add(r, _GEN_POW2[k])
        idx >>= 1
        k += 1
    return r


def subgroup_gen(log_size):
    return 1 << (LOG_ORDER - log_size)


class Coset:
    """core/circle.rs Coset{initial_index, step_size, log_size}."""

    def __init__(self, initial_index, log_size):
        self.initial_index = initial_index & ((1 << 31) - 1)
        self.log_size = log_size
        self.step_size = subgroup_gen(log_size)

    @staticmethod
    def odds(log_size):
        return Coset(subgroup_gen(log_size + 1), log_size)

    @staticmethod
    def half_odds(log_size):
        return Coset(subgroup_gen(log_size + 2), log_size)

    def size(self):
        return 1 << self.log_size

    def index_at(self, i):
        return (self.initial_index + self.step_size * i) & ((1 << 31) - 1)

    def at(self, i):
        return index_to_point(self.index_at(i))

    def double(self):
        c = Coset.__new__(Coset)
        c.initial_index = (self.initial_index * 2) & ((1 << 31) - 1)
        c.step_size = (self.step_size * 2) & ((1 << 31) - 1)
        c.log_size = max(self.log_size - 1, 0)
        return c

    def points(self):
        """All points in coset order as (xs, ys) uint64 arrays (vectorised repeated-doubling build)."""
        n = self.size()
        xs = np.empty(n, dtype=U64)
        ys = np.empty(n, dtype=U64)
        p0 = index_to_point(self.initial_index)
        xs[0], ys[0] = p0
        step = index_to_point(self.step_size)
        m = 1
        while m < n:
            sx, sy = U64(step[0]), U64(step[1])
            xs[m:2 * m] = m_sub(m_mul(xs[:m], sx), m_mul(ys[:m], sy))
            ys[m:2 * m] = m_add(m_mul(xs[:m], sy), m_mul(ys[:m], sx))
            step = pt_double(step)
            m *= 2
        return xs, ys


class CircleDomain:
    """core/poly/circle/domain.rs: half_coset followed by its conjugate."""

    def __init__(self, half_coset):
        self.half_coset = half_coset
        self.log_size = half_coset.log_size + 1

    def size(self):
        return 1 << self.log_size

    def index_at(self, i):
        h = self.half_coset.size()
        if i < h:
            return self.half_coset.index_at(i)
        return (-self.half_coset.index_at(i - h)) & ((1 << 31) - 1)

    def at(self, i):
        return index_to_point(self.index_at(i))

    def points(self):
        xs, ys = self.half_coset.points()
        return np.concatenate([xs, xs]), np.concatenate([ys, m_neg(ys)])

    def points_bitrev(self):
        """Points in storage (bit-reversed) order: memory index j <-> at(bit_reverse(j))."""
        xs, ys = self.points()
        br = bit_reverse_indices(self.log_size)
        return xs[br], ys[br]


def canonic_domain(log_size):
    """CanonicCoset::new(log_size).circle_domain()  (core/poly/circle/canonic.rs)."""
    return CircleDomain(Coset.half_odds(log_size - 1))


def bit_reverse_indices(log_n):
    n = 1 << log_n
    idx = np.arange(n, dtype=np.int64)
    r = np.zeros(n, dtype=np.int64)
    for b in range(log_n):
        r |= ((idx >> b) & 1) << (log_n - 1 - b)
    return r


def bit_reverse_index(i, log_n):
    r = 0
    for b in range(log_n):
        r |= ((i >> b) & 1) << (log_n - 1 - b)
    return r


def double_x(x):
    return (2 * x * x - 1) % P


# ------------------------------------------------------------------------------------------------
# Circle FFT   (stwo prover/backend/cpu/circle.rs interpolate / evaluate; basis 1,y,x,xy,pi(x),...)
# Vectorised over leading axes: values[..., N] in bit-reversed domain order.
# ------------------------------------------------------------------------------------------------
_TW_CACHE = {}


def _layer_twiddles(domain_log):
    """For canonic domain of log size n: list tw[i], i=0..n-1; tw[0][h] = y of half_coset.at(bitrev(h)),
    tw[i>=1][h] = x-coordinate of (half_coset doubled i-1 times).at(bitrev(h)).  (numerators; inverses cached too)"""
    if domain_log in _TW_CACHE:
        return _TW_CACHE[domain_log]
    dom = canonic_domain(domain_log)
    coset = dom.half_coset
    tws = []
    xs, ys = coset.points()
    br = bit_reverse_indices(coset.log_size)
    tws.append(ys[br])
    c = coset
    for i in range(1, domain_log):
        xs, _ = c.points()
        half = c.size() // 2
        br = bit_reverse_indices(c.log_size - 1)
        tws.append(xs[:half][br])
        c = c.double()
    itws = [m_inv(t) for t in tws]
    _TW_CACHE[domain_log] = (tws, itws)
    return _TW_CACHE[domain_log]


def circle_ifft(values):
    """interpolate: values[..., N] (bit-reversed canonic-domain order) -> coefficients[..., N]."""
    v = np.array(values, dtype=U64, copy=True)
    n = v.shape[-1]
    log_n = n.bit_length() - 1
    if log_n == 0:
        return v
    _, itws = _layer_twiddles(log_n)
    lead = v.shape[:-1]
    for i in range(log_n):
        w = v.reshape(lead + (n >> (i + 1), 2, 1 << i))
        a = w[..., 0, :]
        b = w[..., 1, :]
        t = itws[i][:, None]
        s = m_add(a, b)
        d = m_mul(m_sub(a, b), t)
        w[..., 0, :] = s
        w[..., 1, :] = d
    inv_n = U64(pow(n, P - 2, P))
    return m_mul(v, inv_n)


def circle_fft(coeffs, log_domain=None):
    """evaluate: coefficients[..., M] zero-extended to 2^log_domain -> values on canonic domain (bit-reversed)."""
    c = np.asarray(coeffs, dtype=U64)
    m = c.shape[-1]
    log_m = m.bit_length() - 1
    if log_domain is None:
        log_domain = log_m
    n = 1 << log_domain
    v = np.zeros(c.shape[:-1] + (n,), dtype=U64)
    v[..., :m] = c
    if log_domain == 0:
        return v
    tws, _ = _layer_twiddles(log_domain)
    lead = v.shape[:-1]
    for i in reversed(range(log_domain)):
        w = v.reshape(lead + (n >> (i + 1), 2, 1 << i))
        a = w[..., 0, :].copy()
        tb = m_mul(w[..., 1, :], tws[i][:, None])
        w[..., 0, :] = m_add(a, tb)
        w[..., 1, :] = m_sub(a, tb)
    return v


def eval_at_point(coeffs, px, py):
    """CirclePoly::eval_at_point (cpu/circle.rs): coeffs[..., N] M31, point (px,py) QM31 scalars -> QM31 array [...,4]."""
    c = np.asarray(coeffs, dtype=U64)
    n = c.shape[-1]
    log_n = n.bit_length() - 1
    acc = q_from_m31(c)                                  # [..., N, 4]
    maps = []
    if log_n >= 1:
        maps.append(py)
        x = px
        for _ in range(1, log_n):
            maps.append(x)
            x = x * x * 2 - 1
    for i in range(log_n):                               # fold lowest index bit first with maps[i]
        f = maps[i].arr()
        lo = acc[..., 0::2, :]
        hi = acc[..., 1::2, :]
        acc = q_add(lo, q_mul(hi, f))
    return acc[..., 0, :]


# ------------------------------------------------------------------------------------------------
# Blake2s helpers, channel   (stwo core/channel/blake2s.rs, core/vcs/blake2_hash.rs)
# ------------------------------------------------------------------------------------------------
def blake2s(b):
    return hashlib.blake2s(b).digest()


class Blake2sChannel:
    """Blake2sChannel at the pinned rev, as observed in the reference binary (oracle/trace_blake.py):
      mix_*    : digest = H(digest || payload)
      draw     : H(digest || n_sent as 4 LE bytes || 0x00), n_sent += 1      (37-byte preimage)
    """

    def __init__(self):
        self.digest = bytes(32)
        self.n_sent = 0

    def _update(self, d):
        self.digest = d
        self.n_sent = 0

    def mix_root(self, root):
        self._update(blake2s(self.digest + root))

    def mix_u32s(self, words):
        self._update(blake2s(self.digest + struct.pack("<%dI" % len(words), *words)))

    def mix_u64(self, v):
        self.mix_u32s([v & 0xFFFFFFFF, (v >> 32) & 0xFFFFFFFF])

    def mix_felts(self, felts):
        """felts: iterable of QM31 or an array [n,4]."""
        if isinstance(felts, np.ndarray):
            payload = felts.astype("<u4").tobytes()
        else:
            payload = b"".join(struct.pack("<4I", *f.v) for f in felts)
        self._update(blake2s(self.digest + payload))

    def draw_u32s(self):
        h = blake2s(self.digest + struct.pack("<I", self.n_sent) + b"\x00")
        self.n_sent += 1
        return struct.unpack("<8I", h)

    def draw_base_felts(self):
        while True:
            u = self.draw_u32s()
            if all(x < 2 * P for x in u):
                return [x % P for x in u]

    def draw_secure_felt(self):
        f = self.draw_base_felts()
        return QM31(*f[:4])

    def draw_secure_felts(self, n):
        out = []
        buf = []
        while len(out) < n:
            if len(buf) < 4:
                buf += self.draw_base_felts()
            out.append(QM31(*buf[:4]))
            buf = buf[4:]
        return out

    # proof of work (core/channel/blake2s.rs verify_pow_nonce)
    POW_PREFIX = 0x12345678

    def pow_prefixed_digest(self, n_bits):
        return blake2s(struct.pack("<I", self.POW_PREFIX) + bytes(12) + self.digest + struct.pack("<I", n_bits))

    def verify_pow_nonce(self, n_bits, nonce):
        pd = self.pow_prefixed_digest(n_bits)
        res = blake2s(pd + struct.pack("<Q", nonce))
        v = int.from_bytes(res[:16], "little")
        tz = 128 if v == 0 else (v & -v).bit_length() - 1
        return tz >= n_bits

    def grind(self, n_bits):
        nonce = 0
        while not self.verify_pow_nonce(n_bits, nonce):
            nonce += 1
        return nonce


def get_random_point(channel):
    """CirclePoint::<SecureField>::get_random_point (core/circle.rs)."""
    t = channel.draw_secure_felt()
    t2 = t * t
    inv = (t2 + 1).inv()
    x = (QM31(1) - t2) * inv
    y = (t + t) * inv
    return x, y


# ------------------------------------------------------------------------------------------------
# Lifted Merkle tree   (stwo prover/vcs_lifted/prover.rs, backend/{cpu/merkle_lifted,simd/blake2s_lifted}.rs)
# ------------------------------------------------------------------------------------------------
def lifted_index(i, lifting_log, col_log):
    """Row of a size-2^col_log column feeding leaf i of a 2^lifting_log-leaf tree."""
    if col_log == lifting_log:
        return i
    shift = lifting_log - col_log
    return ((i >> (shift + 1)) << 1) + (i & 1)


class MerkleTree:
    """Single leaf layer of 2^L leaves; leaf = Blake2s(LE u32 of every column's (lifted) value, in column order);
    node = Blake2s(left || right); empty tree root = Blake2s("")."""

    def __init__(self, columns, lifting_log=None):
        """columns: list of 1-D uint64 arrays (sizes may differ, powers of two)."""
        self.columns = columns
        if not columns:
            self.layers = [[blake2s(b"")]]
            self.lifting_log = 0
            return
        logs = [len(c).bit_length() - 1 for c in columns]
        L = max(logs) if lifting_log is None else lifting_log
        self.lifting_log = L
        n = 1 << L
        mat = np.empty((n, len(columns)), dtype="<u4")
        idx = np.arange(n)
        # columns of different sizes enter the leaf hash sorted by size, smallest first, original order within a size
        # (pinned against the reference binary on an AES-CTR proof at log 9, whose trees mix log-8 table columns)
        order = sorted(range(len(columns)), key=lambda j: logs[j])
        for j, (c, lg) in enumerate(zip([columns[k] for k in order], [logs[k] for k in order])):
            if lg == L:
                mat[:, j] = c
            else:
                sh = L - lg
                mat[:, j] = np.asarray(c)[((idx >> (sh + 1)) << 1) + (idx & 1)]
        leaves = [blake2s(mat[i].tobytes()) for i in range(n)]
        self.layers = [leaves]
        while len(self.layers[-1]) > 1:
            prev = self.layers[-1]
            self.layers.append([blake2s(prev[2 * i] + prev[2 * i + 1]) for i in range(len(prev) // 2)])

    def root(self):
        return self.layers[-1][0]

    def decommit(self, positions):
        """Hash witness for sorted unique leaf positions: bottom-up, siblings not derivable, in position order."""
        witness = []
        cur = sorted(set(positions))
        for layer in self.layers[:-1]:
            nxt = []
            s = set(cur)
            i = 0
            while i < len(cur):
                p = cur[i]
                sib = p ^ 1
                if sib in s:
                    if sib > p:
                        i += 1  # skip sibling, both known
                else:
                    witness.append(layer[sib])
                nxt.append(p >> 1)
                i += 1
            cur = sorted(set(nxt))
        return witness
