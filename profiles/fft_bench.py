"""Micro-benchmark of the packed-witness LDE transform (cb_lde_packed): per-kernel CUDA-event times for n_words word rows
at a given log size.  Used for ncu captures and kernel iteration; never a bench number."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_symmetric_crypto_b200 as z

L = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n_words = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
be = z.Backend(0)
n, m = 1 << L, 2 << L
rng = np.random.default_rng(1)
words = rng.integers(0, 1 << 32, size=(n_words, n), dtype=np.uint64).astype(np.uint32)
d_w = be.upload(words)
d_t = be.malloc(n_words * 32 * m * 4)
be.set_profile(True)
for r in range(reps):
    be._ck(be.L.cb_lde_packed(be.ctx, 1, d_w, n_words, L, d_t))
    st = be.stage_times()
    cols = n_words * 32
    tot = sum(st.values())
    nbf = cols * (L * n // 2 + L * n)   # butterflies
    print("L=%d words=%d" % (L, n_words), {k: round(v, 3) for k, v in st.items()}, "total %.3f ms" % tot,
          "%.1f GB/s alg (20N B/col)" % (cols * 20 * n / tot / 1e6), "%.2f Gbf/s" % (nbf / tot / 1e6))
