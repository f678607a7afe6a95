"""Micro-benchmark of the streaming Blake2s leaf absorb (cb_merkle_leaves_absorb) on n_cols LDE columns of 2^log rows."""
import sys, os, ctypes, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_symmetric_crypto_b200 as z
L = int(sys.argv[1]) if len(sys.argv) > 1 else 21
ncols = int(sys.argv[2]) if len(sys.argv) > 2 else 384
be = z.Backend(0)
m = 1 << L
rng = np.random.default_rng(1)
d_c = be.upload(rng.integers(0, 2**31 - 1, size=(ncols, m), dtype=np.uint64).astype(np.uint32))
d_s = be.malloc(m * 32); d_o = be.malloc(m * 32)
import torch
for r in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    be.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    be._ck(be.L.cb_merkle_leaves_absorb(be.ctx, d_c, ctypes.c_size_t(m), ncols, L, L, d_s, ctypes.c_uint64(0), 1, 0, d_o))
    be.sync()
    dt = time.perf_counter() - t0
    ncomp = m * ncols / 16
    print("L=%d cols=%d: %.3f ms  %.2f Gcompress/s  %.1f GB/s hashed" % (L, ncols, dt * 1e3, ncomp / dt / 1e9, ncomp * 64 / dt / 1e9))
