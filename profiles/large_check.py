"""Robustness check (not a bench): a log_n_rows = 21 ChaCha20 proof on one GPU (run-time-schedule FFT kernels, partial tile
cache) and an AES-128 log 17 proof, both checked by the host verifier."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, zk_symmetric_crypto_b200 as z
be = z.Backend(0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 21
key, nonce, counter, pt, ct = bench.synth_inputs(L, 0)
ptb, ctb = pt.tobytes(), ct.tobytes()
for i in range(2):
    t = time.time()
    p = be.prove_chacha20_raw(key, nonce, counter, ptb, ctb)
    dt = time.time() - t
print("chacha log", L, "proof", len(p), "bytes,", round(dt, 3), "s", be.counters())
t = time.time()
print("verify:", z.verify_chacha20_raw(p, nonce, counter, ptb, ctb), round(time.time() - t, 3), "s")
akey, anonce, acounter, apt, act = bench.synth_aes_inputs(16, 17, 0)
ap = be.prove_aes_ctr_raw(akey, anonce, acounter, apt, act)
print("aes128 log 17 proof", len(ap), "verify:", z.verify_aes_ctr_raw(ap, anonce, acounter, apt, act))
