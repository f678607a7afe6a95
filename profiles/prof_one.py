"""One ChaCha20 proof at a given log size, for ncu captures (never a bench number)."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, zk_symmetric_crypto_b200 as z
L = int(sys.argv[1]) if len(sys.argv) > 1 else 14
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
key, nonce, counter, pt, ct = bench.synth_inputs(L, 0)
be = z.Backend(0)
for _ in range(reps):
    p = be.prove_chacha20_raw(key, nonce, counter, pt.tobytes(), ct.tobytes())
print("proof bytes", len(p), "launches", be.launch_count())
