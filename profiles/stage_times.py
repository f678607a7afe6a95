"""Per-stage device times (CUDA events around each kernel family) of one ChaCha20 proof; used to compare build variants
(S2C_B200_LIB=<path to a variant .so>).  Not a bench number: profiling mode synchronises around every stage."""
import sys, os, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, zk_symmetric_crypto_b200 as z
if os.environ.get('WITH_TORCH'):
    import torch
    _x = torch.zeros(1 << 20, device='cuda')
L = int(sys.argv[1]) if len(sys.argv) > 1 else 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
key, nonce, counter, pt, ct = bench.synth_inputs(L, 0)
be = z.Backend(0)
p = be.prove_chacha20_raw(key, nonce, counter, pt.tobytes(), ct.tobytes())
be.set_profile(True)
best = None
for _ in range(reps):
    p = be.prove_chacha20_raw(key, nonce, counter, pt.tobytes(), ct.tobytes())
    st = be.stage_times()
    if best is None:
        best = st
    else:
        best = {k: min(best[k], st[k]) for k in st}
print(json.dumps({"lib": os.environ.get("S2C_B200_LIB", "default"), "log": L, "sha": hashlib.sha256(p).hexdigest()[:16],
                  "total": round(sum(best.values()), 2), "counters": be.counters(), **{k: round(v, 2) for k, v in best.items()}}))
