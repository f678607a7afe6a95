"""Where a product-size proof spends its time: device ms (CUDA events) and host wall-clock ms per stage of ONE proof.
usage: small_latency.py [chacha20|aes128|aes256] [n_blocks]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import zk_symmetric_crypto_b200 as z
from make_golden import case_inputs
from make_golden_aes import aes_case_inputs

algo = sys.argv[1] if len(sys.argv) > 1 else "chacha20"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else (2 if algo == "chacha20" else 5)
be = z.Backend(0)
if algo == "chacha20":
    inp = case_inputs(nb, 0)
    call = lambda: be.prove_chacha20_raw(*inp)
else:
    inp = aes_case_inputs(16 if algo == "aes128" else 32, nb, 1)
    call = lambda: be.prove_aes_ctr_raw(*inp)
for _ in range(5):
    call()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); call(); ts.append((time.perf_counter() - t0) * 1e3)
l0 = be.launch_count(); call(); l1 = be.launch_count()
be.set_profile(True)
t0 = time.perf_counter(); call(); tp = (time.perf_counter() - t0) * 1e3
dev, host = be.stage_times(), be.host_times()
print(json.dumps({"algo": algo, "blocks": nb, "ms_min": round(min(ts), 3), "ms_median": round(sorted(ts)[len(ts) // 2], 3), "launches": l1 - l0,
                  "profiled_ms": round(tp, 3), "device_ms": {k: round(v, 3) for k, v in dev.items()},
                  "host_ms": {k: round(v, 3) for k, v in host.items()}, "host_sum": round(sum(host.values()), 3)}))
