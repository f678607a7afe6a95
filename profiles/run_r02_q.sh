#!/bin/bash
# round 2: full GPU suite + product-size latency/throughput after the fused small-tree / basis kernels + default bench line
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python profiles/small_latency.py chacha20 2 | tee gpurun_out/r02q_lat_chacha.json | cut -c1-500
python profiles/small_latency.py aes128 5 | tee gpurun_out/r02q_lat_aes.json | cut -c1-300
python profiles/small_proofs_bench.py chacha20 2048 2>&1 | tee gpurun_out/r02q_small_tp.txt
python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r02q_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["verified"], d["gpu_launches"])
        print(json.dumps(d["roofline"])[:900])
        print(json.dumps(d.get("cpu_baseline"))[:500])
        print(json.dumps(d.get("proof_batches_cfg4"))[:600])
        print({k: v.get("ms_per_proof") for k, v in d.get("aes_ctr", {}).items()})
PY
