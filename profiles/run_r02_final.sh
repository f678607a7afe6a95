#!/bin/bash
# round 2 final check (1 GPU): full -m gpu suite, smoke, default bench line
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-300
python bench.py > gpurun_out/r02final_bench.json 2> gpurun_out/r02final_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/r02final_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["verified"]["ok"], d["gpu_launches"], d["clocks"])
        print(json.dumps(d.get("proof_batches_cfg4"))[:400])
        print({k: v.get("ms_per_proof") for k, v in d.get("aes_ctr", {}).items()})
PY
