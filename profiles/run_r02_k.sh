#!/bin/bash
# round 2: full GPU suite, product-size latency, log 20 bench line (1 GPU)
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python profiles/small_latency.py chacha20 2 | tee gpurun_out/r02k_lat_chacha.json
python profiles/small_latency.py aes128 5 | tee gpurun_out/r02k_lat_aes.json
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/r02k_bench.err > gpurun_out/r02k_bench.json
python - <<'PY'
import json
for l in open("gpurun_out/r02k_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["verified"], json.dumps(d["stage_ms"]))
PY
