#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stages.py -x -q -m gpu -k "aes and not large" 2>&1 | tail -3
python profiles/small_latency.py aes128 5 | tee gpurun_out/r02u_lat_aes.json | cut -c1-700
python profiles/small_latency.py aes256 5 | cut -c1-200
python - <<'PY'
import sys, os, json
sys.path.insert(0, os.getcwd())
import bench, torch
import zk_symmetric_crypto_b200 as z
for kl in (16, 32):
    r = bench.aes_measure(z, torch, 0, kl, 16, 3, 2)
    print(kl, round(r["ms_per_proof"], 2), {k: round(v, 2) for k, v in r["stage_ms"].items() if v > 1})
PY
