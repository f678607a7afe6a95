#!/bin/bash
# round 2: product-size path (materialised LDE, one launch per pass): parity tests + latency + throughput
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for nb in 2 64 1024; do python profiles/small_latency.py chacha20 $nb | tee gpurun_out/r02l_lat_chacha_$nb.json | cut -c1-700; done
python profiles/small_proofs_bench.py chacha20 2048 2>&1 | tee gpurun_out/r02l_small_tp.txt
