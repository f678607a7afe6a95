#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or fixture or oracle or cache or pool or block_air or error" 2>&1 | tail -3
python profiles/small_latency.py chacha20 2 | tee gpurun_out/r02s_lat_chacha.json | cut -c1-900
python profiles/small_latency.py aes128 5 | cut -c1-120
python profiles/small_proofs_bench.py chacha20 2048 2>&1 | tee gpurun_out/r02s_small_tp.txt | tail -3
