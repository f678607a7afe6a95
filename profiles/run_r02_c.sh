#!/bin/bash
# round 2, call C: constraint kernel A/B (v1 integer accumulate vs v2 FP64 accumulate at 4/5/6 blocks per SM), log 20, proof hash must agree
cd "$(dirname "$0")/.."
S2C_CONS_V1=1 python profiles/stage_times.py 20 2 | tee gpurun_out/r02c_v1.json | cut -c1-700
python profiles/stage_times.py 20 2 | tee gpurun_out/r02c_v2_mb5.json | cut -c1-700
S2C_B200_LIB=build/variants/lib_c2mb4.so python profiles/stage_times.py 20 2 | tee gpurun_out/r02c_v2_mb4.json | cut -c1-700
S2C_B200_LIB=build/variants/lib_c2mb6.so python profiles/stage_times.py 20 2 | tee gpurun_out/r02c_v2_mb6.json | cut -c1-700
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or large or oracle_bytes" 2>&1 | tail -3
