#!/bin/bash
# round 2, call A: new large-size parity tests + default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large or log13 or partial_cache or headline" 2>&1 | tail -15 | tee gpurun_out/r02a_tests.log
timeout 900 python bench.py 2> gpurun_out/r02a_bench.err | tee gpurun_out/r02a_bench.json | cut -c1-1500
tail -5 gpurun_out/r02a_bench.err
