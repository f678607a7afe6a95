#!/bin/bash
cd "$(dirname "$0")/.."
python profiles/stage_times.py 20 3 | tee gpurun_out/r02f_stage.json | cut -c1-560
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or large or oracle_bytes or pipeline or sharded" 2>&1 | tail -3
python bench.py --no-extras --steps 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02f_bench.json')); print('value ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['verified'], d['roofline']['int'], d['roofline']['all_kernels'])"
