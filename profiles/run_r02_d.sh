#!/bin/bash
# round 2, call D: new stage-level tests, prove_stream generator, leaves-kernel A/B, then the whole GPU suite
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_parity.py -x -q -m gpu -k "stages or prove_stream or bit_reverse or batch_inverse or barycentric or twiddles or commit_on_layer or quotients_batches or lift_and or finalize_last" 2>&1 | tail -15
S2C_LEAVES_GENERAL=1 python profiles/stage_times.py 20 2 | tee gpurun_out/r02d_leaves_general.json | cut -c1-600
python profiles/stage_times.py 20 2 | tee gpurun_out/r02d_leaves_tiles.json | cut -c1-600
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02d_full_suite.log
