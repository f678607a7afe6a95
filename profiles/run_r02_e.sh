#!/bin/bash
# round 2, call E: constraint kernel with software-pipelined loads (batch size x blocks per SM), bitcol v1/v3
cd "$(dirname "$0")/.."
python profiles/stage_times.py 20 2 | tee gpurun_out/r02e_sb4mb5.json | cut -c1-560
for v in sb2mb5 sb4mb4 sb2mb4 sb2mb6; do S2C_B200_LIB=build/variants/lib_$v.so python profiles/stage_times.py 20 2 | tee gpurun_out/r02e_$v.json | cut -c1-560; done
S2C_BITCOL_V=1 python profiles/stage_times.py 20 2 | tee gpurun_out/r02e_bitcol1.json | cut -c1-560
