#!/bin/bash
# round 2 profile pass 1: one `--set full` launch of each hot kernel inside a log 20 proof, exported as CSV (raw page; source page
# for the constraint kernel), plus the bitcol A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/prof
for k in 'leaves_kernel:40' 'ifft_low12_kernel:40' 'mid12_kernel:40' '^fft_low12_kernel:40' 'constraints_tiles_kernel2:30' 'bitcol_dot_kernel2:0'; do
  name=${k%%:*}; skip=${k##*:}; tag=$(echo $name | tr -d '^')
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$name" -s $skip -c 1 -o /tmp/prof/$tag -f python profiles/prof_one.py 20 1 > /tmp/prof/$tag.log 2>&1
  ncu -i /tmp/prof/$tag.ncu-rep --page raw --csv > gpurun_out/ncu_r02_$tag.csv 2>/dev/null
  ls -la /tmp/prof/$tag.ncu-rep | awk '{print $5, $9}'
done
ncu -i /tmp/prof/constraints_tiles_kernel2.ncu-rep --page source --csv > gpurun_out/ncu_r02_constraints_tiles_kernel2_source.csv 2>/dev/null
ncu -i /tmp/prof/leaves_kernel.ncu-rep --page source --csv > gpurun_out/ncu_r02_leaves_kernel_source.csv 2>/dev/null
S2C_BITCOL_V1=1 python profiles/stage_times.py 20 2 | tee gpurun_out/r02p1_bitcol_v1.json | cut -c1-600
python profiles/stage_times.py 20 2 | tee gpurun_out/r02p1_bitcol_v2.json | cut -c1-600
ls -la gpurun_out/ | tail -12
