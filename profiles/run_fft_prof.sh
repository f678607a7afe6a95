python profiles/fft_bench.py 16 16 2
python profiles/fft_bench.py 20 16 2
ncu --set full --clock-control none --import-source on -k regex:"ifft_low_kernel|mid_kernel|fft_low_kernel" -c 3 -o gpurun_out/prof_fft_r01a python profiles/fft_bench.py 18 16 1 > gpurun_out/ncu_fft.log 2>&1
tail -3 gpurun_out/ncu_fft.log
