ncu --set full --clock-control none --import-source on -k regex:constraints_tiles -s 10 -c 1 -o gpurun_out/prof_cons_r01 python profiles/prof_one.py 18 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:leaves_kernel -s 10 -c 1 -o gpurun_out/prof_leaves_r01 python profiles/prof_one.py 18 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ifft_low12|mid12|fft_low12" -s 30 -c 3 -o gpurun_out/prof_fft_r01b python profiles/prof_one.py 20 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
