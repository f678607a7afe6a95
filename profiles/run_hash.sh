S2C_BLAKE_FMA=1 python profiles/hash_bench.py 21 384
S2C_BLAKE_FMA=0 python profiles/hash_bench.py 21 384
S2C_BLAKE_FMA=1 ncu --set full --clock-control none --import-source on -k regex:leaves_kernel -c 1 -o gpurun_out/prof_hash_fma python profiles/hash_bench.py 21 192 > /dev/null 2>&1
S2C_BLAKE_FMA=0 ncu --set full --clock-control none --import-source on -k regex:leaves_kernel -c 1 -o gpurun_out/prof_hash_alu python profiles/hash_bench.py 21 192 > /dev/null 2>&1
ls gpurun_out/*.ncu-rep
