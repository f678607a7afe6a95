# round artefacts: default bench (both arms), launch list of one proof under ncu
python bench.py > gpurun_out/BENCH_local_cuda.json 2> gpurun_out/BENCH_local_cuda.err; tail -c 300 gpurun_out/BENCH_local_cuda.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 2613 -c 871 --csv --log-file gpurun_out/launches_r01_L20.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
