for L in 12 16 18 20; do
  timeout 600 python bench.py --log-size $L --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s1_L$L.json 2> gpurun_out/bench_s1_L$L.err || echo "L=$L failed rc=$?"
  tail -c 600 gpurun_out/bench_s1_L$L.err
done
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
