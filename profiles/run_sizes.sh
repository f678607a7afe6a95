python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for L in 18 20; do
  timeout 900 python bench.py --log-size $L --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s9_L$L.json 2> gpurun_out/bench_s9_L$L.err || echo "L=$L failed rc=$?"
  tail -c 600 gpurun_out/bench_s9_L$L.err
done
