# round artefacts: default bench (both arms), launch list of one proof under ncu (the 3 warm-up proofs are skipped)
python bench.py > gpurun_out/BENCH_local_cuda.json 2> gpurun_out/BENCH_local_cuda.err; tail -c 300 gpurun_out/BENCH_local_cuda.err
PER=$(python -c "import json; d=json.loads(open('gpurun_out/BENCH_local_cuda.json').read().strip().split(chr(10))[-1]); print(d['gpu_launches'] // d['steps'])")
echo "launches per proof: $PER"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/BENCH_local_ref.json 2> gpurun_out/BENCH_local_ref.err; tail -c 600 gpurun_out/BENCH_local_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * PER)) -c $PER --csv --log-file gpurun_out/launches_r01_L20.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/launches_r01_L20.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
tot = collections.OrderedDict(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    name = re.sub(r"\(.*", "", r[ki]); v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    tot[name] = tot.get(name, 0.0) + v; cnt[name] += 1
s = sum(tot.values())
with open("gpurun_out/launches_r01_L20_by_kernel.csv", "w") as f:
    f.write("kernel,launches,total_ms,share\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write("%s,%d,%.3f,%.3f\n" % (k, cnt[k], v, v / s))
print(open("gpurun_out/launches_r01_L20_by_kernel.csv").read()[:1200])
PY
