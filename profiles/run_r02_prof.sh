#!/bin/bash
# round 2 final profile pass (1 GPU): one `--set full` launch of each hot kernel inside a log 20 proof exported as CSV (raw
# page; source page for the leaf and constraint kernels), and the per-launch time list of one whole proof.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/prof
for k in 'leaves_tiles_kernel:leaves_kernel:40' 'ifft_low12_kernel:ifft_low12_kernel:40' 'mid12_kernel:mid12_kernel:40' '^fft_low12_kernel:fft_low12_kernel:40' 'constraints_tiles_kernel2:constraints_tiles_kernel2:30' 'bitcol_dot_kernel3:bitcol_dot_kernel3:0'; do
  name=${k%%:*}; rest=${k#*:}; tag=${rest%%:*}; skip=${rest##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$name" -s $skip -c 1 -o /tmp/prof/$tag -f python profiles/prof_one.py 20 1 > /tmp/prof/$tag.log 2>&1
  ncu -i /tmp/prof/$tag.ncu-rep --page raw --csv > gpurun_out/ncu_r02_$tag.csv 2>/dev/null
  ls -la /tmp/prof/$tag.ncu-rep | awk '{print $5, $9}'
done
ncu -i /tmp/prof/constraints_tiles_kernel2.ncu-rep --page source --csv > gpurun_out/ncu_r02_constraints_tiles_kernel2_source.csv 2>/dev/null
ncu -i /tmp/prof/leaves_kernel.ncu-rep --page source --csv > gpurun_out/ncu_r02_leaves_kernel_source.csv 2>/dev/null
# launch list of one proof (the 3 warm-up proofs and the device-resident steps are skipped by count)
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
PER=$(python -c "
import json
for l in open('gpurun_out/r02p_bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print(d['gpu_launches'] // d['steps'])")
echo "launches per proof: $PER"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * PER)) -c $PER --csv --log-file gpurun_out/launches_r02_L20.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/launches_r02_L20.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
tot = collections.OrderedDict(); cnt = collections.Counter()
for r in rows[hdr + 1:]:
    name = re.sub(r"\(.*", "", r[ki]); v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
    tot[name] = tot.get(name, 0.0) + v; cnt[name] += 1
s = sum(tot.values())
with open("gpurun_out/launches_r02_L20_by_kernel.csv", "w") as f:
    f.write("kernel,launches,total_ms,share\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write("%s,%d,%.3f,%.4f\n" % (k, cnt[k], v, v / s))
print(open("gpurun_out/launches_r02_L20_by_kernel.csv").read()[:1500])
PY
