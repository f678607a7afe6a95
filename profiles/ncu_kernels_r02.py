#!/usr/bin/env python3
"""Builds profiles/ncu_kernels_r02.json from the committed `ncu --page raw --csv` exports under profiles/r02/ (one
`--set full` launch of each hot kernel inside a log 20 proof, captured by profiles/run_r02_prof.sh).  bench.py reads the JSON
for `roofline.traffic`: DRAM bytes of the captured launch, its algorithmic bytes (derived from the grid), and their ratio,
which bench.py applies to the algorithmic bytes of the proof it has just timed.  Also writes a SASS opcode histogram of the hot
kernels (profiles/r02/sass_histogram.txt) when cuobjdump and build/*.o are present."""
import collections
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ncu_extract

L = 20
N = 1 << L
R = os.path.join(HERE, "r02")


def one(name):
    p = os.path.join(R, "ncu_r02_%s.csv" % name)
    return ncu_extract.read(p)[0] if os.path.exists(p) else None


def entry(d, alg, what):
    dram = d["dram_read_B"] + d["dram_write_B"]
    keep = ("time_us", "regs", "block", "grid", "warps_active_pct", "issue_active_pct", "pipe_alu_pct", "pipe_fma_pct", "pipe_fp64_pct",
            "pipe_lsu_pct", "dram_pct", "l2_pct", "smem_wavefronts", "smem_bank_conflicts", "stall_long_sb", "stall_math_throttle",
            "stall_not_selected", "stall_wait", "stall_no_inst")
    e = {k: d[k] for k in keep if k in d}
    e.update({"kernel": d["kernel"][:100], "dram_bytes": dram, "dram_read_bytes": d["dram_read_B"], "dram_write_bytes": d["dram_write_B"],
              "algorithmic_bytes": alg, "traffic_over_algorithmic": dram / alg, "launch": what})
    return e


out = {"source": "profiles/r02/ncu_r02_*.csv (ncu --set full --clock-control none, one launch each inside a log 20 proof; "
                 "profiles/run_r02_prof.sh)", "log_n_rows": L}
a, b, c = one("ifft_low12_kernel"), one("mid12_kernel"), one("fft_low12_kernel")
if a and b and c:
    jobs = int(round(a["grid"] / (N / 4096) / 8))           # grid = (N/4096 chunks, jobs * 8 column groups)
    cols = 32 * jobs
    alg = cols * N * 20                                      # interpolate 8 N + evaluate 12 N bytes per column
    out["ifft_low12_kernel"] = entry(a, cols * N * (1 / 8 + 4), "pass A of %d packed words (%d columns)" % (jobs, cols))
    out["mid12_kernel"] = entry(b, cols * N * 12, "pass B of the same launch set")
    out["fft_low12_kernel"] = entry(c, cols * N * 16, "pass C of the same launch set")
    dram = sum(x["dram_read_B"] + x["dram_write_B"] for x in (a, b, c))
    out["fft_passes"] = {"dram_bytes": dram, "algorithmic_bytes": alg, "traffic_over_algorithmic": dram / alg,
                         "time_us": a["time_us"] + b["time_us"] + c["time_us"],
                         "launch": "passes A + B + C of one launch set of %d packed words (%d columns)" % (jobs, cols),
                         "source": out["source"]}
d = one("leaves_kernel")
if d:
    # launch 40 = a quarter-round group: 12 words absorbed (8 transformed tiles + 4 adder sums computed and stored)
    alg = 12 * 32 * 2 * N * 4 + 2 * 2 * N * 32
    out["leaves_kernel"] = entry(d, alg, "leaf absorb of one quarter-round group (12 words = 384 columns, 2^21 leaves)")
    out["leaves_kernel"]["source"] = out["source"]
d = one("constraints_tiles_kernel2")
if d:
    # one quarter-round group on rows [0, N): 12 + 4 carried-in operand tiles read, 4 sum tiles written, 4 accumulator columns
    alg = 16 * 32 * N * 4 + 4 * 32 * N * 4 + 2 * 4 * N * 4
    out["constraints_tiles_kernel"] = entry(d, alg, "constraints of one quarter-round group on the half domain (2^20 rows)")
    out["constraints_tiles_kernel"]["source"] = out["source"]
json.dump(out, open(os.path.join(HERE, "ncu_kernels_r02.json"), "w"), indent=1)
print(json.dumps({k: (v.get("traffic_over_algorithmic") if isinstance(v, dict) else v) for k, v in out.items()}, indent=1))

# ---- SASS opcode histogram of the hot kernels
objs = {"kernels_fft2.o": ("ifft_low12_kernel", "mid12_kernel", "fft_low12_kernel"), "kernels_merkle.o": ("leaves_tiles_kernel", "leaves_seq_kernel"),
        "kernels_stream.o": ("constraints_tiles_kernel2", "bitcol_dot_kernel3", "bitrow_lookup_kernel")}
lines = []
for obj, kernels in objs.items():
    p = os.path.join(HERE, "..", "build", obj)
    if not os.path.exists(p):
        continue
    try:
        sass = subprocess.run(["cuobjdump", "-sass", p], capture_output=True, text=True, timeout=300).stdout
    except Exception:
        continue
    cur, hist = None, collections.defaultdict(collections.Counter)
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = next((k for k in kernels if k in m.group(1)), None)
            cur = (cur, m.group(1)) if cur else None
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            hist[cur][m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("IMAD", "LDG", "STG", "LDS", "STS", "LEA")) and "." in m.group(1) else "")] += 1
    for (k, full), h in hist.items():
        tot = sum(h.values())
        lines.append("== %s  (%s)  %d instructions" % (k, full[:90], tot))
        lines.append("   " + "  ".join("%s %d" % kv for kv in h.most_common(18)))
if lines:
    open(os.path.join(R, "sass_histogram.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))
