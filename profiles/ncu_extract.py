#!/usr/bin/env python3
"""Reads `ncu --page raw --csv` exports (one row per profiled launch) and prints / returns the handful of metrics DESIGN.md and
bench.py quote.  Usage: ncu_extract.py file.csv [...]      (writes nothing; profiles/ncu_kernels_r02.py builds the JSON)"""
import csv
import sys

WANT = [
    ("time_us", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"), ("block", "launch__block_size"),
    ("grid", "launch__grid_size"), ("smem_dyn_B", "launch__shared_mem_per_block_dynamic"),
    ("occupancy_limit_regs", "launch__occupancy_limit_registers"), ("occupancy_limit_smem", "launch__occupancy_limit_shared_mem"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("pipe_fmaheavy_pct", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active"),
    ("pipe_fp64_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("pipe_xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("pipe_lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("dram_read_B", "dram__bytes_read.sum"), ("dram_write_B", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall_math_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall_not_selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("stall_lg_throttle", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("stall_mio_throttle", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("stall_dispatch", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
    ("stall_no_inst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "second": 1e6,
              "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}


def read(path):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        d = {"kernel": r[idx["Kernel Name"]]}
        for name, key in WANT:
            if key in idx:
                try:
                    v = float(r[idx[key]].replace(",", ""))
                except ValueError:
                    continue
                u = units[idx[key]]
                if name.endswith("_B") or name == "time_us":
                    v *= UNIT_SCALE.get(u, 1.0)
                d[name] = v
        out.append(d)
    return out


if __name__ == "__main__":
    for p in sys.argv[1:]:
        for d in read(p):
            print("== %s\n   %s" % (p, d["kernel"][:120]))
            items = ["%s=%.4g" % (k, v) for k, v in d.items() if k != "kernel"]
            for i in range(0, len(items), 6):
                print("   " + "  ".join(items[i:i + 6]))
