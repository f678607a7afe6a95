#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers DESIGN.md quotes."""
import csv, subprocess, sys, io
WANT = [
    ("time_ms", "gpu__time_duration.sum"), ("regs", "launch__registers_per_thread"), ("block", "launch__block_size"),
    ("grid", "launch__grid_size"), ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("pipe_lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("dram_read_GB", "dram__bytes_read.sum"), ("dram_write_GB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
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
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("== %s" % rep)
    for r in rows[2:]:
        print("  kernel: %s" % r[idx["Kernel Name"]][:110])
        line = []
        for name, key in WANT:
            if key in idx:
                v = r[idx[key]]
                try:
                    v = "%.4g" % float(v)
                except ValueError:
                    pass
                line.append("%s=%s%s" % (name, v, (" " + units[idx[key]]) if name in ("time_ms",) else ""))
        for i in range(0, len(line), 7):
            print("    " + "  ".join(line[i:i + 7]))
