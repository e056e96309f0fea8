#!/usr/bin/env python3
"""Pipe utilisation, issue and stall breakdown of every launch in an `ncu --set full` report (one column per launch).
usage: ncu_pipes.py <report.ncu-rep> [kernel regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]; rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
kn = idx["Kernel Name"]
data = [r for r in rows[2:] if rx is None or rx.search(r[kn])]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp16.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__warps_eligible.avg.per_cycle_active"]
want += sorted(h for h in hdr if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h)
print("| metric | " + " | ".join(r[kn].split("(")[0].replace("mage::", "")[:28] for r in data) + " |")
print("|---|" + "---:|" * len(data))
for m in want:
    if m not in idx: continue
    cells = []
    for r in data:
        v = r[idx[m]]
        try:
            f = float(v.replace(",", "")); v = "%.3g" % f if abs(f) < 1e5 else "%.4g" % f
        except ValueError: pass
        cells.append(v)
    if m.startswith("smsp__average_warp") and all(c in ("0", "0.0") or (c.replace(".", "").isdigit() and float(c) < 0.15) for c in cells): continue
    print("| %s [%s] | " % (m.replace("smsp__average_warps_issue_stalled_", "stall ").replace("smsp__average_warp_latency_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), units[idx[m]]) + " | ".join(cells) + " |")
