#!/usr/bin/env python3
"""Summarises an `ncu --set full` report: one row per captured kernel launch with the metrics the roofline needs.
usage: ncu_summary.py <report.ncu-rep> <out.md> [traffic.json]"""
import csv, io, json, subprocess, sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "IPC"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("smsp__inst_executed.sum", "warp instr")]

def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kn = idx["Kernel Name"]
    lines = ["| kernel | " + " | ".join(n for _, n in WANT) + " |", "|---|" + "---:|" * len(WANT)]
    traffic = {}
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("mage::", "")
        cells = []
        for m, _ in WANT:
            if m in idx:
                v, u = r[idx[m]], units[idx[m]]
                try:
                    f = float(v.replace(",", ""))
                    v = ("%.3f" % f if abs(f) < 1000 else "%.0f" % f)
                except ValueError:
                    pass
                cells.append("%s %s" % (v, u) if u and u not in ("%",) else v)
            else:
                cells.append("n/a")
        lines.append("| `%s` | " % name + " | ".join(cells) + " |")
        try:
            def b(m):
                v = float(r[idx[m]].replace(",", "")); u = units[idx[m]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            traffic.setdefault(name, []).append(b("dram__bytes_read.sum") + b("dram__bytes_write.sum"))
        except Exception:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    if len(sys.argv) > 3:
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(sys.argv[3], "w"), indent=1)
    print("\n".join(lines))

if __name__ == "__main__":
    main()
