#!/usr/bin/env python3
"""Aggregates tools/ncu_by_line.py output by device function (line ranges of __device__/__global__ definitions).
usage: ncu_by_func.py <report.ncu-rep> <mangled kernel regex> <object.o> <source.cu> <ncu kernel regex>"""
import bisect, re, subprocess, sys
rep, mang, obj, src, kn = sys.argv[1:6]
funcs = []
for no, ln in enumerate(open(src), 1):
    if re.match(r"(template.*)?\s*(__device__|__global__)", ln) or re.match(r"^__(device|global)__", ln):
        m = re.search(r"(\w+)\s*\(", ln.split("__launch_bounds__")[-1] if "__launch_bounds__" not in ln else ln[ln.index(")") + 1:])
        funcs.append((no, m.group(1) if m else ln.strip()[:30]))
starts = [f[0] for f in funcs]
out = subprocess.run([sys.executable, "tools/ncu_by_line.py", rep, mang, obj, "5000", kn], capture_output=True, text=True).stdout
agg = {}
base = src.split("/")[-1]
for ln in out.splitlines():
    m = re.match(r"\s*([\d.]+)% samples\s+([\d.]+)% instr\s+(\S+):(\d+)", ln)
    if not m:
        if ln.startswith("total"): print(ln)
        continue
    s, i, fn, no = float(m.group(1)), float(m.group(2)), m.group(3), int(m.group(4))
    k = funcs[bisect.bisect_right(starts, no) - 1][1] if fn == base and starts and no >= starts[0] else fn
    a = agg.setdefault(k, [0, 0]); a[0] += s; a[1] += i
for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if s >= 0.2 or i >= 0.2: print("%-28s samples %5.1f%%  instr %5.1f%%" % (k, s, i))
