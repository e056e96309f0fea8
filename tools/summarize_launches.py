#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
usage: summarize_launches.py <launches.csv> [title]"""
import collections, csv, sys

def main():
    path = sys.argv[1]; title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, data = r, rows[i + 1:]
            break
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        name = r[kn].split("(")[0].replace("mage::", "")
        v = float(r[mv].replace(",", "")); u = r[mu]
        v = v / 1e3 if u in ("nsecond", "ns") else v * 1e3 if u in ("msecond", "ms") else v
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("## %s\n\n%d launches, %.1f us of kernel time (cold-cache, serialised by ncu: compare SHARES, not absolutes)\n" % (title, len(data), tot))
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k, n, t, t / n, 100 * t / tot))

if __name__ == "__main__":
    main()
