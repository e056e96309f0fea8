#!/usr/bin/env python3
"""Maps an ncu SASS-level source page (ncu -i rep --page source --csv -k regex:<kernel>) onto CUDA source lines using
nvdisasm -g line markers of the same kernel in an object file. Prints stall samples / executed instructions per source line.
usage: ncu_by_line.py <report.ncu-rep> <kernel regex> <object.o> [top N]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

def sass_lines(obj, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    out = []
    for f in os.listdir(d):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        cur_fn, line = None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1); continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2))); continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn and re.search(kernel, cur_fn):
                out.append((int(m.group(1), 16), line, m.group(2)))
    return out

def main():
    rep, kernel, obj = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    sl = sass_lines(obj, kernel)
    kernel_ncu = sys.argv[5] if len(sys.argv) > 5 else kernel
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kernel_ncu], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(csvtxt)))
    hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
    hdr = rows[hi]; si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
    stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[hi + 1:] if len(r) > si and r[si].isdigit()]
    if len(data) % len(sl) == 0 and len(data) != len(sl):
        data = data[:len(sl)]                     # several captured launches are concatenated: use the first
    assert len(data) == len(sl), (len(data), len(sl))
    agg = collections.OrderedDict()
    for (addr, line, ins), r in zip(sl, data):
        a = agg.setdefault(line, [0, 0, collections.Counter()]); a[0] += int(r[si]); a[1] += int(r[ie])
        for i, h in stall_cols:
            if r[i].isdigit(): a[2][h] += int(r[i])
    ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
    print("total samples %d, warp instructions %d" % (ts, ti))
    src = {}
    allst = collections.Counter()
    for a in agg.values(): allst.update(a[2])
    print("stall reasons overall:", ", ".join("%s %.1f%%" % (h, 100 * v / max(1, sum(allst.values()))) for h, v in allst.most_common(8)))
    for line, (s, i, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if line is None:
            print("%6.2f%% samples %6.2f%% instr  <no line info>" % (100 * s / ts, 100 * i / ti)); continue
        fn, no = line
        if fn not in src:
            for root in ("mageslam_b200/csrc", "."):
                pth = os.path.join(root, fn)
                if os.path.exists(pth):
                    src[fn] = open(pth).read().splitlines(); break
            else:
                src[fn] = []
        text = src[fn][no - 1].strip()[:110] if 0 < no <= len(src[fn]) else ""
        why = " ".join("%s:%d%%" % (h, 100 * v / max(1, sum(st.values()))) for h, v in st.most_common(2))
        print("%6.2f%% samples %6.2f%% instr  %s:%d  [%s]  %s" % (100 * s / ts, 100 * i / ti, fn, no, why, text))

if __name__ == "__main__":
    main()
