#!/usr/bin/env python3
"""Regenerates profiles/ from the files tools/collect_round_artifacts.sh left in gpurun_out/ (run here, after the gpurun call):
ncu summaries, launch list, DRAM-traffic table, pipe-utilisation table, bench lines and profiles/README.md.
usage: python tools/refresh_profiles.py [round tag, default r01]"""
import csv, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def run(*a):
    return subprocess.run([sys.executable] + list(a), cwd=ROOT, capture_output=True, text=True).stdout


def pipe_table(reps):
    cols = [("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"), ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %")]
    lines, seen = [], set()
    for rep in reps:
        if not os.path.exists(rep):
            continue
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        h = rows[0]; kn = h.index("Kernel Name")
        for r in rows[2:]:
            name = r[kn].split("(")[0].replace("mage::", "")
            if name in seen:
                continue
            seen.add(name)
            lines.append("| `%s` | " % name + " | ".join("%.0f" % float(r[h.index(c)].replace(",", "")) if c in h else "n/a" for c, _ in cols) + " |")
    return "| kernel | " + " | ".join(n for _, n in cols) + " |\n|---|" + "---:|" * len(cols) + "\n" + "\n".join(lines) + "\n"


def main():
    t = lambda n: os.path.join(P, "%s_%s" % (TAG, n))
    run("tools/ncu_summary.py", os.path.join(G, "orb_full.ncu-rep"), t("orb_ncu_full.md"), "/tmp/_traffic_orb.json")
    run("tools/ncu_summary.py", os.path.join(G, "match_full.ncu-rep"), "/tmp/_match.md", "/tmp/_traffic_match.json")
    with open(t("orb_ncu_full.md"), "a") as f:
        f.write("".join(open("/tmp/_match.md").read().splitlines(True)[2:]))
    run("tools/ncu_summary.py", os.path.join(G, "ba_full.ncu-rep"), t("ba_ncu_full.md"))
    run("tools/ncu_summary.py", os.path.join(G, "ba_many_full.ncu-rep"), t("ba_batched_ncu_full.md"))
    run("tools/ncu_summary.py", os.path.join(G, "fast_tma_full.ncu-rep"), t("fast_tma_ncu_full.md"))
    open(t("launches_summary.md"), "w").write(run("tools/summarize_launches.py", os.path.join(G, "launches.csv"),
                                                   "%s launch list: bench.py --steps 2 --warmup 3 --ba-problems 8 --cpu-frames 2" % TAG))
    shutil.copy(os.path.join(G, "launches.csv"), t("launches.csv"))
    shutil.copy(os.path.join(G, "bench_n1.json"), t("bench_n1.json"))
    # DRAM traffic per kernel in the format bench.py reads
    a = json.load(open("/tmp/_traffic_orb.json")); a.update(json.load(open("/tmp/_traffic_match.json")))
    resize = 0.0
    for ln in open(t("orb_ncu_full.md")).read().splitlines():          # sum of the first 7 level launches
        m = re.match(r"\| `k_resize4` \| [^|]+\| ([\d.]+) Mbyte \| ([\d.]+) [MK]?byte", ln)
        if m and resize < 7e9 and ln.count("k_resize4") and a.setdefault("_n", 0) < 7:
            a["_n"] += 1; resize += float(m.group(1)) * 1e6
    a.pop("_n", None); a.pop("k_resize4", None)
    a["k_resize"] = resize
    a["k_blur"] = a.get("k_blur7f", 0) + a.get("k_blur7f_edges", 0)
    json.dump({"frames_per_launch": 32, "source": "ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum), tools/quick_bench.py 32 3 (%s_orb_ncu_full.md); "
               "k_resize = sum of the 7 level launches, k_blur = k_blur7f + k_blur7f_edges" % TAG, "kernels": a}, open(t("traffic.json"), "w"), indent=1)
    by_func = "\n".join(run("tools/ncu_by_func.py", os.path.join(G, "ba_many_full.ncu-rep"), "9k_ba_stepE", "mageslam_b200/csrc/ba.o",
                            "mageslam_b200/csrc/ba.cu", "k_ba_step").splitlines()[:14])
    pipes = pipe_table([os.path.join(G, n) for n in ("orb_full.ncu-rep", "match_full.ncu-rep", "fast_tma_full.ncu-rep", "ba_many_full.ncu-rep")])
    d = json.load(open(t("bench_n1.json")))
    rows = "".join("| `%s` | %.3f | %.1f%% | %.0f |\n" % (n, v["ms"], 100 * v["share"], v["GBps"]) for n, v in d["roofline"]["kernels"].items())
    readme = open(os.path.join(P, "README.md")).read()
    readme = re.sub(r"(\| kernel \| ms per 128-frame step \| share \| algorithmic GB/s \|\n\|---\|---:\|---:\|---:\|\n)(\|.*\n)+", lambda m: m.group(1) + rows, readme)
    readme = re.sub(r"(## Pipe utilisation per kernel[^\n]*\n\n[^\n]*\n\n)(\|.*\n)+", lambda m: m.group(1) + pipes, readme)
    readme = re.sub(r"```\n(total samples.*?)```", lambda m: "```\n" + by_func + "\n```", readme, flags=re.S)
    open(os.path.join(P, "README.md"), "w").write(readme)
    print("value %.0f  e2e %.0f  ba %.0f" % (d["value"], d["e2e"]["value"], d["ba"]["value"]))
    print("profiles/ refreshed; update the headline bullets of profiles/README.md and DESIGN.md section 6 by hand if the numbers moved")


if __name__ == "__main__":
    main()
