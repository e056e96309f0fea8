#!/usr/bin/env python3
"""Regenerates profiles/ from the files tools/collect_round_artifacts.sh left in gpurun_out/ (run here, after the gpurun call):
ncu summaries, launch list, DRAM-traffic table, pipe-utilisation table, bench lines and profiles/README.md.
usage: python tools/refresh_profiles.py [round tag, default r02]   (parts whose reports are missing in gpurun_out/ are skipped)"""
import csv, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"


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


def ncu_md(rep, out, traffic_json=None, title=None, extra=None):
    """ncu_summary table (+ optional sections) -> out; returns False when the report is missing"""
    if not os.path.exists(rep):
        return False
    args = ["tools/ncu_summary.py", rep, out] + ([traffic_json] if traffic_json else [])
    run(*args)
    if title or extra:
        body = open(out).read()
        open(out, "w").write((("# %s\n\n" % title) if title else "") + body + (extra or ""))
    return True


def by_line(rep, kernel, obj, top=18):
    if not os.path.exists(rep):
        return ""
    return "\n".join(run("tools/ncu_by_line.py", rep, kernel, obj, str(top)).splitlines()[:top + 2])


def main():
    t = lambda n: os.path.join(P, "%s_%s" % (TAG, n))
    g = lambda n: os.path.join(G, n)
    have_orb = os.path.exists(g("orb_full.ncu-rep")) and os.path.getmtime(g("orb_full.ncu-rep")) > os.path.getmtime(os.path.join(P, "r01_orb_ncu_full.md"))
    traffic = {}
    if os.path.exists(t("traffic.json")):
        traffic = json.load(open(t("traffic.json")))
    if have_orb:
        run("tools/ncu_summary.py", g("orb_full.ncu-rep"), t("orb_ncu_full.md"), "/tmp/_traffic_orb.json")
        run("tools/ncu_summary.py", g("match_full.ncu-rep"), "/tmp/_match.md", "/tmp/_traffic_match.json")
        with open(t("orb_ncu_full.md"), "a") as f:
            f.write("".join(open("/tmp/_match.md").read().splitlines(True)[2:]))
        ncu_md(g("fast_reg_full.ncu-rep"), t("fast_reg_ncu_full.md"), title="k_fast, register-staged variant (MAGE_FAST_TMA=0)")
        open(t("launches_summary.md"), "w").write(run("tools/summarize_launches.py", g("launches.csv"),
                                                       "%s launch list: bench.py --steps 2 --warmup 3 --ba-problems 8 --cpu-frames 2" % TAG))
        shutil.copy(g("launches.csv"), t("launches.csv"))
        shutil.copy(g("bench_n1.json"), t("bench_n1.json"))
        a = json.load(open("/tmp/_traffic_orb.json")); a.update(json.load(open("/tmp/_traffic_match.json")))
        resize = 0.0
        for ln in open(t("orb_ncu_full.md")).read().splitlines():          # sum of the first 7 level launches
            m = re.match(r"\| `k_resize4` \| [^|]+\| ([\d.]+) Mbyte \| ([\d.]+) [MK]?byte", ln)
            if m and a.setdefault("_n", 0) < 7:
                a["_n"] += 1; resize += float(m.group(1)) * 1e6
        a.pop("_n", None); a.pop("k_resize4", None)
        for k in list(a):                                                 # template instances: void k_fast_tma<8, 8, 0> -> k_fast
            m = re.match(r"(?:void )?(k_[a-z0-9_]+?)(?:_tma)?<", k)
            if m:
                a[m.group(1)] = a.pop(k)
        a["k_resize"] = resize
        a["k_blur"] = a.get("k_blur7f", 0) + a.get("k_blur7f_edges", 0)
        inst = {}                                                         # executed warp instructions per launch (last column of the summary table)
        for ln in open(t("orb_ncu_full.md")).read().splitlines():
            m = re.match(r"\| `(?:void )?(k_[a-z0-9_]+?)(?:_tma)?(?:<[^`]*)?` \|.*\| ([\d.]+) inst \|$", ln)
            if m and m.group(1) not in inst:
                inst[m.group(1)] = float(m.group(2))
        if "k_blur7f" in inst:
            inst["k_blur"] = inst.get("k_blur7f", 0) + inst.get("k_blur7f_edges", 0)
        traffic["inst_executed"] = inst
        traffic.update({"frames_per_launch": 32, "source": "ncu --set full (dram__bytes_read.sum + dram__bytes_write.sum), tools/quick_bench.py 32 3 (%s_orb_ncu_full.md); "
                        "k_resize = sum of the 7 level launches, k_blur = k_blur7f + k_blur7f_edges" % TAG, "kernels": a})
    # ---- bundle adjustment
    ncu_md(g("ba_full.ncu-rep"), t("ba_ncu_full.md"), title="k_ba_step_coop, one local window (10 KF / 2000 pts / 8000 obs), tools/ba_one_call.py")
    if ncu_md(g("ba_many_full.ncu-rep"), t("ba_batched_ncu_full.md"), "/tmp/_traffic_ba.json", title="k_ba_step, 296 local windows x 10 LM iterations in one launch, tools/ba_many_call.py 296 10"):
        tb = json.load(open("/tmp/_traffic_ba.json"))
        key = [k for k in tb if "k_ba_step" in k][0]
        traffic["ba"] = {"k_ba_step_bytes_per_trial": tb[key] / (296 * 10.0), "windows": 296, "trials": 10,
                         "source": "ncu --set full of the batched launch (%s_ba_batched_ncu_full.md): DRAM read + write bytes / (296 windows x 10 lambda trials)" % TAG}
    phases = lambda n: ("\n## Phase times seen by CTA 0 (in-kernel %globaltimer, run without the profiler)\n\n```\n" + open(g(n)).read().strip() + "\n```\n") if os.path.exists(g(n)) else ""
    reps = [g("ba_global_full.ncu-rep"), g("dense_full.ncu-rep")]
    ncu_md(reps[0], t("ba_global_ncu.md"), title="k_ba_step_coop at BASELINE config 4 (500 KF / 50 000 pts / 400 000 obs, n = 2988): one LM step, tools/ba_global_call.py",
           extra="\n## Pipe utilisation\n\n" + pipe_table([reps[0]]) + phases("ba_global_phases.log") +
                 "\n## Stall samples by source line (tools/ncu_by_line.py)\n\n```\n" + by_line(reps[0], "k_ba_step_coop", "mageslam_b200/csrc/ba.o") + "\n```\n" +
                 ("\n## Parity and time against the compiled reference (tools/global_ba_check.py)\n\n```\n" + open(g("global_ba_check.log")).read().strip() + "\n```\n" if os.path.exists(g("global_ba_check.log")) else ""))
    ncu_md(reps[1], t("dense_ncu.md"), title="k_dense_debug: the dense solver of the reduced camera system alone (n = 2988; blocked LDL^T with the trailing update on tcgen05 kind::i8), tools/dense_call.py",
           extra="\n## Pipe utilisation\n\n" + pipe_table([reps[1]]) + phases("dense_phases.log") +
                 "\n## Stall samples by source line (tools/ncu_by_line.py)\n\n```\n" + by_line(reps[1], "k_dense_debug", "mageslam_b200/csrc/ba.o") + "\n```\n")
    if os.path.exists(g("quick_bench_ba.log")):
        shutil.copy(g("quick_bench_ba.log"), t("quick_bench_ba.log"))
    json.dump(traffic, open(t("traffic.json"), "w"), indent=1)
    if have_orb:
        pipes = pipe_table([g(n) for n in ("orb_full.ncu-rep", "match_full.ncu-rep", "fast_reg_full.ncu-rep", "ba_many_full.ncu-rep", "ba_global_full.ncu-rep", "dense_full.ncu-rep")])
        d = json.load(open(t("bench_n1.json")))
        rows = "".join("| `%s` | %.3f | %.1f%% | %.0f |\n" % (n, v["ms"], 100 * v["share"], v["GBps"]) for n, v in d["roofline"]["kernels"].items())
        readme = open(os.path.join(P, "README.md")).read()
        readme = re.sub(r"(\| kernel \| ms per 128-frame step \| share \| algorithmic GB/s \|\n\|---\|---:\|---:\|---:\|\n)(\|.*\n)+", lambda m: m.group(1) + rows, readme)
        readme = re.sub(r"(## Pipe utilisation per kernel[^\n]*\n\n[^\n]*\n\n)(\|.*\n)+", lambda m: m.group(1) + pipes, readme)
        open(os.path.join(P, "README.md"), "w").write(readme)
        print("value %.0f  e2e %.0f  ba %.0f" % (d["value"], d["e2e"]["value"], d["ba"]["value"]))
    print("profiles/ refreshed (%s); update the headline bullets of profiles/README.md and DESIGN.md section 6 by hand if the numbers moved" % TAG)


if __name__ == "__main__":
    main()
