#!/usr/bin/env python3
"""Measures the FP64 FMA rate and the instruction-issue rates of this pool's B200 with tools/fp64_peak.cu and tools/pipe_probe.cu
(both built here for sm_100a, run on the GPU box) and writes profiles/pipe_peaks.json -- the denominators bench.py uses for the
"fp64" and "alu" rooflines (MEASURED_PEAKS.json, driver-written, only holds the HBM and bf16 tensor peaks).
usage (under gpurun):  python tools/measure_pipe_peaks.py [--out profiles/pipe_peaks.json]"""
import argparse, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def build(name):
    exe = os.path.join(ROOT, "tools", name)
    if not os.path.exists(exe):
        subprocess.run(["nvcc", *ARCH, "-O3", "-o", exe, exe + ".cu"], check=True)
    return exe


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "pipe_peaks.json"))
    ap.add_argument("--from-logs", nargs=2, metavar=("FP64_LOG", "PIPE_LOG"), help="parse earlier outputs of the two probes instead of running them")
    a = ap.parse_args()
    if a.from_logs:
        fp64_txt, pipe_txt = (open(p).read() for p in a.from_logs)
    else:
        fp64_txt = subprocess.run([build("fp64_peak")], capture_output=True, text=True, check=True).stdout
        pipe_txt = subprocess.run([build("pipe_probe")], capture_output=True, text=True, check=True).stdout
    fp64 = max(float(m) for m in re.findall(r"DFMA:.*?->\s*([0-9.]+) TFLOP/s", fp64_txt))
    rates = {m[0].strip(): float(m[1]) for m in re.findall(r"^(.+?)\s+[0-9.]+ ms\s+([0-9.]+) warp-instr/clk/SM", pipe_txt, re.M)}
    out = {
        "fp64_tflops": fp64,
        # one pipe (ALU: packed 16-bit min/max/add, LOP3, SHF, PRMT; or FMA: IMAD, FFMA, HFMA2) issues 2 warp instructions / clk / SM,
        # the four schedulers together 4 (measured with an ALU + FMA mix)
        "alu_pipe_warp_inst_per_clk_sm": max(rates.get("VIMNMX3.S16x2", 0.0), rates.get("VIADD.16x2", 0.0), rates.get("VIMNMX.S16x2 max+min", 0.0)),
        "issue_warp_inst_per_clk_sm": max(rates.get("IMAD (mad.lo x1)", 0.0), rates.get("VIMNMX3.S16x2 + IMAD", 0.0)),
        "int8_tops": 4500.0,
        "rates_warp_inst_per_clk_sm": rates,
        "source": "tools/measure_pipe_peaks.py (tools/fp64_peak.cu, tools/pipe_probe.cu) on a B200 of this pool at 1965 MHz; int8_tops is the nominal dense figure",
    }
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
