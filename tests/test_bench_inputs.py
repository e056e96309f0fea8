"""The tracked files bench.py reads for its roofline objects (profiles/pipe_peaks.json, profiles/r0N_traffic.json) carry every key it
looks up: a missing key would silently turn `roofline.bound` into the HBM fallback. CPU-only."""
import json
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pipe_peaks_file_has_the_measured_rates():
    pp = bench.pipe_peaks()
    assert "fallback" not in str(pp.get("source", ""))
    assert 20.0 < pp["fp64_tflops"] < 60.0 and 1.0 < pp["issue_warp_inst_per_clk_sm"] <= 4.0 and pp.get("int8_tops", 0) > 0


def test_traffic_file_covers_every_kernel_of_the_step():
    tj = bench.profile_counters()
    assert tj is not None and tj["file"].startswith("profiles/r0")
    alg = bench.algorithmic_bytes_per_frame()
    for k in alg:
        assert tj["kernels"].get(k, 0) > 0, "no DRAM traffic for %s in %s" % (k, tj["file"])
    for k in ("k_fast", "k_blur", "k_resize4", "k_orient_describe", "k_select", "k_match_dir"):
        assert tj["inst_executed"].get(k, 0) > 0, "no instruction count for %s" % k
    # DRAM traffic of the streaming kernels stays within 10 % of their algorithmic bytes (no re-reads): 32 frames per captured launch
    # (k_resize reads the level the launch before it wrote: part of that is still in L2, so it may stay below)
    for k, lo in (("k_fast", 0.8), ("k_resize", 0.4)):
        ratio = tj["kernels"][k] / (alg[k] * tj["frames_per_launch"])
        assert lo < ratio < 1.1, (k, ratio)
    assert 0.5e6 < tj["ba"]["k_ba_step_bytes_per_trial"] < 10e6


def test_committed_bench_line_is_a_contract_line():
    path = os.path.join(ROOT, "profiles", "r02_bench_n1.json")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["roofline"]["bound"] in ("alu", "hbm", "tensor") and 0 < d["roofline"]["frac"] <= 1.0 and d["roofline"]["traffic"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    assert d["global_ba"]["parity_after_2_steps"]["relF_positions"] < 1e-4 and d["config5"]["value"] > 0
