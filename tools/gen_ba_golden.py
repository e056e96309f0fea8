#!/usr/bin/env python3
"""Generates tests/golden/ba_golden.npz by running the REFERENCE's own BundlerLib + g2o (oracle/_ref/libbundler_ref.so,
compiled from /root/reference by oracle/Makefile) on small seeded problems. Run in the build container only.
Each case stores, after every StepBundleAdjustment call: poses (pos, rot), points, lambda, mean error, outlier indices."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mageslam_b200 import synth
from tests.oracle_ba import BaOracle

CASES = {
    # name: (problem kwargs, points_fixed, huber per call, max_err_sq, calls)
    "local_small": (dict(K=6, P=150, obs_per_point=3, seed=11), False, [1.8], 1e9, 6),
    "local_outliers": (dict(K=6, P=200, obs_per_point=4, seed=12, outlier_frac=0.06), False, [1.8, 1.8], 7.25, 4),
    "pose_only": (dict(K=1, P=120, obs_per_point=1, seed=13, n_fixed=0, pose_sigma=0.03), True, [2.0, 2.0, 2.0], 25.0, 2),
    "info_weights": (dict(K=5, P=120, obs_per_point=3, seed=14, info_mode="confidence"), False, [0.9], 1e9, 5),
    # tether edges (BundlerLib.cpp:24-90, :311-350): fixed-distance, relative-rotation and relative-transform constraints
    "tether_distance": (dict(K=6, P=150, obs_per_point=3, seed=15, tethers=dict(seed=1, n_distance=4, n_rotation=0, n_transform=0, noise=1e-2)), False, [1.8, 1.8], 1e9, 3),
    "tether_rotation": (dict(K=6, P=150, obs_per_point=3, seed=16, tethers=dict(seed=2, n_distance=0, n_rotation=4, n_transform=0, noise=1e-2)), False, [1.8, 1.8], 1e9, 3),
    "tether_transform": (dict(K=6, P=150, obs_per_point=3, seed=17, tethers=dict(seed=3, n_distance=0, n_rotation=0, n_transform=4, noise=1e-2)), False, [1.8, 1.8], 1e9, 3),
    "tether_all_outliers": (dict(K=7, P=200, obs_per_point=4, seed=18, outlier_frac=0.05, tethers=dict(seed=4, n_distance=3, n_rotation=3, n_transform=3, noise=1e-2)), False, [1.8, 1.8], 7.25, 3),
}


def build_problem(kw):
    """CASES kwargs -> synth problem dict (the optional 'tethers' entry goes to synth.ba_add_tethers)."""
    kw = dict(kw)
    teth = kw.pop("tethers", None)
    prob = synth.ba_problem(**kw)
    return synth.ba_add_tethers(prob, **teth) if teth else prob


def main():
    out = {}
    for name, (kw, pf, hub, mx, calls) in CASES.items():
        prob = build_problem(kw)
        ref = BaOracle("ref", pf).load(prob)
        for c in range(calls):
            mean, outl = ref.StepBundleAdjustment(hub, mx)
            pos, rot = ref.poses()
            out["%s/%d/pos" % (name, c)] = pos; out["%s/%d/rot" % (name, c)] = rot
            out["%s/%d/pts" % (name, c)] = ref.points()
            out["%s/%d/scalars" % (name, c)] = np.array([mean, ref.GetCurrentLambda()], np.float64)
            out["%s/%d/outliers" % (name, c)] = outl.astype(np.int64)
        print(name, "final mean", mean, "lambda", ref.GetCurrentLambda())
    np.savez_compressed("tests/golden/ba_golden.npz", **out)
    print("wrote tests/golden/ba_golden.npz", sum(v.nbytes for v in out.values()), "bytes raw")

if __name__ == "__main__":
    main()
