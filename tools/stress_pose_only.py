"""Seeded sweep of the tracking thread's pose-only bundle adjustment beyond the test suite's cases: random problems (5 - 600 map points,
mild to gross pose errors, 0 - 20 % gross outliers, unit or confidence-scaled information, 1 - 8 iterations, Huber widths and error
bounds over the reference's range) through mage_optimize_camera_pose AND through the BundlerLib-shaped handle path, each against the
compiled reference (oracle/_ref) run the way TrackLocalMap::OptimizeCameraPose runs it (ref Tracking/TrackLocalMap.cpp:421-501).
usage: python tools/stress_pose_only.py [first_seed count]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib, BundlerParameters
from mageslam_b200.tracking import OptimizeCameraPose
from tests.ba_checks import TOL, best_checker
from tests.oracle_ba import have_ref, rel_frobenius

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 300
t0 = time.time()
worst, ran, flagged, diverged = 0.0, 0, 0, 0
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    kw = dict(K=1, P=int(rng.integers(5, 600)), obs_per_point=1, n_fixed=0, pose_sigma=float(rng.choice([0.005, 0.03, 0.08, 0.2])),
              outlier_frac=float(rng.choice([0.0, 0.05, 0.2])), info_mode=str(rng.choice(["one", "confidence"])), seed=seed)
    prob = synth.ba_problem(**kw)
    iters = int(rng.integers(1, 9)); hub = float(rng.choice([0.5, 1.0, 2.0, 3.0])); mx = float(rng.choice([2.0, 7.25, 25.0, 1e9]))
    pos, rot, outl, mean = OptimizeCameraPose(prob["cam_pos"][0], prob["cam_rot"][0], prob["intrinsics"][0], prob["points"], prob["obs_uv"], prob["obs_info"], iters, mx, hub)
    h = BundlerLib(BundlerParameters(True)).load(prob)
    mh = h.StepBundleAdjustment([hub] * iters, mx)
    ph, rh = h.poses()
    assert np.array_equal(pos, ph[0]) and np.array_equal(rot, rh[0].reshape(9)) and list(outl) == list(h.last_outliers), ("one call vs handle path", seed, kw)
    chk = best_checker(True).load(prob)
    mc, oc = chk.StepBundleAdjustment([hub] * iters, mx)
    pc, rc = chk.poses()
    e = max(rel_frobenius(pos, pc[0]), rel_frobenius(rot, rc[0].reshape(9)))
    same = list(outl) == list(oc)
    if not (e < TOL and same):
        # a start far outside the basin (pose_sigma 0.2 with few points) may leave the two LM runs on different sides of an accept / reject
        # decision: count it, and require the closed-form quantities of the FIRST iteration to agree instead
        p1, r1, o1, _ = OptimizeCameraPose(prob["cam_pos"][0], prob["cam_rot"][0], prob["intrinsics"][0], prob["points"], prob["obs_uv"], prob["obs_info"], 1, 1e9, hub)
        c1 = best_checker(True).load(prob); c1.StepBundleAdjustment([hub], 1e9); pc1, rc1 = c1.poses()
        e1 = max(rel_frobenius(p1, pc1[0]), rel_frobenius(r1, rc1[0].reshape(9)))
        assert e1 < TOL, ("first iteration differs", seed, kw, e1)
        diverged += 1
        print("  seed %d: differs after %d iterations (relF %.2e, outliers equal %s) but not after one (%.2e): %s" % (seed, iters, e, same, e1, kw))
    else:
        worst = max(worst, e)
    ran += 1; flagged += len(outl)
print("pose-only BA: %d random problems (%s), one call == handle path to the last float in all of them; %d within %.1e of the reference with equal outlier lists "
      "(%d observations removed in total), %d differing only after the first iteration, %.0f s" % (ran, "compiled reference" if have_ref() else "oracle port", ran - diverged, worst, flagged, diverged, time.time() - t0))
