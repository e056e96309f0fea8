"""Latency of the tracking thread's pose-only bundle adjustment (ref TrackLocalMap::OptimizeCameraPose, TrackLocalMap.cpp:421-501:
a new BundlerLib per call, ArePointsFixed, one camera, 3 then 4 LM iterations) on the GPU and with the compiled reference on one host
core. usage: python tools/probe_pose_only.py [points]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib, BundlerParameters
from tests.oracle_ba import BaOracle, have_ref

P = int(sys.argv[1]) if len(sys.argv) > 1 else 300
probs = [synth.ba_problem(K=1, P=P, obs_per_point=1, n_fixed=0, pose_sigma=0.03, seed=5 + i) for i in range(8)]


def run(make, reps):
    ts = {"load": [], "step3": [], "step4": [], "pose": [], "total": []}
    for r in range(reps):
        t0 = time.perf_counter(); b = make().load(probs[r % 8])
        t1 = time.perf_counter(); b.StepBundleAdjustment([2.0] * 3, 25.0)
        t2 = time.perf_counter(); b.StepBundleAdjustment([2.0] * 4, 25.0)
        t3 = time.perf_counter(); b.poses()
        t4 = time.perf_counter()
        for k, v in zip(ts, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0)):
            ts[k].append(v)
    return {k: 1e6 * float(np.median(v[len(v) // 4:])) for k, v in ts.items()}


g = run(lambda: BundlerLib(BundlerParameters(True)), 200)
print("GPU  pose-only BA, %d points, fresh instance per call (us): %s" % (P, {k: round(v, 1) for k, v in g.items()}))
c = run(lambda: BaOracle("ref" if have_ref() else "port", True), 40)
print("CPU  %s (one core)                                    (us): %s" % ("compiled reference" if have_ref() else "oracle port", {k: round(v, 1) for k, v in c.items()}))
