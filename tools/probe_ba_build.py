import sys, time, os
sys.path.insert(0, ".")
import torch
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib
prob = synth.ba_problem(seed=1)
b0 = BundlerLib().load(prob); b0.StepBundleAdjustment([1.8], 1e9)
for r in range(2):
    t0 = time.perf_counter(); b = BundlerLib().load(prob); t1 = time.perf_counter(); b.StepBundleAdjustment([1.8], 1e9); t2 = time.perf_counter()
    print("load %.2f ms  first step %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), file=sys.stderr)
