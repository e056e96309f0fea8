import os, sys, time
sys.path.insert(0, ".")
import torch
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib
prob = synth.ba_problem()
for G in (24, 32, 36, 40, 48, 64, 96, 148):
    os.environ["MAGE_BA_COOP_BLOCKS"] = str(G)
    ts = []
    for rep in range(4):
        b = BundlerLib().load(prob); b.StepBundleAdjustment([1.8], 1e9)
        torch.cuda.synchronize(); t0 = time.perf_counter(); b.StepBundleAdjustment([1.8] * 10, 1e9); ts.append(time.perf_counter() - t0)
    print("G=%3d: %.3f ms per 10 iterations -> %.0f it/s" % (G, min(ts) * 1e3, 10 / min(ts)))
