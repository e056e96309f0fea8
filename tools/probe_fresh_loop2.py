import sys, time, os
sys.path.insert(0, ".")
import torch
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib, StepMany
probs = [synth.ba_problem(seed=1 + i) for i in range(8)]
hub = [1.8] * 10
bs = [BundlerLib().load(probs[i % 8]) for i in range(64)]
StepMany(bs, [1.8], 1e9); StepMany(bs, hub, 1e9); torch.cuda.synchronize()
del bs
for r in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b = BundlerLib(); t1 = time.perf_counter()
    b.load(probs[0]); t2 = time.perf_counter()
    b.StepBundleAdjustment([1.8], 1e9); t3 = time.perf_counter()
    b.StepBundleAdjustment(hub[:9], 1e9); t4 = time.perf_counter()
    del b; t5 = time.perf_counter()
    print("create %.2f load %.2f first step %.2f nine more %.2f destroy %.2f ms" % tuple((y - x) * 1e3 for x, y in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5))), file=sys.stderr)
