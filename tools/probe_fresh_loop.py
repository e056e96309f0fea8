import sys, time
sys.path.insert(0, ".")
import torch
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib, StepMany
probs = [synth.ba_problem(seed=1 + i) for i in range(8)]
hub = [1.8] * 10
# like bench_ba: single, fresh, then batched sets, then fresh again
def fresh(n):
    out = []
    for r in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        b = BundlerLib().load(probs[0]); b.StepBundleAdjustment(hub, 1e9)
        out.append((time.perf_counter() - t0) * 1e3); del b
    return ["%.2f" % x for x in out]
print("fresh before batched:", fresh(6))
bs = [BundlerLib().load(probs[i % 8]) for i in range(64)]
StepMany(bs, [1.8], 1e9); StepMany(bs, hub, 1e9); torch.cuda.synchronize()
print("fresh while 64 handles alive:", fresh(6))
del bs
print("fresh after batched:", fresh(6))
