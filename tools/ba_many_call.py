"""One batched StepMany call on 296 local windows (for ncu captures of k_ba_step)."""
import sys
sys.path.insert(0, ".")
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib, StepMany
import torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 296
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prob = synth.ba_problem()
bs = [BundlerLib().load(prob) for _ in range(N)]
StepMany(bs, [1.8], 1e9)
torch.cuda.synchronize()
m = StepMany(bs, [1.8] * iters, 1e9)
torch.cuda.synchronize()
print(float(m[0]), bs[0].stats())
