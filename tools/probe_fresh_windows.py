"""Host-side cost of a batch of FRESH local-BA windows (config 5 steps one every 20 frames): load + StepMany(10 iterations) on 12 new
BundlerLib instances, repeated; MAGE_BA_SERIAL_PREPARE=1 switches the parallel structure build off for comparison."""
import sys, time
sys.path.insert(0, ".")
import torch
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib, StepMany
probs = [synth.ba_problem(seed=100 + w) for w in range(4)]
StepMany([BundlerLib().load(probs[0])], [1.8] * 10, 1e9)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ws = [BundlerLib().load(probs[w % 4]) for w in range(12)]
    t1 = time.perf_counter()
    StepMany(ws, [1.8] * 10, 1e9)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("12 fresh windows: load %.2f ms, StepMany (structure builds + 10 iterations) %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    del ws
