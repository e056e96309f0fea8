"""Early timing probe for local BA (not the contract bench)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib, StepMany
import torch
prob = synth.ba_problem()
hub = [1.8] * 10
for rep in range(3):
    b = BundlerLib().load(prob)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = b.StepBundleAdjustment(hub, 1e9)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    m2 = b.StepBundleAdjustment(hub, 1e9)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("single problem: first call (incl. structure) %.3f ms, second call %.3f ms -> %.0f LM it/s; mean %.4f stats %s" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, 10 / (t2 - t1), m2, b.stats()))
for N in (16, 148, 296):
    bs = [BundlerLib().load(prob) for _ in range(N)]
    StepMany(bs, hub, 1e9)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    StepMany(bs, hub, 1e9)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("batched %d problems: %.3f ms/call -> %.0f LM it/s aggregate" % (N, (t1 - t0) * 1e3, N * 10 / (t1 - t0)))
