"""Two LM steps of the global BA problem (config 4) without the CPU reference (for ncu captures of k_ba_step_coop at full size)."""
import sys, time
sys.path.insert(0, ".")
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib
prob = synth.ba_problem(K=500, P=50000, obs_per_point=8, seed=2, loop=True)
gpu = BundlerLib().load(prob)
for s in range(2):
    t0 = time.perf_counter(); m = gpu.StepBundleAdjustment([1.8], 1e9); print("step", s, (time.perf_counter() - t0) * 1e3, "ms", m)
