"""LM steps of the global BA problem (config 4) without the CPU reference (for ncu captures of k_ba_step_coop at full size and
the per-phase times of the cooperative kernel)."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np
from mageslam_b200 import synth, _lib
from mageslam_b200.bundler import BundlerLib
L = _lib.lib()
prob = synth.ba_problem(K=500, P=50000, obs_per_point=8, seed=2, loop=True)
gpu = BundlerLib().load(prob)
names = ["errors+chi2", "build", "schur_pts", "schur_prod", "solve", "sync", "backsub+update", "errors+scale", "assemble",
         "dense:diag", "dense:panel", "dense:update", "dense:barriers", "dense:back", "dense:subfactor", "dense:rows"]
prev = np.zeros(16, np.int64)
for s in range(3):
    t0 = time.perf_counter(); m = gpu.StepBundleAdjustment([1.8], 1e9); dt = (time.perf_counter() - t0) * 1e3
    ph = np.zeros(16, np.int64)
    L.mage_ba_debug_phase_ns(gpu._h, ph.ctypes.data_as(C.c_void_p))
    d = ph - prev; prev = ph
    print("step %d %.3f ms mean %.6f |" % (s, dt, m), " ".join("%s %d" % (n, v / 1e3) for n, v in zip(names, d)), "(us)")
