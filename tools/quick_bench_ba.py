"""Timing probe for local BA (not the contract bench): kernel time of the 10-iteration call on a fresh window, per coop grid size."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, ".")
from mageslam_b200 import synth, _lib
from mageslam_b200.bundler import BundlerLib, StepMany
import torch
L = _lib.lib()
L.mage_profile_get.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
def kernel_ms(slot=7):
    L.mage_profile_collect()
    t, n = C.c_double(0), C.c_longlong(0)
    L.mage_profile_get(slot, C.byref(t), C.byref(n))
    return t.value, n.value
prob = synth.ba_problem()
hub = [1.8] * 10
for G in (0, 32):
    os.environ["MAGE_BA_COOP_BLOCKS"] = str(G)
    res = []
    for rep in range(3):
        b = BundlerLib().load(prob)
        b.StepBundleAdjustment([1.8], 1e9)
        L.mage_profile_reset(); L.mage_profile_enable(1)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        m = b.StepBundleAdjustment(hub, 1e9)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        L.mage_profile_enable(0)
        kms, n = kernel_ms()
        st = b.stats()
        ph = np.zeros(16, np.int64)
        L.mage_ba_debug_phase_ns(b._h, ph.ctypes.data_as(C.c_void_p))
        res.append(((t1 - t0) * 1e3, kms))
    print("G=%2d: call %.3f ms (kernel %.3f ms) -> %.0f LM it/s end-to-end, %.0f it/s kernel-only; iterations %d trials %d mean %.4f" % (
        G, min(r[0] for r in res), min(r[1] for r in res), 10 / (min(r[0] for r in res) * 1e-3), 10 / (min(r[1] for r in res) * 1e-3), st["lm_iterations"], st["lambda_trials"], m))
    print("      phase us (errors+chi2, build, schur_pts, schur_prod, assemble+ldlt, sync, backsub+update, errors+scale):", [round(float(x) / 1e3, 1) for x in ph[:9]], '(last = assemble only)')
os.environ["MAGE_BA_COOP_BLOCKS"] = "32"
for N in (296,):
    bs = [BundlerLib().load(prob) for _ in range(N)]
    StepMany(bs, [1.8], 1e9)
    L.mage_profile_reset(); L.mage_profile_enable(1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    StepMany(bs, hub, 1e9)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    L.mage_profile_enable(0)
    kms, n = kernel_ms()
    print("batched %d problems: %.3f ms/call (kernel %.3f ms) -> %.0f LM it/s aggregate" % (N, (t1 - t0) * 1e3, kms, N * 10 / (t1 - t0)))
