"""BASELINE config 4 at full size: global BA 500 keyframes / 50 000 points / 400 000 observations, GPU vs the reference's own
BundlerLib+g2o (oracle/_ref) for a few LM steps. Prints timings and the relative Frobenius errors (tolerance 1e-4)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib
from tests.oracle_ba import BaOracle, have_ref, rel_frobenius

K, P, D, STEPS = 500, 50000, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 3
t0 = time.time(); prob = synth.ba_problem(K=K, P=P, obs_per_point=D, seed=2, loop=True); print("generated in %.1f s, %d observations" % (time.time() - t0, len(prob["obs_uv"])))
gpu = BundlerLib().load(prob)
t0 = time.time(); chk = BaOracle("ref" if have_ref() else "port").load(prob); print("oracle (%s) loaded in %.1f s" % (chk.kind, time.time() - t0))
for s in range(STEPS):
    t0 = time.perf_counter(); mg = gpu.StepBundleAdjustment([1.8], 1e9); t1 = time.perf_counter()
    mr, _ = chk.StepBundleAdjustment([1.8], 1e9); t2 = time.perf_counter()
    pc, rc = gpu.poses(); pr, rr = chk.poses()
    print("step %d: gpu %.1f ms  cpu %.1f ms  mean %.6f / %.6f  lambda %.6g / %.6g  relF pos %.2e rot %.2e pts %.2e  stats %s" % (
        s, (t1 - t0) * 1e3, (t2 - t1) * 1e3, mg, mr, gpu.GetCurrentLambda(), chk.GetCurrentLambda(), rel_frobenius(pc, pr), rel_frobenius(rc, rr),
        rel_frobenius(gpu.points(), chk.points()), gpu.stats()))
    import ctypes as _C, numpy as _np
    from mageslam_b200 import _lib as _L
    ph = _np.zeros(16, _np.int64)
    _L.lib().mage_ba_debug_phase_ns(gpu._h, ph.ctypes.data_as(_C.c_void_p))
    print('      cumulative phase us (errors+chi2, build, schur_pts, schur_prod, solve(after assemble), sync, backsub+update, errors+scale, assemble):', [round(float(x) / 1e3, 1) for x in ph[:9]], ' LDLT (diag, panel, update, barriers):', [round(float(x) / 1e3, 1) for x in ph[9:13]])
