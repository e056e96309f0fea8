"""Extended seeded parity sweep (beyond the seeds the test-suite runs): ORB configurations vs the oracle (bit-exact), matcher cases vs
the oracle (exact pairs), local-BA windows vs the compiled reference (1e-4 relative Frobenius), dense solves vs numpy.
usage: python tools/stress_parity.py [first seed] [count]   -> one summary line per family; exits 1 on the first mismatch"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import pytest
from _pytest.monkeypatch import MonkeyPatch
from tests import test_orb_gpu, test_ba_gpu, test_match_gpu, oracle_orb as orc
from tests.test_dense_gpu import solve, spd
from mageslam_b200.matcher import Match

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 150
only = sys.argv[3] if len(sys.argv) > 3 else "all"          # all | global
t0 = time.time()
ran = skipped = 0
for seed in range(first, first + (count if only == "all" else 0)):
    mp = MonkeyPatch()
    try:
        test_orb_gpu.test_randomised_configurations_bit_exact(seed, mp); ran += 1
    except pytest.skip.Exception:
        skipped += 1
    finally:
        mp.undo()
print("ORB: %d random configurations bit-exact vs the oracle (%d degenerate pyramids skipped), %.0f s" % (ran, skipped, time.time() - t0))
t0 = time.time()
for seed in range(first, first + (count if only == "all" else 0)):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 2300)); flips = int(rng.integers(1, 100)); maxd = int(rng.choice([0, 10, 30, 40, 64, 65, 100])); mind = int(rng.integers(0, 4))
    A = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    B = test_match_gpu.noisy_copy(rng, A, flips)[: max(1, n - int(rng.integers(0, n // 3 + 1)))]
    got, ref = Match(A, B, None, None, maxd, mind), orc.match(A, B, maxd, mind)
    assert test_match_gpu.as_tuples(got) == test_match_gpu.as_tuples(ref, "query", "train"), (seed, n, flips, maxd, mind)
print("Match: %d random cases equal the oracle's pairs, %.0f s" % (count if only == "all" else 0, time.time() - t0))
t0 = time.time()
nba = max(count // 5, 1) if only == "all" else 0
knife = 0
for seed in range(first, first + nba):
    try:
        test_ba_gpu.test_randomised_windows_match_reference(seed)
    except AssertionError as e:
        # The one way a window may differ: once LM has converged, the gain ratio rho = (chi2 - chi2_trial) / scale is a difference of two
        # numbers that agree to ~1e-9, its SIGN (accept / reject, hence lambda x 1/3 or x 2) is decided by summation order. The states must
        # still agree to the tolerance; only lambda may differ, and only where the call changed the returned mean error by < 1e-5 relative.
        if "lambda" not in str(e):
            raise
        from mageslam_b200 import synth
        from mageslam_b200.bundler import BundlerLib
        from tests.ba_checks import TOL, best_checker
        from tests.oracle_ba import rel_frobenius
        rng = np.random.default_rng(500 + seed)
        K = int(rng.integers(3, 14)); d = int(rng.integers(2, min(K, 7) + 1))
        kw = dict(K=K, P=int(rng.integers(60, 1200)), obs_per_point=d, seed=600 + seed, n_fixed=int(rng.integers(1, 3)),
                  outlier_frac=float(rng.choice([0.0, 0.0, 0.05])), info_mode=str(rng.choice(["one", "confidence"])))
        prob = synth.ba_problem(**kw)
        hub = [float(x) for x in np.linspace(2.5, 1.0, int(rng.integers(1, 6)))]
        mx = 7.25 if kw["outlier_frac"] > 0 else 1e9
        g, r = BundlerLib().load(prob), best_checker().load(prob)
        prev = None
        for c in range(3):
            mg = g.StepBundleAdjustment(hub, mx); mr, outr = r.StepBundleAdjustment(hub, mx)
            pg, rg = g.poses(); pr, rr = r.poses()
            assert max(rel_frobenius(pg, pr), rel_frobenius(rg, rr), rel_frobenius(g.points(), r.points())) < TOL and abs(mg - mr) <= 1e-4 * abs(mr), (seed, c)
            lam_g, lam_r = g.GetCurrentLambda(), r.GetCurrentLambda()
            if abs(lam_g - lam_r) > 1e-3 * abs(lam_r):
                assert prev is not None and abs(mg - prev) <= 1e-5 * abs(prev), ("lambda differs before convergence", seed, c, lam_g, lam_r)
                knife += 1
                break
            prev = mg
print("BA: %d random windows (solo + batched) within 1e-4 of the reference, call by call; lambda, outliers and mean equal in all but %d "
      "window(s) where LM had converged (mean error unchanged to 1e-5) and the sign of the gain ratio -- a difference of two sums that "
      "agree to ~1e-9 -- fell the other way: states still within 1e-4, lambda off by the accept / reject factor; %.0f s" % (nba, knife, time.time() - t0))
t0 = time.time()
ng = max(count // 25, 1) if only == "all" else count
from mageslam_b200 import synth as _synth
from mageslam_b200.bundler import BundlerLib as _BL
from tests.ba_checks import TOL as _TOL, best_checker as _chk, run_side_by_side as _side
worst_g = 0.0
ran_g = 0
for seed in range(first, first + ng):
    rng = np.random.default_rng(9000 + seed)
    K = int(rng.integers(16, 140)); d = int(rng.integers(3, 9))
    kw = dict(K=K, P=int(rng.integers(40 * K // 4, 60 * K)), obs_per_point=d, seed=seed, loop=True, outlier_frac=float(rng.choice([0.0, 0.0, 0.03])))
    try:
        prob = _synth.ba_problem(**kw)
    except RuntimeError:
        continue                                  # the generator could not place the points of this shape
    hub = [1.8] * int(rng.integers(1, 4))
    mx = 7.25 if kw["outlier_frac"] > 0 else 1e9
    rep = _side(_BL().load(prob), _chk().load(prob), hub, mx, int(rng.integers(1, 4)), tag="global %s" % kw)
    worst_g = max(worst_g, max(max(r) for r in rep))
    ran_g += 1
    assert worst_g < _TOL
print("global BA: %d random loop problems (16-140 key frames: reduced systems of 84-828 unknowns through the tcgen05 dense solver, 1-3 Huber widths per "
      "call, outliers in a third) within %.1e of the reference, lambda / outliers / mean equal, %.0f s" % (ran_g, worst_g, time.time() - t0))
t0 = time.time()
worst = 0.0
for seed in range(first, first + max(count // 5, 1)):
    rng = np.random.default_rng(seed)
    n = 2 * int(rng.integers(1, 900))
    A = spd(n, seed, spread=float(rng.uniform(0, 2.5))); b = rng.standard_normal(n)
    x, _, ok, _ = solve(A, b, want_factor=False)
    xr = np.linalg.solve(A, b)
    err = float(np.linalg.norm(x - xr) / np.linalg.norm(xr))
    assert ok == 1 and err < 1e-8, (seed, n, err)
    worst = max(worst, err)
print("dense solver: %d random SPD systems (n = 2 .. 1800), worst relative error vs numpy %.2e, %.0f s" % (max(count // 5, 1), worst, time.time() - t0))
