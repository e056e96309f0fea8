"""Reproduces the one mismatch of the extended parity sweep (tools/stress_parity.py, BA seed 2137): lambda after the third call."""
import sys
sys.path.insert(0, ".")
import numpy as np
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib, StepMany
from tests.oracle_ba import BaOracle, rel_frobenius
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2137
rng = np.random.default_rng(500 + seed)
K = int(rng.integers(3, 14)); d = int(rng.integers(2, min(K, 7) + 1))
kw = dict(K=K, P=int(rng.integers(60, 1200)), obs_per_point=d, seed=600 + seed, n_fixed=int(rng.integers(1, 3)),
          outlier_frac=float(rng.choice([0.0, 0.0, 0.05])), info_mode=str(rng.choice(["one", "confidence"])))
prob = synth.ba_problem(**kw)
hub = [float(x) for x in np.linspace(2.5, 1.0, int(rng.integers(1, 6)))]
mx = 7.25 if kw["outlier_frac"] > 0 else 1e9
print(kw, hub, mx)
g = BundlerLib().load(prob); r = BaOracle("ref").load(prob); q = BaOracle("port").load(prob)
for c in range(3):
    mg = g.StepBundleAdjustment(hub, mx); mr, outr = r.StepBundleAdjustment(hub, mx); mq, outq = q.StepBundleAdjustment(hub, mx)
    pg, rg = g.poses(); pr, rr = r.poses(); pq, rq = q.poses()
    print("call %d: gpu mean %.9g lambda %.9g outliers %d stats %s | ref mean %.9g lambda %.9g outliers %d | port mean %.9g lambda %.9g outliers %d | relF gpu-ref pos %.2e pts %.2e, port-ref pos %.2e pts %.2e" % (
        c, mg, g.GetCurrentLambda(), len(g.last_outliers), g.stats(), mr, r.GetCurrentLambda(), len(outr), mq, q.GetCurrentLambda(), len(outq),
        rel_frobenius(pg, pr), rel_frobenius(g.points(), r.points()), rel_frobenius(pq, pr), rel_frobenius(q.points(), r.points())))
