import sys
sys.path.insert(0, ".")
import numpy as np
from tests.test_dense_gpu import solve, spd
for n in (1324, 2604, 2860, 2944, 2946, 2976, 2988, 3000):
    A = spd(n, 7, spread=1.5); b = np.random.default_rng(8).standard_normal(n)
    xr = np.linalg.solve(A, b)
    errs = []
    for rep in range(3):
        x, _, ok, ns = solve(A, b, want_factor=False)
        d = np.abs(x - xr) / np.linalg.norm(xr)
        bad = np.nonzero(d > 1e-9)[0]
        errs.append((float(np.linalg.norm(x - xr) / np.linalg.norm(xr)), len(bad), int(bad.max()) if len(bad) else -1))
    print(n, errs)
