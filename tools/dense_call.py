"""Solves of a 2988 x 2988 SPD system through the dense solver's test entry (for ncu captures of k_dense_debug and phase timing)."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from tests.test_dense_gpu import solve, spd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2988
A = spd(n, 7, spread=1.5); b = np.random.default_rng(8).standard_normal(n)
for _ in range(3):
    x, _, ok, ns = solve(A, b, want_factor=False)
print("ok", ok, "kernel %.0f us;" % (ns[15] / 1e3), "phases (us): diag %d panel %d update %d barriers %d back %d | sub-factor %d rows %d diag-update %d | epilogue wait %d (unused %d) store %d" % tuple(round(v / 1e3) for v in ns[:11]),
      "| MMA warp: wait operands %d us, wait accumulator release %d us, tiles %d | factor CTA: waits for its tiles %d us, factors %d us" % (ns[11] / 1e3, ns[12] / 1e3, ns[13], ns[14] / 1e3, ns[9] / 1e3))
