"""Matcher-only run for compute-sanitizer memcheck: the tcgen05 kernel (tile edges, masks, hit-list overflow) and the xor/popc kernel."""
import sys
import numpy as np
sys.path.insert(0, ".")
from mageslam_b200.matcher import Match
rng = np.random.default_rng(0)
for n, maxd in ((333, 30), (129, 64), (600, 100)):
    A = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    B = A[: n - 7].copy(); B[:, 0] ^= 1
    mA = (rng.random(n) > 0.1).astype(np.uint8)
    print(n, maxd, len(Match(A, B, None, None, maxd, 1)), len(Match(A, B, mA, None, maxd, 1)))
base = rng.integers(0, 256, (8, 32), dtype=np.uint8)
print(len(Match(np.repeat(base, 128, axis=0), np.tile(base, (128, 1)), None, None, 30, 0)))
