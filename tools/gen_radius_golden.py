#!/usr/bin/env python3
"""Generates tests/golden/radius_golden.npz from the RadiusMatch oracle built on the REAL boost R*-tree
(oracle/_ref/libradius_ref.so). Run in the build container (needs /root/reference for the vendored boost headers)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import oracle_orb as orc
from tests.test_radius_oracle import CASES, features

(k0, d0), (k1, d1) = features(33)
out = {"order": orc.rtree_order(k1)}
for i, (radius, maxh, mind) in enumerate(CASES):
    m = orc.radius_match_ref(k0, d0, k1, d1, radius, maxh, mind)
    out["case%d" % i] = m.view(np.uint8).reshape(len(m), 12)
    print(radius, maxh, mind, len(m))
np.savez_compressed("tests/golden/radius_golden.npz", **out)
