"""CPU tests that PIN the RadiusMatch oracle (restated packed-R*-tree enumeration order + matching loops) against the REAL
boost R*-tree the reference uses (oracle/_ref/libradius_ref.so, compiled from the vendored boost headers), and against golden
vectors generated from it (tests/golden/radius_golden.npz) for boxes without the reference tree."""
import ctypes as C
import os

import numpy as np
import pytest

from mageslam_b200 import synth
from tests import oracle_orb as orc

GOLD_PATH = os.path.join(os.path.dirname(__file__), "golden", "radius_golden.npz")


def features(seed, n=2):
    p = orc.tier_params(nfeatures=1200, nlevels=6)
    vid = synth.video_frames(n, 480, 360, seed=seed)
    return [orc.detect_and_compute(p, f, 1) for f in vid]


CASES = [(12.0, 30, 1), (24.0, 40, 2), (36.0, 64, 0), (8.0, 256, 5)]


@pytest.mark.skipif(orc.radius_ref() is None, reason="oracle/_ref/libradius_ref.so not built (needs /root/reference)")
def test_enumeration_order_equals_real_boost_rtree():
    (k0, d0), (k1, d1) = features(31)
    R = orc.radius_ref()
    order = orc.rtree_order(k1)
    assert sorted(order) == list(range(len(k1)))
    idx = R.rmref_index_create(k1.ctypes.data_as(C.c_void_p), len(k1))
    rank = np.empty(len(k1), np.int64); rank[order] = np.arange(len(k1))
    out = np.zeros(len(k1), np.int32)
    for q in range(0, len(k0), 5):
        for radius in (6.0, 30.0, 1000.0):
            m = R.rmref_query(idx, float(k0[q]["x"]), float(k0[q]["y"]), int(k0[q]["octave"]), radius, out.ctypes.data_as(C.c_void_p), len(out))
            ref = out[:m]
            assert np.all(np.diff(rank[ref]) > 0), "real R-tree results must be a subsequence of the restated enumeration order"
            x, y = np.float32(k0[q]["x"]), np.float32(k0[q]["y"]); r = np.float32(radius)
            box = (k1["octave"] == k0[q]["octave"]) & (x - r <= k1["x"]) & (k1["x"] <= x + r) & (y - r <= k1["y"]) & (k1["y"] <= y + r)
            assert set(ref.tolist()) == set(np.nonzero(box)[0].tolist())
    R.rmref_index_destroy(idx)


@pytest.mark.skipif(orc.radius_ref() is None, reason="oracle/_ref/libradius_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("radius,maxh,mind", CASES)
def test_port_equals_real_rtree_radius_match(radius, maxh, mind):
    (k0, d0), (k1, d1) = features(32)
    rng = np.random.default_rng(1)
    qmask = (rng.random(len(k0)) < 0.8).astype(np.uint8); tmask = (rng.random(len(k1)) < 0.9).astype(np.uint8)
    qpos = np.stack([k0["x"], k0["y"]], 1) + rng.normal(0, 2.0, (len(k0), 2)).astype(np.float32)
    for kw in ({}, {"qmask": qmask, "tmask": tmask}, {"qpos": qpos.astype(np.float32)}):
        a = orc.radius_match_ref(k0, d0, k1, d1, radius, maxh, mind, **kw)
        b = orc.radius_match(k0, d0, k1, d1, radius, maxh, mind, **kw)
        assert a.tobytes() == b.tobytes() and len(a) > 0


def test_port_reproduces_golden_vectors():
    gold = np.load(GOLD_PATH)
    (k0, d0), (k1, d1) = features(33)
    for i, (radius, maxh, mind) in enumerate(CASES):
        m = orc.radius_match(k0, d0, k1, d1, radius, maxh, mind)
        assert np.array_equal(m.view(np.uint8).reshape(len(m), 12), gold["case%d" % i])
    assert np.array_equal(orc.rtree_order(k1), gold["order"])


def test_order_dependence_is_real():
    """The acceptance of a query can depend on the enumeration order (a close competitor enumerated AFTER the best is never seen
    as 'second best'): check the restatement implements that literal behaviour, not the symmetric best/second-best rule."""
    (k0, d0), (k1, d1) = features(34)
    lit = orc.radius_match(k0, d0, k1, d1, 36.0, 64, 3)
    # symmetric rule: accept iff (true second-smallest - smallest) > minDiff
    sym = 0
    for q in range(len(k0)):
        box = (k1["octave"] == k0[q]["octave"]) & (np.abs(k1["x"] - k0[q]["x"]) <= 36) & (np.abs(k1["y"] - k0[q]["y"]) <= 36)
        ds = sorted(int(np.unpackbits(d0[q] ^ d1[t]).sum()) for t in np.nonzero(box)[0])
        ds = [d for d in ds if d <= 64]
        if ds and ((ds[1] if len(ds) > 1 else 65) - ds[0]) > 3:
            sym += 1
    assert len(lit) <= sym + 200 and len(lit) > 0      # loose sanity: same order of magnitude; they are NOT required to be equal
