"""GPU parity tests of RadiusMatch: CUDA path (C ABI) vs the oracle on the real boost R*-tree when oracle/_ref travelled with
the repo, else the pinned restatement. Exact: same matches (query, train, distance), same order."""
import numpy as np
import pytest

from mageslam_b200 import synth
from mageslam_b200.matcher import KeypointSpatialIndex, RadiusMatch
from tests import oracle_orb as orc
from tests.test_orb_gpu import make_detector

pytestmark = pytest.mark.gpu


def checker(*a, **kw):
    return orc.radius_match_ref(*a, **kw) if orc.radius_ref() is not None else orc.radius_match(*a, **kw)


def tuples(m, q="query_idx", t="train_idx"):
    return [(int(a), int(b), float(d)) for a, b, d in zip(m[q], m[t], m["distance"])]


@pytest.fixture(scope="module")
def feats():
    p = orc.tier_params()
    det = make_detector(p)
    vid = synth.video_frames(2, 640, 480, seed=41)
    return [det.DetectAndCompute(f) for f in vid]


def test_rank_equals_oracle_enumeration_order(feats):
    (k0, d0), (k1, d1) = feats
    rank = KeypointSpatialIndex(k1).Rank()
    order = orc.rtree_order(k1.view(orc.KP_DTYPE))
    exp = np.empty(len(k1), np.int32); exp[order] = np.arange(len(k1), dtype=np.int32)
    assert np.array_equal(rank, exp)


@pytest.mark.parametrize("radius,maxh,mind", [(12.0, 30, 1), (24.0, 30, 1), (36.0, 30, 1), (24.0, 64, 3), (8.0, 256, 0)])
def test_radius_match_equals_oracle(feats, radius, maxh, mind):
    # PoseEstimator's cascade uses radii 12 / 24 / 36 px (reference PoseEstimator.cpp:502-567)
    (k0, d0), (k1, d1) = feats
    ix = KeypointSpatialIndex(k1)
    rng = np.random.default_rng(int(radius))
    qmask = (rng.random(len(k0)) < 0.8).astype(np.uint8); tmask = (rng.random(len(k1)) < 0.9).astype(np.uint8)
    qpos = (np.stack([k0["x"], k0["y"]], 1) + rng.normal(0, 2.0, (len(k0), 2))).astype(np.float32)
    ok0, ok1 = k0.view(orc.KP_DTYPE), k1.view(orc.KP_DTYPE)
    for kw_gpu, kw_orc in (((None, None, None), {}), ((None, qmask, tmask), {"qmask": qmask, "tmask": tmask}), ((qpos, None, None), {"qpos": qpos})):
        got = RadiusMatch(k0, kw_gpu[0], kw_gpu[1], d0, ix, kw_gpu[2], d1, radius, maxh, mind)
        ref = checker(ok0, d0, ok1, d1, radius, maxh, mind, **kw_orc)
        assert tuples(got) == tuples(ref, "query", "train") and len(ref) > 50


def test_single_query_overload_and_empty_inputs(feats):
    (k0, d0), (k1, d1) = feats
    ix = KeypointSpatialIndex(k1)
    ok0, ok1 = k0.view(orc.KP_DTYPE), k1.view(orc.KP_DTYPE)
    hits = 0
    for q in range(0, 200, 7):          # TrackLocalMap matches one projected map point at a time (TrackLocalMap.cpp:586)
        got = RadiusMatch(k0[q:q + 1], None, None, d0[q:q + 1], ix, None, d1, 20.0, 30, 1)
        ref = checker(ok0[q:q + 1], d0[q:q + 1], ok1, d1, 20.0, 30, 1)
        assert tuples(got) == tuples(ref, "query", "train")
        hits += len(got)
    assert hits > 5
    assert len(RadiusMatch(k0[:0], None, None, d0[:0], ix, None, d1, 20.0, 30, 1)) == 0
    empty = KeypointSpatialIndex(k1[:0])
    assert len(RadiusMatch(k0, None, None, d0, empty, None, d1[:0], 20.0, 30, 1)) == 0
    assert len(RadiusMatch(k0, None, np.zeros(len(k0), np.uint8), d0, ix, None, d1, 20.0, 30, 1)) == 0
