"""CPU tests of the map-point projection oracle (oracle/tracking_oracle.cpp): against the reference's own MappingMath.h
(oracle/_ref/libtracking_ref.so, when built) and against an independent float32 numpy restatement."""
import ctypes as C

import numpy as np
import pytest

from mageslam_b200 import synth
from mageslam_b200.tracking import PROJ_GOOD_CANDIDATE, PROJ_PREDICTED, make_params

from tests import oracle_tracking as ot


def scene_params(sc, angle=60.0, border=16.0):
    return make_params(sc["view"], sc["K"], sc["position"], sc["forward"], angle, border, sc["width"], sc["height"], sc["scale"], sc["levels"])


def test_compute_octave_equals_the_reference_header():
    R = ot.ref()
    if R is None:
        pytest.skip("oracle/_ref/libtracking_ref.so not built (needs /root/reference)")
    rng = np.random.default_rng(3)
    d = rng.uniform(0.05, 40.0, 20000).astype(np.float32)
    dmin = rng.uniform(0.05, 10.0, 20000).astype(np.float32)
    for sf in (1.2, 1.5, 2.0):
        a = [ot.lib().trk_compute_octave(float(x), float(y), sf) for x, y in zip(d, dmin)]
        b = [R.trkref_compute_octave(float(x), float(y), sf) for x, y in zip(d, dmin)]
        assert a == b
    # the scene generator's scale-invariance ranges follow ComputeDMin / ComputeDMax of the same header
    assert abs(R.trkref_compute_dmin(3.0, 2, 1.2) - 3.0 * 1.2 ** -2.5) < 1e-5
    assert abs(R.trkref_compute_dmax(3.0, 2, 8, 1.2) - 3.0 * 1.2 ** 5.5) < 1e-4


def numpy_projection(p, pts):
    """Independent float32 restatement with numpy scalars (every intermediate rounded to f32)."""
    f = np.float32
    V = np.array(list(p.view), f).reshape(3, 4)
    out = []
    for m in pts:
        X, Y, Z = (f(v) for v in m["position"])
        cs = []
        for r in range(3):
            s = f(0)
            for a, b in zip(V[r], (X, Y, Z, f(1))):
                s = f(s + f(a * b))
            cs.append(s)
        depth = cs[2]
        div = depth if depth != 0 else f(1)
        u = f(f(f(cs[0] / div) * f(p.fx)) + f(p.cx))
        v = f(f(f(cs[1] / div) * f(p.fy)) + f(p.cy))
        b = f(p.image_border)
        good = not (depth < 0) and b <= u and b <= v and u < f(f(p.width) - b) and v < f(f(p.height) - b)
        dot = f(0)
        for a, c in zip(m["mean_view_dir"], list(p.frame_forward)):
            dot = f(dot + f(f(a) * f(c)))
        good = good and not (dot < f(p.min_cos_view_angle))
        dx, dy, dz = (f(f(a) - f(c)) for a, c in zip(m["position"], list(p.frame_position)))
        d2 = f(f(f(dx * dx) + f(dy * dy)) + f(dz * dz))
        good = good and not (d2 < f(m["dmin"] * m["dmin"])) and not (f(m["dmax"] * m["dmax"]) < d2)
        out.append((u, v, depth, good))
    return out


def test_oracle_equals_numpy_restatement_and_covers_every_branch():
    sc = synth.local_map_scene(3000, seed=5)
    p = scene_params(sc)
    kps, depth, flags = ot.project_map_points(p, sc["points"])
    want = numpy_projection(p, sc["points"])
    assert np.array_equal(kps["x"], np.array([w[0] for w in want], np.float32))
    assert np.array_equal(kps["y"], np.array([w[1] for w in want], np.float32))
    assert np.array_equal(depth, np.array([w[2] for w in want], np.float32))
    assert np.array_equal((flags & PROJ_GOOD_CANDIDATE) != 0, np.array([w[3] for w in want]))
    good = (flags & PROJ_GOOD_CANDIDATE) != 0
    pred = (flags & PROJ_PREDICTED) != 0
    assert 0.05 < good.mean() < 0.9 and (depth < 0).any() and pred.sum() > 50
    assert not (pred & ~good).any()
    assert (kps["size"] == -1).all() and (kps["class_id"] == -1).all() and (kps["angle"] == 0).all()
    assert (kps["octave"][~good] == 0).all()
    # octave of the good candidates = ComputeOctave(sqrt(d2), dmin, scale)
    for i in np.flatnonzero(good)[:500]:
        m = sc["points"][i]
        d = np.float32(0)
        for a, c in zip(m["position"], sc["position"]):
            t = np.float32(np.float32(a) - np.float32(c))
            d = np.float32(d + np.float32(t * t))
        assert kps["octave"][i] == ot.lib().trk_compute_octave(float(np.sqrt(d, dtype=np.float32)), float(m["dmin"]), sc["scale"])


def test_empty_and_degenerate_inputs():
    sc = synth.local_map_scene(8, seed=1)
    p = scene_params(sc)
    kps, depth, flags = ot.project_map_points(p, sc["points"][:0])
    assert len(kps) == 0
    pts = sc["points"][:2].copy()
    pts["position"][0] = sc["position"]                       # on the camera centre: depth 0 -> divides by 1
    kps, depth, flags = ot.project_map_points(p, pts)
    assert depth[0] == 0 or abs(depth[0]) < 1e-6
    assert np.isfinite(kps["x"][0]) and np.isfinite(kps["y"][0])


def test_undistort_keypoints_oracle_equals_cv2_bit_for_bit():
    """A16: the restated cv::undistortPoints (five fixed-point iterations in double, icdist < 0 escape, projective map by the
    new camera matrix) against stock OpenCV on the same inputs -- float outputs must be identical, only pt may change."""
    cv2 = pytest.importorskip("cv2")
    for name, kps, Kd, D, Ku in ot.undistort_cases():
        got = ot.undistort_keypoints(kps, Kd, D, Ku)
        src = np.stack([kps["x"], kps["y"]], 1).reshape(-1, 1, 2).copy()
        ref = cv2.undistortPoints(src, Kd.astype(np.float64), D.astype(np.float64) if len(D) else None, None, Ku.astype(np.float64)).reshape(-1, 2)
        mine = np.stack([got["x"], got["y"]], 1)
        same = (mine.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(mine) & np.isnan(ref))
        assert same.all(), (name, int((~same).sum()))
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(got[f], kps[f])
