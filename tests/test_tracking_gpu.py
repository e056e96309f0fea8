"""GPU parity of the map-point projection step (mage_project_map_points) against the oracle, through the C ABI."""
import numpy as np
import pytest

from mageslam_b200 import synth
from mageslam_b200.tracking import (MAP_POINT_DTYPE, PROJ_GOOD_CANDIDATE, PROJ_PREDICTED, ProjectMapPoints, ProjectMapPointsDevice,
                                    ProjectPoints, make_params)
from mageslam_b200._lib import KEYPOINT_DTYPE

from tests import oracle_tracking as ot

pytestmark = pytest.mark.gpu


def scene_params(sc, angle=60.0, border=16.0):
    return make_params(sc["view"], sc["K"], sc["position"], sc["forward"], angle, border, sc["width"], sc["height"], sc["scale"], sc["levels"])


def check_against_oracle(sc, p, kps, depth, flags):
    okps, odepth, oflags = ot.project_map_points(p, sc["points"])
    # projection, depth and the IsGoodCandidate verdict are pure IEEE f32 in a fixed order: bit-exact
    assert np.array_equal(kps["x"].view(np.uint32), okps["x"].view(np.uint32))
    assert np.array_equal(kps["y"].view(np.uint32), okps["y"].view(np.uint32))
    assert np.array_equal(depth.view(np.uint32), odepth.view(np.uint32))
    assert np.array_equal(flags & PROJ_GOOD_CANDIDATE, oflags & PROJ_GOOD_CANDIDATE)
    for name in ("size", "angle", "response", "class_id"):
        assert np.array_equal(kps[name], okps[name])
    # the octave goes through log2f (CUDA's vs the host libm's, both <= 1 ulp): it may only differ when the pre-rounding
    # value sits within 2e-6 of a rounding boundary (x.5)
    bad = np.flatnonzero((kps["octave"] != okps["octave"]) | (flags != oflags))
    for i in bad:
        m = sc["points"][i]
        d = np.linalg.norm(m["position"].astype(np.float64) - sc["position"].astype(np.float64))
        t = ot.octave_real(d, m["dmin"], sc["scale"])
        assert abs((t - np.floor(t)) - 0.5) < 2e-6, "octave differs away from a rounding boundary at point %d" % i
    assert len(bad) <= max(1, len(kps) // 10000)
    return len(bad)


@pytest.mark.parametrize("seed,n", [(0, 4000), (7, 257), (11, 1)])
def test_project_map_points_equals_oracle(seed, n):
    sc = synth.local_map_scene(n, seed=seed)
    p = scene_params(sc)
    kps, depth, flags = ProjectMapPoints(p, sc["points"])
    check_against_oracle(sc, p, kps, depth, flags)
    if n > 1000:
        assert ((flags & PROJ_PREDICTED) != 0).sum() > 100


def test_other_settings_and_full_size():
    # BASELINE-size local map: 200k points in one call, default tracking angle and a 1280x720 frame
    sc = synth.local_map_scene(200000, seed=21, width=1280, height=720, scale=1.5, levels=4)
    p = scene_params(sc, angle=45.0, border=31.0)
    kps, depth, flags = ProjectMapPoints(p, sc["points"])
    check_against_oracle(sc, p, kps, depth, flags)


def test_device_variant_and_empty():
    import torch
    sc = synth.local_map_scene(1000, seed=3)
    p = scene_params(sc)
    kps0, depth0, flags0 = ProjectMapPoints(p, sc["points"][:0])
    assert len(kps0) == 0
    d_pts = torch.from_numpy(sc["points"].view(np.uint8).reshape(-1, 32)).cuda()
    d_kps = torch.zeros((1000, 28), dtype=torch.uint8, device="cuda")
    d_depth = torch.zeros(1000, dtype=torch.float32, device="cuda")
    d_flags = torch.zeros(1000, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    ProjectMapPointsDevice(p, d_pts, d_kps, d_depth, d_flags, 1000, s)
    s.synchronize()
    kps = d_kps.cpu().numpy().view(KEYPOINT_DTYPE).reshape(-1)
    check_against_oracle(sc, p, kps, d_depth.cpu().numpy(), d_flags.cpu().numpy())


def test_project_points_mirror():
    sc = synth.local_map_scene(500, seed=9)
    pts2d, depth = ProjectPoints(sc["points"]["position"], sc["view"], sc["K"])
    p = scene_params(sc)
    okps, odepth, _ = ot.project_map_points(p, sc["points"])
    assert np.array_equal(pts2d[:, 0], okps["x"]) and np.array_equal(pts2d[:, 1], okps["y"]) and np.array_equal(depth, odepth)


def test_feeds_radius_match():
    """The projected keypoints are valid RadiusMatch queries: end to end against the oracle's RadiusMatch."""
    from mageslam_b200.matcher import KeypointSpatialIndex, RadiusMatch
    from tests import oracle_orb as orc
    rng = np.random.default_rng(2)
    sc = synth.local_map_scene(1500, seed=13)
    p = scene_params(sc)
    kps, depth, flags = ProjectMapPoints(p, sc["points"])
    pred = ((flags & PROJ_PREDICTED) != 0).astype(np.uint8)
    # current-frame keypoints scattered around the projections, random octaves / descriptors
    nt = 1800
    tk = np.zeros(nt, KEYPOINT_DTYPE)
    src = rng.integers(0, len(kps), nt)
    tk["x"] = kps["x"][src] + rng.normal(0, 4, nt).astype(np.float32)
    tk["y"] = kps["y"][src] + rng.normal(0, 4, nt).astype(np.float32)
    tk["octave"] = np.clip(kps["octave"][src] + rng.integers(-1, 2, nt), 0, 7)
    qdesc = rng.integers(0, 256, (len(kps), 32), dtype=np.uint8)
    tdesc = qdesc[src].copy()
    flip = rng.integers(0, 256, (nt, 3))
    for j in range(3):
        tdesc[np.arange(nt), flip[:, j] // 8] ^= (1 << (flip[:, j] % 8)).astype(np.uint8)
    ix = KeypointSpatialIndex(tk)
    got = RadiusMatch(kps, None, pred, qdesc, ix, None, tdesc, 8.0, 50, 2)
    want = orc.radius_match(kps.astype(orc.KP_DTYPE), qdesc, tk.astype(orc.KP_DTYPE), tdesc, 8.0, 50, 2, qmask=pred)
    assert len(got) == len(want) and len(got) > 50
    assert np.array_equal(got["query_idx"], want["query"]) and np.array_equal(got["train_idx"], want["train"])


def test_undistort_keypoints_equals_oracle_bit_for_bit():
    """A16 on the GPU (mage_undistort_keypoints through the python mirror of OrbFeatureDetector::UndistortKeypoints)."""
    from mageslam_b200.orb import CameraCalibration, UndistortKeypoints
    for name, kps, Kd, D, Ku in ot.undistort_cases(seed=5):
        want = ot.undistort_keypoints(kps, Kd, D, Ku)
        dist = CameraCalibration(Kd[0, 0], Kd[1, 1], Kd[0, 2], Kd[1, 2], D)
        und = CameraCalibration(Ku[0, 0], Ku[1, 1], Ku[0, 2], Ku[1, 2])
        got = UndistortKeypoints(np.ascontiguousarray(kps.astype(KEYPOINT_DTYPE)).copy(), dist, und)
        assert got.tobytes() == want.astype(KEYPOINT_DTYPE).tobytes(), name
    assert len(UndistortKeypoints(np.zeros(0, KEYPOINT_DTYPE), dist, und)) == 0        # empty span: no-op (ref :39-42)


def test_process_with_distorted_calibration_undistorts_the_detected_keypoints():
    from mageslam_b200 import synth
    from mageslam_b200.orb import CameraCalibration, FeatureExtractorSettings, OrbFeatureDetector
    s = FeatureExtractorSettings.tier(500, 4, 1.2, 10)
    img = synth.video_frames(1, 320, 240, seed=2)[0]
    det = OrbFeatureDetector(s)
    dist = CameraCalibration(260, 262, 160, 120, [0.12, -0.06, 0.001, -0.002, 0.01])
    und = CameraCalibration(250, 250, 160, 120)
    k0, d0 = det.Process(img)
    k1, d1 = det.Process(img, dist, und)
    assert np.array_equal(d0, d1) and len(k0) == len(k1) > 100
    want = ot.undistort_keypoints(k0, dist.camera_matrix, dist.dist_coeffs, und.camera_matrix)
    assert np.ascontiguousarray(k1).tobytes() == want.astype(KEYPOINT_DTYPE).tobytes()
    k2, _ = det.Process(img, und, und)                  # equal calibrations: keypoints untouched (ref :97)
    assert np.ascontiguousarray(k2).tobytes() == np.ascontiguousarray(k0).tobytes()


# ---- TrackLocalMap::OptimizeCameraPose in one call (mage_optimize_camera_pose)
def _pose_problem(P, seed, pose_sigma=0.03, outlier_frac=0.0):
    from mageslam_b200 import synth
    return synth.ba_problem(K=1, P=P, obs_per_point=1, n_fixed=0, pose_sigma=pose_sigma, outlier_frac=outlier_frac, seed=seed)


@pytest.mark.parametrize("P,iters,max_err,outlier_frac,seed", [(300, 3, 25.0, 0.0, 5), (300, 4, 25.0, 0.05, 6), (60, 3, 7.25, 0.1, 7), (1, 4, 25.0, 0.0, 8), (900, 10, 2.0, 0.02, 9)])
def test_optimize_camera_pose_equals_bundlerlib_sequence(P, iters, max_err, outlier_frac, seed):
    """one call == the reference's sequence (ref TrackLocalMap.cpp:421-501: new BundlerLib(ArePointsFixed), one camera, one observation per
    map point, ONE StepBundleAdjustment, GetPose(0)): against the compiled reference within 1e-4, against this library's own handle
    path (same kernel) to the last float, outlier lists equal"""
    from mageslam_b200.bundler import BundlerLib, BundlerParameters
    from mageslam_b200.tracking import OptimizeCameraPose
    from tests.ba_checks import TOL, best_checker
    from tests.oracle_ba import rel_frobenius
    prob = _pose_problem(P, seed, outlier_frac=outlier_frac)
    hub = 2.0
    pos, rot, outl, mean = OptimizeCameraPose(prob["cam_pos"][0], prob["cam_rot"][0], prob["intrinsics"][0], prob["points"], prob["obs_uv"], prob["obs_info"], iters, max_err, hub)
    h = BundlerLib(BundlerParameters(True)).load(prob)
    mh = h.StepBundleAdjustment([hub] * iters, max_err)
    ph, rh = h.poses()
    assert np.array_equal(pos, ph[0]) and np.array_equal(rot, rh[0].reshape(9)) and list(outl) == list(h.last_outliers)
    assert (np.isnan(mean) and np.isnan(mh)) or mean == np.float32(mh)
    chk = best_checker(True).load(prob)
    mc, oc = chk.StepBundleAdjustment([hub] * iters, max_err)
    pc, rc = chk.poses()
    assert rel_frobenius(pos, pc[0]) < TOL and rel_frobenius(rot, rc[0].reshape(9)) < TOL and list(outl) == list(oc)


def test_optimize_camera_pose_edge_cases():
    from mageslam_b200.tracking import OptimizeCameraPose
    prob = _pose_problem(40, 3)
    # no map point: the pose comes back as given (through the same quaternion round trip as GetPose), no outliers, NaN mean
    pos, rot, outl, mean = OptimizeCameraPose(prob["cam_pos"][0], prob["cam_rot"][0], prob["intrinsics"][0], np.zeros((0, 3)), np.zeros((0, 2)), np.zeros(0), 3, 25.0, 2.0)
    assert np.allclose(pos, prob["cam_pos"][0]) and np.allclose(rot, prob["cam_rot"][0].reshape(9), atol=1e-6) and len(outl) == 0 and np.isnan(mean)
    # zero iterations: nothing moves and no error is ever computed, so nothing is flagged (the handle path behaves the same)
    pos, rot, outl, mean = OptimizeCameraPose(prob["cam_pos"][0], prob["cam_rot"][0], prob["intrinsics"][0], prob["points"], prob["obs_uv"], prob["obs_info"], 0, 1e-6, 2.0)
    assert np.allclose(pos, prob["cam_pos"][0]) and len(outl) == 0
    with pytest.raises(Exception):
        OptimizeCameraPose(prob["cam_pos"][0], prob["cam_rot"][0], prob["intrinsics"][0], prob["points"], prob["obs_uv"], prob["obs_info"], 65, 25.0, 2.0)
    # many calls from several threads (pooled contexts): same answer every time
    import threading
    ref = OptimizeCameraPose(prob["cam_pos"][0], prob["cam_rot"][0], prob["intrinsics"][0], prob["points"], prob["obs_uv"], prob["obs_info"], 4, 25.0, 2.0)
    bad = []
    def work():
        for _ in range(50):
            r = OptimizeCameraPose(prob["cam_pos"][0], prob["cam_rot"][0], prob["intrinsics"][0], prob["points"], prob["obs_uv"], prob["obs_info"], 4, 25.0, 2.0)
            if not (np.array_equal(r[0], ref[0]) and np.array_equal(r[1], ref[1]) and list(r[2]) == list(ref[2])):
                bad.append(1)
    ts = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not bad
