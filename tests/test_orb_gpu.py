"""GPU parity tests of the ORB front-end: CUDA path (through the C ABI) vs the CPU oracle, bit-exact.

Integer/byte work => the bar is exact equality of every intermediate (pyramid levels, blurred levels, FAST candidates)
and of the final keypoints (all 7 fields, bit patterns of the floats) and descriptors.
"""
import numpy as np
import pytest

from mageslam_b200 import synth
from mageslam_b200.orb import FeatureExtractorSettings, OrbDetector, OrbFeatureDetector
from tests import oracle_orb as orc

pytestmark = pytest.mark.gpu


def make_detector(p, max_batch=1):
    return OrbDetector(p.gaussian_kernel_size, p.nfeatures, p.scale_factor, p.nlevels, p.patch_size, p.fast_threshold,
                       p.use_orientation, p.feature_factor, p.feature_strength, p.strong_response, p.min_robust_factor,
                       p.max_robust_factor, p.num_cells_x, p.num_cells_y, max_batch=max_batch)


def assert_same_features(gk, gd, ok, od, tag=""):
    assert len(gk) == len(ok), "%s count %d vs oracle %d" % (tag, len(gk), len(ok))
    for name in ("octave", "class_id"):
        assert np.array_equal(gk[name], ok[name]), (tag, name)
    for name in ("x", "y", "size", "angle", "response"):
        a = np.ascontiguousarray(gk[name]).view(np.uint32); b = np.ascontiguousarray(ok[name]).view(np.uint32)
        bad = np.nonzero(a != b)[0]
        assert len(bad) == 0, "%s field %s differs at %s: gpu %s oracle %s" % (tag, name, bad[:5], gk[name][bad[:5]], ok[name][bad[:5]])
    bad = np.nonzero((gd != od).any(axis=1))[0]
    assert len(bad) == 0, "%s descriptors differ at %s" % (tag, bad[:8])


@pytest.fixture(scope="module")
def frames():
    return {"noise": synth.noise_frame(0), "video": synth.video_frames(4, 640, 480, seed=0)}


def test_pyramid_and_blur_levels_bit_exact(frames):
    p = orc.tier_params()
    det = make_detector(p)
    img = frames["video"][0]
    det.DetectAndCompute(img)
    ref = orc.build_pyramid(p, img)
    for l in range(p.nlevels):
        got = det.DebugLevel(0, l, blurred=False)
        assert np.array_equal(got, ref[l]), "pyramid level %d: %d px differ" % (l, int((got != ref[l]).sum()))
        gotb = det.DebugLevel(0, l, blurred=True)
        refb = orc.level_blur(ref[l], 7, p.nlevels, img.shape[1])       # 8 levels: every ROI is a submatrix => cv::GaussianBlur's float path
        assert np.array_equal(gotb, refb), "blurred level %d: %d px differ" % (l, int((gotb != refb).sum()))


@pytest.mark.parametrize("ksize,nlevels,width", [(7, 1, 320), (7, 1, 330), (5, 3, 320), (9, 2, 200), (3, 4, 333), (15, 2, 256)])
def test_blur_paths_bit_exact(ksize, nlevels, width):
    """fixed-point path only when a single level fills the packed buffer (width % 16 == 0); float path otherwise; every kernel size"""
    p = orc.tier_params(nfeatures=300, nlevels=nlevels)
    p.gaussian_kernel_size = ksize
    det = make_detector(p)
    img = synth.video_frames(1, width, 200, seed=ksize)[0]
    det.DetectAndCompute(img)
    ref = orc.build_pyramid(p, img)
    for l in range(nlevels):
        gotb = det.DebugLevel(0, l, blurred=True)
        refb = orc.level_blur(ref[l], ksize, nlevels, width)
        assert np.array_equal(gotb, refb), "ksize %d level %d: %d px differ" % (ksize, l, int((gotb != refb).sum()))


@pytest.mark.parametrize("which,thr", [("noise", 10), ("video", 10), ("video", 4), ("noise", 35)])
def test_fast_candidates_bit_exact(frames, which, thr):
    p = orc.tier_params(fast_threshold=thr)
    det = make_detector(p)
    img = frames[which] if which == "noise" else frames[which][1]
    det.DetectAndCompute(img)
    ref_levels = orc.build_pyramid(p, img)
    border = 22
    for l in range(p.nlevels):
        lvl = ref_levels[l]
        h, w = lvl.shape
        k = orc.fast9(lvl, thr)
        keep = (k["x"] >= border) & (k["x"] < w - border) & (k["y"] >= border) & (k["y"] < h - border)
        k = k[keep]
        ref = (k["response"].astype(np.uint32) << 24) | (k["y"].astype(np.uint32) * w + k["x"].astype(np.uint32))
        got = det.DebugCandidates(0, l)
        assert len(got) == len(ref), "level %d: %d candidates vs oracle %d" % (l, len(got), len(ref))
        assert np.array_equal(got, ref), "level %d candidates differ" % l


@pytest.mark.parametrize("cfg", ["tier", "tier_thr20", "default", "three_level", "no_orient_31", "small_budget", "k5"])
def test_detect_and_compute_matches_oracle(frames, cfg):
    if cfg == "tier":
        p, imgs = orc.tier_params(), [frames["noise"], frames["video"][0], frames["video"][3]]
    elif cfg == "tier_thr20":
        p, imgs = orc.tier_params(fast_threshold=19), [frames["noise"], frames["video"][2]]
    elif cfg == "default":          # reference defaults at the reference's tracking resolution
        p, imgs = orc.default_params(), [synth.video_frames(2, 320, 180, seed=3)[1], synth.noise_frame(4, 320, 180)]
    elif cfg == "three_level":
        p, imgs = orc.tier_params(nfeatures=1000, nlevels=3, scale_factor=1.5), [frames["video"][1]]
    elif cfg == "no_orient_31":
        p = orc.tier_params(nfeatures=800, nlevels=4); p.use_orientation = 0
        imgs = [frames["video"][1]]
    elif cfg == "small_budget":     # far more candidates than budget: heavy retain-best + ANMS
        p, imgs = orc.tier_params(nfeatures=120, nlevels=2, fast_threshold=5), [frames["noise"]]
    else:
        p = orc.tier_params(nfeatures=1500, nlevels=5); p.gaussian_kernel_size = 5
        imgs = [frames["video"][2]]
    det = make_detector(p)
    for i, img in enumerate(imgs):
        gk, gd = det.DetectAndCompute(img)
        ok, od = orc.detect_and_compute(p, img, mode=1)
        assert len(ok) > 50
        assert_same_features(gk, gd, ok, od, "%s[%d]" % (cfg, i))


def test_few_candidates_keeps_raster_order():
    # flat image with a handful of corners: n <= n_l on every level => raster order, no suppression
    img = np.full((480, 640), 90, np.uint8)
    rng = np.random.default_rng(7)
    for _ in range(25):
        x, y = int(rng.integers(40, 600)), int(rng.integers(40, 440))
        img[y:y + 9, x:x + 9] = 200
    p = orc.tier_params()
    gk, gd = make_detector(p).DetectAndCompute(img)
    ok, od = orc.detect_and_compute(p, img, 1)
    assert 10 < len(ok) < 400
    assert_same_features(gk, gd, ok, od, "sparse")


def test_empty_image_gives_zero_features():
    p = orc.tier_params()
    gk, gd = make_detector(p).DetectAndCompute(np.full((480, 640), 128, np.uint8))
    assert len(gk) == 0 and len(gd) == 0


def test_capacity_truncates_like_image_data_insert(frames):
    p = orc.tier_params()
    det = make_detector(p)
    gk, gd = det.DetectAndCompute(frames["noise"], capacity=777)
    ok, od = orc.detect_and_compute(p, frames["noise"], 1, capacity=777)
    assert len(ok) == 777
    assert_same_features(gk, gd, ok, od, "capacity")


def test_batch_equals_single_and_is_deterministic(frames):
    p = orc.tier_params()
    vid = frames["video"]
    det1 = make_detector(p)
    detb = make_detector(p, max_batch=4)
    kps, desc, counts = detb.DetectAndComputeBatch(vid)
    kps2, desc2, counts2 = detb.DetectAndComputeBatch(vid)
    assert np.array_equal(counts, counts2) and np.array_equal(desc, desc2) and kps.tobytes() == kps2.tobytes()
    for f in range(4):
        gk, gd = det1.DetectAndCompute(vid[f])
        assert counts[f] == len(gk)
        assert kps[f, :counts[f]].tobytes() == gk.tobytes()
        assert np.array_equal(desc[f, :counts[f]], gd)


def test_device_resident_variant_equals_host_variant(frames):
    import torch
    p = orc.tier_params()
    vid = frames["video"]
    det = make_detector(p, max_batch=4)
    kps, desc, counts = det.DetectAndComputeBatch(vid)
    d_img = torch.from_numpy(vid).cuda()
    cap = 2000
    d_kps = torch.zeros((4, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((4, cap, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(4, dtype=torch.int32, device="cuda")
    det.ExtractDevice(d_img, d_kps, d_desc, d_cnt, cap, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    assert np.array_equal(d_cnt.cpu().numpy(), counts)
    for f in range(4):
        n = int(counts[f])
        assert d_kps[f, :n].cpu().numpy().tobytes() == kps[f, :n].tobytes()
        assert np.array_equal(d_desc[f, :n].cpu().numpy(), desc[f, :n])


def test_unsupported_configurations_are_rejected():
    from mageslam_b200._lib import MageError
    p = orc.tier_params(); p.patch_size = 129           # pattern coordinates are kept in 8 bits
    with pytest.raises(MageError):
        make_detector(p).DetectAndCompute(np.zeros((480, 640), np.uint8))
    with pytest.raises(TypeError):                       # CV_Assert(type == CV_8UC1)
        make_detector(orc.tier_params()).DetectAndCompute(np.zeros((480, 640), np.float32))


@pytest.mark.parametrize("patch,orient,nlevels", [(25, True, 4), (19, True, 3), (36, True, 2), (21, False, 3), (9, True, 1)])
def test_generic_pattern_patch_sizes(patch, orient, nlevels):
    """A15: patch sizes without a pre-rotated table -- cv::RNG pattern, runtime float32 rotation (ref :452-492, :551-560, :878-885)"""
    p = orc.tier_params(nfeatures=600, nlevels=nlevels)
    p.patch_size = patch; p.use_orientation = 1 if orient else 0
    det = make_detector(p)
    for img in synth.video_frames(2, 400, 300, seed=patch):
        gk, gd = det.DetectAndCompute(img)
        ok, od = orc.detect_and_compute(p, img, 1)
        assert len(ok) > 100
        assert_same_features(gk, gd, ok, od, "patch %d" % patch)


def test_tma_staged_fast_kernel_is_bit_exact(frames, monkeypatch):
    """the opt-in TMA variant of FAST (cp.async.bulk.tensor tile fetch, persistent double-buffered CTAs) against the oracle, host and device input"""
    import torch
    monkeypatch.setenv("MAGE_FAST_TMA", "1")
    p = orc.tier_params()
    det = make_detector(p, max_batch=3)
    vid = frames["video"][:3]
    kps, desc, counts = det.DetectAndComputeBatch(np.ascontiguousarray(vid))
    d_img = torch.from_numpy(np.ascontiguousarray(vid)).cuda()
    d_kps = torch.zeros((3, 2000, 28), dtype=torch.uint8, device="cuda"); d_desc = torch.zeros((3, 2000, 32), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(3, dtype=torch.int32, device="cuda")
    det.ExtractDevice(d_img, d_kps, d_desc, d_cnt, 2000, stream=torch.cuda.current_stream())
    torch.cuda.synchronize()
    for f in range(3):
        ok, od = orc.detect_and_compute(p, vid[f], 1)
        n = int(counts[f])
        assert_same_features(kps[f, :n], desc[f, :n], ok, od, "tma host frame %d" % f)
        assert d_kps[f, :n].cpu().numpy().tobytes() == kps[f, :n].tobytes() and np.array_equal(d_desc[f, :n].cpu().numpy(), desc[f, :n])
    cands = det.DebugCandidates(0, 3)
    monkeypatch.setenv("MAGE_FAST_TMA", "0")
    det2 = make_detector(p, max_batch=3)
    det2.DetectAndComputeBatch(np.ascontiguousarray(vid))
    assert np.array_equal(cands, det2.DebugCandidates(0, 3))


def test_orb_feature_detector_process(frames):
    s = FeatureExtractorSettings.tier()
    gk, gd = OrbFeatureDetector(s).Process(frames["video"][0])
    ok, od = orc.detect_and_compute(orc.tier_params(), frames["video"][0], 1)
    assert_same_features(gk, gd, ok, od, "process")


def test_full_size_properties_1280x720():
    # BASELINE config 5 geometry: no oracle-size limit here, check domain properties + oracle equality on one frame
    img = synth.video_frames(1, 1280, 720, seed=10)[0]
    p = orc.tier_params()
    gk, gd = make_detector(p).DetectAndCompute(img)
    assert len(gk) == 2000
    ws = np.array([1280 / (1.2 ** o) for o in range(8)])
    assert (gk["x"] >= 0).all() and (gk["x"] < 1280).all() and (gk["y"] >= 0).all() and (gk["y"] < 720).all()
    assert (np.diff(gk["octave"]) >= 0).all()                      # levels are appended in order
    assert np.array_equal(np.bincount(gk["octave"], minlength=8), [434, 362, 302, 251, 209, 175, 145, 122])
    assert len({(float(k["x"]), float(k["y"]), int(k["octave"])) for k in gk}) == 2000   # no duplicates
    ok, od = orc.detect_and_compute(p, img, 1)
    assert_same_features(gk, gd, ok, od, "720p")


@pytest.mark.parametrize("seed", range(24))
def test_randomised_configurations_bit_exact(seed, monkeypatch):
    """seeded sweep over image sizes, pyramid shapes, thresholds, patch sizes (pre-rotated and generic), blur sizes, orientation on/off,
    FAST kernels (register-staged / TMA): every configuration must reproduce the oracle's keypoints and descriptors exactly"""
    rng = np.random.default_rng(1000 + seed)
    w = int(rng.integers(120, 700)); h = int(rng.integers(100, 520))
    nlevels = int(rng.integers(1, 7))
    scale = float(np.float32(rng.choice([1.2, 1.25, 1.5, 2.0, 1.1])))
    patch = int(rng.choice([31, 31, 15, 15, 9, 21, 25]))
    p = orc.tier_params(nfeatures=int(rng.integers(50, 1200)), nlevels=nlevels, scale_factor=scale, fast_threshold=int(rng.integers(4, 40)))
    p.patch_size = patch
    p.use_orientation = int(rng.integers(0, 2))
    p.gaussian_kernel_size = int(rng.choice([7, 7, 7, 5, 9, 3, 0]))
    p.strong_response = max(int(p.fast_threshold) + 1 + int(rng.integers(0, 30)), int(p.strong_response))
    p.num_cells_x = int(rng.choice([8, 16, 32])); p.num_cells_y = int(rng.choice([8, 16, 32]))
    # every level must keep a usable interior (the reference asserts on degenerate pyramids)
    border = int(np.ceil(patch // 2 * np.sqrt(2.0))) if p.use_orientation else patch // 2
    smallest = min(w, h) / scale ** (nlevels - 1)
    if smallest < 2 * border + 24:
        pytest.skip("smallest level would be inside the border")
    monkeypatch.setenv("MAGE_FAST_TMA", str(seed & 1))
    kind = seed % 3
    img = synth.noise_frame(seed, w, h) if kind == 0 else synth.video_frames(1, w, h, seed=seed)[0] if kind == 1 else \
        np.clip(synth.video_frames(1, w, h, seed=seed)[0].astype(np.int32) // 3 + rng.integers(0, 40, (h, w)), 0, 255).astype(np.uint8)
    det = make_detector(p)
    gk, gd = det.DetectAndCompute(img)
    ok, od = orc.detect_and_compute(p, img, 1)
    assert_same_features(gk, gd, ok, od, "seed %d (%dx%d L%d s%.2f patch %d orient %d blur %d thr %d)" % (
        seed, w, h, nlevels, scale, patch, p.use_orientation, p.gaussian_kernel_size, p.fast_threshold))


# ---------------------------------------------------------------------------------------------------------------------------------
# Straight against the REFERENCE'S OWN CODE (OpenCVModified.cpp compiled unmodified, oracle/_ref/liborb_ref.so, tests/test_orb_ref.py).
# The compiled reference orders a level by libstdc++'s std::nth_element, the CUDA path by the canonical conforming order, so the two
# are compared as per-octave sets of (key point, descriptor) records: bit-identical records, same multiplicity.
def _records(k, d):
    rec = np.concatenate([np.ascontiguousarray(k).view(np.uint8).reshape(len(k), 28), d], axis=1)
    return rec[np.lexsort(rec.T[::-1])]


needs_ref = pytest.mark.skipif(orc.orb_ref() is None, reason="oracle/_ref/liborb_ref.so not built")


@needs_ref
@pytest.mark.parametrize("cfg", ["tier_video", "tier_noise", "defaults", "patch15_orient", "generic21", "720p"])
def test_gpu_equals_compiled_reference(frames, cfg):
    if cfg == "tier_video":
        p, img = orc.tier_params(), frames["video"][2]
    elif cfg == "tier_noise":
        p, img = orc.tier_params(), frames["noise"]
    elif cfg == "defaults":
        p, img = orc.default_params(), synth.video_frames(1, 320, 180, seed=5)[0]
    elif cfg == "patch15_orient":
        p = orc.tier_params(nfeatures=500, nlevels=1); p.patch_size = 15
        img = synth.video_frames(1, 400, 300, seed=12)[0]
    elif cfg == "generic21":
        p = orc.tier_params(nfeatures=500, nlevels=3); p.patch_size = 21
        img = synth.video_frames(1, 400, 300, seed=13)[0]
    else:
        p, img = orc.tier_params(), synth.video_frames(1, 1280, 720, seed=10)[0]
    gk, gd = make_detector(p).DetectAndCompute(img)
    rk, rd = orc.detect_and_compute_ref(p, img)
    assert len(gk) == len(rk) > 100
    gs = {bytes(r) for r in _records(gk, gd)}; rs = {bytes(r) for r in _records(rk, rd)}
    assert len(gs) == len(gk) and len(rs) == len(rk)
    only_g, only_r = gs - rs, rs - gs
    # a (suppression radius, strength) tie straddling a level's cut may pick a different member of the tie (the
    # compiled reference takes whatever libstdc++'s nth_element leaves, the CUDA path the lowest raster index): a few records per
    # level at most, and the swapped records carry the same (octave, response) multiset
    assert len(only_g) == len(only_r) <= 0.02 * len(gk), "%s: %d of %d records differ from the compiled reference" % (cfg, len(only_g), len(gk))
    if only_g:
        key = lambda x: sorted((bytes(r[16:24])) for r in (np.frombuffer(b, np.uint8) for b in x))      # response (16:20) + octave (20:24)
        assert key(only_g) == key(only_r)


@needs_ref
@pytest.mark.parametrize("mode", [orc.BLUR_FLOAT_FUSED, orc.BLUR_FLOAT_UNFUSED, orc.BLUR_FIXED])
def test_blur_mode_selection(mode):
    """mage_orb_set_blur_mode: every arithmetic variant of the reference's GaussianBlur call, level pixels and features bit-exact"""
    p = orc.tier_params(nfeatures=600, nlevels=4)
    img = synth.video_frames(1, 416, 300, seed=17)[0]
    det = make_detector(p)
    det.SetBlurMode(mode)
    gk, gd = det.DetectAndCompute(img)
    ok, od = orc.detect_and_compute(p, img, 1, blur_mode=mode)
    assert_same_features(gk, gd, ok, od, "blur mode %d" % mode)
    ref = orc.build_pyramid(p, img)
    for l in range(p.nlevels):
        want = orc.blur(ref[l], 7) if mode == orc.BLUR_FIXED else orc.blur_submatrix(ref[l], 7, fused=(mode == orc.BLUR_FLOAT_FUSED))
        assert np.array_equal(det.DebugLevel(0, l, blurred=True), want), "level %d" % l
    rk, rd = orc.detect_and_compute_ref(p, img, mode)
    assert np.array_equal(_records(gk, gd), _records(rk, rd))
    # and back to the default on the same handle (captured launch graphs are dropped)
    det.SetBlurMode(orc.BLUR_AUTO)
    gk, gd = det.DetectAndCompute(img)
    ok, od = orc.detect_and_compute(p, img, 1)
    assert_same_features(gk, gd, ok, od, "blur mode auto again")
    # other kernel sizes through the generic float kernel, unfused
    p5 = orc.tier_params(nfeatures=300, nlevels=2); p5.gaussian_kernel_size = 5
    d5 = make_detector(p5); d5.SetBlurMode(mode)
    gk, gd = d5.DetectAndCompute(img)
    ok, od = orc.detect_and_compute(p5, img, 1, blur_mode=mode)
    assert_same_features(gk, gd, ok, od, "k5 blur mode %d" % mode)


def test_retain_best_keeps_fewer_than_the_budget():
    """feature_strength > 1 or feature_factor < 1: RetainBestFeatures leaves K < n_l key points and ANMS returns early
    (ref OpenCVModified.cpp:181-184); the K survivors are emitted, nothing else (round-1 advisor finding)"""
    img = synth.noise_frame(4, 320, 240)
    for strength, factor in ((1.6, 1.5), (0.9, 0.5), (1.3, 0.8)):
        p = orc.tier_params(nfeatures=800, nlevels=2)
        p.feature_strength, p.feature_factor = strength, factor
        gk, gd = make_detector(p).DetectAndCompute(img)
        ok, od = orc.detect_and_compute(p, img, 1)
        assert 0 < len(ok) < 800
        assert_same_features(gk, gd, ok, od, "strength %.1f factor %.1f" % (strength, factor))


# ---------------------------------------------------------------------------------------------------------------------------------
# k_fast internals (round 2): border-inset tile grid, sparse / dense scoring modes, the half-precision share of the min/max network.
@pytest.mark.parametrize("thr", [10, 25])
def test_fast_sparse_mode_on_camera_like_frames(thr):
    """a frame with camera statistics (about 2 % corners): most pairs fail the high-speed test, so the warps run in sparse mode
    (queued survivors); candidates per level and the final features equal the oracle"""
    img = synth.natural_frames(2, 640, 480, seed=3)[1]
    p = orc.tier_params(fast_threshold=thr)
    det = make_detector(p)
    gk, gd = det.DetectAndCompute(img)
    ok, od = orc.detect_and_compute(p, img, 1)
    assert len(ok) > 200
    assert_same_features(gk, gd, ok, od, "natural thr %d" % thr)
    ref_levels = orc.build_pyramid(p, img)
    for l in range(p.nlevels):
        h, w = ref_levels[l].shape
        k = orc.fast9(ref_levels[l], thr)
        k = k[(k["x"] >= 22) & (k["x"] < w - 22) & (k["y"] >= 22) & (k["y"] < h - 22)]
        ref = (k["response"].astype(np.uint32) << 24) | (k["y"].astype(np.uint32) * w + k["x"].astype(np.uint32))
        assert np.array_equal(det.DebugCandidates(0, l), ref), "level %d candidates differ" % l


def test_fast_mixed_density_frame():
    """left half corner-dense chart, right half flat with a few shapes: warps switch between the two modes inside a tile row"""
    a = synth.video_frames(1, 640, 480, seed=4)[0]
    b = synth.natural_frames(1, 640, 480, seed=5)[0]
    img = np.where(np.arange(640)[None, :] < 300, a, b).astype(np.uint8)
    img[200:280, :] = 128                                   # a flat band: no pair passes the test
    p = orc.tier_params()
    gk, gd = make_detector(p).DetectAndCompute(img)
    ok, od = orc.detect_and_compute(p, img, 1)
    assert_same_features(gk, gd, ok, od, "mixed density")


@pytest.mark.parametrize("patch,orient", [(2, False), (4, False), (6, True), (8, False), (9, False)])
def test_fast_small_borders(patch, orient):
    """borders of 1 .. 4 pixels: the scored rectangle reaches the columns / rows FAST is not defined on (x < 3, x > w - 4) and the
    first tile starts left of the image"""
    img = synth.video_frames(1, 200, 150, seed=21)[0]
    p = orc.tier_params(nfeatures=400, nlevels=2)
    p.patch_size, p.use_orientation = patch, orient
    gk, gd = make_detector(p).DetectAndCompute(img)
    ok, od = orc.detect_and_compute(p, img, 1)
    assert len(ok) > 50
    assert_same_features(gk, gd, ok, od, "patch %d" % patch)


def test_fast_levels_narrower_than_the_border():
    """upper pyramid levels not wider than twice the border hold no key point (ref :706-707) and get no tile"""
    img = synth.video_frames(1, 96, 80, seed=22)[0]
    p = orc.tier_params(nfeatures=300, nlevels=6, scale_factor=1.3)
    gk, gd = make_detector(p).DetectAndCompute(img)
    ok, od = orc.detect_and_compute(p, img, 1)
    assert len(ok) > 0
    assert_same_features(gk, gd, ok, od, "narrow levels")


@pytest.mark.parametrize("tma", [0, 1])
def test_fast_every_network_split_is_identical(frames, monkeypatch, tma):
    """MAGE_FAST_VARIANT moves part of the min/max network from packed-integer (ALU pipe) to half2 (FMA pipe) instructions: same bits"""
    monkeypatch.setenv("MAGE_FAST_TMA", str(tma))
    p = orc.tier_params()
    imgs = np.ascontiguousarray(np.stack([frames["video"][1], synth.natural_frames(1, 640, 480, seed=9)[0]]))
    want = None
    for v in range(8):
        monkeypatch.setenv("MAGE_FAST_VARIANT", str(v))
        det = make_detector(p, max_batch=2)
        kps, desc, counts = det.DetectAndComputeBatch(imgs)
        got = (kps.tobytes(), desc.tobytes(), counts.tobytes(), [det.DebugCandidates(f, l).tobytes() for f in range(2) for l in range(p.nlevels)])
        if want is None:
            want = got
            ok, od = orc.detect_and_compute(p, imgs[1], 1)
            assert_same_features(kps[1, :int(counts[1])], desc[1, :int(counts[1])], ok, od, "variant 0")
        assert got == want, "variant %d differs" % v
