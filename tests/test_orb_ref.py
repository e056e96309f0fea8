"""Pins the restated ORB oracle to the REFERENCE'S OWN CODE: Core/MAGESLAM/Source/Image/OpenCVModified.cpp compiled unmodified
from /root/reference behind oracle/cvshim (oracle/_ref/liborb_ref.so, built by `make -C oracle ref`; the .so travels to the GPU
box, /root/reference does not). Order mode 0 of the oracle (literal libstdc++ std::nth_element, like the compiled reference)
must reproduce the reference bit for bit: key points including their order, angles, responses, octaves and descriptors --
i.e. ref OpenCVModified.cpp:144-360 (ANMS), :399-437 (ICAngles), :502-549 / :452-492 (descriptors), :571-617 (RetainBest),
:642-761 (ComputeKeyPoints), :771-886 (DetectAndCompute) and :1224-1512 (FAST, both its SSE2 and its scalar branches) are
executed here. cv::resize / GaussianBlur / fastAtan2 are not in the reference tree; the shim forwards them to the restatements
that tests/test_oracle_vs_cv2.py pins against OpenCV 4.13."""
import numpy as np
import pytest

from mageslam_b200 import synth
from tests import oracle_orb as orc
from tools.gen_orb_golden import cases

pytestmark = pytest.mark.skipif(orc.orb_ref() is None, reason="oracle/_ref/liborb_ref.so not built (needs /root/reference at build time)")


def assert_ref_equals_oracle(p, img, blur_mode=orc.BLUR_AUTO, sse=True, capacity=None, what=""):
    rk, rd = orc.detect_and_compute_ref(p, img, blur_mode, sse, capacity)
    ok, od = orc.detect_and_compute(p, img, 0, capacity, blur_mode)
    assert len(rk) == len(ok), "%s: %d reference vs %d oracle key points" % (what, len(rk), len(ok))
    assert np.array_equal(rk.view(np.uint8), ok.view(np.uint8)), "%s: key points differ" % what
    assert np.array_equal(rd, od), "%s: descriptors differ" % what
    return rk, rd


@pytest.mark.parametrize("sse", [True, False])
@pytest.mark.parametrize("name", sorted(cases()))
def test_golden_cases_equal_the_compiled_reference(name, sse):
    p, img = cases()[name]
    k, _ = assert_ref_equals_oracle(p, img, sse=sse, what=name)
    assert len(k) == p.nfeatures


@pytest.mark.parametrize("scene", ["video", "noise"])
def test_tier_config_640x480(scene):
    """BASELINE config 1/2: 640x480, 2000 features, 8 levels, scale 1.2, patch 31, FAST threshold 10"""
    p = orc.tier_params()
    img = synth.video_frames(1, 640, 480, seed=3)[0] if scene == "video" else synth.noise_frame(0)
    k, _ = assert_ref_equals_oracle(p, img, what="tier " + scene)
    assert len(k) == 2000 and set(k["octave"]) == set(range(8))
    if scene == "video":
        assert_ref_equals_oracle(p, img, sse=False, what="tier video, scalar FAST")


def test_tier_config_1280x720():
    """BASELINE config 5 frame size"""
    assert_ref_equals_oracle(orc.tier_params(), synth.video_frames(1, 1280, 720, seed=10)[0], what="720p")


def test_reference_defaults_320x180():
    """MageSettings.h:151-167: 440 features, one level, patch 15, no orientation, threshold 4 (fixed-point blur: whole-buffer view)"""
    img = synth.video_frames(1, 320, 180, seed=5)[0]
    k, _ = assert_ref_equals_oracle(orc.default_params(), img, what="defaults")
    assert np.all(k["angle"] == 0)


@pytest.mark.parametrize("blur_mode", [orc.BLUR_FLOAT_FUSED, orc.BLUR_FLOAT_UNFUSED, orc.BLUR_FIXED])
def test_blur_arithmetic_variants(blur_mode):
    img = synth.video_frames(1, 320, 240, seed=8)[0]
    assert_ref_equals_oracle(orc.tier_params(nfeatures=500, nlevels=4), img, blur_mode, what="blur mode %d" % blur_mode)
    assert_ref_equals_oracle(orc.default_params(), synth.video_frames(1, 320, 180, seed=9)[0], blur_mode, what="defaults, blur mode %d" % blur_mode)


@pytest.mark.parametrize("patch", [9, 15, 19, 25, 36])
@pytest.mark.parametrize("orient", [0, 1])
def test_patch_sizes_and_orientation(patch, orient):
    """patch 15: the second pre-rotated table; other sizes: MakeRandomPattern + run-time rotation (ref :452-492, :551-560)"""
    p = orc.tier_params(nfeatures=400, nlevels=3)
    p.patch_size = patch
    p.use_orientation = orient
    assert_ref_equals_oracle(p, synth.video_frames(1, 400, 300, seed=30 + patch)[0], what="patch %d orient %d" % (patch, orient))


def test_selection_corner_cases():
    img = synth.noise_frame(4, 320, 240)
    # RetainBestFeatures keeps FEWER than n_l (feature_strength > 1 lifts the cut): ANMS returns early (ref :181-184)
    p = orc.tier_params(nfeatures=800, nlevels=2)
    p.feature_strength = 1.6
    k, _ = assert_ref_equals_oracle(p, img, what="strength 1.6")
    assert len(k) < 800
    # feature_factor < 1: maxNum below n_l
    p = orc.tier_params(nfeatures=800, nlevels=2)
    p.feature_factor = 0.5
    assert_ref_equals_oracle(p, img, what="factor 0.5")
    # fewer candidates than the budget: no selection at all, FAST's raster order survives
    p = orc.tier_params(nfeatures=3000, nlevels=3, fast_threshold=60)
    k, _ = assert_ref_equals_oracle(p, synth.video_frames(1, 320, 240, seed=2)[0], what="under budget")
    assert len(k) < 3000
    # capacity smaller than the budget: ImageData::Insert truncates (ref Image/ImageData.h:65-70)
    assert_ref_equals_oracle(orc.tier_params(nfeatures=600, nlevels=4), img, capacity=450, what="capacity 450")
    # a flat image: no key points
    k, _ = assert_ref_equals_oracle(orc.tier_params(nfeatures=100, nlevels=2), np.full((120, 160), 77, np.uint8), what="flat")
    assert len(k) == 0
    # ANMS grid shapes and robustness range
    p = orc.tier_params(nfeatures=500, nlevels=2)
    p.num_cells_x, p.num_cells_y, p.min_robust_factor, p.max_robust_factor, p.strong_response = 7, 45, 1.0, 3.5, 40
    assert_ref_equals_oracle(p, img, what="7x45 cells")


@pytest.mark.parametrize("seed", range(12))
def test_randomised_settings(seed):
    rng = np.random.default_rng(1000 + seed)
    w, h = int(rng.integers(200, 420)), int(rng.integers(160, 320))
    p = orc.OrbParams(int(rng.choice([1, 3, 5, 7, 9])), int(rng.integers(50, 900)), float(rng.choice([1.2, 1.3, 1.5, 2.0])),
                      int(rng.integers(1, 5)), int(rng.choice([31, 31, 15, 21])), int(rng.integers(4, 30)), int(rng.integers(0, 2)),
                      float(rng.uniform(1.0, 2.5)), float(rng.uniform(0.6, 1.0)), int(rng.integers(31, 60)),
                      float(rng.uniform(1.0, 1.5)), float(rng.uniform(1.5, 3.0)), int(rng.integers(4, 40)), int(rng.integers(4, 40)))
    img = synth.noise_frame(seed, w, h) if seed % 3 == 0 else synth.video_frames(1, w, h, seed=seed)[0]
    assert_ref_equals_oracle(p, img, sse=bool(seed % 2), what="random %d" % seed)


def test_strided_input_view():
    """the reference wraps caller pixels with an arbitrary row stride (MAGESlam.cpp:123)"""
    big = synth.video_frames(1, 480, 300, seed=6)[0]
    view = big[10:250, 40:360]
    p = orc.tier_params(nfeatures=400, nlevels=3)
    a = orc.detect_and_compute_ref(p, np.ascontiguousarray(view))
    R = orc.orb_ref()
    import ctypes as C
    kps = np.zeros(400, orc.KP_DTYPE); desc = np.zeros((400, 32), np.uint8); cnt = C.c_int(0)
    assert R.ref_orb_detect_and_compute(C.byref(p), C.c_void_p(view.ctypes.data), 320, 240, big.strides[0], 0, kps.ctypes.data_as(C.c_void_p),
                                        desc.ctypes.data_as(C.c_void_p), 400, C.byref(cnt)) == 0
    assert cnt.value == len(a[0]) and np.array_equal(kps[:cnt.value].view(np.uint8), a[0].view(np.uint8)) and np.array_equal(desc[:cnt.value], a[1])
