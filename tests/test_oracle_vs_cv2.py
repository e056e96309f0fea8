"""Pins the ORB/Match oracle against stock OpenCV 4.13 primitives (the reference ships no tests for this path).

CPU only. Skipped if cv2 is not importable. The reference calls cv::resize / GaussianBlur / fastAtan2 / BFMatcher and
its FAST_t<16> shares OpenCV's FAST lineage.
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from mageslam_b200 import synth
from tests import oracle_orb as orc


def test_cv_round_half_even():
    for v, e in ((0.5, 0), (1.5, 2), (2.5, 2), (-0.5, 0), (-1.5, -2), (3.49, 3), (1e6 + 0.5, 1000000)):
        assert orc.lib().orc_cv_round_f(v) == e


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(0)
    ys = rng.integers(-200000, 200001, 20000).astype(np.float32)
    xs = rng.integers(-200000, 200001, 20000).astype(np.float32)
    ys[:8] = [0, 0, -1, 1, 1, -1, 0, 5]; xs[:8] = [0, -1, 0, 1, 0, -1, 7, 0]
    # the scalar cv::fastAtan2 is what the reference calls (OpenCVModified.cpp:435); cv2.phase's vectorised
    # variant rounds differently and is NOT the reference arithmetic
    ref = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    got = np.array([orc.fast_atan2(y, x) for y, x in zip(ys, xs)], np.float32)
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))
    assert abs(orc.fast_atan2(1, 1) - 44.990456) < 1e-5


@pytest.mark.parametrize("seed", range(4))
def test_resize_matches_cv2(seed):
    rng = np.random.default_rng(seed)
    for _ in range(6):
        sw, sh = int(rng.integers(40, 700)), int(rng.integers(40, 500))
        sc = float(rng.uniform(1.05, 2.2))
        dw, dh = max(4, int(round(sw / sc))), max(4, int(round(sh / sc)))
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ref, orc.resize(src, dw, dh))


def test_pyramid_chain_matches_cv2():
    p = orc.tier_params()
    img = synth.noise_frame(3)
    levels = orc.build_pyramid(p, img)
    sizes, scales, nfeat = orc.level_layout(p, 640, 480)
    assert int(nfeat.sum()) == 2000 and len(levels) == 8
    prev = img
    for l in range(1, 8):
        prev = cv2.resize(prev, (int(sizes[l][0]), int(sizes[l][1])), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(prev, levels[l]), l
    # sizes follow cvRound(cols / (float)pow(sf, l))
    for l in range(8):
        s = np.float32(np.float64(np.float32(1.2)) ** l)
        assert sizes[l][0] == int(np.rint(np.float32(640) / s)) and sizes[l][1] == int(np.rint(np.float32(480) / s))


@pytest.mark.parametrize("ksize", [3, 5, 7, 9, 11, 13, 15])
def test_gaussian_blur_matches_cv2(ksize):
    rng = np.random.default_rng(ksize)
    for shape in ((480, 640), (37, 53), (134, 179), (16, 9)):
        src = rng.integers(0, 256, shape, dtype=np.uint8)
        ref = cv2.GaussianBlur(src, (ksize, ksize), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, orc.blur(src, ksize)), (ksize, shape)


@pytest.mark.parametrize("ksize", [3, 5, 7, 9, 11, 13, 15])
def test_submatrix_gaussian_blur_matches_cv2_generic_path(ksize):
    """cv::GaussianBlur of a SUBMATRIX (the reference's in-place blur of a pyramid-level ROI, OpenCVModified.cpp:853-865) does not
    take the fixed-point path: it is sepFilter2D with the CV_32F Gaussian kernel. Python cannot hand cv2 a submatrix, so the
    same code path is reached through cv2.sepFilter2D; test_stock_orb_descriptors_use_the_submatrix_blur closes the loop."""
    rng = np.random.default_rng(100 + ksize)
    k = cv2.getGaussianKernel(ksize, 2, cv2.CV_32F)
    worst = 0
    for shape in [(480, 640), (400, 533), (134, 179), (37, 53), (16, 16)]:
        src = rng.integers(0, 256, shape).astype(np.uint8)
        ref = cv2.sepFilter2D(src, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101)
        fused, unfused = orc.blur_submatrix(src, ksize, True), orc.blur_submatrix(src, ksize, False)
        # this cv2 build contracts multiply-adds (AVX2 unit): the fused variant must be identical; a build without FMA gives the
        # unfused one. Accept either as a whole, never a mixture.
        assert np.array_equal(ref, fused) or np.array_equal(ref, unfused), (ksize, shape)
        d = fused.astype(int) - unfused.astype(int)
        assert np.abs(d).max(initial=0) <= 1
        worst = max(worst, int((d != 0).sum()) / src.size)
    assert worst < 2e-4            # the two evaluations differ only on rounding ties


def test_stock_orb_descriptors_use_the_submatrix_blur():
    """Stock cv::ORB blurs its level ROI with the same GaussianBlur(7x7, 2, 2, REFLECT_101) call on a submatrix as the reference.
    With provided keypoints (angle 0, learned 31-pattern) its descriptors are reproduced exactly from the oracle's submatrix blur and
    NOT from the fixed-point whole-image blur -- evidence that a submatrix source takes the float path."""
    rng = np.random.default_rng(1)
    img = cv2.GaussianBlur(rng.integers(0, 256, (240, 320)).astype(np.uint8), (0, 0), 1.5)
    orb = cv2.ORB_create(nfeatures=500, scaleFactor=1.2, nlevels=1, edgeThreshold=31, firstLevel=0, WTA_K=2, patchSize=31, fastThreshold=10)
    kps = [cv2.KeyPoint(float(x), float(y), 31.0, 0.0, 1.0, 0, -1) for x, y in zip(rng.integers(40, 280, 300), rng.integers(40, 200, 300))]
    kps, desc = orb.compute(img, kps)
    pat = orc.brief_pattern(31)[0].astype(np.int32).reshape(-1, 2)          # row 0 of the pre-rotated table = the learned pattern

    def describe(blurred):
        out = np.zeros((len(kps), 32), np.uint8)
        for n, kp in enumerate(kps):
            cy, cx = int(kp.pt[1]), int(kp.pt[0])
            a = blurred[cy + pat[0::2, 1], cx + pat[0::2, 0]].astype(np.int32)
            b = blurred[cy + pat[1::2, 1], cx + pat[1::2, 0]].astype(np.int32)
            out[n] = np.packbits((a < b).reshape(32, 8), axis=1, bitorder="little").ravel()
        return out

    assert np.array_equal(describe(orc.blur_submatrix(img, 7)), desc)
    assert not np.array_equal(describe(orc.blur(img, 7)), desc)



@pytest.mark.parametrize("thr,seed", [(10, 0), (20, 1), (4, 2), (40, 3)])
def test_fast_matches_cv2(thr, seed):
    img = synth.noise_frame(seed, 320, 240) if seed % 2 == 0 else synth.video_frames(1, 320, 240, seed)[0]
    det = cv2.FastFeatureDetector_create(thr, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    ref = det.detect(img)
    got = orc.fast9(img, thr)
    assert len(ref) == len(got) and len(got) > 50
    r = np.array([(k.pt[0], k.pt[1], k.response) for k in ref], np.float32)
    g = np.stack([got["x"], got["y"], got["response"]], 1)
    assert np.array_equal(r, g)         # same keypoints, same raster order, same score


def test_match_matches_cv2_pipeline():
    rng = np.random.default_rng(5)
    for n, flips, maxd, mind in ((300, 10, 30, 1), (500, 24, 40, 3), (64, 4, 30, 1), (200, 60, 64, 2)):
        A = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        B = A.copy()[rng.permutation(n)]
        # flip a few random bits per descriptor, and duplicate some rows to create ties
        for i in range(n):
            for b in rng.integers(0, 256, int(rng.integers(0, flips))):
                B[i, b // 8] ^= np.uint8(1 << (b % 8))
        B[: n // 20] = B[n // 20: 2 * (n // 20)]
        bf = cv2.BFMatcher(cv2.NORM_HAMMING, False)
        fwd = bf.radiusMatch(A, B, float(maxd), None, False)
        bwd = bf.radiusMatch(B, A, float(maxd), None, False)

        def best(rows):
            out = {}
            for q, row in enumerate(rows):
                if len(row) == 0:
                    continue
                row = sorted(row, key=lambda m: m.distance)
                if len(row) > 1 and row[1].distance - row[0].distance < mind:
                    continue
                out[q] = (row[0].trainIdx, row[0].distance)
            return out
        fb, bb = best(fwd), best(bwd)
        ref = [(a, t, d) for a, (t, d) in sorted(fb.items()) if bb.get(t, (-1, 0))[0] == a]
        got = orc.match(A, B, maxd, mind)
        assert [(int(m["query"]), int(m["train"]), float(m["distance"])) for m in got] == ref
        assert len(ref) > 0


def test_descriptor_distance_is_popcount():
    rng = np.random.default_rng(9)
    for _ in range(200):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert orc.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())
        assert orc.descriptor_distance(a, b) == int(cv2.norm(a, b, cv2.NORM_HAMMING))


@pytest.mark.parametrize("patch", [19, 25, 36])
def test_generic_pattern_and_rotation_match_stock_orb(patch):
    """A15: MakeRandomPattern (cv::RNG restated) + the float32 rotation / cvRound of ComputeOrbDescriptors against stock cv::ORB, which
    builds the same pattern for patchSize != 31 and samples it the same way. Ramp images (value = x, value = y) make the blur the
    identity, so every descriptor bit is an ordering of two rotated, rounded pattern coordinates: 2 x 200 random angles x 256 tests."""
    rng = np.random.default_rng(patch)
    H, W = 160, 230
    ramps = {"x": np.tile(np.arange(W, dtype=np.uint8), (H, 1)), "y": np.tile(np.arange(H, dtype=np.uint8)[:, None], (1, W))}
    orb = cv2.ORB_create(nfeatures=500, scaleFactor=1.2, nlevels=1, edgeThreshold=patch, firstLevel=0, WTA_K=2, patchSize=patch, fastThreshold=10)
    m = patch + 6
    xya = np.stack([rng.integers(m, W - m, 200), rng.integers(m, H - m, 200), rng.uniform(0, 360, 200)], 1).astype(np.float32)
    kps = [cv2.KeyPoint(float(x), float(y), float(patch), float(a), 1.0, 0, -1) for x, y, a in xya]
    for name, img in ramps.items():
        kk, desc = orb.compute(img, kps)
        assert len(kk) == len(kps)
        assert np.array_equal(orc.blur_submatrix(img, 7)[8:-8, 8:-8], img[8:-8, 8:-8])         # the blur is the identity on a ramp
        assert np.array_equal(orc.generic_descriptors(img, xya, patch), desc), (patch, name)
    pat = orc.random_pattern(patch)
    assert pat.min() == -(patch // 2) and pat.max() == patch // 2
