"""CPU tests of the ORB/Match oracle itself: golden vectors, order modes, selection edge cases, pattern regeneration."""
import json
import os
import zlib

import numpy as np
import pytest

from mageslam_b200 import synth
from tests import oracle_orb as orc
from tools.gen_orb_golden import cases

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "orb_golden.npz"))


@pytest.mark.parametrize("name", sorted(cases()))
def test_oracle_reproduces_golden_vectors(name):
    p, img = cases()[name]
    for mode in (0, 1):
        k, d = orc.detect_and_compute(p, img, mode)
        assert np.array_equal(k.view(np.uint8).reshape(len(k), 28), GOLD["%s/mode%d/kps" % (name, mode)])
        assert np.array_equal(d, GOLD["%s/mode%d/desc" % (name, mode)])


def test_match_golden():
    m = orc.match(GOLD["tier4_video_next/mode1/desc"], GOLD["tier4_video/mode1/desc"], 30, 1)
    assert np.array_equal(m.view(np.uint8).reshape(len(m), 12), GOLD["match/tier4"])


def test_brief_tables_regenerate_to_the_reference_crc():
    crc = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "brief_pattern_crc.json")))
    assert zlib.crc32(orc.brief_pattern(31).tobytes()) & 0xFFFFFFFF == crc["bit_pattern_31_rotated"]
    assert zlib.crc32(orc.brief_pattern(15).tobytes()) & 0xFFFFFFFF == crc["bit_pattern_15_rotated"]
    assert int(np.abs(orc.brief_pattern(31)).max()) == 18 and int(np.abs(orc.brief_pattern(15)).max()) == 10


def test_umax_and_level_budget():
    assert list(orc.umax(15)[:16]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert list(orc.umax(7)[:8]) == [7, 7, 7, 6, 6, 5, 4, 2]
    sizes, scales, nfeat = orc.level_layout(orc.tier_params(), 640, 480)
    assert [tuple(s) for s in sizes] == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
    assert list(nfeat) == [434, 362, 302, 251, 209, 175, 145, 122] and int(nfeat.sum()) == 2000


def test_order_modes_agree_as_sets_unless_a_tie_straddles_the_cut():
    """Mode A = literal libstdc++ nth_element, mode B = canonical order (what the GPU implements). The kept SET can only differ
    where several candidates share (r, strength) at the cut (hard part 1)."""
    p = orc.tier_params()
    tie_levels = total = 0
    for seed in range(4):
        img = synth.noise_frame(seed) if seed % 2 else synth.video_frames(1, 640, 480, seed)[0]
        levels = orc.build_pyramid(p, img)
        _, _, nfeat = orc.level_layout(p, 640, 480)
        for l, lvl in enumerate(levels):
            h, w = lvl.shape
            k = orc.fast9(lvl, p.fast_threshold)
            k = k[(k["x"] >= 22) & (k["x"] < w - 22) & (k["y"] >= 22) & (k["y"] < h - 22)]
            if len(k) <= nfeat[l]:
                continue
            a = orc.select_level(p, k, nfeat[l], 0); b = orc.select_level(p, k, nfeat[l], 1)
            assert len(a) == len(b) == nfeat[l]
            sa = {(float(q["x"]), float(q["y"])) for q in a}; sb = {(float(q["x"]), float(q["y"])) for q in b}
            total += 1
            if sa != sb:
                tie_levels += 1
                # the differing keypoints must all have the (r, strength) of the last kept element of mode B
                # (mode B output is sorted: its last element is the cut)
                diff = sa ^ sb
                assert len(diff) <= 2 * 40
                strengths = {float(q["response"]) for q in list(a) + list(b) if (float(q["x"]), float(q["y"])) in diff}
                assert len(strengths) == 1
            # canonical order is sorted by construction: keys never increase
    assert total >= 16


def test_selection_edge_cases():
    p = orc.tier_params(nfeatures=50, nlevels=1)
    rng = np.random.default_rng(0)
    # all candidates share one score: RetainBest keeps the whole bin, ANMS falls back to raster order on ties
    n = 300
    k = np.zeros(n, orc.KP_DTYPE)
    xs = rng.permutation(500)[:n] + 30; ys = rng.permutation(400)[:n] + 30
    order = np.lexsort((xs, ys))
    k["x"], k["y"], k["response"], k["class_id"] = xs[order], ys[order], 25.0, -1
    b = orc.select_level(p, k, 50, 1)
    assert len(b) == 50
    # with equal strengths nobody is "stronger" (strict >), so every radius is the global maximum and raster order decides
    assert np.array_equal(b["x"], k["x"][:50]) and np.array_equal(b["y"], k["y"][:50])
    # fewer candidates than budget: untouched, same order
    few = orc.select_level(p, k[:20], 50, 1)
    assert few.tobytes() == k[:20].tobytes()
    # radii are independent of the input order
    r1 = orc.anms_radii(p, k, 50)
    perm = rng.permutation(n)
    r2 = orc.anms_radii(p, k[perm], 50)
    assert np.array_equal(r1[perm], r2)
