#!/usr/bin/env python3
"""Generates tests/golden/orb_golden.npz from the ORB/Match oracle AFTER it has been pinned against cv2 4.13
(tests/test_oracle_vs_cv2.py): small seeded frames -> keypoints, descriptors, matches. Guards the oracle (and through it the
CUDA path) against regressions on boxes without cv2. The reference itself ships no fixture for this path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mageslam_b200 import synth
from tests import oracle_orb as orc

def cases():
    p1 = orc.tier_params(nfeatures=600, nlevels=4)
    p2 = orc.default_params()
    p3 = orc.tier_params(nfeatures=300, nlevels=3, scale_factor=1.5, fast_threshold=20)
    vid = synth.video_frames(2, 320, 240, seed=21)
    return {"tier4_video": (p1, vid[0]), "tier4_video_next": (p1, vid[1]), "default_320x180": (p2, synth.video_frames(1, 320, 180, seed=22)[0]),
            "three_level_noise": (p3, synth.noise_frame(23, 320, 240))}

def main():
    out = {}
    for name, (p, img) in cases().items():
        for mode in (0, 1):
            k, d = orc.detect_and_compute(p, img, mode)
            out["%s/mode%d/kps" % (name, mode)] = k.view(np.uint8).reshape(len(k), 28)
            out["%s/mode%d/desc" % (name, mode)] = d
        print(name, len(k))
    a = out["tier4_video/mode1/desc"]; b = out["tier4_video_next/mode1/desc"]
    m = orc.match(b, a, 30, 1)
    out["match/tier4"] = m.view(np.uint8).reshape(len(m), 12)
    print("matches", len(m))
    np.savez_compressed("tests/golden/orb_golden.npz", **out)

if __name__ == "__main__":
    main()
