"""Latency of the per-frame calls a reference integration makes: OrbDetector::DetectAndCompute on one 640x480 frame and Match vs the previous frame."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from mageslam_b200 import synth
from mageslam_b200.orb import FeatureExtractorSettings, OrbFeatureDetector
from mageslam_b200.matcher import Match
vid = synth.video_frames(8, 640, 480, seed=0)
det = OrbFeatureDetector(FeatureExtractorSettings.tier())
for f in vid[:3]: det.Process(f)
ts = []
for r in range(200):
    t0 = time.perf_counter(); k, d = det.Process(vid[r % 8]); ts.append(time.perf_counter() - t0)
print("DetectAndCompute 640x480, one frame per call: median %.1f us, p10 %.1f us" % (1e6 * np.median(ts), 1e6 * np.percentile(ts, 10)))
k0, d0 = det.Process(vid[0]); k1, d1 = det.Process(vid[1])
for _ in range(3): Match(d1, d0, None, None, 30, 1)
ts = []
for r in range(200):
    t0 = time.perf_counter(); m = Match(d1, d0, None, None, 30, 1); ts.append(time.perf_counter() - t0)
print("Match 2000 x 2000, one pair per call: median %.1f us (%d matches)" % (1e6 * np.median(ts), len(m)))
