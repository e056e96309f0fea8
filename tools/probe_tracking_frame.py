"""Latency of the calls the reference's tracking thread makes per frame, one after the other through the host-buffer C ABI
(ref Tasks/ImageAnalyzer.cpp:119 -> OrbFeatureDetector::Process; Image/KeypointSpatialIndex.cpp; Tracking/TrackLocalMap.cpp:325-388
projection, :586 RadiusMatch, :421-501 OptimizeCameraPose twice): medians over 100 frames, host clock."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from mageslam_b200 import synth
from mageslam_b200.orb import FeatureExtractorSettings, OrbFeatureDetector
from mageslam_b200.matcher import KeypointSpatialIndex, RadiusMatch
from mageslam_b200.tracking import OptimizeCameraPose, ProjectMapPoints, make_params

vid = synth.video_frames(8, 640, 480, seed=0)
det = OrbFeatureDetector(FeatureExtractorSettings.tier())
sc = synth.local_map_scene(4000, seed=0)
params = make_params(sc["view"], sc["K"], sc["position"], sc["forward"], 60.0, 16.0, sc["width"], sc["height"], sc["scale"], sc["levels"])
pose = synth.ba_problem(K=1, P=300, obs_per_point=1, n_fixed=0, pose_sigma=0.03, outlier_frac=0.05, seed=5)
names = ["DetectAndCompute (640x480 -> 2000 key points)", "KeypointSpatialIndex (host R*-tree order + upload)", "ProjectMapPoints (4000 map points)",
         "RadiusMatch (2000 projected points vs the frame's index)", "OptimizeCameraPose x 2 (300 points, 3 + 4 iterations)"]
ts = [[] for _ in names]
kq, dq = det.Process(vid[7])
for r in range(110):
    t = [time.perf_counter()]
    k, d = det.Process(vid[r % 8]); t.append(time.perf_counter())
    ix = KeypointSpatialIndex(k); t.append(time.perf_counter())
    ProjectMapPoints(params, sc["points"]); t.append(time.perf_counter())
    m = RadiusMatch(kq, None, None, dq, ix, None, d, 24.0, 30, 1); t.append(time.perf_counter())
    for it in (3, 4):
        OptimizeCameraPose(pose["cam_pos"][0], pose["cam_rot"][0], pose["intrinsics"][0], pose["points"], pose["obs_uv"], pose["obs_info"], it, 25.0, 2.0)
    t.append(time.perf_counter())
    ix.close()
    if r >= 10:
        for i in range(len(names)):
            ts[i].append(t[i + 1] - t[i])
tot = 0.0
for n, v in zip(names, ts):
    print("%-58s %7.1f us" % (n, 1e6 * np.median(v))); tot += np.median(v)
print("%-58s %7.1f us per tracked frame (%.0f frames/s on one stream)" % ("sum", 1e6 * tot, 1.0 / tot))
