"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): ORB (two configs), matcher, front-end, BA (both kernels + big path)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from mageslam_b200 import synth
from mageslam_b200.orb import FeatureExtractorSettings, OrbFeatureDetector
from mageslam_b200.matcher import Match
from mageslam_b200.frontend import FrontEnd
from mageslam_b200.bundler import BundlerLib, BundlerParameters, StepMany
vid = synth.video_frames(3, 333, 241, seed=1)            # odd sizes: partial words / tiles everywhere
for s in (FeatureExtractorSettings.tier(500, 4), FeatureExtractorSettings()):
    det = OrbFeatureDetector(s)
    k0, d0 = det.Process(vid[0]); k1, d1 = det.Process(vid[1])
    print(len(k0), len(k1), len(Match(d0, d1)))
fe = FrontEnd(FeatureExtractorSettings.tier(500, 4), 333, 241, batch=3, chunk=2)
outs = fe.alloc_outputs(pinned=True)
print(fe.Process(torch.from_numpy(vid).pin_memory(), outs)[4])
prob = synth.ba_problem(K=6, P=300, obs_per_point=4, seed=3, outlier_frac=0.05)
b = BundlerLib().load(prob); print(b.StepBundleAdjustment([1.8] * 3, 7.25), len(b.last_outliers)); print(b.StepBundleAdjustment([1.8] * 2, 7.25))
bs = [BundlerLib().load(synth.ba_problem(K=5, P=100, obs_per_point=3, seed=i)) for i in range(3)]
print(StepMany(bs, [1.8] * 2, 1e9))
big = BundlerLib().load(synth.ba_problem(K=60, P=600, obs_per_point=6, seed=4, loop=True)); print(big.StepBundleAdjustment([1.8] * 2, 1e9))
from mageslam_b200.sharded import ShardedGlobalBA
shd = ShardedGlobalBA(synth.ba_problem(K=60, P=600, obs_per_point=6, seed=4, loop=True)); print('sharded stages (one rank):', [shd.StepBundleAdjustment([1.8]) for _ in range(2)], shd.trials)
from mageslam_b200.tracking import OptimizeCameraPose
_pp = synth.ba_problem(K=1, P=77, obs_per_point=1, n_fixed=0, pose_sigma=0.03, outlier_frac=0.1, seed=6)
print('optimize_camera_pose:', [OptimizeCameraPose(_pp['cam_pos'][0], _pp['cam_rot'][0], _pp['intrinsics'][0], _pp['points'], _pp['obs_uv'], _pp['obs_info'], it, 7.25, 2.0)[2].tolist() for it in (3, 4)])
po = BundlerLib(BundlerParameters(True)).load(synth.ba_problem(K=1, P=100, obs_per_point=1, n_fixed=0, seed=5)); print(po.StepBundleAdjustment([2.0] * 3, 25.0))
# paths added later in the round: TMA variant of FAST, generic BRIEF pattern, single-level fixed-point blur, undistortion, pipelined front-end,
# general (materialised) one-CTA BA path, tether edges
import os
from mageslam_b200.orb import CameraCalibration, UndistortKeypoints
os.environ["MAGE_FAST_TMA"] = "1"
det = OrbFeatureDetector(FeatureExtractorSettings.tier(400, 3)); print(len(det.Process(vid[2])[0]))
os.environ["MAGE_FAST_TMA"] = "0"
sg = FeatureExtractorSettings.tier(300, 3); sg.PatchSize = 21
kg, dg = OrbFeatureDetector(sg).Process(vid[0]); print(len(kg))
s1 = FeatureExtractorSettings.tier(300, 1)
print(len(OrbFeatureDetector(s1).Process(synth.video_frames(1, 320, 200, seed=2)[0])[0]))
print(UndistortKeypoints(np.ascontiguousarray(kg), CameraCalibration(260, 262, 160, 120, [0.1, -0.05, 0.001, -0.002, 0.01]), CameraCalibration(250, 250, 160, 120))[:2]["x"])
fp = FrontEnd(FeatureExtractorSettings.tier(300, 3), 333, 241, batch=3, chunk=2)
o2 = [fp.alloc_outputs(pinned=True), fp.alloc_outputs(pinned=True)]
hv = torch.from_numpy(vid).pin_memory()
fp.Submit(hv, o2[0]); fp.Submit(hv[:2], o2[1]); fp.Wait(); fp.Wait(); print(fp.views(o2[1])[4][:2])
gen = [BundlerLib().load(synth.ba_problem(K=13, P=150, obs_per_point=4, seed=7))]; print(StepMany(gen, [1.8] * 2, 1e9))
from tools.gen_ba_golden import CASES, build_problem
for name in sorted(CASES):
    kw, pf, hub, mx, calls = CASES[name]
    if "tethers" in kw:
        tb = BundlerLib(BundlerParameters(pf)).load(build_problem(kw)); print(name, tb.StepBundleAdjustment(hub, mx)); break
# round 2: the dense solver of the reduced camera system alone (tcgen05 update, look-ahead factorisation, backward substitution)
from tests.test_dense_gpu import solve, spd
for n in (90, 300):
    A = spd(n, n, spread=1.0); bb = np.random.default_rng(n).standard_normal(n)
    xx, _, okk, _ = solve(A, bb); print("dense", n, okk, float(np.linalg.norm(A @ xx - bb) / np.linalg.norm(bb)))
