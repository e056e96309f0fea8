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
po = BundlerLib(BundlerParameters(True)).load(synth.ba_problem(K=1, P=100, obs_per_point=1, n_fixed=0, seed=5)); print(po.StepBundleAdjustment([2.0] * 3, 25.0))
