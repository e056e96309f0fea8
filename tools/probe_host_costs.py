import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from mageslam_b200 import synth
from mageslam_b200.bundler import BundlerLib, StepMany
from mageslam_b200.frontend import FrontEnd
from mageslam_b200.orb import FeatureExtractorSettings
prob = synth.ba_problem(seed=1)
b0 = BundlerLib().load(prob); b0.StepBundleAdjustment([1.8], 1e9)   # warm
t0 = time.perf_counter(); bs = [BundlerLib().load(prob) for _ in range(12)]; t1 = time.perf_counter()
StepMany(bs, [1.8], 1e9); torch.cuda.synchronize(); t2 = time.perf_counter()
StepMany(bs, [1.8] * 10, 1e9); torch.cuda.synchronize(); t3 = time.perf_counter()
b = BundlerLib().load(prob); t4 = time.perf_counter(); b.StepBundleAdjustment([1.8] * 10, 1e9); t5 = time.perf_counter()
print("load 12 windows %.1f ms; first StepMany (structure build + 1 it) %.1f ms; 10 its %.1f ms; single fresh window 10 its incl. build %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t5 - t4) * 1e3))
W, H, B = 1280, 720, 32
vid = synth.video_frames(32, W, H, seed=10)
frames = torch.from_numpy(np.concatenate([vid] * 8, 0)).pin_memory()
fe = FrontEnd(FeatureExtractorSettings.tier(), W, H, B, chunk=B)
outs = fe.alloc_outputs(pinned=True)
fe.Process(frames[:B], outs); fe.Reset()
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(0, 256, B): fe.Process(frames[i:i + B], outs)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("256 frames 1280x720 sync Process: %.1f ms -> %.0f fps" % ((t1 - t0) * 1e3, 256 / (t1 - t0)))
o2 = [outs, fe.alloc_outputs(pinned=True)]
t0 = time.perf_counter()
for k, i in enumerate(range(0, 256, B)):
    fe.Submit(frames[i:i + B], o2[k & 1])
    if k: fe.Wait()
fe.Wait(); torch.cuda.synchronize(); t1 = time.perf_counter()
print("256 frames 1280x720 pipelined: %.1f ms -> %.0f fps" % ((t1 - t0) * 1e3, 256 / (t1 - t0)))
