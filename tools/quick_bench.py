"""Early timing probe (not the contract bench): device-resident ORB extract + match throughput, CUDA events."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from mageslam_b200 import synth
from mageslam_b200.orb import FeatureExtractorSettings, OrbFeatureDetector
from mageslam_b200.matcher import Matcher

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
vid = synth.video_frames(B, 640, 480, seed=0)
det = OrbFeatureDetector(FeatureExtractorSettings.tier(), max_batch=B).m_detector
cap = 2000
d_img = torch.from_numpy(vid).cuda()
d_kps = torch.zeros((B + 1, cap, 28), dtype=torch.uint8, device="cuda")
d_desc = torch.zeros((B + 1, cap, 32), dtype=torch.uint8, device="cuda")
d_cnt = torch.zeros(B + 1, dtype=torch.int32, device="cuda")
d_m = torch.zeros((B, cap, 12), dtype=torch.uint8, device="cuda")
d_mc = torch.zeros(B, dtype=torch.int32, device="cuda")
m = Matcher(cap, B)
s = torch.cuda.current_stream()
a_idx = list(range(1, B + 1)); b_idx = list(range(0, B))
def step(match=True):
    det.ExtractDevice(d_img, d_kps[1:], d_desc[1:], d_cnt[1:], cap, s)
    if match:
        m.MatchDevice(d_desc, d_cnt, cap * 32, a_idx, b_idx, d_m, cap, d_mc, 30, 1, s)
for _ in range(3): step()
torch.cuda.synchronize()
for name, mt in (("extract only", False), ("extract+match", True)):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters): step(mt)
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%s: batch %d  %.3f ms/batch  %.1f fps" % (name, B, ms, B / ms * 1e3))
print("counts", d_cnt[1:5].cpu().numpy(), "matches", d_mc[:5].cpu().numpy())
