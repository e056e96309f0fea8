"""Per-kernel timing of the ORB chain for every k_fast variant (MAGE_FAST_VARIANT, MAGE_FAST_TMA): CUDA events of the library's own
profiler around each launch, 128 frames per launch from a ring larger than L2. Tuning aid, not the contract bench.
usage: python tools/fast_sweep.py [variants, e.g. 0,1,2] [tma: 0/1/both]"""
import ctypes as C, os, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from mageslam_b200 import synth, _lib
from mageslam_b200.orb import FeatureExtractorSettings, OrbFeatureDetector

variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,1,2,3,4,5,6,7").split(",")]
tmas = {"0": [0], "1": [1], "both": [0, 1]}[sys.argv[2] if len(sys.argv) > 2 else "both"]
scene = sys.argv[3] if len(sys.argv) > 3 else "video"
B, RING, cap = 128, 512, 2000
if scene == "video":
    base = synth.video_frames(32, 640, 480, seed=10)
else:
    base = synth.natural_frames(32, 640, 480, seed=10)
ring = np.concatenate([base, base[:, ::-1], base[:, :, ::-1], base[:, ::-1, ::-1]] * (RING // 128), 0)
d_ring = torch.from_numpy(np.ascontiguousarray(ring)).cuda()
L = _lib.lib()
L.mage_profile_get.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
L.mage_profile_name.restype = C.c_char_p
s = torch.cuda.current_stream()
d_kps = torch.zeros((B, cap, 28), dtype=torch.uint8, device="cuda")
d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
d_cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
ref = None
for tma in tmas:
    for v in variants:
        os.environ["MAGE_FAST_VARIANT"] = str(v); os.environ["MAGE_FAST_TMA"] = str(tma)
        det = OrbFeatureDetector(FeatureExtractorSettings.tier(), max_batch=B).m_detector
        for k in range(3):
            det.ExtractDevice(d_ring[k * B:(k + 1) * B], d_kps, d_desc, d_cnt, cap, s)
        torch.cuda.synchronize()
        sig = (d_kps[:8].cpu().numpy().tobytes(), d_desc[:8].cpu().numpy().tobytes())
        if ref is None: ref = sig
        L.mage_profile_reset(); L.mage_profile_enable(1)
        for it in range(12):
            det.ExtractDevice(d_ring[(it % 4) * B:(it % 4 + 1) * B], d_kps, d_desc, d_cnt, cap, s)
        L.mage_profile_collect(); L.mage_profile_enable(0)
        out = {}
        for sl in range(L.mage_profile_slots()):
            tot, n = C.c_double(0), C.c_longlong(0)
            L.mage_profile_get(sl, C.byref(tot), C.byref(n))
            if n.value: out[L.mage_profile_name(sl).decode()] = tot.value / n.value
        print("variant %d tma %d same_output %s  " % (v, tma, sig == ref) + "  ".join("%s %.4f" % kv for kv in out.items()), flush=True)
        del det
