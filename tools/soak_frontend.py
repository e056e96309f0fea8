"""Soak test of the pipelined front-end: random call sizes, chunk sizes and interleavings of submit / wait / process / reset over a
few thousand calls, every result compared with the single-frame path (bit-exact key points / descriptors, exact matches).
usage: python tools/soak_frontend.py [calls] [seed]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from mageslam_b200 import synth
from mageslam_b200.frontend import FrontEnd
from mageslam_b200.orb import FeatureExtractorSettings, OrbFeatureDetector
from mageslam_b200.matcher import Match

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
s = FeatureExtractorSettings.tier(num_features=500, num_levels=4)
W, H, NF = 416, 300, 48
vid = synth.video_frames(NF, W, H, seed=3)
det = OrbFeatureDetector(s)
singles = [det.Process(f) for f in vid]
match_cache = {}
def ref_match(a, b):
    if (a, b) not in match_cache:
        m = Match(singles[a][1], singles[b][1], None, None, 30, 1)
        match_cache[(a, b)] = sorted(zip(m["query_idx"].tolist(), m["train_idx"].tolist()))
    return match_cache[(a, b)]
h = torch.from_numpy(vid).pin_memory()
t0 = time.time(); checked = 0
for trial in range(max(calls // 60, 1)):
    batch = int(rng.integers(1, 9)); chunk = int(rng.integers(1, batch + 1))
    fe = FrontEnd(s, W, H, batch=batch, chunk=chunk)
    outs = [fe.alloc_outputs(pinned=bool(rng.integers(0, 2))), fe.alloc_outputs(pinned=True)]
    prev = None                                   # index of the frame the next first frame is matched against
    pending = []                                  # (slot, frame indices, prev) of calls in flight
    def check(slot, idx, prv):
        global checked
        kps, desc, cnt, mt, mc = fe.views(outs[slot])
        for i, g in enumerate(idx):
            sk, sd = singles[g]
            assert cnt[i] == len(sk) and kps[i, :cnt[i]].tobytes() == sk.tobytes() and np.array_equal(desc[i, :cnt[i]], sd), (trial, g)
            p = prv if i == 0 else idx[i - 1]
            if p is None:
                assert mc[i] == 0, (trial, g)
            else:
                got = sorted(zip(mt[i, :mc[i]]["query_idx"].tolist(), mt[i, :mc[i]]["train_idx"].tolist()))
                assert got == ref_match(g, p), (trial, g, p)
            checked += 1
    for c in range(60):
        n = int(rng.integers(1, batch + 1))
        idx = [int(x) for x in rng.integers(0, NF, n)]
        frames = h[idx] if n > 1 else h[idx[0]:idx[0] + 1]
        frames = frames.contiguous().pin_memory() if n > 1 else frames
        op = rng.random()
        if op < 0.08:
            while pending:
                fe.Wait(); check(*pending.pop(0))
            fe.Reset(); prev = None
            continue
        if op < 0.45:                             # synchronous call (waits for what is in flight first)
            while pending:
                fe.Wait(); check(*pending.pop(0))
            slot = int(rng.integers(0, 2))
            fe.Process(frames, outs[slot]); check(slot, idx, prev)
        else:
            if len(pending) == 2:
                fe.Wait(); check(*pending.pop(0))
            slot = 0 if not pending else 1 - pending[-1][0]
            fe.Submit(frames, outs[slot]); pending.append((slot, idx, prev))
        prev = idx[-1]
    while pending:
        fe.Wait(); check(*pending.pop(0))
print("front-end soak: %d frames through random submit / wait / process / reset sequences, all equal to the single-frame path, %.0f s" % (checked, time.time() - t0))
