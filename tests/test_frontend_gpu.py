"""GPU tests of the video front-end (mage_frontend_*): batch extract + match-vs-previous equals the single-call APIs and the oracle."""
import numpy as np
import pytest

from mageslam_b200 import synth
from mageslam_b200.frontend import FrontEnd
from mageslam_b200.matcher import Match
from mageslam_b200.orb import FeatureExtractorSettings, OrbFeatureDetector
from tests import oracle_orb as orc

pytestmark = pytest.mark.gpu


def tuples(m, q="query_idx", t="train_idx"):
    return [(int(a), int(b), float(d)) for a, b, d in zip(m[q], m[t], m["distance"])]


@pytest.mark.parametrize("chunk", [0, 3])
def test_host_and_device_variants_match_single_calls(chunk):
    import torch
    s = FeatureExtractorSettings.tier()
    vid = synth.video_frames(10, 640, 480, seed=2)
    det = OrbFeatureDetector(s)
    singles = [det.Process(f) for f in vid]
    fe = FrontEnd(s, 640, 480, batch=5, chunk=chunk)
    outs = fe.alloc_outputs(pinned=True)
    h = torch.from_numpy(vid).pin_memory()
    d = h.cuda()
    fe_dev = FrontEnd(s, 640, 480, batch=5, chunk=chunk)
    for call in range(2):                       # second call matches its first frame against the last frame of the first call
        kps, desc, cnt, mt, mc = fe.Process(h[call * 5:(call + 1) * 5], outs)
        fe_dev.ProcessDevice(d[call * 5:(call + 1) * 5], torch.cuda.current_stream())
        dk, dd, dc, dm, dmc = fe_dev.ReadDeviceResults()
        for i in range(5):
            g = call * 5 + i
            sk, sd = singles[g]
            assert cnt[i] == len(sk) == dc[i]
            assert kps[i, :cnt[i]].tobytes() == sk.tobytes() == dk[i, :dc[i]].tobytes()
            assert np.array_equal(desc[i, :cnt[i]], sd) and np.array_equal(dd[i, :dc[i]], sd)
            if g == 0:
                assert mc[i] == 0 and dmc[i] == 0          # no predecessor
            else:
                ref = Match(sd, singles[g - 1][1], None, None, 30, 1)
                assert tuples(mt[i, :mc[i]]) == tuples(ref) == tuples(dm[i, :dmc[i]])
                assert mc[i] > 100
    # oracle cross-check of one pair end to end
    o1 = orc.detect_and_compute(orc.tier_params(), vid[7], 1)[1]; o0 = orc.detect_and_compute(orc.tier_params(), vid[6], 1)[1]
    assert tuples(mt[2, :mc[2]]) == tuples(orc.match(o1, o0, 30, 1), "query", "train")


def test_reset_forgets_predecessor():
    import torch
    s = FeatureExtractorSettings.tier(num_features=500, num_levels=4)
    vid = torch.from_numpy(synth.video_frames(4, 640, 480, seed=4)).pin_memory()
    fe = FrontEnd(s, 640, 480, batch=2)
    outs = fe.alloc_outputs()
    fe.Process(vid[:2], outs)
    _, _, _, _, mc = fe.Process(vid[2:], outs)
    assert mc[0] > 0
    fe.Reset()
    _, _, _, _, mc = fe.Process(vid[2:], outs)
    assert mc[0] == 0 and mc[1] > 0


def test_pipelined_submit_wait_equals_synchronous_process():
    """two calls in flight (upload of call j+1 and download of call j-1 under the compute of call j): same results as Process"""
    import torch
    s = FeatureExtractorSettings.tier(num_features=800, num_levels=5)
    vid = torch.from_numpy(synth.video_frames(16, 640, 480, seed=6)).pin_memory()
    sync = FrontEnd(s, 640, 480, batch=4, chunk=2)
    so = sync.alloc_outputs()
    want = []
    for c in range(4):
        want.append([np.copy(a) for a in sync.Process(vid[4 * c:4 * c + 4], so)])
    pipe = FrontEnd(s, 640, 480, batch=4, chunk=2)
    po = [pipe.alloc_outputs(), pipe.alloc_outputs()]
    got = []
    pipe.Submit(vid[0:4], po[0])
    for c in range(1, 4):
        pipe.Submit(vid[4 * c:4 * c + 4], po[c & 1])
        pipe.Wait()
        got.append([np.copy(a) for a in pipe.views(po[(c - 1) & 1])])
    pipe.Wait()
    got.append([np.copy(a) for a in pipe.views(po[1])])
    with pytest.raises(Exception):
        pipe.Wait()                                   # nothing in flight
    for w, g in zip(want, got):
        kps, desc, cnt, mt, mc = w
        gk, gd, gc, gm, gmc = g
        assert np.array_equal(cnt, gc) and np.array_equal(mc, gmc)
        for i in range(4):
            assert kps[i, :cnt[i]].tobytes() == gk[i, :cnt[i]].tobytes() and np.array_equal(desc[i, :cnt[i]], gd[i, :cnt[i]])
            assert mt[i, :mc[i]].tobytes() == gm[i, :mc[i]].tobytes()
    assert want[1][4][0] > 50                         # the first frame of a later call is matched against the previous call's last frame


def test_pipelined_calls_with_varying_frame_counts():
    """calls of different sizes (fewer frames than the batch, different chunk counts per call) through submit / wait and process"""
    import torch
    s = FeatureExtractorSettings.tier(num_features=600, num_levels=4)
    vid = synth.video_frames(14, 640, 480, seed=8)
    det = OrbFeatureDetector(s)
    singles = [det.Process(f) for f in vid]
    h = torch.from_numpy(vid).pin_memory()
    fe = FrontEnd(s, 640, 480, batch=5, chunk=2)
    outs = [fe.alloc_outputs(), fe.alloc_outputs()]
    sizes = [5, 3, 1, 5]                      # 14 frames
    starts = np.cumsum([0] + sizes)
    results = []
    fe.Submit(h[starts[0]:starts[1]], outs[0])
    for c in range(1, len(sizes)):
        fe.Submit(h[starts[c]:starts[c + 1]], outs[c & 1])
        fe.Wait()
        results.append([np.copy(a) for a in fe.views(outs[(c - 1) & 1])])
    fe.Wait()
    results.append([np.copy(a) for a in fe.views(outs[(len(sizes) - 1) & 1])])
    for c, (kps, desc, cnt, mt, mc) in enumerate(results):
        for i in range(sizes[c]):
            g = starts[c] + i
            sk, sd = singles[g]
            assert cnt[i] == len(sk) and kps[i, :cnt[i]].tobytes() == sk.tobytes() and np.array_equal(desc[i, :cnt[i]], sd), (c, i)
            if g == 0:
                assert mc[i] == 0
            else:
                ref = Match(sd, singles[g - 1][1], None, None, 30, 1)
                assert tuples(mt[i, :mc[i]]) == tuples(ref), (c, i)
    # the synchronous call after pipelined ones continues the same sequence
    fe2 = FrontEnd(s, 640, 480, batch=5, chunk=2)
    o = fe2.alloc_outputs()
    fe2.Process(h[0:5], o)
    kps, desc, cnt, mt, mc = fe2.Process(h[5:8], o)
    assert tuples(mt[0, :mc[0]]) == tuples(Match(singles[5][1], singles[4][1], None, None, 30, 1))
