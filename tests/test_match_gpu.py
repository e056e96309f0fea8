"""GPU parity tests of the brute-force matcher: CUDA path (C ABI) vs the CPU oracle, exact (indices and distances)."""
import numpy as np
import pytest

from mageslam_b200 import synth
from mageslam_b200.matcher import GetDescriptorDistance, Match, Matcher
from tests import oracle_orb as orc

pytestmark = pytest.mark.gpu


def as_tuples(m, q="query_idx", t="train_idx"):
    return [(int(a), int(b), float(d)) for a, b, d in zip(m[q], m[t], m["distance"])]


def noisy_copy(rng, A, flips, dup_frac=0.05):
    n = len(A)
    B = A.copy()[rng.permutation(n)]
    for i in range(n):
        for b in rng.integers(0, 256, int(rng.integers(0, flips))):
            B[i, b // 8] ^= np.uint8(1 << (b % 8))
    k = int(n * dup_frac)
    if k:
        B[:k] = B[k:2 * k]          # exact duplicates => ties for best => rejected by min-diff
    return B


# maxd <= 64 runs the tcgen05 kernel (sizes straddle its 128-query / 256-train tiles), larger radii the xor/popc kernel
@pytest.mark.parametrize("n,flips,maxd,mind", [(2000, 12, 30, 1), (777, 30, 40, 3), (33, 4, 30, 1), (1500, 60, 64, 2), (256, 8, 30, 0),
                                               (129, 6, 30, 1), (300, 6, 0, 0), (600, 40, 100, 2), (1000, 90, 65, 1)])
def test_match_equals_oracle(n, flips, maxd, mind):
    rng = np.random.default_rng(n)
    A = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    B = noisy_copy(rng, A, flips)[: n - n // 7]         # ragged: nA != nB
    got = Match(A, B, None, None, maxd, mind)
    ref = orc.match(A, B, maxd, mind)
    assert as_tuples(got) == as_tuples(ref, "query", "train")
    assert len(ref) > 0


@pytest.mark.parametrize("mind", [0, 1])
def test_match_many_hits_per_query(mind):
    """8 distinct descriptors repeated 128 times: 128 x 128 in-radius pairs per query block overflow the kernel's shared-memory hit list
    (the overflow takes the direct atomic path); ties resolve to the lowest index and are rejected when minDiff >= 1."""
    rng = np.random.default_rng(11)
    base = rng.integers(0, 256, (8, 32), dtype=np.uint8)
    A = np.repeat(base, 128, axis=0)
    B = np.tile(base, (128, 1))
    got = Match(A, B, None, None, 30, mind)
    ref = orc.match(A, B, 30, mind)
    assert as_tuples(got) == as_tuples(ref, "query", "train")


def test_match_with_masks_and_empty_sides():
    rng = np.random.default_rng(3)
    A = rng.integers(0, 256, (600, 32), dtype=np.uint8)
    B = noisy_copy(rng, A, 10)
    mA = (rng.random(600) < 0.7).astype(np.uint8); mB = (rng.random(600) < 0.6).astype(np.uint8)
    got = Match(A, B, mA, mB, 30, 1)
    ref = orc.match(A, B, 30, 1, mA, mB)
    assert as_tuples(got) == as_tuples(ref, "query", "train") and len(ref) > 0
    assert len(Match(A[:0], B)) == 0 and len(Match(A, B[:0])) == 0
    # all-false mask behaves like an empty side (reference returns 0 and clears goodMatches)
    assert len(Match(A, B, np.zeros(600, np.uint8), None)) == 0


def test_match_is_symmetric_under_swap():
    rng = np.random.default_rng(11)
    A = rng.integers(0, 256, (900, 32), dtype=np.uint8)
    B = noisy_copy(rng, A, 16)
    ab = Match(A, B); ba = Match(B, A)
    assert sorted((int(q), int(t)) for q, t in zip(ab["query_idx"], ab["train_idx"])) == \
           sorted((int(t), int(q)) for q, t in zip(ba["query_idx"], ba["train_idx"]))
    self_m = Match(A, A, maxHammingDist=0, minHammingDifference=1)      # identity when descriptors are distinct
    assert np.array_equal(self_m["query_idx"], self_m["train_idx"]) and len(self_m) == 900


def test_real_descriptors_self_match():
    # frame vs shifted+noisy copy of itself
    from tests.test_orb_gpu import make_detector
    p = orc.tier_params()
    img = synth.video_frames(1, 640, 480, seed=5)[0]
    img2 = synth.shifted_noisy(img)
    det = make_detector(p)
    k1, d1 = det.DetectAndCompute(img); k2, d2 = det.DetectAndCompute(img2)
    got = Match(d1, d2, None, None, 30, 1)
    ref = orc.match(d1, d2, 30, 1)
    assert as_tuples(got) == as_tuples(ref, "query", "train")
    assert len(ref) > 100
    # matched keypoints should mostly be displaced by the (3, 2) shift
    dx = k2["x"][got["train_idx"]] - k1["x"][got["query_idx"]]
    dy = k2["y"][got["train_idx"]] - k1["y"][got["query_idx"]]
    assert np.median(np.abs(dx - 3)) < 2.5 and np.median(np.abs(dy - 2)) < 2.5


def test_device_batched_pairs_equal_host_api():
    import torch
    rng = np.random.default_rng(21)
    slots, cap = 5, 2000
    desc = np.zeros((slots, cap, 32), np.uint8)
    counts = np.array([2000, 1873, 1999, 640, 2000], np.int32)
    base = rng.integers(0, 256, (cap, 32), dtype=np.uint8)
    for s in range(slots):
        desc[s, :counts[s]] = noisy_copy(rng, base, 10 + 3 * s)[:counts[s]]
    d_desc = torch.from_numpy(desc).cuda(); d_counts = torch.from_numpy(counts).cuda()
    a_idx = [1, 2, 3, 4]; b_idx = [0, 1, 2, 3]
    d_matches = torch.zeros((4, cap, 12), dtype=torch.uint8, device="cuda")
    d_mcounts = torch.zeros(4, dtype=torch.int32, device="cuda")
    m = Matcher(cap, 4)
    for _ in range(2):          # second call exercises the cached job table
        m.MatchDevice(d_desc, d_counts, cap * 32, a_idx, b_idx, d_matches, cap, d_mcounts, 30, 1, torch.cuda.current_stream())
    torch.cuda.synchronize()
    mc = d_mcounts.cpu().numpy()
    from mageslam_b200._lib import DMATCH_DTYPE
    for p in range(4):
        got = d_matches[p].cpu().numpy().reshape(-1).view(DMATCH_DTYPE)[: mc[p]]
        ref = orc.match(desc[a_idx[p], :counts[a_idx[p]]], desc[b_idx[p], :counts[b_idx[p]]], 30, 1)
        assert as_tuples(got) == as_tuples(ref, "query", "train") and len(ref) > 0


def test_descriptor_distance():
    import torch
    rng = np.random.default_rng(2)
    a = rng.integers(0, 256, (500, 32), dtype=np.uint8); b = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    got = GetDescriptorDistance(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    ref = np.array([orc.descriptor_distance(x, y) for x, y in zip(a, b)])
    assert np.array_equal(got, ref)
