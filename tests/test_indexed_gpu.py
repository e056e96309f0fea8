"""GPU parity of mage_indexed_match (IndexedMatch, ref Tracking/FeatureMatcher.cpp:192-268) against the oracle, through the C ABI."""
import numpy as np
import pytest

from mageslam_b200.matcher import IndexedMatch, Match, Matcher

from tests import oracle_orb as orc
from tests.indexed_cases import csr, make_case

pytestmark = pytest.mark.gpu


def same(got, want):
    return (len(got) == len(want) and np.array_equal(got["query_idx"], want["query"]) and np.array_equal(got["train_idx"], want["train"])
            and np.array_equal(got["distance"], want["distance"]))


@pytest.mark.parametrize("seed,nA,nB,mh,md", [(0, 600, 700, 30, 1), (1, 2000, 2000, 50, 2), (2, 37, 1500, 30, 0), (3, 1200, 33, 64, 1)])
def test_indexed_match_equals_oracle(seed, nA, nB, mh, md):
    descA, descB, a2b, b2a, maskA, maskB = make_case(nA, nB, seed=seed)
    got = IndexedMatch(a2b, b2a, descA, descB, maskA, maskB, mh, md)
    want = orc.indexed_match(descA, descB, csr(a2b), csr(b2a), mh, md, maskA, maskB)
    assert same(got, want)
    if nA >= 600 and nB >= 600:
        assert len(got) > 50
    # CSR input form and no masks
    got2 = IndexedMatch(csr(a2b), csr(b2a), descA, descB, None, None, mh, md)
    assert same(got2, orc.indexed_match(descA, descB, csr(a2b), csr(b2a), mh, md))


def test_full_lists_equal_brute_force_match():
    descA, descB, _, _, maskA, maskB = make_case(800, 900, seed=9)
    a2b = csr([np.arange(len(descB), dtype=np.int32)] * len(descA))
    b2a = csr([np.arange(len(descA), dtype=np.int32)] * len(descB))
    got = IndexedMatch(a2b, b2a, descA, descB, maskA, maskB, 30, 1)
    bf = Match(descA, descB, maskA, maskB, 30, 1)
    assert len(got) > 100 and np.array_equal(got, bf)


def test_ties_resolve_to_the_first_list_entry_when_min_diff_is_zero():
    rng = np.random.default_rng(5)
    descB = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    descB[40] = descB[7]                                   # two identical candidates
    descA = descB[[7]].copy()
    a2b = [np.array([50, 40, 3, 7], np.int32)]             # 40 comes first in QueryFeatures order
    b2a = [np.array([0], np.int32)] * 64
    got = IndexedMatch(a2b, b2a, descA, descB, None, None, 30, 0)
    want = orc.indexed_match(descA, descB, csr(a2b), csr(b2a), 30, 0)
    assert same(got, want) and len(got) == 1 and got["train_idx"][0] == 40
    assert len(IndexedMatch(a2b, b2a, descA, descB, None, None, 30, 1)) == 0      # best - second = 0 < minDiff


def test_empty_inputs_and_errors():
    from mageslam_b200._lib import MageError
    descA, descB, a2b, b2a, maskA, maskB = make_case(50, 60, seed=1)
    assert len(IndexedMatch(a2b[:0], b2a, descA[:0], descB)) == 0
    assert len(IndexedMatch([np.zeros(0, np.int32)] * 50, b2a, descA, descB)) == 0
    assert len(IndexedMatch(a2b, b2a, descA, descB, np.zeros(50, np.uint8), maskB)) == 0
    bad = [np.array([9999], np.int32)] * 50                # out-of-range candidates are ignored, not dereferenced
    assert len(IndexedMatch(bad, b2a, descA, descB)) == 0
    m = Matcher(40, 1)
    with pytest.raises(MageError):
        m.IndexedMatch(a2b, b2a, descA, descB)
