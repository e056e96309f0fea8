"""CPU tests of the IndexedMatch oracle (oracle/orb_oracle.cpp, ref Tracking/FeatureMatcher.cpp:192-268)."""
import numpy as np

from tests import oracle_orb as orc
from tests.indexed_cases import csr, make_case


def test_full_candidate_lists_reduce_to_match():
    """With every feature of the other image as candidate (ascending) and minDiff >= 1, IndexedMatch is Match."""
    descA, descB, _, _, maskA, maskB = make_case(300, 350, seed=2)
    a2b = csr([np.arange(len(descB), dtype=np.int32)] * len(descA))
    b2a = csr([np.arange(len(descA), dtype=np.int32)] * len(descB))
    for mh, md in ((30, 1), (50, 3)):
        got = orc.indexed_match(descA, descB, a2b, b2a, mh, md, maskA, maskB)
        want = orc.match(descA, descB, mh, md, maskA, maskB)
        assert len(got) > 20 and np.array_equal(got, want)


def test_gated_lists_properties():
    descA, descB, a2b, b2a, maskA, maskB = make_case(500, 600, seed=4)
    got = orc.indexed_match(descA, descB, csr(a2b), csr(b2a), 30, 1, maskA, maskB)
    assert len(got) > 50
    assert np.all(np.diff(got["query"]) > 0)                       # ascending idxA, each A at most once
    for m in got[:200]:
        a, b = int(m["query"]), int(m["train"])
        assert maskA[a] and maskB[b] and b in a2b[a] and a in b2a[b]
        assert orc.descriptor_distance(descA[a], descB[b]) == int(m["distance"]) <= 30
    # empty masks / empty lists
    assert len(orc.indexed_match(descA, descB, csr(a2b), csr(b2a), 30, 1, np.zeros(len(descA), np.uint8), maskB)) == 0
    empty = csr([np.zeros(0, np.int32)] * len(descA))
    assert len(orc.indexed_match(descA, descB, empty, csr(b2a), 30, 1, maskA, maskB)) == 0
