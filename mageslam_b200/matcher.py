"""Host-side mirror of the reference's matcher surface (Tracking/FeatureMatcher.h:68-77, :134-136) over the C ABI."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import DMATCH_DTYPE, KEYPOINT_DTYPE, check, lib, ptr, stream_ptr


@dataclass
class OrbMatcherSettings:
    """Defaults = reference MageSettings.h:36-39."""
    MaxHammingDistance: int = 30
    MinHammingDifference: int = 1


class Matcher:
    """Owns the device workspace of mage_matcher_t (capacity in descriptors / batched pairs)."""

    def __init__(self, max_descriptors=4096, max_pairs=1):
        self._h = C.c_void_p()
        self.max_descriptors, self.max_pairs = int(max_descriptors), int(max_pairs)
        check(lib().mage_matcher_create(self.max_descriptors, self.max_pairs, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().mage_matcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Match(self, descA, descB, maskA=None, maskB=None, maxHammingDist=30, minHammingDifference=1):
        """Mirror of Match(imageA, imageB, imageAMask, imageBMask, ..., maxHammingDist, minHammingDifference, goodMatches):
        returns goodMatches as a DMATCH_DTYPE array (queryIdx = index in A, trainIdx = index in B), ascending queryIdx."""
        descA = np.ascontiguousarray(descA, np.uint8).reshape(-1, 32)
        descB = np.ascontiguousarray(descB, np.uint8).reshape(-1, 32)
        nA, nB = len(descA), len(descB)
        if maskA is not None:
            maskA = np.ascontiguousarray(maskA, np.uint8)
            assert len(maskA) == nA
        if maskB is not None:
            maskB = np.ascontiguousarray(maskB, np.uint8)
            assert len(maskB) == nB
        out = np.zeros(max(nA, 1), DMATCH_DTYPE)
        cnt = C.c_int(0)
        check(lib().mage_match_bf(self._h, ptr(descA) if nA else None, nA, ptr(maskA), ptr(descB) if nB else None, nB, ptr(maskB),
                                  int(maxHammingDist), int(minHammingDifference), ptr(out), C.byref(cnt), None))
        return out[:cnt.value]

    def IndexedMatch(self, closeMatchesAtoB, closeMatchesBtoA, descA, descB, maskA=None, maskB=None, maxHammingDist=30,
                     minHammingDifference=1):
        """Mirror of IndexedMatch (reference Tracking/FeatureMatcher.h:30-45, .cpp:192-268). closeMatchesAtoB[i] is what
        QueryFeatures(descA[i], idB / matcherB) returned (indices into B, in that order), closeMatchesBtoA[j] likewise for
        descB[j]; either a list of index lists or a CSR pair (offsets int32[n+1], candidates int32[...])."""
        descA = np.ascontiguousarray(descA, np.uint8).reshape(-1, 32)
        descB = np.ascontiguousarray(descB, np.uint8).reshape(-1, 32)
        nA, nB = len(descA), len(descB)
        o0, c0 = _as_csr(closeMatchesAtoB, nA)
        o1, c1 = _as_csr(closeMatchesBtoA, nB)
        maskA = None if maskA is None else np.ascontiguousarray(maskA, np.uint8)
        maskB = None if maskB is None else np.ascontiguousarray(maskB, np.uint8)
        out = np.zeros(max(nA, 1), DMATCH_DTYPE)
        cnt = C.c_int(0)
        check(lib().mage_indexed_match(self._h, ptr(descA) if nA else None, nA, ptr(maskA), ptr(descB) if nB else None, nB, ptr(maskB),
                                       ptr(o0), ptr(c0) if len(c0) else None, ptr(o1), ptr(c1) if len(c1) else None,
                                       int(maxHammingDist), int(minHammingDifference), ptr(out), C.byref(cnt), None))
        return out[:cnt.value]

    def MatchDevice(self, d_desc, d_counts, slot_stride, a_index, b_index, d_matches, capacity, d_match_counts,
                    maxHammingDist=30, minHammingDifference=1, stream=None):
        a_index = np.ascontiguousarray(a_index, np.int32); b_index = np.ascontiguousarray(b_index, np.int32)
        check(lib().mage_match_bf_device(self._h, ptr(d_desc), ptr(d_counts), int(slot_stride), ptr(a_index), ptr(b_index), len(a_index),
                                         int(maxHammingDist), int(minHammingDifference), ptr(d_matches), int(capacity),
                                         ptr(d_match_counts), stream_ptr(stream)))


def _as_csr(lists, n):
    if isinstance(lists, tuple) and len(lists) == 2:
        off, cand = (np.ascontiguousarray(a, np.int32) for a in lists)
    else:
        off = np.zeros(n + 1, np.int32)
        off[1:] = np.cumsum([len(l) for l in lists], dtype=np.int64)
        cand = np.ascontiguousarray(np.concatenate([np.asarray(l, np.int32) for l in lists]) if n and off[-1] else np.zeros(0, np.int32), np.int32)
    assert len(off) == n + 1
    return off, cand


_default = None


def Match(descA, descB, maskA=None, maskB=None, maxHammingDist=30, minHammingDifference=1):
    """Free-function form, as in the reference (a process-wide workspace is created on first use)."""
    global _default
    n = max(len(descA), len(descB), 1)
    if _default is None or _default.max_descriptors < n:
        _default = Matcher(max(4096, n), 1)
    return _default.Match(descA, descB, maskA, maskB, maxHammingDist, minHammingDifference)


def IndexedMatch(closeMatchesAtoB, closeMatchesBtoA, descA, descB, maskA=None, maskB=None, maxHammingDist=30, minHammingDifference=1):
    """Free-function form of Matcher.IndexedMatch."""
    global _default
    n = max(len(descA), len(descB), 1)
    if _default is None or _default.max_descriptors < n:
        _default = Matcher(max(4096, n), 1)
    return _default.IndexedMatch(closeMatchesAtoB, closeMatchesBtoA, descA, descB, maskA, maskB, maxHammingDist, minHammingDifference)


def GetDescriptorDistance(d0, d1):
    """Mirror of GetDescriptorDistance (reference FeatureMatcher.cpp:453-504) for arrays of device descriptors (torch)."""
    import torch
    n = d0.shape[0]
    out = torch.empty(n, dtype=torch.int32, device=d0.device)
    check(lib().mage_descriptor_distance_device(ptr(d0), ptr(d1), n, ptr(out), stream_ptr(torch.cuda.current_stream())))
    return out


class KeypointSpatialIndex:
    """Mirror of mage::KeypointSpatialIndex (reference Image/KeypointSpatialIndex.h:22-38): built from an image's keypoints."""

    def __init__(self, keypoints):
        self.keypoints = np.ascontiguousarray(keypoints, KEYPOINT_DTYPE)
        self._h = C.c_void_p()
        check(lib().mage_spatial_index_create(ptr(self.keypoints) if len(self.keypoints) else None, len(self.keypoints), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().mage_spatial_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Rank(self):
        r = np.zeros(len(self.keypoints), np.int32)
        if len(r):
            check(lib().mage_spatial_index_rank(self._h, ptr(r)))
        return r


def RadiusMatch(queryKeypoints, queryKeypointPositionOverrides, queryKeypointsMask, queryDescriptors, targetKeypointsIndex,
                targetKeypointsMask, targetDescriptors, radius, maxHammingDist, minHammingDifference):
    """Mirror of RadiusMatch (reference Tracking/FeatureMatcher.h:92-106): returns goodMatches (DMATCH_DTYPE, ascending queryIdx).
    targetKeypoints are the keypoints targetKeypointsIndex was built from."""
    qk = np.ascontiguousarray(queryKeypoints, KEYPOINT_DTYPE)
    qd = np.ascontiguousarray(queryDescriptors, np.uint8).reshape(-1, 32)
    td = np.ascontiguousarray(targetDescriptors, np.uint8).reshape(-1, 32)
    assert len(qd) == len(qk) and len(td) == len(targetKeypointsIndex.keypoints)
    qp = None if queryKeypointPositionOverrides is None else np.ascontiguousarray(queryKeypointPositionOverrides, np.float32).reshape(-1, 2)
    qm = None if queryKeypointsMask is None else np.ascontiguousarray(queryKeypointsMask, np.uint8)
    tm = None if targetKeypointsMask is None else np.ascontiguousarray(targetKeypointsMask, np.uint8)
    out = np.zeros(max(len(qk), 1), DMATCH_DTYPE)
    cnt = C.c_int(0)
    check(lib().mage_radius_match(targetKeypointsIndex._h, ptr(qk) if len(qk) else None, len(qk), ptr(qp), ptr(qm), ptr(qd) if len(qk) else None,
                                  ptr(tm), ptr(td) if len(td) else None, float(radius), int(maxHammingDist), int(minHammingDifference),
                                  ptr(out), C.byref(cnt), None))
    return out[:cnt.value]
