"""ctypes binding of oracle/liborb_oracle.so (TEST INFRASTRUCTURE: the CPU checker, never the product path)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
DM_DTYPE = np.dtype([("query", "<i4"), ("train", "<i4"), ("distance", "<f4")])
assert KP_DTYPE.itemsize == 28 and DM_DTYPE.itemsize == 12


class OrbParams(C.Structure):
    """The 14 OrbDetector ctor scalars (reference Image/OpenCVModified.h:68-82)."""
    _fields_ = [("gaussian_kernel_size", C.c_uint32), ("nfeatures", C.c_uint32), ("scale_factor", C.c_float),
                ("nlevels", C.c_uint32), ("patch_size", C.c_uint32), ("fast_threshold", C.c_uint32),
                ("use_orientation", C.c_int32), ("feature_factor", C.c_float), ("feature_strength", C.c_float),
                ("strong_response", C.c_int32), ("min_robust_factor", C.c_float), ("max_robust_factor", C.c_float),
                ("num_cells_x", C.c_int32), ("num_cells_y", C.c_int32)]


def tier_params(nfeatures=2000, nlevels=8, scale_factor=1.2, fast_threshold=10):
    """SURVEY 8(d) config 1/2 extractor settings."""
    return OrbParams(7, nfeatures, scale_factor, nlevels, 31, fast_threshold, 1, 1.5, 0.9, 20, 1.1, 2.0, 32, 32)


def default_params():
    """Reference defaults, MageSettings.h:151-167."""
    return OrbParams(7, 440, 1.5, 1, 15, 4, 0, 1.5, 0.9, 20, 1.1, 2.0, 32, 32)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liborb_oracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
        L = C.CDLL(path)
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_cv_round_f.argtypes = [C.c_float]
        _LIB = L
    return _LIB


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


def resize(src, dw, dh):
    src, sp = _u8(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(sp, src.shape[1], src.shape[0], src.strides[0], dst.ctypes.data_as(C.c_void_p), dw, dh, dw)
    return dst


def blur(src, ksize=7):
    src, sp = _u8(src)
    dst = np.empty_like(src)
    rc = lib().orc_gaussian_blur_u8(sp, src.shape[1], src.shape[0], src.strides[0], dst.ctypes.data_as(C.c_void_p),
                                    dst.strides[0], ksize)
    assert rc == 0
    return dst


def blur_submatrix(src, ksize=7, fused=True):
    """cv::GaussianBlur of a SUBMATRIX source (the reference's in-place blur of a level ROI): the generic separable float path"""
    src, sp = _u8(src)
    dst = np.empty_like(src)
    rc = lib().orc_gaussian_blur_submatrix_u8(sp, src.shape[1], src.shape[0], src.strides[0], dst.ctypes.data_as(C.c_void_p),
                                              dst.strides[0], ksize, 1 if fused else 0)
    assert rc == 0
    return dst


def level_blur(level, ksize, nlevels, width):
    """the blur DetectAndCompute applies to a level: float path when the level ROI is a proper submatrix of the packed buffer"""
    return blur_submatrix(level, ksize) if (nlevels > 1 or (width & 15) != 0) else blur(level, ksize)


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def fast9(img, threshold):
    img, ip = _u8(img)
    cap = img.size // 4 + 16
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_fast9_nms(ip, img.shape[1], img.shape[0], img.strides[0], int(threshold), out.ctypes.data_as(C.c_void_p), cap)
    assert n <= cap
    return out[:n]


def score_map(img, threshold):
    img, ip = _u8(img)
    out = np.zeros_like(img)
    lib().orc_fast9_score_map(ip, img.shape[1], img.shape[0], img.strides[0], int(threshold), out.ctypes.data_as(C.c_void_p), out.strides[0])
    return out


def level_layout(params, w, h):
    n = params.nlevels
    sizes = np.zeros(2 * n, np.int32); scales = np.zeros(n, np.float32); nfeat = np.zeros(n, np.int32)
    lib().orc_level_layout(C.byref(params), w, h, sizes.ctypes.data_as(C.c_void_p), scales.ctypes.data_as(C.c_void_p),
                           nfeat.ctypes.data_as(C.c_void_p))
    return sizes.reshape(n, 2), scales, nfeat


def build_pyramid(params, img):
    img, ip = _u8(img)
    sizes, _, _ = level_layout(params, img.shape[1], img.shape[0])
    levels = [np.zeros((int(hh), int(ww)), np.uint8) for ww, hh in sizes]
    ptrs = (C.c_void_p * len(levels))(*[l.ctypes.data for l in levels])
    lib().orc_build_pyramid(C.byref(params), ip, img.shape[1], img.shape[0], img.strides[0], ptrs)
    return levels


def select_level(params, kps, n_keep, mode):
    buf = np.array(kps, dtype=KP_DTYPE, copy=True)
    n = lib().orc_select_level(C.byref(params), buf.ctypes.data_as(C.c_void_p), len(buf), int(n_keep), int(mode))
    return buf[:n]


def anms_radii(params, kps, n_keep):
    buf = np.ascontiguousarray(kps, dtype=KP_DTYPE)
    r = np.zeros(len(buf), np.int32)
    lib().orc_anms_radii(C.byref(params), buf.ctypes.data_as(C.c_void_p), len(buf), int(n_keep), r.ctypes.data_as(C.c_void_p))
    return r


def brief_pattern(patch):
    out = np.zeros(30 * 1024, np.int8)
    rc = lib().orc_brief_pattern(int(patch), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out.reshape(30, 256, 4)


def random_pattern(patch):
    out = np.zeros(1024, np.int32)
    lib().orc_random_pattern(int(patch), out.ctypes.data_as(C.c_void_p), 512)
    return out.reshape(512, 2)


def generic_descriptors(img, xya, patch):
    """ComputeOrbDescriptors (generic pattern) on one blurred level: xya = float32 [n, 3] (x, y, angle in degrees)"""
    img, sp = _u8(img)
    xya = np.ascontiguousarray(xya, np.float32)
    desc = np.zeros((len(xya), 32), np.uint8)
    lib().orc_generic_descriptors(sp, img.strides[0], xya.ctypes.data_as(C.c_void_p), len(xya), int(patch), desc.ctypes.data_as(C.c_void_p))
    return desc


def umax(half_patch):
    out = np.zeros(half_patch + 2, np.int32)
    lib().orc_umax(int(half_patch), out.ctypes.data_as(C.c_void_p))
    return out


BLUR_AUTO, BLUR_FLOAT_FUSED, BLUR_FLOAT_UNFUSED, BLUR_FIXED = 0, 1, 2, 3


def detect_and_compute(params, img, mode=1, capacity=None, blur_mode=BLUR_AUTO):
    img, ip = _u8(img)
    cap = int(capacity if capacity is not None else params.nfeatures)
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    cnt = C.c_int(0)
    rc = lib().orc_orb_detect_and_compute_ex(C.byref(params), ip, img.shape[1], img.shape[0], img.strides[0], int(mode), int(blur_mode),
                                             kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap, C.byref(cnt))
    if rc != 0:
        raise ValueError("orb oracle: unsupported configuration (%d)" % rc)
    return kps[:cnt.value].copy(), desc[:cnt.value].copy()


_OREF = {}


def orb_ref(sse=True):
    """The REFERENCE's own OrbDetector (OpenCVModified.cpp compiled unmodified behind oracle/cvshim, oracle/_ref/liborb_ref.so;
    sse=False: the build without its CV_SSE2 branches). None when it was not built (needs /root/reference at build time)."""
    name = "liborb_ref.so" if sse else "liborb_ref_nosse.so"
    if name not in _OREF:
        path = os.path.join(ROOT, "oracle", "_ref", name)
        _OREF[name] = None
        if os.path.exists(path):
            lib()                                   # liborb_oracle.so first: the shim's resize / blur / fastAtan2 live there
            R = C.CDLL(path)
            assert R.ref_orb_sse2() == (1 if sse else 0)
            _OREF[name] = R
    return _OREF[name]


def detect_and_compute_ref(params, img, blur_mode=BLUR_AUTO, sse=True, capacity=None):
    """DetectAndCompute of the compiled reference; raises ValueError where the reference throws (CV_Assert)."""
    R = orb_ref(sse)
    img, ip = _u8(img)
    cap = int(capacity if capacity is not None else params.nfeatures)
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    cnt = C.c_int(0)
    rc = R.ref_orb_detect_and_compute(C.byref(params), ip, img.shape[1], img.shape[0], img.strides[0], int(blur_mode),
                                      kps.ctypes.data_as(C.c_void_p), desc.ctypes.data_as(C.c_void_p), cap, C.byref(cnt))
    if rc != 0:
        raise ValueError("reference OrbDetector threw (%d)" % rc)
    return kps[:cnt.value].copy(), desc[:cnt.value].copy()


def match(descA, descB, max_hamming=30, min_diff=1, maskA=None, maskB=None):
    descA, ap = _u8(descA); descB, bp = _u8(descB)
    nA, nB = len(descA), len(descB)
    out = np.zeros(max(nA, 1), DM_DTYPE)
    cnt = C.c_int(0)
    ma = mb = None
    if maskA is not None:
        maskA = np.ascontiguousarray(maskA, np.uint8); ma = maskA.ctypes.data_as(C.c_void_p)
    if maskB is not None:
        maskB = np.ascontiguousarray(maskB, np.uint8); mb = maskB.ctypes.data_as(C.c_void_p)
    lib().orc_match(ap, nA, ma, bp, nB, mb, int(max_hamming), int(min_diff), out.ctypes.data_as(C.c_void_p), C.byref(cnt))
    return out[:cnt.value].copy()


def indexed_match(descA, descB, a2b, b2a, max_hamming=30, min_diff=1, maskA=None, maskB=None):
    """a2b / b2a: CSR pairs (offsets int32[n+1], candidates int32[...])."""
    descA, ap = _u8(descA); descB, bp = _u8(descB)
    nA, nB = len(descA), len(descB)
    o0, c0 = (np.ascontiguousarray(x, np.int32) for x in a2b)
    o1, c1 = (np.ascontiguousarray(x, np.int32) for x in b2a)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    maskA = None if maskA is None else np.ascontiguousarray(maskA, np.uint8)
    maskB = None if maskB is None else np.ascontiguousarray(maskB, np.uint8)
    out = np.zeros(max(nA, 1), DM_DTYPE)
    L = lib()
    L.orc_indexed_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    n = L.orc_indexed_match(ap, nA, p(maskA), bp, nB, p(maskB), p(o0), p(c0), p(o1), p(c1), int(max_hamming), int(min_diff), p(out))
    return out[:n].copy()


def descriptor_distance(a, b):
    a, ap = _u8(a); b, bp = _u8(b)
    return lib().orc_descriptor_distance(ap, bp)


def rtree_order(kps):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    order = np.zeros(len(kps), np.int32)
    n = lib().orc_rtree_order(kps.ctypes.data_as(C.c_void_p), len(kps), order.ctypes.data_as(C.c_void_p))
    return order[:n]


def radius_match(qk, qdesc, tk, tdesc, radius, max_hamming, min_diff, qpos=None, qmask=None, tmask=None):
    qk = np.ascontiguousarray(qk, KP_DTYPE); tk = np.ascontiguousarray(tk, KP_DTYPE)
    qdesc, qdp = _u8(qdesc); tdesc, tdp = _u8(tdesc)
    out = np.zeros(max(len(qk), 1), DM_DTYPE)
    vp = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
    qpos, qmask, tmask = vp(qpos, np.float32), vp(qmask, np.uint8), vp(tmask, np.uint8)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    L = lib()
    L.orc_radius_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_float, C.c_int, C.c_int, C.c_void_p]
    n = L.orc_radius_match(p(qk), len(qk), p(qpos), p(qmask), qdp, p(tk), len(tk), p(tmask), tdp, float(radius), int(max_hamming), int(min_diff), p(out))
    return out[:n].copy()


_RREF = None


def radius_ref():
    """The RadiusMatch oracle on the REAL boost R*-tree (oracle/_ref/libradius_ref.so); None when it was not built."""
    global _RREF
    path = os.path.join(ROOT, "oracle", "_ref", "libradius_ref.so")
    if _RREF is None and os.path.exists(path):
        R = C.CDLL(path)
        R.rmref_index_create.restype = C.c_void_p
        R.rmref_index_create.argtypes = [C.c_void_p, C.c_int]
        R.rmref_index_destroy.argtypes = [C.c_void_p]
        R.rmref_query.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_int]
        R.rmref_radius_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_float, C.c_int, C.c_int, C.c_void_p]
        _RREF = R
    return _RREF


def radius_match_ref(qk, qdesc, tk, tdesc, radius, max_hamming, min_diff, qpos=None, qmask=None, tmask=None):
    R = radius_ref()
    qk = np.ascontiguousarray(qk, KP_DTYPE); tk = np.ascontiguousarray(tk, KP_DTYPE)
    qdesc, qdp = _u8(qdesc); tdesc, tdp = _u8(tdesc)
    vp = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
    qpos, qmask, tmask = vp(qpos, np.float32), vp(qmask, np.uint8), vp(tmask, np.uint8)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    idx = R.rmref_index_create(p(tk), len(tk))
    out = np.zeros(max(len(qk), 1), DM_DTYPE)
    n = R.rmref_radius_match(idx, p(qk), len(qk), p(qpos), p(qmask), qdp, len(tk), p(tmask), tdp, float(radius), int(max_hamming), int(min_diff), p(out))
    R.rmref_index_destroy(idx)
    return out[:n].copy()
