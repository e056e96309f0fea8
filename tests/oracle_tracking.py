"""ctypes binding of oracle/libtracking_oracle.so and oracle/_ref/libtracking_ref.so (TEST INFRASTRUCTURE, never the product path)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAP_POINT_DTYPE = np.dtype([("position", "<f4", 3), ("mean_view_dir", "<f4", 3), ("dmin", "<f4"), ("dmax", "<f4")])
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
_LIB = None
_REF = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "libtracking_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "libtracking_oracle.so"], check=True)
        L = C.CDLL(path)
        L.trk_compute_octave.argtypes = [C.c_float, C.c_float, C.c_float]
        L.trk_octave_real.argtypes = [C.c_float, C.c_float, C.c_float]
        L.trk_octave_real.restype = C.c_float
        L.trk_project_map_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.trk_project_map_points.restype = None
        _LIB = L
    return _LIB


def ref():
    """The reference's own Map/MappingMath.h compiled into oracle/_ref/libtracking_ref.so; None when it was not built."""
    global _REF
    path = os.path.join(ROOT, "oracle", "_ref", "libtracking_ref.so")
    if _REF is None and os.path.exists(path):
        R = C.CDLL(path)
        R.trkref_compute_octave.argtypes = [C.c_float, C.c_float, C.c_float]
        R.trkref_compute_dmax.argtypes = [C.c_float, C.c_int, C.c_int, C.c_float]
        R.trkref_compute_dmax.restype = C.c_float
        R.trkref_compute_dmin.argtypes = [C.c_float, C.c_int, C.c_float]
        R.trkref_compute_dmin.restype = C.c_float
        _REF = R
    return _REF


def project_map_points(params, pts):
    """params: any ctypes struct with the mage_projection_params layout."""
    pts = np.ascontiguousarray(pts, MAP_POINT_DTYPE)
    n = len(pts)
    kps = np.zeros(n, KP_DTYPE); depth = np.zeros(n, np.float32); flags = np.zeros(n, np.uint8)
    lib().trk_project_map_points(C.byref(params), pts.ctypes.data_as(C.c_void_p), n, kps.ctypes.data_as(C.c_void_p),
                                 depth.ctypes.data_as(C.c_void_p), flags.ctypes.data_as(C.c_void_p))
    return kps, depth, flags


def octave_real(distance, dmin, scale):
    return lib().trk_octave_real(float(distance), float(dmin), float(scale))
