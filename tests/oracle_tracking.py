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
        L.trk_undistort_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.trk_undistort_keypoints.restype = None
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


class Calibration(C.Structure):
    """trk_calibration == mage_camera_calibration"""
    _fields_ = [("camera_matrix", C.c_float * 9), ("dist_coeffs", C.c_float * 8), ("n_dist_coeffs", C.c_int32)]


def calibration(K, dist=()):
    c = Calibration()
    for i, v in enumerate(np.asarray(K, np.float32).ravel()):
        c.camera_matrix[i] = float(v)
    for i, v in enumerate(np.asarray(dist, np.float32).ravel()):
        c.dist_coeffs[i] = float(v)
    c.n_dist_coeffs = len(dist)
    return c


def undistort_keypoints(kps, K_dist, dist, K_undist):
    """oracle of OrbFeatureDetector::UndistortKeypoints; returns a modified copy"""
    out = np.ascontiguousarray(kps, KP_DTYPE).copy()
    d, u = calibration(K_dist, dist), calibration(K_undist)
    lib().trk_undistort_keypoints(out.ctypes.data_as(C.c_void_p), len(out), C.byref(d), C.byref(u))
    return out


def undistort_cases(seed=0, n=1500):
    """(name, keypoints, K_dist, dist, K_undist): Poly3k / Rational6k, mild and strong (icdist < 0 escape) distortion, identity P"""
    rng = np.random.default_rng(seed)
    cases = []
    for t, (name, nco, scale) in enumerate([("poly3k", 5, 1.0), ("rational6k", 8, 1.0), ("poly3k_strong", 5, 8.0), ("rational6k_strong", 8, 8.0),
                                             ("radial_only", 4, 1.0), ("no_distortion", 0, 1.0)]):
        fx, fy = rng.uniform(300, 700, 2).astype(np.float32)
        cx, cy = np.float32(320 + rng.normal(0, 10)), np.float32(240 + rng.normal(0, 10))
        Kd = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
        Ku = np.array([[fx * 0.95, 0, cx + 2], [0, fy * 0.97, cy - 1], [0, 0, 1]], np.float32)
        D = (rng.normal(0, 1, nco) * np.array([0.2, 0.1, 0.003, 0.003, 0.05, 0.1, 0.05, 0.02][:nco]) * scale).astype(np.float32)
        kps = np.zeros(n, KP_DTYPE)
        kps["x"] = rng.uniform(0, 640, n).astype(np.float32); kps["y"] = rng.uniform(0, 480, n).astype(np.float32)
        kps["size"] = 31; kps["angle"] = rng.uniform(0, 360, n).astype(np.float32); kps["response"] = rng.integers(10, 200, n)
        kps["octave"] = rng.integers(0, 8, n); kps["class_id"] = -1
        cases.append((name, kps, Kd, D, Ku))
    return cases
