"""Loader for libmage_b200.so (the C ABI of include/mage_b200.h).

The CUDA library is the product: there is no CPU fallback. If the shared object is missing this module raises at import
of the symbol table, and every compute entry point returns MAGE_ERR_CUDA when no device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmage_b200.so")

MAGE_OK = 0
MAGE_ERR_INVALID = -1
MAGE_ERR_UNSUPPORTED = -2
MAGE_ERR_CUDA = -3
MAGE_ERR_OVERFLOW = -4

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4")])      # cv::KeyPoint, 28 bytes
DMATCH_DTYPE = np.dtype([("query_idx", "<i4"), ("train_idx", "<i4"), ("distance", "<f4")])
assert KEYPOINT_DTYPE.itemsize == 28 and DMATCH_DTYPE.itemsize == 12


class MageError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mage_b200 error %d: %s" % (code, msg))
        self.code = code


class OrbParams(C.Structure):
    """mage_orb_params: the 14 OrbDetector ctor scalars (reference Image/OpenCVModified.h:68-82)."""
    _fields_ = [("gaussian_kernel_size", C.c_uint32), ("nfeatures", C.c_uint32), ("scale_factor", C.c_float),
                ("nlevels", C.c_uint32), ("patch_size", C.c_uint32), ("fast_threshold", C.c_uint32),
                ("use_orientation", C.c_int32), ("feature_factor", C.c_float), ("feature_strength", C.c_float),
                ("strong_response", C.c_int32), ("min_robust_factor", C.c_float), ("max_robust_factor", C.c_float),
                ("num_cells_x", C.c_int32), ("num_cells_y", C.c_int32)]


class CameraCalibrationC(C.Structure):
    """mage_camera_calibration: GetCameraMatrix() (row-major 3x3) + GetCVDistortionCoeffs() (reference Device/CameraCalibration.h:44-71)."""
    _fields_ = [("camera_matrix", C.c_float * 9), ("dist_coeffs", C.c_float * 8), ("n_dist_coeffs", C.c_int32)]


_lib = None


def lib():
    """Returns the loaded CDLL; raises if the CUDA extension has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("mageslam_b200: %s is missing -- build the CUDA extension first (make, or __graft_entry__.build()). "
                          "There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.mage_last_error.restype = C.c_char_p
    L.mage_version.restype = C.c_char_p
    vp, ci, cf, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    L.mage_orb_create.argtypes = [C.POINTER(OrbParams), ci, ci, ci, C.POINTER(vp)]
    L.mage_orb_destroy.argtypes = [vp]
    L.mage_orb_destroy.restype = None
    L.mage_orb_detect_and_compute.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci, C.POINTER(ci), vp]
    L.mage_orb_detect_and_compute_batch.argtypes = [vp, vp, ci, ci, ci, ci, sz, vp, vp, ci, vp, vp]
    L.mage_orb_extract_device.argtypes = [vp, vp, ci, ci, ci, ci, sz, vp, vp, ci, vp, vp]
    L.mage_orb_level_info.argtypes = [vp, vp, vp, vp, vp]
    L.mage_dense_debug_solve.argtypes = [ci, vp, vp, vp, vp, vp, vp]
    L.mage_orb_set_blur_mode.argtypes = [vp, ci]
    L.mage_orb_debug_get_level.argtypes = [vp, ci, ci, ci, vp]
    L.mage_orb_debug_get_candidates.argtypes = [vp, ci, ci, vp, ci, C.POINTER(ci)]
    L.mage_matcher_create.argtypes = [ci, ci, C.POINTER(vp)]
    L.mage_matcher_destroy.argtypes = [vp]
    L.mage_matcher_destroy.restype = None
    L.mage_match_bf.argtypes = [vp, vp, ci, vp, vp, ci, vp, ci, ci, vp, C.POINTER(ci), vp]
    L.mage_match_bf_device.argtypes = [vp, vp, vp, sz, vp, vp, ci, ci, ci, vp, ci, vp, vp]
    L.mage_descriptor_distance_device.argtypes = [vp, vp, ci, vp, vp]
    L.mage_indexed_match.argtypes = [vp, vp, ci, vp, vp, ci, vp, vp, vp, vp, vp, ci, ci, vp, C.POINTER(ci), vp]
    L.mage_spatial_index_create.argtypes = [vp, ci, C.POINTER(vp)]
    L.mage_spatial_index_destroy.argtypes = [vp]
    L.mage_spatial_index_destroy.restype = None
    L.mage_spatial_index_rank.argtypes = [vp, vp]
    L.mage_radius_match.argtypes = [vp, vp, ci, vp, vp, vp, vp, vp, cf, ci, ci, vp, C.POINTER(ci), vp]
    L.mage_project_map_points.argtypes = [vp, vp, ci, vp, vp, vp, vp]
    L.mage_optimize_camera_pose.argtypes = [vp, vp, vp, ci, vp, vp, vp, ci, cf, cf, vp, vp, vp, ci, C.POINTER(ci), C.POINTER(cf)]
    L.mage_project_map_points_device.argtypes = [vp, vp, ci, vp, vp, vp, vp]
    L.mage_undistort_keypoints.argtypes = [vp, ci, vp, vp, vp]
    L.mage_undistort_keypoints_device.argtypes = [vp, vp, ci, ci, vp, vp, vp]
    if hasattr(L, "mage_ba_create"):
        L.mage_ba_create.argtypes = [ci, C.POINTER(vp)]
        L.mage_ba_destroy.argtypes = [vp]
        L.mage_ba_destroy.restype = None
        for name in ("mage_ba_alloc_cameras", "mage_ba_alloc_points", "mage_ba_alloc_observations"):
            getattr(L, name).argtypes = [vp, ci]
        L.mage_ba_set_camera.argtypes = [vp, ci, vp, vp, vp, ci]
        L.mage_ba_fix_camera.argtypes = [vp, ci, ci]
        L.mage_ba_set_point.argtypes = [vp, ci, vp]
        L.mage_ba_set_observation.argtypes = [vp, ci, vp, ci, ci, cf]
        L.mage_ba_set_cameras_bulk.argtypes = [vp, ci, vp, vp, vp, vp]
        L.mage_ba_set_points_bulk.argtypes = [vp, ci, vp]
        L.mage_ba_set_observations_bulk.argtypes = [vp, ci, vp, vp, vp, vp]
        L.mage_ba_set_lambda.argtypes = [vp, cf]
        for name in ("mage_ba_alloc_fixed_distance_constraints", "mage_ba_alloc_relative_rotation_constraints", "mage_ba_alloc_relative_transform_constraints"):
            getattr(L, name).argtypes = [vp, ci]
        L.mage_ba_set_fixed_distance_constraint.argtypes = [vp, ci, ci, ci, cf, cf]
        L.mage_ba_set_relative_rotation_constraint.argtypes = [vp, ci, ci, ci, vp, cf]
        L.mage_ba_set_relative_transform_constraint.argtypes = [vp, ci, ci, ci, vp, vp, cf]
        L.mage_ba_get_lambda.argtypes = [vp, C.POINTER(cf)]
        L.mage_ba_step.argtypes = [vp, vp, ci, cf, vp, ci, C.POINTER(ci), C.POINTER(cf)]
        L.mage_ba_get_pose.argtypes = [vp, ci, vp, vp]
        L.mage_ba_get_point.argtypes = [vp, ci, vp]
        L.mage_ba_get_poses_bulk.argtypes = [vp, vp, vp]
        L.mage_ba_get_points_bulk.argtypes = [vp, vp]
        L.mage_ba_get_state_f64.argtypes = [vp, vp, vp]
        L.mage_ba_get_stats.argtypes = [vp, vp]
        L.mage_ba_step_many.argtypes = [vp, ci, vp, ci, cf, vp]
        L.mage_ba_last_outliers.argtypes = [vp, vp, ci, C.POINTER(ci)]
        L.mage_ba_last_outlier_counts.argtypes = [vp, ci, vp]
    _lib = L
    return L


def check(rc):
    if rc != MAGE_OK:
        raise MageError(rc, lib().mage_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a numpy array (must be C-contiguous) or a torch tensor, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


def stream_ptr(stream=None):
    if stream is None:
        return None
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))
