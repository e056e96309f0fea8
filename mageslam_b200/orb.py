"""Host-side mirror of the reference's ORB plugin surface over the C ABI (include/mage_b200.h).

  OrbDetector            <- reference Core/MAGESLAM/Source/Image/OpenCVModified.h:64-172 (ctor + DetectAndCompute)
  FeatureExtractorSettings / OrbFeatureDetector
                         <- reference MageSettings.h:151-167 and Image/OrbFeatureDetector.h:20-50
Same argument names, order and meaning; status codes surface as MageError. All arithmetic runs in the CUDA library.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import KEYPOINT_DTYPE, OrbParams, check, lib, ptr, stream_ptr


@dataclass
class FeatureExtractorSettings:
    """Defaults = reference MageSettings.h:151-167."""
    GaussianKernelSize: int = 7
    NumFeatures: int = 440
    ScaleFactor: float = 1.5
    NumLevels: int = 1
    PatchSize: int = 15
    FastThreshold: int = 4
    UseOrientation: bool = False
    FeatureFactor: float = 1.5
    FeatureStrength: float = 0.9
    StrongResponse: int = 20
    MinRobustnessFactor: float = 1.1
    MaxRobustnessFactor: float = 2.0
    NumCellsX: int = 32
    NumCellsY: int = 32

    @staticmethod
    def tier(num_features=2000, num_levels=8, scale_factor=1.2, fast_threshold=10):
        """2000 features, 8 levels x 1.2, patch 31, oriented, FAST threshold 10."""
        return FeatureExtractorSettings(7, num_features, scale_factor, num_levels, 31, fast_threshold, True)


class OrbDetector:
    """Mirror of class OrbDetector (reference OpenCVModified.h:64-172). The handle is sized for one image geometry;
    it is (re)created lazily when DetectAndCompute sees a different width x height."""

    def __init__(self, gaussianKernelSize, nfeatures, scaleFactor, nlevels, patchSize, fastThreshold, useOrientation,
                 featureFactorANMS, featureStrengthANMS, strongResponseANMS, minRobustFactor, maxRobustFactor,
                 numCellsX, numCellsY, max_batch=1):
        self.params = OrbParams(int(gaussianKernelSize), int(nfeatures), float(scaleFactor), int(nlevels), int(patchSize),
                                int(fastThreshold), 1 if useOrientation else 0, float(featureFactorANMS),
                                float(featureStrengthANMS), int(strongResponseANMS), float(minRobustFactor),
                                float(maxRobustFactor), int(numCellsX), int(numCellsY))
        self.max_batch = int(max_batch)
        self._h = None
        self._wh = None
        self._blur_mode = 0

    # -- handle management
    def _ensure(self, w, h):
        if self._h is not None and self._wh == (w, h):
            return
        self.close()
        hnd = C.c_void_p()
        check(lib().mage_orb_create(C.byref(self.params), w, h, self.max_batch, C.byref(hnd)))
        self._h, self._wh = hnd, (w, h)
        if self._blur_mode:
            check(lib().mage_orb_set_blur_mode(self._h, self._blur_mode))

    def SetBlurMode(self, mode):
        """mage_orb_set_blur_mode: 0 auto (OpenCV 4.13 behaviour), 1 float fused, 2 float unfused, 3 fixed point"""
        self._blur_mode = int(mode)
        if self._h is not None:
            check(lib().mage_orb_set_blur_mode(self._h, self._blur_mode))

    def close(self):
        if self._h is not None:
            lib().mage_orb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference surface
    def DetectAndCompute(self, image, capacity=None):
        """image: uint8 [h, w] (CV_8UC1). Returns (keypoints[KEYPOINT_DTYPE], descriptors uint8 [n, 32]).
        capacity = ImageData::maxFeatures (defaults to nfeatures, reference MAGESlam.cpp:81-87)."""
        if image.dtype != np.uint8 or image.ndim != 2:
            raise TypeError("image must be CV_8UC1 (uint8, 2-D)")      # CV_Assert(srcImage.type() == CV_8UC1)
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        self._ensure(w, h)
        cap = int(capacity if capacity is not None else self.params.nfeatures)
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        cnt = C.c_int(0)
        check(lib().mage_orb_detect_and_compute(self._h, C.c_void_p(image.ctypes.data), w, h, image.strides[0], ptr(kps), ptr(desc),
                                                cap, C.byref(cnt), None))
        return kps[:cnt.value], desc[:cnt.value]

    def DetectAndComputeBatch(self, images, capacity=None, out=None):
        """images: uint8 [n, h, w] host array. Returns (kps [n, cap], desc [n, cap, 32], counts [n])."""
        n, h, w = images.shape
        assert n <= self.max_batch and images.dtype == np.uint8 and images.flags["C_CONTIGUOUS"]
        self._ensure(w, h)
        cap = int(capacity if capacity is not None else self.params.nfeatures)
        if out is None:
            out = (np.zeros((n, cap), KEYPOINT_DTYPE), np.zeros((n, cap, 32), np.uint8), np.zeros(n, np.int32))
        kps, desc, counts = out
        check(lib().mage_orb_detect_and_compute_batch(self._h, ptr(images), n, w, h, w, w * h, ptr(kps), ptr(desc), cap, ptr(counts), None))
        return kps, desc, counts

    def ExtractDevice(self, d_images, d_kps, d_desc, d_counts, capacity, stream=None):
        """Device-resident variant: torch uint8 [n, h, w] in, torch buffers out (kps as uint8 [n, cap, 28]). Asynchronous."""
        n, h, w = d_images.shape
        self._ensure(w, h)
        check(lib().mage_orb_extract_device(self._h, ptr(d_images), n, w, h, d_images.stride(1), d_images.stride(0), ptr(d_kps),
                                            ptr(d_desc), int(capacity), ptr(d_counts), stream_ptr(stream)))

    def LevelInfo(self, w, h):
        self._ensure(w, h)
        n = self.params.nlevels
        ws = np.zeros(n, np.int32); hs = np.zeros(n, np.int32); sc = np.zeros(n, np.float32); nf = np.zeros(n, np.int32)
        check(lib().mage_orb_level_info(self._h, ptr(ws), ptr(hs), ptr(sc), ptr(nf)))
        return ws, hs, sc, nf

    # -- inspection taps for the stage parity tests
    def DebugLevel(self, frame, level, blurred):
        ws, hs, _, _ = self.LevelInfo(*self._wh)
        out = np.zeros((int(hs[level]), int(ws[level])), np.uint8)
        check(lib().mage_orb_debug_get_level(self._h, frame, level, 1 if blurred else 0, ptr(out)))
        return out

    def DebugCandidates(self, frame, level):
        ws, hs, _, _ = self.LevelInfo(*self._wh)
        cap = ((int(ws[level]) + 1) // 2) * ((int(hs[level]) + 1) // 2)
        out = np.zeros(cap, np.uint32)
        cnt = C.c_int(0)
        check(lib().mage_orb_debug_get_candidates(self._h, frame, level, ptr(out), cap, C.byref(cnt)))
        return out[:cnt.value]


class CameraCalibration:
    """The part of mage::CameraCalibration the detector reads (reference Device/CameraCalibration.h:44-71): the 3x3 camera
    matrix and the OpenCV-ordered distortion coefficients k1 k2 p1 p2 k3 [k4 k5 k6] (none = DistortionType::None)."""

    def __init__(self, fx, fy, cx, cy, dist_coeffs=()):
        self.camera_matrix = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32)
        self.dist_coeffs = np.asarray(dist_coeffs, np.float32).ravel()
        assert len(self.dist_coeffs) in (0, 4, 5, 8)

    def __eq__(self, other):          # reference CameraCalibration::operator== (.cpp:95-108)
        return (isinstance(other, CameraCalibration) and np.array_equal(self.camera_matrix, other.camera_matrix)
                and np.array_equal(self.dist_coeffs, other.dist_coeffs))

    def __ne__(self, other):
        return not self == other

    def c_struct(self):
        from ._lib import CameraCalibrationC
        c = CameraCalibrationC()
        for i, v in enumerate(self.camera_matrix.ravel()):
            c.camera_matrix[i] = float(v)
        for i, v in enumerate(self.dist_coeffs):
            c.dist_coeffs[i] = float(v)
        c.n_dist_coeffs = len(self.dist_coeffs)
        return c


def UndistortKeypoints(inoutKeypoints, distortedCalibration, undistortedCalibration):
    """Mirror of OrbFeatureDetector::UndistortKeypoints (reference Image/OrbFeatureDetector.cpp:30-62): in place on a
    KEYPOINT_DTYPE array."""
    assert inoutKeypoints.dtype == KEYPOINT_DTYPE and inoutKeypoints.flags["C_CONTIGUOUS"]
    d, u = distortedCalibration.c_struct(), undistortedCalibration.c_struct()
    check(lib().mage_undistort_keypoints(ptr(inoutKeypoints), len(inoutKeypoints), C.byref(d), C.byref(u), None))
    return inoutKeypoints


class OrbFeatureDetector:
    """Mirror of mage::OrbFeatureDetector (reference Image/OrbFeatureDetector.cpp:64-100). Process() runs
    DetectAndCompute and, when the distorted and undistorted calibrations differ, UndistortKeypoints (:97-99)."""

    def __init__(self, settings: FeatureExtractorSettings, max_batch=1):
        s = settings
        self.settings = s
        self.m_detector = OrbDetector(s.GaussianKernelSize, s.NumFeatures, s.ScaleFactor, s.NumLevels, s.PatchSize, s.FastThreshold,
                                      s.UseOrientation, s.FeatureFactor, s.FeatureStrength, s.StrongResponse, s.MinRobustnessFactor,
                                      s.MaxRobustnessFactor, s.NumCellsX, s.NumCellsY, max_batch=max_batch)

    def Process(self, image, distortedCalibration=None, undistortedCalibration=None):
        kps, desc = self.m_detector.DetectAndCompute(image, capacity=self.settings.NumFeatures)
        if distortedCalibration is not None and distortedCalibration != undistortedCalibration:
            kps = UndistortKeypoints(np.ascontiguousarray(kps), distortedCalibration, undistortedCalibration)
        return kps, desc
