"""Host-side mirror of mage::BundlerLib (reference Dependencies/BundlerLib/Include/BundlerLib.h:20-66) over the C ABI.

Same method names, argument meaning and call protocol (allocate once, set, step many times, read back); the arithmetic
runs in the CUDA library (mageslam_b200/csrc/ba.cu). Eigen::Map arguments become numpy float32 arrays.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import check, lib, ptr


@dataclass
class BundlerParameters:
    ArePointsFixed: bool = False        # True if map points should not be optimized


def _f32(a, n=None):
    a = np.ascontiguousarray(a, np.float32)
    if n is not None:
        assert a.size == n, "expected %d floats" % n
    return a


class BundlerLib:
    def __init__(self, bundlerParameters: BundlerParameters = BundlerParameters()):
        self._h = C.c_void_p()
        check(lib().mage_ba_create(1 if bundlerParameters.ArePointsFixed else 0, C.byref(self._h)))
        self._K = self._P = self._E = 0

    def close(self):
        if self._h:
            lib().mage_ba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference surface --------------------------------------------------------------------------------------------
    def AllocateCameras(self, count):
        check(lib().mage_ba_alloc_cameras(self._h, int(count))); self._K = int(count)

    def SetCameraPose(self, idx, position, orientation, intrinsics, isFixed):
        """position[3]; orientation 3x3 column-major (Eigen::Matrix3f storage) = world->camera rotation; intrinsics (cx, cy, fx, fy)."""
        check(lib().mage_ba_set_camera(self._h, int(idx), ptr(_f32(position, 3)), ptr(_f32(orientation, 9)), ptr(_f32(intrinsics, 4)),
                                       1 if isFixed else 0))

    def FixCameraPose(self, idx, value):
        check(lib().mage_ba_fix_camera(self._h, int(idx), 1 if value else 0))

    def AllocateMapPoints(self, count):
        check(lib().mage_ba_alloc_points(self._h, int(count))); self._P = int(count)

    def SetMapPoint(self, idx, point):
        check(lib().mage_ba_set_point(self._h, int(idx), ptr(_f32(point, 3))))

    def AllocateObservations(self, count):
        check(lib().mage_ba_alloc_observations(self._h, int(count))); self._E = int(count)

    def SetObservation(self, idx, position, cameraIndex, mapPointIndex, informationMatrixScalar):
        check(lib().mage_ba_set_observation(self._h, int(idx), ptr(_f32(position, 2)), int(cameraIndex), int(mapPointIndex),
                                            float(informationMatrixScalar)))

    # tether constraints between two cameras (reference BundlerLib.h:41-48); deltaRotation = (x, y, z, w)
    def AllocateFixedDistanceConstraints(self, count):
        check(lib().mage_ba_alloc_fixed_distance_constraints(self._h, int(count)))

    def SetFixedDistanceConstraint(self, idx, cameraIndex1, cameraIndex2, distance=1.0, weight=1.0):
        check(lib().mage_ba_set_fixed_distance_constraint(self._h, int(idx), int(cameraIndex1), int(cameraIndex2), float(distance), float(weight)))

    def AllocateRelativeRotationConstraints(self, count):
        check(lib().mage_ba_alloc_relative_rotation_constraints(self._h, int(count)))

    def SetRelativeRotationConstraint(self, idx, cameraIndex1, cameraIndex2, deltaRotation, weight=1.0):
        check(lib().mage_ba_set_relative_rotation_constraint(self._h, int(idx), int(cameraIndex1), int(cameraIndex2), ptr(_f32(deltaRotation, 4)),
                                                             float(weight)))

    def AllocateRelativeTransformConstraints(self, count):
        check(lib().mage_ba_alloc_relative_transform_constraints(self._h, int(count)))

    def SetRelativeTransformConstraint(self, idx, cameraIndex1, cameraIndex2, deltaPosition, deltaRotation, weight):
        check(lib().mage_ba_set_relative_transform_constraint(self._h, int(idx), int(cameraIndex1), int(cameraIndex2), ptr(_f32(deltaPosition, 3)),
                                                              ptr(_f32(deltaRotation, 4)), float(weight)))

    def SetCurrentLambda(self, userLambda):
        check(lib().mage_ba_set_lambda(self._h, float(userLambda)))

    def GetCurrentLambda(self):
        v = C.c_float(0)
        check(lib().mage_ba_get_lambda(self._h, C.byref(v)))
        return float(v.value)

    def StepBundleAdjustment(self, huberWidthPerIteration, maxErrorSquare, outliers=None):
        """Runs one solver iteration per Huber width; appends the indices of removed observations to `outliers`
        (a list, like the reference's std::vector<unsigned>&) and returns the average squared error of the inliers."""
        hub = _f32(huberWidthPerIteration)
        buf = np.zeros(max(self._E, 1), np.uint32)
        n = C.c_int(0); mean = C.c_float(0)
        check(lib().mage_ba_step(self._h, ptr(hub) if len(hub) else None, len(hub), float(maxErrorSquare), ptr(buf), len(buf),
                                 C.byref(n), C.byref(mean)))
        if outliers is not None:
            outliers.extend(int(v) for v in buf[:n.value])
        self.last_outliers = buf[:n.value].copy()
        return float(mean.value)

    def GetPose(self, idx):
        pos = np.zeros(3, np.float32); rot = np.zeros(9, np.float32)
        check(lib().mage_ba_get_pose(self._h, int(idx), ptr(pos), ptr(rot)))
        return pos, rot

    def GetPoint(self, idx):
        xyz = np.zeros(3, np.float32)
        check(lib().mage_ba_get_point(self._h, int(idx), ptr(xyz)))
        return xyz

    # -- bulk conveniences (avoid thousands of ABI crossings) ------------------------------------------------------------
    def load(self, prob):
        """prob: dict from mageslam_b200.synth.ba_problem (same order of calls as BuildDataForG2O, BundleAdjust.cpp:25-193)."""
        K, P, E = len(prob["cam_pos"]), len(prob["points"]), len(prob["obs_uv"])
        self.AllocateCameras(K); self.AllocateMapPoints(P); self.AllocateObservations(E)
        check(lib().mage_ba_set_cameras_bulk(self._h, K, ptr(_f32(prob["cam_pos"])), ptr(_f32(prob["cam_rot"])), ptr(_f32(prob["intrinsics"])),
                                             ptr(np.ascontiguousarray(prob["fixed"], np.int32))))
        check(lib().mage_ba_set_points_bulk(self._h, P, ptr(_f32(prob["points"]))))
        check(lib().mage_ba_set_observations_bulk(self._h, E, ptr(_f32(prob["obs_uv"])), ptr(np.ascontiguousarray(prob["obs_cam"], np.int32)),
                                                  ptr(np.ascontiguousarray(prob["obs_pt"], np.int32)), ptr(_f32(prob["obs_info"]))))
        dist, rot, xf = prob.get("tether_distance", []), prob.get("tether_rotation", []), prob.get("tether_transform", [])
        if len(dist) or len(rot) or len(xf):               # synth.ba_add_tethers, same call order as the reference's setters
            self.AllocateFixedDistanceConstraints(len(dist)); self.AllocateRelativeRotationConstraints(len(rot))
            self.AllocateRelativeTransformConstraints(len(xf))
            for i, (c1, c2, d, w) in enumerate(dist):
                self.SetFixedDistanceConstraint(i, c1, c2, d, w)
            for i, (c1, c2, q, w) in enumerate(rot):
                self.SetRelativeRotationConstraint(i, c1, c2, q, w)
            for i, (c1, c2, t, q, w) in enumerate(xf):
                self.SetRelativeTransformConstraint(i, c1, c2, t, q, w)
        return self

    def poses(self):
        pos = np.zeros((self._K, 3), np.float32); rot = np.zeros((self._K, 9), np.float32)
        check(lib().mage_ba_get_poses_bulk(self._h, ptr(pos), ptr(rot)))
        return pos, rot

    def points(self):
        pts = np.zeros((self._P, 3), np.float32)
        check(lib().mage_ba_get_points_bulk(self._h, ptr(pts)))
        return pts

    def state_f64(self):
        cams = np.zeros((self._K, 7)); pts = np.zeros((self._P, 3))
        check(lib().mage_ba_get_state_f64(self._h, ptr(cams), ptr(pts)))
        return cams, pts

    def stats(self):
        s = np.zeros(4, np.int64)
        check(lib().mage_ba_get_stats(self._h, ptr(s)))
        return dict(lm_iterations=int(s[0]), lambda_trials=int(s[1]), kernel_launches=int(s[2]), structure_builds=int(s[3]))


def StepMany(bundlers, huberWidthPerIteration, maxErrorSquare):
    """Steps many independent problems with one kernel launch (one CTA per problem). Returns the mean squared errors."""
    hub = _f32(huberWidthPerIteration)
    arr = (C.c_void_p * len(bundlers))(*[b._h for b in bundlers])
    means = np.zeros(len(bundlers), np.float32)
    check(lib().mage_ba_step_many(arr, len(bundlers), ptr(hub), len(hub), float(maxErrorSquare), ptr(means)))
    counts = np.zeros(len(bundlers), np.int32)      # the `outliers` vector of each problem's StepBundleAdjustment
    check(lib().mage_ba_last_outlier_counts(arr, len(bundlers), ptr(counts)))
    empty = np.zeros(0, np.uint32)
    for b, cnt in zip(bundlers, counts):
        if cnt == 0:
            b.last_outliers = empty
            continue
        buf = np.zeros(int(cnt), np.uint32); n = C.c_int(0)
        check(lib().mage_ba_last_outliers(b._h, ptr(buf), int(cnt), C.byref(n)))
        b.last_outliers = buf
    return means
