"""Host-side mirror of the reference's map-point projection step over the C ABI (include/mage_b200.h).

  ProjectMapPoints  <- the prologue of TrackLocalMap::ProjectMapPointIntoCurrentFrame
                       (reference Core/MAGESLAM/Source/Tracking/TrackLocalMap.cpp:325-366): ProjectUndistorted
                       (Tracking/Reprojection.cpp:26-43), IsGoodCandidate (TrackLocalMap.cpp:519-554), ComputeOctave
                       (Map/MappingMath.h:13-16), for a whole local map per call.
  ProjectPoints     <- reference Tracking/Reprojection.cpp:16-24.
  OptimizeCameraPose <- TrackLocalMap::OptimizeCameraPose (TrackLocalMap.cpp:421-501): the pose-only bundle adjustment of the
                       tracking thread in one call (mage_optimize_camera_pose).
The returned keypoints are the `mapPointKp` the reference hands to RadiusMatch (TrackLocalMap.cpp:371), so they go straight
into matcher.RadiusMatch as queries (mask = predicted). All arithmetic runs in the CUDA library.
"""
import ctypes as C
import math

import numpy as np

from ._lib import KEYPOINT_DTYPE, check, lib, ptr, stream_ptr

MAP_POINT_DTYPE = np.dtype([("position", "<f4", 3), ("mean_view_dir", "<f4", 3), ("dmin", "<f4"), ("dmax", "<f4")])
assert MAP_POINT_DTYPE.itemsize == 32

PROJ_GOOD_CANDIDATE = 1
PROJ_PREDICTED = 2


class ProjectionParams(C.Structure):
    """mage_projection_params."""
    _fields_ = [("view", C.c_float * 12), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("frame_position", C.c_float * 3), ("frame_forward", C.c_float * 3), ("min_cos_view_angle", C.c_float),
                ("image_border", C.c_float), ("width", C.c_uint32), ("height", C.c_uint32), ("pyramid_scale", C.c_float),
                ("num_levels", C.c_uint32)]


def make_params(viewMatrix, cameraCalibrationMatrix, framePosition, frameForward, minDegreesBetweenCurrentViewAndMapPointView,
                imageBorder, width, height, pyramidScale, numLevels):
    """Argument names follow TrackLocalMap::ProjectMapPointIntoCurrentFrame (reference TrackLocalMap.cpp:325-340)."""
    v = np.asarray(viewMatrix, np.float32).reshape(3, 4)
    K = np.asarray(cameraCalibrationMatrix, np.float32).reshape(3, 3)
    p = ProjectionParams()
    p.view[:] = v.ravel().tolist()
    p.fx, p.fy, p.cx, p.cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    p.frame_position[:] = [float(x) for x in framePosition]
    p.frame_forward[:] = [float(x) for x in frameForward]
    # std::cos(mira::deg2rad(angle)) in float (reference TrackLocalMap.cpp:541, arcana/math.h:85-90)
    deg2rad = np.float32(minDegreesBetweenCurrentViewAndMapPointView) * (np.float32(math.pi) / np.float32(180.0))
    p.min_cos_view_angle = float(np.cos(np.float32(deg2rad), dtype=np.float32))
    p.image_border = float(imageBorder)
    p.width, p.height = int(width), int(height)
    p.pyramid_scale = float(pyramidScale)
    p.num_levels = int(numLevels)
    return p


def ProjectMapPoints(params, mapPoints):
    """mapPoints: array of MAP_POINT_DTYPE. Returns (keypoints[KEYPOINT_DTYPE], depth f32[n], flags u8[n])."""
    pts = np.ascontiguousarray(mapPoints, MAP_POINT_DTYPE)
    n = len(pts)
    kps = np.zeros(n, KEYPOINT_DTYPE)
    depth = np.zeros(n, np.float32)
    flags = np.zeros(n, np.uint8)
    check(lib().mage_project_map_points(C.byref(params), ptr(pts), n, ptr(kps), ptr(depth), ptr(flags), None))
    return kps, depth, flags


def ProjectMapPointsDevice(params, d_points, d_kps, d_depth, d_flags, n, stream=None):
    """Device-resident variant (torch tensors / raw device pointers); asynchronous on `stream`."""
    check(lib().mage_project_map_points_device(C.byref(params), ptr(d_points), int(n), ptr(d_kps), ptr(d_depth) if d_depth is not None else None,
                                               ptr(d_flags), stream_ptr(stream)))


def ProjectPoints(points3D, cameraPose, calibrationMatrix):
    """Mirror of mage::ProjectPoints (reference Tracking/Reprojection.cpp:16-24): returns (points2D f32[n,2], depth f32[n])."""
    pts = np.zeros(len(points3D), MAP_POINT_DTYPE)
    pts["position"] = np.asarray(points3D, np.float32).reshape(-1, 3)
    pts["dmin"] = 1.0
    pts["dmax"] = 1.0
    p = make_params(cameraPose, calibrationMatrix, (0, 0, 0), (0, 0, 1), 0.0, 0.0, 1, 1, 2.0, 1)
    kps, depth, _ = ProjectMapPoints(p, pts)
    return np.stack([kps["x"], kps["y"]], axis=1), depth


def OptimizeCameraPose(position, rotation, intrinsics, mapPoints, projections, information, numIterations, maxOutlierErrorSquared, huberWidth):
    """TrackLocalMap::OptimizeCameraPose (reference TrackLocalMap.cpp:421-501). position[3] + rotation[9, column-major] = the frame's view
    transform, intrinsics = (cx, cy, fx, fy), mapPoints[n,3] / projections[n,2] / information[n] as the reference hands them to
    BundlerLib (information = MapPointRefinementConfidence). Returns (position f32[3], rotation f32[9], outlierIndices u32[], mean error)."""
    pos = np.ascontiguousarray(position, np.float32).reshape(3); rot = np.ascontiguousarray(rotation, np.float32).reshape(9)
    intr = np.ascontiguousarray(intrinsics, np.float32).reshape(4)
    pts = np.ascontiguousarray(mapPoints, np.float32).reshape(-1, 3); uv = np.ascontiguousarray(projections, np.float32).reshape(-1, 2)
    info = np.ascontiguousarray(information, np.float32).reshape(-1)
    n = len(pts)
    assert len(uv) == n and len(info) == n
    opos = np.zeros(3, np.float32); orot = np.zeros(9, np.float32); out = np.zeros(max(n, 1), np.uint32)
    cnt = C.c_int(0); mean = C.c_float(0)
    check(lib().mage_optimize_camera_pose(ptr(pos), ptr(rot), ptr(intr), n, ptr(pts), ptr(uv), ptr(info), int(numIterations), C.c_float(huberWidth),
                                          C.c_float(maxOutlierErrorSquared), ptr(opos), ptr(orot), ptr(out), len(out), C.byref(cnt), C.byref(mean)))
    return opos, orot, out[:cnt.value].copy(), float(mean.value)
