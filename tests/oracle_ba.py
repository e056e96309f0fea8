"""ctypes bindings of the two BA checkers (TEST INFRASTRUCTURE):
  kind="ref"  -> oracle/_ref/libbundler_ref.so : the reference's own BundlerLib.cpp + g2o + Eigen, compiled unmodified
  kind="port" -> oracle/libba_oracle.so        : the FP64 restatement (oracle/ba_oracle.cpp)
Both expose the same calls, so a test can drive either (or the CUDA path) with one script."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_libs = {}


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libbundler_ref.so"))


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    path = os.path.join(ROOT, "oracle", "_ref", "libbundler_ref.so") if kind == "ref" else os.path.join(ROOT, "oracle", "libba_oracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "-j8"], check=True)
    L = C.CDLL(path)
    pre = "refba_" if kind == "ref" else "baorc_"
    fn = lambda n: getattr(L, pre + n)
    fn("create").restype = C.c_void_p
    fn("create").argtypes = [C.c_int]
    fn("destroy").argtypes = [C.c_void_p]
    fn("alloc").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    fn("set_camera").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    fn("set_point").argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    fn("set_observation").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float]
    fn("fix_camera").argtypes = [C.c_void_p, C.c_int, C.c_int]
    fn("set_lambda").argtypes = [C.c_void_p, C.c_float]
    fn("get_lambda").argtypes = [C.c_void_p]
    fn("get_lambda").restype = C.c_float
    fn("step").argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    fn("step").restype = C.c_float
    fn("alloc_tethers").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    fn("set_fixed_distance").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
    fn("set_relative_rotation").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float]
    fn("set_relative_transform").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float]
    fn("get_pose").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    fn("get_point").argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    _libs[kind] = (L, pre)
    return _libs[kind]


class BaOracle:
    """Same method names as mage::BundlerLib (reference Dependencies/BundlerLib/Include/BundlerLib.h:20-66)."""

    def __init__(self, kind="ref", are_points_fixed=False):
        self.L, self.pre = _load(kind)
        self.kind = kind
        self.h = self._f("create")(1 if are_points_fixed else 0)
        self.K = self.P = self.E = 0

    def _f(self, n):
        return getattr(self.L, self.pre + n)

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, prob):
        """prob: dict from mageslam_b200.synth.ba_problem"""
        K, P, E = len(prob["cam_pos"]), len(prob["points"]), len(prob["obs_uv"])
        self.K, self.P, self.E = K, P, E
        self._f("alloc")(self.h, K, P, E)
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        for k in range(K):
            self._f("set_camera")(self.h, k, f32(prob["cam_pos"][k]).ctypes.data, f32(prob["cam_rot"][k]).ctypes.data,
                                  f32(prob["intrinsics"][k]).ctypes.data, int(prob["fixed"][k]))
        for i in range(P):
            self._f("set_point")(self.h, i, f32(prob["points"][i]).ctypes.data)
        for e in range(E):
            self._f("set_observation")(self.h, e, f32(prob["obs_uv"][e]).ctypes.data, int(prob["obs_cam"][e]), int(prob["obs_pt"][e]),
                                       float(prob["obs_info"][e]))
        self.load_tethers(prob)
        return self

    def load_tethers(self, prob):
        """Optional tether edges of a synth.ba_problem(..., tethers=True) dict, in BundlerLib's own call order."""
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        dist, rot, xf = prob.get("tether_distance", []), prob.get("tether_rotation", []), prob.get("tether_transform", [])
        if not (len(dist) or len(rot) or len(xf)):
            return
        self._f("alloc_tethers")(self.h, len(dist), len(rot), len(xf))
        for i, (c1, c2, d, w) in enumerate(dist):
            self._f("set_fixed_distance")(self.h, i, int(c1), int(c2), float(d), float(w))
        for i, (c1, c2, q, w) in enumerate(rot):
            self._f("set_relative_rotation")(self.h, i, int(c1), int(c2), f32(q).ctypes.data, float(w))
        for i, (c1, c2, t, q, w) in enumerate(xf):
            self._f("set_relative_transform")(self.h, i, int(c1), int(c2), f32(t).ctypes.data, f32(q).ctypes.data, float(w))

    def SetCurrentLambda(self, l):
        self._f("set_lambda")(self.h, float(l))

    def GetCurrentLambda(self):
        return float(self._f("get_lambda")(self.h))

    def FixCameraPose(self, idx, value):
        self._f("fix_camera")(self.h, int(idx), 1 if value else 0)

    def StepBundleAdjustment(self, huber, max_error_square):
        hub = np.ascontiguousarray(huber, np.float32)
        out = np.zeros(max(self.E, 1), np.uint32)
        n = C.c_int(0)
        r = self._f("step")(self.h, hub.ctypes.data, len(hub), float(max_error_square), out.ctypes.data, len(out), C.byref(n))
        return float(r), out[:n.value].copy()

    def poses(self):
        pos = np.zeros((self.K, 3), np.float32); rot = np.zeros((self.K, 9), np.float32)
        for k in range(self.K):
            self._f("get_pose")(self.h, k, pos[k].ctypes.data, rot[k].ctypes.data)
        return pos, rot

    def points(self):
        pts = np.zeros((self.P, 3), np.float32)
        for i in range(self.P):
            self._f("get_point")(self.h, i, pts[i].ctypes.data)
        return pts

    def state_f64(self):
        assert self.kind == "port"
        cams = np.zeros((self.K, 7)); pts = np.zeros((self.P, 3))
        self.L.baorc_get_state_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.baorc_get_state_f64(self.h, cams.ctypes.data, pts.ctypes.data)
        return cams, pts


def rel_frobenius(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
