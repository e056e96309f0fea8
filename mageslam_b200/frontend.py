"""Video-stream front-end over the C ABI: ORB extract of a batch of frames + Match of each frame against its predecessor
(mage_frontend_* in include/mage_b200.h). This is the call a user makes per batch of frames; bench.py's end-to-end
number goes through FrontEnd.Process with pinned host buffers."""
import ctypes as C

import numpy as np

from ._lib import DMATCH_DTYPE, KEYPOINT_DTYPE, OrbParams, check, lib, ptr, stream_ptr
from .matcher import OrbMatcherSettings
from .orb import FeatureExtractorSettings


def _params(s: FeatureExtractorSettings):
    return OrbParams(int(s.GaussianKernelSize), int(s.NumFeatures), float(s.ScaleFactor), int(s.NumLevels), int(s.PatchSize),
                     int(s.FastThreshold), 1 if s.UseOrientation else 0, float(s.FeatureFactor), float(s.FeatureStrength),
                     int(s.StrongResponse), float(s.MinRobustnessFactor), float(s.MaxRobustnessFactor), int(s.NumCellsX), int(s.NumCellsY))


class FrontEnd:
    def __init__(self, settings: FeatureExtractorSettings, width, height, batch, chunk=0, matcher: OrbMatcherSettings = OrbMatcherSettings()):
        L = lib()
        L.mage_frontend_create.argtypes = [C.POINTER(OrbParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.mage_frontend_destroy.argtypes = [C.c_void_p]; L.mage_frontend_destroy.restype = None
        L.mage_frontend_reset.argtypes = [C.c_void_p]
        L.mage_frontend_capacity.argtypes = [C.c_void_p]
        L.mage_frontend_process.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t] + [C.c_void_p] * 5
        L.mage_frontend_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t] + [C.c_void_p] * 5
        L.mage_frontend_wait.argtypes = [C.c_void_p]
        L.mage_frontend_process_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p]
        L.mage_frontend_join.argtypes = [C.c_void_p, C.c_void_p]
        L.mage_frontend_device_buffers.argtypes = [C.c_void_p] + [C.POINTER(C.c_void_p)] * 5 + [C.POINTER(C.c_int)]
        self.width, self.height, self.batch = int(width), int(height), int(batch)
        self._p = _params(settings)
        self._h = C.c_void_p()
        check(L.mage_frontend_create(C.byref(self._p), self.width, self.height, self.batch, int(chunk), int(matcher.MaxHammingDistance),
                                     int(matcher.MinHammingDifference), C.byref(self._h)))
        self.capacity = L.mage_frontend_capacity(self._h)

    def close(self):
        if self._h:
            lib().mage_frontend_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Reset(self):
        check(lib().mage_frontend_reset(self._h))

    def alloc_outputs(self, pinned=True):
        """Host output buffers for Process (pinned torch tensors viewed as numpy when torch is available)."""
        n, cap = self.batch, self.capacity
        shapes = [((n, cap, 28), np.uint8), ((n, cap, 32), np.uint8), ((n,), np.int32), ((n, cap, 12), np.uint8), ((n,), np.int32)]
        outs = []
        for shp, dt in shapes:
            if pinned:
                import torch
                t = torch.zeros(shp, dtype=torch.uint8 if dt == np.uint8 else torch.int32).pin_memory()
                outs.append(t)
            else:
                outs.append(np.zeros(shp, dt))
        return outs

    def Process(self, images, outs):
        """images: host uint8 [n, h, w] (numpy or pinned torch); outs from alloc_outputs. Synchronous.
        Returns (kps, desc, counts, matches, match_counts) as numpy views of `outs`."""
        n = images.shape[0]
        stride = images.stride(1) if hasattr(images, "stride") else images.strides[1]
        fstride = images.stride(0) if hasattr(images, "stride") else images.strides[0]
        check(lib().mage_frontend_process(self._h, ptr(images), n, stride, fstride, *[ptr(o) for o in outs]))
        arr = [o.numpy() if hasattr(o, "numpy") else o for o in outs]
        kps = arr[0].reshape(self.batch, -1).view(KEYPOINT_DTYPE)
        matches = arr[3].reshape(self.batch, -1).view(DMATCH_DTYPE)
        return kps, arr[1], arr[2], matches, arr[4]

    def Submit(self, images, outs):
        """Asynchronous Process: enqueue one call (at most two in flight); its results are in `outs` after the matching Wait()."""
        n = images.shape[0]
        stride = images.stride(1) if hasattr(images, "stride") else images.strides[1]
        fstride = images.stride(0) if hasattr(images, "stride") else images.strides[0]
        check(lib().mage_frontend_submit(self._h, ptr(images), n, stride, fstride, *[ptr(o) for o in outs]))

    def Wait(self):
        """Blocks until the oldest submitted call has delivered its results."""
        check(lib().mage_frontend_wait(self._h))

    def views(self, outs):
        arr = [o.numpy() if hasattr(o, "numpy") else o for o in outs]
        return arr[0].reshape(self.batch, -1).view(KEYPOINT_DTYPE), arr[1], arr[2], arr[3].reshape(self.batch, -1).view(DMATCH_DTYPE), arr[4]

    def ProcessDevice(self, d_images, stream=None):
        """d_images: torch uint8 [n, h, w] on the device. Asynchronous; results via DeviceBuffers()."""
        n = d_images.shape[0]
        check(lib().mage_frontend_process_device(self._h, ptr(d_images), n, d_images.stride(1), d_images.stride(0), stream_ptr(stream)))

    def Join(self, stream=None):
        """Makes `stream` wait for everything ProcessDevice has enqueued so far (the matches run on an internal stream)."""
        check(lib().mage_frontend_join(self._h, stream_ptr(stream)))

    def DeviceBuffers(self):
        ps = [C.c_void_p() for _ in range(5)]
        cap = C.c_int(0)
        check(lib().mage_frontend_device_buffers(self._h, *[C.byref(p) for p in ps], C.byref(cap)))
        return [p.value for p in ps], cap.value

    def ReadDeviceResults(self):
        """Copies the device-resident results of the last ProcessDevice call to host numpy arrays (test helper)."""
        import torch
        torch.cuda.synchronize()
        (p_kps, p_desc, p_cnt, p_m, p_mc), cap = self.DeviceBuffers()
        n = self.batch
        L = lib()
        L.mage_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        d2h = lambda dst, src: check(L.mage_memcpy_d2h(C.c_void_p(dst.ctypes.data), C.c_void_p(src), dst.nbytes))
        kps = np.zeros((n, cap), KEYPOINT_DTYPE); desc = np.zeros((n, cap, 32), np.uint8); cnt = np.zeros(n, np.int32)
        mt = np.zeros((n, cap), DMATCH_DTYPE); mc = np.zeros(n, np.int32)
        for dst, src in ((kps, p_kps), (desc, p_desc), (cnt, p_cnt), (mt, p_m), (mc, p_mc)):
            d2h(dst, src)
        return kps, desc, cnt, mt, mc
