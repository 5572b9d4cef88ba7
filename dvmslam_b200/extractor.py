"""Python mirror of ORB_SLAM3::ORBextractor (O3/include/ORBextractor.h:44-96) over the C-ABI.

Same constructor arguments, same call semantics: `extractor(image, lap)` returns what
`ORBextractor::operator()` writes -- the keypoint array (cv::KeyPoint layout), the [N,32] uint8
descriptor matrix and the monoIndex return value.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

_u8p = C.POINTER(C.c_uint8)
_ip = C.POINTER(C.c_int)


def _bind(L):
    if getattr(L, "_orb_bound", False):
        return
    L.dvm_orb_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int]
    L.dvm_orb_destroy.argtypes = [C.c_void_p]
    L.dvm_orb_destroy.restype = None
    L.dvm_orb_tables.argtypes = [C.c_void_p, _ip] + [C.c_void_p] * 5
    L.dvm_orb_max_keypoints.argtypes = [C.c_void_p]
    L.dvm_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, _ip, _ip]
    L.dvm_orb_extract_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.dvm_orb_sync.argtypes = [C.c_void_p]
    L.dvm_orb_result_device.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p)]
    L.dvm_orb_stream.argtypes = [C.c_void_p]
    L.dvm_orb_stream.restype = C.c_void_p
    L.dvm_orb_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.dvm_orb_get_profile.argtypes = [C.c_void_p, _ip, C.c_void_p]
    L.dvm_orb_debug_level_size.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
    L.dvm_orb_debug_level_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.dvm_orb_debug_level_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_int, _ip]
    L._orb_bound = True


class ORBextractor:
    """ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) on one B200."""

    def __init__(self, nfeatures: int, scaleFactor: float = 1.2, nlevels: int = 8, iniThFAST: int = 20,
                 minThFAST: int = 7, max_width: int = 1280, max_height: int = 720, device: int = 0):
        self.L = lib()
        _bind(self.L)
        self.h = C.c_void_p()
        self.nfeatures, self.nlevels = nfeatures, nlevels
        check(self.L.dvm_orb_create(C.byref(self.h), device, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST,
                                    max_width, max_height))
        self.cap = self.L.dvm_orb_max_keypoints(self.h)
        self._kps = np.zeros(self.cap, KP_DTYPE)
        self._desc = np.zeros((self.cap, 32), np.uint8)

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.dvm_orb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (GetLevels / GetScaleFactors / ...) ----
    def tables(self):
        n = C.c_int()
        sc, inv, s2, is2 = (np.zeros(self.nlevels, np.float32) for _ in range(4))
        per = np.zeros(self.nlevels, np.int32)
        check(self.L.dvm_orb_tables(self.h, C.byref(n), *(a.ctypes.data for a in (sc, inv, s2, is2, per))))
        return dict(nlevels=n.value, scale=sc, inv_scale=inv, sigma2=s2, inv_sigma2=is2, per_level=per)

    def GetLevels(self):
        return self.nlevels

    def GetScaleFactors(self):
        return self.tables()["scale"]

    def GetInverseScaleFactors(self):
        return self.tables()["inv_scale"]

    def GetScaleSigmaSquares(self):
        return self.tables()["sigma2"]

    def GetInverseScaleSigmaSquares(self):
        return self.tables()["inv_sigma2"]

    # ---- operator() ----
    def __call__(self, image: np.ndarray, vLappingArea=(0, 1000), copy: bool = True):
        """Host image in, host results out (synchronous), like the reference's call.
        Returns (keypoints, descriptors, monoIndex); monoIndex == -1 for an empty image."""
        if image is None or image.size == 0:
            n, mono = C.c_int(), C.c_int()
            check(self.L.dvm_orb_extract(self.h, None, 0, 0, 0, vLappingArea[0], vLappingArea[1], None, None, 0,
                                         C.byref(n), C.byref(mono)))
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8), mono.value
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 expected"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        n, mono = C.c_int(), C.c_int()
        check(self.L.dvm_orb_extract(self.h, image.ctypes.data, w, h, image.strides[0], vLappingArea[0],
                                     vLappingArea[1], self._kps.ctypes.data, self._desc.ctypes.data, self.cap,
                                     C.byref(n), C.byref(mono)))
        k, d = self._kps[:n.value], self._desc[:n.value]
        return (k.copy(), d.copy(), mono.value) if copy else (k, d, mono.value)

    def extract_device(self, dev_ptr: int, width: int, height: int, stride: int, vLappingArea=(0, 1000)):
        """Image already in HBM; enqueue only (results stay on the device)."""
        check(self.L.dvm_orb_extract_device(self.h, C.c_void_p(dev_ptr), width, height, stride, vLappingArea[0],
                                            vLappingArea[1]))

    def sync(self):
        check(self.L.dvm_orb_sync(self.h))

    def stream(self) -> int:
        return int(self.L.dvm_orb_stream(self.h) or 0)

    def result_device(self):
        k, d, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self.L.dvm_orb_result_device(self.h, C.byref(k), C.byref(d), C.byref(c)))
        return k.value, d.value, c.value

    def set_profiling(self, enable: bool):
        check(self.L.dvm_orb_set_profiling(self.h, int(enable)))

    def get_profile(self):
        """(frames, summed ms per stage [pyramid, fast, octree, describe]); resets the counters."""
        n = C.c_int()
        ms = np.zeros(4, np.float32)
        check(self.L.dvm_orb_get_profile(self.h, C.byref(n), ms.ctypes.data))
        return n.value, ms

    # ---- stage read-back used by the parity tests ----
    def level_image(self, level: int, blurred: bool = False) -> np.ndarray:
        w, h = C.c_int(), C.c_int()
        check(self.L.dvm_orb_debug_level_size(self.h, level, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        check(self.L.dvm_orb_debug_level_image(self.h, level, int(blurred), out.ctypes.data))
        return out

    def level_keypoints(self, level: int, which: int):
        cap = 1 << 18
        xs, ys, rs = (np.zeros(cap, np.int32) for _ in range(3))
        n = C.c_int()
        check(self.L.dvm_orb_debug_level_keypoints(self.h, level, which, xs.ctypes.data, ys.ctypes.data,
                                                   rs.ctypes.data, cap, C.byref(n)))
        m = min(n.value, cap)
        return np.stack([xs[:m], ys[:m], rs[:m]], 1)
