"""ORB extractor oracle bindings -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

`OrbOracle` wraps oracle/orb_oracle.cpp (restatement of O3/src/ORBextractor.cc).
`extract_with_cv2` re-runs the same pipeline with the *real* OpenCV primitives
(cv2 4.13.0: FastFeatureDetector per cell, resize, GaussianBlur, fastAtan2)
substituted for the C models, re-using only the in-tree logic (octree, moments,
rBRIEF) from the C++ restatement; it is the pinning cross-check.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

_u8p = C.POINTER(C.c_uint8)


def _p(a, t=_u8p):
    return a.ctypes.data_as(t)


class OrbOracle:
    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        L = lib()
        L.orbo_create.restype = C.c_void_p
        L.orbo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbo_destroy.argtypes = [C.c_void_p]
        L.orbo_extract.restype = C.c_int
        L.orbo_extract.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, _u8p, C.c_int, C.POINTER(C.c_int)]
        L.orbo_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.orbo_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orbo_level_image.argtypes = [C.c_void_p, C.c_int, C.c_int, _u8p]
        for f in (L.orbo_level_candidates, L.orbo_level_selected):
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self.L = L
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.ini_th, self.min_th = ini_th, min_th
        self.h = L.orbo_create(nfeatures, scale, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orbo_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, inv, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        per = np.zeros(n, np.int32)
        umax = np.zeros(16, np.int32)
        self.L.orbo_tables(self.h, *(a.ctypes.data for a in (sc, inv, s2, is2, per, umax)))
        return dict(scale=sc, inv_scale=inv, sigma2=s2, inv_sigma2=is2, per_level=per, umax=umax)

    def extract(self, img: np.ndarray, lap=(0, 1000)):
        """Returns (keypoints[KP_DTYPE], descriptors[N,32] u8, mono_index)."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = max(self.nfeatures * 2, 64)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        mono = C.c_int(0)
        n = self.L.orbo_extract(self.h, _p(img), w, h, img.strides[0], lap[0], lap[1],
                                kps.ctypes.data, _p(desc), cap, C.byref(mono))
        if n < 0:
            return n, None, -1
        return kps[:n].copy(), desc[:n].copy(), mono.value

    # ---- stage accessors (valid after extract) ----
    def level_image(self, level, blurred=False):
        w, h = C.c_int(), C.c_int()
        self.L.orbo_level_size(self.h, level, C.byref(w), C.byref(h))
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.orbo_level_image(self.h, level, int(blurred), _p(out))
        return out

    def level_candidates(self, level):
        buf = np.zeros(1 << 17, KP_DTYPE)
        n = self.L.orbo_level_candidates(self.h, level, buf.ctypes.data, len(buf))
        return buf[:n].copy()

    def level_selected(self, level):
        buf = np.zeros(1 << 15, KP_DTYPE)
        n = self.L.orbo_level_selected(self.h, level, buf.ctypes.data, len(buf))
        return buf[:n].copy()


def distribute(pts: np.ndarray, min_x, max_x, min_y, max_y, n):
    L = lib()
    L.orbo_distribute.restype = C.c_int
    L.orbo_distribute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    pts = np.ascontiguousarray(pts, KP_DTYPE)
    out = np.zeros(max(len(pts), 1), KP_DTYPE)
    m = L.orbo_distribute(pts.ctypes.data, len(pts), min_x, max_x, min_y, max_y, n, out.ctypes.data)
    return out[:m].copy()


def extract_with_cv2(img: np.ndarray, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, lap=(0, 1000)):
    """ORBextractor::operator() (O3/src/ORBextractor.cc:876-955) with cv2 supplying the
    external primitives exactly where the reference calls them."""
    import cv2

    cv2.setNumThreads(1)
    L = lib()
    L.orbo_ic_moments.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orbo_descriptor.argtypes = [_u8p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _u8p]
    orc = OrbOracle(nfeatures, scale, nlevels, ini_th, min_th)
    T = orc.tables()
    umax = T["umax"]
    h0, w0 = img.shape
    pyr = [np.ascontiguousarray(img)]
    for l in range(1, nlevels):
        s = T["inv_scale"][l]
        dw = int(np.rint(np.float32(w0) * s))
        dh = int(np.rint(np.float32(h0) * s))
        pyr.append(cv2.resize(pyr[l - 1], (dw, dh), interpolation=cv2.INTER_LINEAR))
    det_ini = cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_min = cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    sel = []
    for l in range(nlevels):
        im = pyr[l]
        hh, ww = im.shape
        minbx = minby = 16
        maxbx, maxby = ww - 16, hh - 16
        width, height = np.float32(maxbx - minbx), np.float32(maxby - minby)
        ncols, nrows = int(width / np.float32(35)), int(height / np.float32(35))
        wcell = int(np.ceil(width / np.float32(ncols)))
        hcell = int(np.ceil(height / np.float32(nrows)))
        cand = []
        for i in range(nrows):
            iy = minby + i * hcell
            my = iy + hcell + 6
            if iy >= maxby - 3:
                continue
            my = min(my, maxby)
            for j in range(ncols):
                ix = minbx + j * wcell
                mx = ix + wcell + 6
                if ix >= maxbx - 6:
                    continue
                mx = min(mx, maxbx)
                cell = im[iy:my, ix:mx]
                k = det_ini.detect(cell)
                if len(k) == 0:
                    k = det_min.detect(cell)
                for p in k:
                    cand.append((p.pt[0] + j * wcell, p.pt[1] + i * hcell, p.size, p.angle, p.response,
                                 p.octave, p.class_id))
        cand = np.array(cand, KP_DTYPE) if cand else np.zeros(0, KP_DTYPE)
        s_l = distribute(cand, minbx, maxbx, minby, maxby, int(T["per_level"][l])) if len(cand) else cand
        s_l["x"] += minbx
        s_l["y"] += minby
        s_l["octave"] = l
        s_l["size"] = np.float32(int(np.float32(31) * T["scale"][l]))
        for q in s_l:
            m01, m10 = C.c_int(), C.c_int()
            L.orbo_ic_moments(_p(im), ww, hh, int(np.rint(q["x"])), int(np.rint(q["y"])), umax.ctypes.data,
                              C.byref(m01), C.byref(m10))
            q["angle"] = cv2.fastAtan2(float(m01.value), float(m10.value))
        sel.append(s_l)
    nk = sum(len(s) for s in sel)
    kps = np.zeros(nk, KP_DTYPE)
    desc = np.zeros((nk, 32), np.uint8)
    mono, stereo = 0, nk - 1
    d = np.zeros(32, np.uint8)
    for l in range(nlevels):
        if len(sel[l]) == 0:
            continue
        work = cv2.GaussianBlur(pyr[l].copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        hh, ww = work.shape
        sc = T["scale"][l]
        for q in sel[l]:
            L.orbo_descriptor(_p(work), ww, hh, float(q["x"]), float(q["y"]), float(q["angle"]), _p(d))
            o = q.copy()
            if l != 0:
                o["x"] = np.float32(o["x"]) * sc
                o["y"] = np.float32(o["y"]) * sc
            if lap[0] <= o["x"] <= lap[1]:
                dst = stereo
                stereo -= 1
            else:
                dst = mono
                mono += 1
            kps[dst] = o
            desc[dst] = d
    return kps, desc, mono
