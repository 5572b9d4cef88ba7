"""Bindings of oracle/_ref/libref_matcher.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The library is the reference's own ORBmatcher.cc (and the Frame / KeyFrame / MapPoint / Pinhole member bodies it calls),
compiled unmodified from /root/reference against oracle/slamshim (see oracle/Makefile, target `ref`).  It exists only
where /root/reference is present (or where the built .so travelled); tests that need it skip otherwise.  Call
signatures mirror oracle.track / oracle.bow so that the same inputs go to both."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .orb import KP_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_matcher.so")
_vp = C.c_void_p
_LIB = None


def available() -> bool:
    return os.path.exists(SO) or os.path.isdir("/root/reference/src/slam_system/orb_slam3")


def _L():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO):
            import oracle

            oracle.build(ref=True)
        L = C.CDLL(SO)
        L.refm_frame_create.restype = _vp
        L.refm_frame_create.argtypes = [_vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp]
        L.refm_frame_destroy.argtypes = [_vp]
        L.refm_grid_cell.argtypes = [_vp, C.c_int, C.c_int, _vp, C.c_int]
        L.refm_features_in_area.argtypes = [_vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp, C.c_int]
        L.refm_kf_features_in_area.argtypes = [_vp, C.c_float, C.c_float, C.c_float, _vp, C.c_int]
        L.refm_descriptor_distance.argtypes = [_vp, _vp]
        L.refm_search_by_projection_last.argtypes = [_vp, _vp, _vp, C.c_int] + [_vp] * 7 + [C.c_float, C.c_int, _vp]
        L.refm_is_in_frustum.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_int] + [_vp] * 5 + [C.c_float] + [_vp] * 5
        L.refm_is_in_frustum.restype = None
        L.refm_search_by_projection_map.argtypes = [_vp, C.c_int] + [_vp] * 6 + [C.c_float, C.c_float, _vp, _vp]
        L.refm_search_by_bow.argtypes = [C.c_int] + [C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp] * 2 + [C.c_float, C.c_int, _vp, _vp]
        L.refm_search_for_initialization.argtypes = [C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_int, _vp]
        L.refm_search_for_triangulation.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_int, C.c_int, _vp]
        L.refm_frame_set_featvec.argtypes = [_vp, C.c_int, _vp, _vp, _vp]
        L.refm_frame_set_featvec.restype = None
        L.refm_fuse.argtypes = [_vp, _vp, _vp, C.c_int] + [_vp] * 6 + [C.c_float, _vp]
        L.refm_search_by_projection_sim3.argtypes = [_vp, _vp, _vp, C.c_int] + [_vp] * 7 + [C.c_int, C.c_float, _vp]
        L.refm_fuse_sim3.argtypes = [_vp, _vp, _vp, C.c_int] + [_vp] * 6 + [C.c_float, _vp]
        L.refm_search_by_sim3.argtypes = [_vp] * 18 + [C.c_float, _vp]
        _LIB = L
    return _LIB


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def _p(a):
    return a.ctypes.data if a is not None else None


class RefFrame:
    """A reference Frame (and, on demand, the KeyFrame made from it) over flat arrays."""

    def __init__(self, kps, desc, bounds, scale_factors, K=(500.0, 500.0, 320.0, 240.0)):
        self.L = _L()
        self.kps, self.desc = _c(kps, KP_DTYPE), _c(desc, np.uint8)
        self.sf, self.b, self.K = _c(scale_factors, np.float32), _c(bounds, np.float32), _c(K, np.float32)
        self.n = len(self.kps)
        self.h = self.L.refm_frame_create(_p(self.kps), _p(self.desc), self.n, _p(self.b), _p(self.sf), len(self.sf), _p(self.K))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.refm_frame_destroy(self.h)
            self.h = None

    def grid_cell(self, ix, iy):
        buf = np.zeros(max(self.n, 1), np.int32)
        return buf[:self.L.refm_grid_cell(self.h, ix, iy, _p(buf), len(buf))].copy()

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        buf = np.zeros(max(self.n, 1), np.int32)
        return buf[:self.L.refm_features_in_area(self.h, x, y, r, min_level, max_level, _p(buf), len(buf))].copy()

    def kf_features_in_area(self, x, y, r):
        buf = np.zeros(max(self.n, 1), np.int32)
        return buf[:self.L.refm_kf_features_in_area(self.h, x, y, r, _p(buf), len(buf))].copy()

    def set_feature_vector(self, fv):
        from .bow import _csr

        nid, st, idx = _csr(fv)
        self.L.refm_frame_set_featvec(self.h, len(nid), _p(nid), _p(st), _p(idx))

    def search_by_projection_last(self, q, t, has_mp, outlier, Xw, mp_desc, obs_pos, last_octave, last_angle, th, check_ori=True):
        cur_mp = np.full(max(self.n, 1), -1, np.int32)
        q, t = _c(q, np.float32), _c(t, np.float32)
        b = [_c(has_mp, np.uint8), _c(outlier, np.uint8), _c(Xw, np.float32), _c(mp_desc, np.uint8), _c(obs_pos, np.uint8),
             _c(last_octave, np.int32), _c(last_angle, np.float32)]
        n = self.L.refm_search_by_projection_last(self.h, _p(q), _p(t), len(b[0]), *(_p(x) for x in b), float(th), int(check_ori),
                                                  _p(cur_mp))
        return n, cur_mp[:self.n]

    def search_by_projection_map(self, projX, projY, level, view_cos, mp_desc, obs_pos, th, nnratio, cur_blocked):
        cur_mp = np.full(max(self.n, 1), -1, np.int32)
        b = [_c(projX, np.float32), _c(projY, np.float32), _c(level, np.int32), _c(view_cos, np.float32), _c(mp_desc, np.uint8),
             _c(obs_pos, np.uint8)]
        blk = _c(cur_blocked, np.uint8)
        n = self.L.refm_search_by_projection_map(self.h, len(b[0]), *(_p(x) for x in b), float(th), float(nnratio), _p(blk), _p(cur_mp))
        return n, cur_mp[:self.n]

    def fuse(self, q, t, xw, normal, min_dist, max_dist, mp_desc, skip=None, th=3.0):
        """min_dist / max_dist: mfMinDistance / mfMaxDistance of the map points (not the 0.8 / 1.2 scaled ones)."""
        a = [_c(xw, np.float32), _c(normal, np.float32), _c(min_dist, np.float32), _c(max_dist, np.float32), _c(mp_desc, np.uint8)]
        m = len(a[2])
        sk = _c(skip, np.uint8) if skip is not None else None
        bi = np.full(max(m, 1), -1, np.int32)
        q, t = _c(q, np.float32), _c(t, np.float32)
        n = self.L.refm_fuse(self.h, _p(q), _p(t), m, *(_p(x) for x in a), _p(sk), float(th), _p(bi))
        return n, bi[:m]


def _pts(xw, normal, min_dist, max_dist, mp_desc, skip):
    a = [_c(xw, np.float32), _c(normal, np.float32), _c(min_dist, np.float32), _c(max_dist, np.float32), _c(mp_desc, np.uint8)]
    return a, (_c(skip, np.uint8) if skip is not None else None), len(a[2])


def search_by_projection_sim3(kf: RefFrame, sq, st, xw, normal, min_dist, max_dist, mp_desc, skip, kp_matched, th, ratio_hamming=1.0):
    a, sk, m = _pts(xw, normal, min_dist, max_dist, mp_desc, skip)
    km = _c(kp_matched, np.uint8)
    out = np.full(max(kf.n, 1), -1, np.int32)
    sq, st = _c(sq, np.float32), _c(st, np.float32)
    n = _L().refm_search_by_projection_sim3(kf.h, _p(sq), _p(st), m, *(_p(x) for x in a), _p(sk), _p(km), int(th), float(ratio_hamming), _p(out))
    return n, out[:kf.n]


def fuse_sim3(kf: RefFrame, sq, st, xw, normal, min_dist, max_dist, mp_desc, skip=None, th=3.0):
    a, sk, m = _pts(xw, normal, min_dist, max_dist, mp_desc, skip)
    bi = np.full(max(m, 1), -1, np.int32)
    sq, st = _c(sq, np.float32), _c(st, np.float32)
    n = _L().refm_fuse_sim3(kf.h, _p(sq), _p(st), m, *(_p(x) for x in a), _p(sk), float(th), _p(bi))
    return n, bi[:m]


def search_by_sim3(f1: RefFrame, f2: RefFrame, q1, t1, q2, t2, s12q, s12t, side1, side2, th=7.5):
    p = [_c(x, np.float32) for x in (q1, t1, q2, t2, s12q, s12t)]
    sides = []
    for sk, xw, mn, mx, d in (side1, side2):
        sides += [_c(sk, np.uint8), _c(xw, np.float32), _c(mn, np.float32), _c(mx, np.float32), _c(d, np.uint8)]
    m12 = np.full(max(f1.n, 1), -1, np.int32)
    n = _L().refm_search_by_sim3(f1.h, f2.h, *(_p(x) for x in p), *(_p(x) for x in sides), float(th), _p(m12))
    return n, m12[:f1.n]


def descriptor_distance(a, b):
    return _L().refm_descriptor_distance(_p(_c(a, np.uint8)), _p(_c(b, np.uint8)))


def is_in_frustum(q, t, K, bounds, nlevels, scale_factor, xw, normal, min_dist, max_dist, skip=None, cos_limit=0.5):
    """min_dist / max_dist: mfMinDistance / mfMaxDistance."""
    L = _L()
    xw, normal = _c(xw, np.float32), _c(normal, np.float32)
    m = len(xw)
    inv, px, py = np.zeros(max(m, 1), np.uint8), np.zeros(max(m, 1), np.float32), np.zeros(max(m, 1), np.float32)
    lv, vc = np.zeros(max(m, 1), np.int32), np.zeros(max(m, 1), np.float32)
    sk = _c(skip, np.uint8) if skip is not None else None
    a = [_c(q, np.float32), _c(t, np.float32), _c(K, np.float32), _c(bounds, np.float32)]
    d = [_c(min_dist, np.float32), _c(max_dist, np.float32)]
    L.refm_is_in_frustum(*(_p(x) for x in a), nlevels, float(scale_factor), m, _p(xw), _p(normal), _p(d[0]), _p(d[1]), _p(sk),
                         float(cos_limit), _p(inv), _p(px), _p(py), _p(lv), _p(vc))
    return inv[:m], px[:m], py[:m], lv[:m], vc[:m]


def search_by_bow(kf_kf, desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2, nnratio=0.6, check_ori=True):
    from .bow import _csr

    L = _L()
    sides = []
    for d, a, v, fv in ((desc1, angle1, valid1, fv1), (desc2, angle2, valid2, fv2)):
        d, a = _c(d, np.uint8), _c(a, np.float32)
        v = None if v is None else _c(v, np.uint8)
        nid, st, idx = _csr(fv)
        sides.append((len(a), d, a, v, len(nid), nid, st, idx))
    m12 = np.full(max(sides[0][0], 1), -1, np.int32)
    m21 = np.full(max(sides[1][0], 1), -1, np.int32)
    args = []
    for n, d, a, v, nn, nid, st, idx in sides:
        args += [n, _p(d), _p(a), _p(v), nn, _p(nid), _p(st), _p(idx)]
    n = L.refm_search_by_bow(int(kf_kf), *args, float(nnratio), int(check_ori), _p(m12), _p(m21))
    return n, m12[:sides[0][0]], m21[:sides[1][0]]


def search_for_initialization(kps1, desc1, frame2: RefFrame, prev_matched, window=100, nnratio=0.9, check_ori=True):
    L = _L()
    k1, d1 = _c(kps1, KP_DTYPE), _c(desc1, np.uint8)
    pm = _c(prev_matched, np.float32).copy()
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = L.refm_search_for_initialization(len(k1), _p(k1), _p(d1), frame2.h, _p(pm), int(window), float(nnratio), int(check_ori), _p(m12))
    return n, m12[:len(k1)], pm


def search_for_triangulation(f1: RefFrame, has_mp1, q1, t1, f2: RefFrame, has_mp2, q2, t2, nnratio=0.6, check_ori=False, coarse=False):
    """f1 / f2 carry the keypoints, descriptors, intrinsics and (set_feature_vector) the feature vectors."""
    L = _L()
    m12 = np.full(max(f1.n, 1), -1, np.int32)
    a = [_c(has_mp1, np.uint8), _c(q1, np.float32), _c(t1, np.float32), _c(has_mp2, np.uint8), _c(q2, np.float32), _c(t2, np.float32)]
    n = L.refm_search_for_triangulation(f1.h, _p(a[0]), _p(a[1]), _p(a[2]), f2.h, _p(a[3]), _p(a[4]), _p(a[5]), float(nnratio),
                                        int(check_ori), int(coarse), _p(m12))
    return n, m12[:f1.n]
