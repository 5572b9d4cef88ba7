"""Tracking-operator oracle bindings (oracle/track_oracle.cpp) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib
from .orb import KP_DTYPE

_vp = C.c_void_p


def _L():
    L = lib()
    if getattr(L, "_trk_bound", False):
        return L
    L.trko_frame_create.restype = _vp
    L.trko_frame_create.argtypes = [_vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _vp, C.c_int]
    L.trko_frame_destroy.argtypes = [_vp]
    L.trko_grid_cell.argtypes = [_vp, C.c_int, C.c_int, _vp, C.c_int]
    L.trko_features_in_area.argtypes = [_vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp, C.c_int]
    L.trko_descriptor_distance.argtypes = [_vp, _vp]
    L.trko_search_by_projection_last.argtypes = [_vp, _vp, _vp, _vp, C.c_int] + [_vp] * 7 + [C.c_float, C.c_int, _vp]
    L.trko_search_by_projection_map.argtypes = [_vp, C.c_int] + [_vp] * 6 + [C.c_float, C.c_float, _vp, _vp]
    L.trko_pose_optimization.argtypes = [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]
    L._trk_bound = True
    return L


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


class FrameOracle:
    """Frame (mono): undistorted keypoints, descriptors, image bounds, 64x48 grid."""

    def __init__(self, kps, desc, bounds, scale_factors):
        self.L = _L()
        self.kps = _c(kps, KP_DTYPE)
        self.desc = _c(desc, np.uint8)
        self.sf = _c(scale_factors, np.float32)
        self.n = len(self.kps)
        self.h = self.L.trko_frame_create(self.kps.ctypes.data, self.desc.ctypes.data, self.n, *map(float, bounds),
                                          self.sf.ctypes.data, len(self.sf))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.trko_frame_destroy(self.h)
            self.h = None

    def grid_cell(self, ix, iy):
        buf = np.zeros(max(self.n, 1), np.int32)
        n = self.L.trko_grid_cell(self.h, ix, iy, buf.ctypes.data, len(buf))
        return buf[:n].copy()

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        buf = np.zeros(max(self.n, 1), np.int32)
        n = self.L.trko_features_in_area(self.h, x, y, r, min_level, max_level, buf.ctypes.data, len(buf))
        return buf[:n].copy()

    def search_by_projection_last(self, Rcw, tcw, K, has_mp, outlier, Xw, mp_desc, obs_pos, last_octave, last_angle,
                                  th, check_ori=True):
        cur_mp = np.full(max(self.n, 1), -1, np.int32)
        a = [_c(Rcw, np.float32), _c(tcw, np.float32), _c(K, np.float32)]
        b = [_c(has_mp, np.uint8), _c(outlier, np.uint8), _c(Xw, np.float32), _c(mp_desc, np.uint8),
             _c(obs_pos, np.uint8), _c(last_octave, np.int32), _c(last_angle, np.float32)]
        n = self.L.trko_search_by_projection_last(self.h, *(x.ctypes.data for x in a), len(b[0]),
                                                  *(x.ctypes.data for x in b), float(th), int(check_ori),
                                                  cur_mp.ctypes.data)
        return n, cur_mp[:self.n]

    def search_by_projection_map(self, projX, projY, level, view_cos, mp_desc, obs_pos, th, nnratio, cur_blocked):
        cur_mp = np.full(max(self.n, 1), -1, np.int32)
        b = [_c(projX, np.float32), _c(projY, np.float32), _c(level, np.int32), _c(view_cos, np.float32),
             _c(mp_desc, np.uint8), _c(obs_pos, np.uint8)]
        blk = _c(cur_blocked, np.uint8)
        n = self.L.trko_search_by_projection_map(self.h, len(b[0]), *(x.ctypes.data for x in b), float(th),
                                                 float(nnratio), blk.ctypes.data, cur_mp.ctypes.data)
        return n, cur_mp[:self.n]


def descriptor_distance(a, b):
    return _L().trko_descriptor_distance(_c(a, np.uint8).ctypes.data, _c(b, np.uint8).ctypes.data)


def pose_optimization(q, t, K, Xw, kp_xy, inv_sigma2):
    """Returns (n_inliers, q[4] f32, t[3] f32, outlier[n] u8, (lm_iterations, lm_trials))."""
    L = _L()
    q = _c(q, np.float32).copy()
    t = _c(t, np.float32).copy()
    Xw, kp_xy, w = _c(Xw, np.float32), _c(kp_xy, np.float32), _c(inv_sigma2, np.float32)
    n = len(w)
    out = np.zeros(max(n, 1), np.uint8)
    stats = np.zeros(2, np.int32)
    r = L.trko_pose_optimization(q.ctypes.data, t.ctypes.data, _c(K, np.float32).ctypes.data, n, Xw.ctypes.data,
                                 kp_xy.ctypes.data, w.ctypes.data, out.ctypes.data, stats.ctypes.data)
    return r, q, t, out[:n], tuple(stats)
