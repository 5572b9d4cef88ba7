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
    L.trko_velocity_prior.argtypes = [_vp, _vp, _vp]
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

    def search_by_projection_last(self, qcw, tcw, K, has_mp, outlier, Xw, mp_desc, obs_pos, last_octave, last_angle,
                                  th, check_ori=True):
        """qcw (x, y, z, w), tcw: the current frame's pose as its SE3f holds it."""
        cur_mp = np.full(max(self.n, 1), -1, np.int32)
        a = [_c(qcw, np.float32), _c(tcw, np.float32), _c(K, np.float32)]
        assert a[0].shape == (4,)
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


def is_in_frustum(q, t, K, bounds, nlevels, scale_factor, xw, normal, min_dist, max_dist, skip=None, cos_limit=0.5):
    L = _L()
    L.trko_is_in_frustum.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_int] + [_vp] * 5 + [C.c_float] + [_vp] * 5
    xw, normal = _c(xw, np.float32), _c(normal, np.float32)
    m = len(xw)
    inv, px, py = np.zeros(max(m, 1), np.uint8), np.zeros(max(m, 1), np.float32), np.zeros(max(m, 1), np.float32)
    lv, vc = np.zeros(max(m, 1), np.int32), np.zeros(max(m, 1), np.float32)
    sk = _c(skip, np.uint8) if skip is not None else None
    L.trko_is_in_frustum(_c(q, np.float32).ctypes.data, _c(t, np.float32).ctypes.data, _c(K, np.float32).ctypes.data,
                         _c(bounds, np.float32).ctypes.data, nlevels, float(scale_factor), m, xw.ctypes.data,
                         normal.ctypes.data, _c(min_dist, np.float32).ctypes.data, _c(max_dist, np.float32).ctypes.data,
                         sk.ctypes.data if sk is not None else None, float(cos_limit), inv.ctypes.data, px.ctypes.data,
                         py.ctypes.data, lv.ctypes.data, vc.ctypes.data)
    return inv[:m], px[:m], py[:m], lv[:m], vc[:m]


def descriptor_distance(a, b):
    return _L().trko_descriptor_distance(_c(a, np.uint8).ctypes.data, _c(b, np.uint8).ctypes.data)


def velocity_prior(last_q, last_t, prev_q, prev_t):
    """mVelocity * mLastFrame.GetPose() with mVelocity = last * prev^-1 (Tracking.cc:1990-1991,2598), float32 Sophus order.
    Returns (q[4] f32, t[3] f32)."""
    L = _L()
    last = np.concatenate([_c(last_q, np.float32), _c(last_t, np.float32)])
    prev = np.concatenate([_c(prev_q, np.float32), _c(prev_t, np.float32)])
    out = np.zeros(7, np.float32)
    L.trko_velocity_prior(last.ctypes.data, prev.ctypes.data, out.ctypes.data)
    return out[:4].copy(), out[4:].copy()


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


class TrackerOracle:
    """The per-frame chain of Tracking::TrackWithMotionModel + TrackLocalMap (O3/src/Tracking.cc:2584-2768,
    3041-3106) composed from the oracle operators, for a map snapshot given as flat arrays.  Mirrors
    dvm_tracker step by step so that the GPU pipeline can be compared frame by frame."""

    def __init__(self, extract, tables, K, bounds, world_map):
        self.extract, self.T, self.K, self.bounds, self.map = extract, tables, np.asarray(K, np.float32), bounds, world_map
        self.last = None

    def _local_map_search(self, F, q, t, cur_map, th, nnratio, seen):
        M = self.map
        inv, px, py, lv, vc = is_in_frustum(q, t, self.K, self.bounds, len(self.T["scale"]), self.T["scale"][1],
                                            M["xw"], M["normal"], M["min_dist"], M["max_dist"], seen, 0.5)
        qi = np.nonzero(inv)[0]
        blocked = (cur_map >= 0).astype(np.uint8)
        n, m2 = F.search_by_projection_map(px[qi], py[qi], lv[qi], vc[qi], M["desc"][qi], np.ones(len(qi), np.uint8), th,
                                           nnratio, blocked)
        out = cur_map.copy()
        take = (out < 0) & (m2 >= 0)
        out[take] = qi[m2[take]]
        return out, n

    def bootstrap(self, img, q, t):
        kps, desc, _ = self.extract(img)
        F = FrameOracle(kps, desc, self.bounds, self.T["scale"])
        cur_map = np.full(len(kps), -1, np.int64)
        cur_map, n = self._local_map_search(F, q, t, cur_map, 3.0, 0.8, np.zeros(len(self.map["xw"]), np.uint8))
        self.last = dict(kps=kps, mp=cur_map, outlier=np.zeros(len(kps), np.uint8), q=np.asarray(q, np.float32),
                         t=np.asarray(t, np.float32))
        return n

    def track(self, img, prior_q, prior_t):
        M, L = self.map, self.last
        kps, desc, _ = self.extract(img)
        F = FrameOracle(kps, desc, self.bounds, self.T["scale"])
        lm = L["mp"]
        has = (lm >= 0).astype(np.uint8)
        lidx = np.where(lm >= 0, lm, 0)
        args = (prior_q, prior_t, self.K, has, L["outlier"], M["xw"][lidx], M["desc"][lidx], np.ones(len(lm), np.uint8),
                L["kps"]["octave"], L["kps"]["angle"])
        nm, cur_mp = F.search_by_projection_last(*args, 15.0)
        self.retried = nm < 20   # Tracking.cc:2614-2621
        if nm < 20:
            nm, cur_mp = F.search_by_projection_last(*args, 30.0)
        cur_map = np.where(cur_mp >= 0, lm[np.where(cur_mp >= 0, cur_mp, 0)], -1).astype(np.int64)
        idx = np.nonzero(cur_map >= 0)[0]
        xy = np.stack([kps["x"], kps["y"]], 1)
        w = self.T["inv_sigma2"][kps["octave"]]
        r1, q1, t1, o1, _ = pose_optimization(prior_q, prior_t, self.K, M["xw"][cur_map[idx]], xy[idx], w[idx])
        seen = np.zeros(len(M["xw"]), np.uint8)
        seen[cur_map[idx]] = 1
        cur_map[idx[o1 != 0]] = -1
        cur_map, n2 = self._local_map_search(F, q1, t1, cur_map, 1.0, 0.8, seen)
        idx = np.nonzero(cur_map >= 0)[0]
        r2, q2, t2, o2, _ = pose_optimization(q1, t1, self.K, M["xw"][cur_map[idx]], xy[idx], w[idx])
        outl = np.zeros(len(kps), np.uint8)
        outl[idx] = o2
        inl = int(((cur_map >= 0) & (outl == 0)).sum())
        self.last = dict(kps=kps, mp=cur_map, outlier=outl, q=q2, t=t2)
        return dict(q=q2, t=t2, counts=(len(kps), nm, r1, inl), cur_map=cur_map, outlier=outl)
