"""Python mirrors of the per-frame tracking operators over the C-ABI: Frame (grid),
ORBmatcher.SearchByProjection (both overloads) and Optimizer.PoseOptimization.  Argument meaning
follows the reference (O3/include/ORBmatcher.h:40-95, O3/include/Optimizer.h:59); the pointer-graph
arguments (Frame&, vector<MapPoint*>) are flattened to arrays as include/dvmslam_b200.h documents."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib
from .extractor import KP_DTYPE

_vp = C.c_void_p
_ip = C.POINTER(C.c_int)


def _bind(L):
    if getattr(L, "_trk_bound", False):
        return
    L.dvm_frame_create.argtypes = [C.POINTER(_vp), C.c_int, _vp, C.c_int, C.c_int, _vp, _vp]
    L.dvm_frame_destroy.argtypes = [_vp]
    L.dvm_frame_destroy.restype = None
    L.dvm_frame_assign.argtypes = [_vp, _vp, _vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
    L.dvm_frame_assign_from_orb.argtypes = [_vp, _vp, C.c_float, C.c_float, C.c_float, C.c_float]
    L.dvm_frame_features_in_area.argtypes = [_vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp, C.c_int, _ip]
    L.dvm_frame_grid_cell.argtypes = [_vp, C.c_int, C.c_int, _vp, C.c_int, _ip]
    L.dvm_match_by_projection_last.argtypes = [_vp, _vp, _vp, _vp, C.c_int] + [_vp] * 7 + [C.c_float, C.c_int, _vp, _ip]
    L.dvm_match_by_projection_map.argtypes = [_vp, C.c_int] + [_vp] * 6 + [C.c_float, C.c_float, _vp, _vp, _ip]
    L.dvm_match_last_rounds.argtypes = [_vp]
    L.dvm_pose_optimization.argtypes = [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _ip, _vp]
    L.dvm_undistort_keypoints.argtypes = [_vp, _vp, C.c_int, _vp, _vp]
    L.dvm_frame_set_distortion.argtypes = [_vp, _vp, _vp]
    L.dvm_image_bounds.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, _vp]
    L._trk_bound = True


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


class Frame:
    """Mono Frame: undistorted keypoints + descriptors + 64x48 grid, resident on one B200."""

    def __init__(self, max_keypoints: int, scale_factors, inv_level_sigma2, device: int = 0, stream: int = 0):
        self.L = lib()
        _bind(self.L)
        self.h = _vp()
        sf, isg = _c(scale_factors, np.float32), _c(inv_level_sigma2, np.float32)
        check(self.L.dvm_frame_create(C.byref(self.h), device, _vp(stream) if stream else None, max_keypoints,
                                      len(sf), sf.ctypes.data, isg.ctypes.data))
        self.cap = (max_keypoints + 3) & ~3
        self.n = 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.dvm_frame_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def assign(self, kps_un, desc, bounds):
        k, d = _c(kps_un, KP_DTYPE), _c(desc, np.uint8)
        self.n = len(k)
        check(self.L.dvm_frame_assign(self.h, k.ctypes.data, d.ctypes.data, self.n, *map(float, bounds)))

    def assign_from_orb(self, extractor, bounds):
        check(self.L.dvm_frame_assign_from_orb(self.h, extractor.h, *map(float, bounds)))
        self.n = self.cap

    def UndistortKeyPoints(self, kps, K, dist5):
        """Frame::UndistortKeyPoints -> mvKeysUn (a copy; the identity when k1 == 0)."""
        k = _c(kps, KP_DTYPE).copy()
        check(self.L.dvm_undistort_keypoints(self.h, k.ctypes.data, len(k), _c(K, np.float32).ctypes.data,
                                             _c(dist5, np.float32).ctypes.data))
        return k

    def ComputeImageBounds(self, K, dist5, width, height):
        b = np.zeros(4, np.float32)
        check(self.L.dvm_image_bounds(self.h, _c(K, np.float32).ctypes.data, _c(dist5, np.float32).ctypes.data, int(width),
                                      int(height), b.ctypes.data))
        return tuple(float(x) for x in b)

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        out = np.zeros(self.cap, np.int32)
        n = C.c_int()
        check(self.L.dvm_frame_features_in_area(self.h, x, y, r, minLevel, maxLevel, out.ctypes.data, self.cap,
                                                C.byref(n)))
        return out[:n.value].copy()

    def grid_cell(self, ix, iy):
        out = np.zeros(self.cap, np.int32)
        n = C.c_int()
        check(self.L.dvm_frame_grid_cell(self.h, ix, iy, out.ctypes.data, self.cap, C.byref(n)))
        return out[:n.value].copy()


class ORBmatcher:
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True):
        self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri
        self.L = lib()
        _bind(self.L)

    def SearchByProjectionLast(self, cur: Frame, qcw, tcw, K, has_mp, outlier, Xw, mp_desc, obs_pos, last_octave,
                               last_angle, th):
        """SearchByProjection(CurrentFrame, LastFrame, th, bMono=True) -> (nmatches, cur_mp).  qcw (x, y, z, w) / tcw:
        CurrentFrame.GetPose() as its SE3f holds it."""
        a = [_c(qcw, np.float32), _c(tcw, np.float32), _c(K, np.float32)]
        assert a[0].shape == (4,), "the pose is a quaternion (x, y, z, w), not a rotation matrix"
        b = [_c(has_mp, np.uint8), _c(outlier, np.uint8), _c(Xw, np.float32), _c(mp_desc, np.uint8),
             _c(obs_pos, np.uint8), _c(last_octave, np.int32), _c(last_angle, np.float32)]
        cur_mp = np.full(cur.cap, -1, np.int32)
        n = C.c_int()
        check(self.L.dvm_match_by_projection_last(cur.h, *(x.ctypes.data for x in a), len(b[0]),
                                                  *(x.ctypes.data for x in b), float(th),
                                                  int(self.mbCheckOrientation), cur_mp.ctypes.data, C.byref(n)))
        return n.value, cur_mp[:cur.n]

    def SearchByProjectionMap(self, cur: Frame, projX, projY, level, view_cos, mp_desc, obs_pos, th, cur_blocked=None):
        """SearchByProjection(F, vpMapPoints, th) -> (nmatches, cur_mp)."""
        b = [_c(projX, np.float32), _c(projY, np.float32), _c(level, np.int32), _c(view_cos, np.float32),
             _c(mp_desc, np.uint8), _c(obs_pos, np.uint8)]
        blk = _c(cur_blocked, np.uint8) if cur_blocked is not None else None
        cur_mp = np.full(cur.cap, -1, np.int32)
        n = C.c_int()
        check(self.L.dvm_match_by_projection_map(cur.h, len(b[0]), *(x.ctypes.data for x in b), float(th),
                                                 float(self.mfNNratio), blk.ctypes.data if blk is not None else None,
                                                 cur_mp.ctypes.data, C.byref(n)))
        return n.value, cur_mp[:cur.n]

    def rounds(self, cur: Frame) -> int:
        return self.L.dvm_match_last_rounds(cur.h)


def PoseOptimization(frame: Frame, q, t, K, Xw, kp_xy, inv_sigma2):
    """Optimizer::PoseOptimization -> (n_inliers, q, t, outlier, (lm_iterations, lm_trials))."""
    L = lib()
    _bind(L)
    q, t = _c(q, np.float32).copy(), _c(t, np.float32).copy()
    Xw, kp_xy, w = _c(Xw, np.float32), _c(kp_xy, np.float32), _c(inv_sigma2, np.float32)
    n = len(w)
    out = np.zeros(max(n, 1), np.uint8)
    stats = np.zeros(2, np.int32)
    ninl = C.c_int()
    check(L.dvm_pose_optimization(frame.h, q.ctypes.data, t.ctypes.data, _c(K, np.float32).ctypes.data, n,
                                  Xw.ctypes.data, kp_xy.ctypes.data, w.ctypes.data, out.ctypes.data, C.byref(ninl),
                                  stats.ctypes.data))
    return ninl.value, q, t, out[:n], tuple(int(s) for s in stats)


def is_in_frustum(frame: Frame, q, t, K, xw, normal, min_dist, max_dist, skip=None, cos_limit=0.5):
    """Frame::isInFrustum + MapPoint::PredictScale over a batch -> (in_view, projX, projY, level, viewCos)."""
    L = lib()
    _bind(L)
    L.dvm_frame_is_in_frustum.argtypes = [_vp, _vp, _vp, _vp, C.c_int] + [_vp] * 5 + [C.c_float] + [_vp] * 5
    xw, normal = _c(xw, np.float32), _c(normal, np.float32)
    m = len(xw)
    inv, px, py = np.zeros(max(m, 1), np.uint8), np.zeros(max(m, 1), np.float32), np.zeros(max(m, 1), np.float32)
    lv, vc = np.zeros(max(m, 1), np.int32), np.zeros(max(m, 1), np.float32)
    sk = _c(skip, np.uint8) if skip is not None else None
    check(L.dvm_frame_is_in_frustum(frame.h, _c(q, np.float32).ctypes.data, _c(t, np.float32).ctypes.data,
                                    _c(K, np.float32).ctypes.data, m, xw.ctypes.data, normal.ctypes.data,
                                    _c(min_dist, np.float32).ctypes.data, _c(max_dist, np.float32).ctypes.data,
                                    sk.ctypes.data if sk is not None else None, float(cos_limit), inv.ctypes.data,
                                    px.ctypes.data, py.ctypes.data, lv.ctypes.data, vc.ctypes.data))
    return inv[:m], px[:m], py[:m], lv[:m], vc[:m]


class Tracker:
    """dvm_tracker: ExtractORB -> Frame -> TrackWithMotionModel -> TrackLocalMap chained on the GPU."""

    def __init__(self, extractor, K, bounds, world_map, dist_coef=None):
        self.L = lib()
        _bind(self.L)
        L = self.L
        L.dvm_tracker_create.argtypes = [C.POINTER(_vp), _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]
        L.dvm_tracker_destroy.argtypes = [_vp]
        L.dvm_tracker_destroy.restype = None
        L.dvm_tracker_bootstrap.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _ip]
        L.dvm_tracker_track.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp]
        L.dvm_tracker_result.argtypes = [_vp, _vp, _vp]
        L.dvm_tracker_result_lag.argtypes = [_vp, C.c_int, _vp, _vp]
        L.dvm_tracker_prefetch.argtypes = [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.dvm_tracker_stream.argtypes = [_vp]
        L.dvm_tracker_stream.restype = _vp
        L.dvm_tracker_debug_matches.argtypes = [_vp, _vp, _vp, C.c_int, _ip]
        self.ext = extractor
        self.h = _vp()
        m = world_map
        self._keep = [_c(K, np.float32), _c(bounds, np.float32), _c(m["xw"], np.float32), _c(m["desc"], np.uint8),
                      _c(m["normal"], np.float32), _c(m["min_dist"], np.float32), _c(m["max_dist"], np.float32)]
        k = self._keep
        check(L.dvm_tracker_create(C.byref(self.h), extractor.h, k[0].ctypes.data, k[1].ctypes.data, len(k[2]),
                                   k[2].ctypes.data, k[3].ctypes.data, k[4].ctypes.data, k[5].ctypes.data,
                                   k[6].ctypes.data))
        if dist_coef is not None:   # mDistCoef (k1, k2, p1, p2, k3): frames are undistorted on the device
            L.dvm_tracker_set_distortion.argtypes = [_vp, _vp]
            check(L.dvm_tracker_set_distortion(self.h, _c(dist_coef, np.float32).ctypes.data))
        self._pose = np.zeros(7, np.float32)
        self._counts = np.zeros(4, np.int32)

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.dvm_tracker_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bootstrap(self, img, q, t):
        img = np.ascontiguousarray(img, np.uint8)
        n = C.c_int()
        check(self.L.dvm_tracker_bootstrap(self.h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                                           _c(q, np.float32).ctypes.data, _c(t, np.float32).ctypes.data, C.byref(n)))
        return n.value

    @staticmethod
    def _image_args(img):
        if img is None:
            return None, 0, 0, 0, 0
        if isinstance(img, tuple):
            ptr, w, h, stride = img
            return ptr, w, h, stride, 1
        assert img.dtype == np.uint8 and img.strides[1] == 1, "CV_8UC1 rows expected"
        return img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], 0

    def stream(self) -> int:
        """cudaStream_t of the tracking chain (the extractor runs on its own stream)."""
        return int(self.L.dvm_tracker_stream(self.h) or 0)

    def prefetch(self, img):
        """Enqueue upload + ExtractORB of the NEXT frame; the next track() call then takes img=None."""
        ptr, w, h, stride, dev = self._image_args(img)
        check(self.L.dvm_tracker_prefetch(self.h, _vp(ptr), dev, w, h, stride))

    def track(self, img, prior_q=None, prior_t=None, sync=True):
        """img: host uint8 array, (device_ptr, width, height, stride) for an image already in HBM, or None
        when the frame was handed over with prefetch()."""
        ptr, w, h, stride, dev = self._image_args(img)
        pq = _c(prior_q, np.float32) if prior_q is not None else None
        pt = _c(prior_t, np.float32) if prior_t is not None else None
        check(self.L.dvm_tracker_track(self.h, _vp(ptr), dev, w, h, stride, pq.ctypes.data if pq is not None else None,
                                       pt.ctypes.data if pt is not None else None, int(sync), self._pose.ctypes.data,
                                       self._counts.ctypes.data))
        if sync:
            return self._pose[:4].copy(), self._pose[4:].copy(), tuple(int(c) for c in self._counts)
        return None

    def set_profiling(self, on: bool):
        self.L.dvm_tracker_set_profiling.argtypes = [_vp, C.c_int]
        check(self.L.dvm_tracker_set_profiling(self.h, int(on)))

    def get_profile(self):
        """(frames, [ms per segment]): prior/reset, SearchByProjection(last), PoseOptimization, SearchLocalPoints, PoseOptimization."""
        seg = np.zeros(5, np.float64)
        n = C.c_longlong()
        self.L.dvm_tracker_get_profile.argtypes = [_vp, _vp, _vp]
        check(self.L.dvm_tracker_get_profile(self.h, seg.ctypes.data, C.addressof(n)))
        return int(n.value), seg

    def result(self, lag=0):
        """Pose and counts of the frame `lag` frames before the last one handed to track(); waits for that frame only."""
        check(self.L.dvm_tracker_result_lag(self.h, int(lag), self._pose.ctypes.data, self._counts.ctypes.data))
        return self._pose[:4].copy(), self._pose[4:].copy(), tuple(int(c) for c in self._counts)

    def debug_matches(self):
        cap = self.ext.cap
        cm, ol = np.zeros(cap, np.int32), np.zeros(cap, np.uint8)
        n = C.c_int()
        check(self.L.dvm_tracker_debug_matches(self.h, cm.ctypes.data, ol.ctypes.data, cap, C.byref(n)))
        return cm[:n.value].copy(), ol[:n.value].copy()
