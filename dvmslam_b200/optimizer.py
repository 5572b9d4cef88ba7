"""Python mirror of Optimizer::LocalBundleAdjustment (O3/include/Optimizer.h:56-57) over the C-ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib

_vp = C.c_void_p
_ip = C.POINTER(C.c_int)


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


class LocalBA:
    """One solver context per agent/GPU (dvm_lba)."""

    def __init__(self, max_free_cameras: int = 64, device: int = 0):
        self.L = lib()
        L = self.L
        L.dvm_lba_create.argtypes = [C.POINTER(_vp), C.c_int, C.c_int]
        L.dvm_lba_destroy.argtypes = [_vp]
        L.dvm_lba_destroy.restype = None
        L.dvm_local_ba.argtypes = [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int,
                                   _vp, _vp, _vp, _vp, _ip]
        L.dvm_bundle_adjustment.argtypes = [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int,
                                            C.c_float, _vp, _vp, _vp, _vp, _ip]
        L.dvm_merge_ba.argtypes = [_vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _ip]
        L.dvm_lba_last_kernel_ms.argtypes = [_vp]
        L.dvm_lba_last_kernel_ms.restype = C.c_float
        self.h = _vp()
        check(L.dvm_lba_create(C.byref(self.h), device, max_free_cameras))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.dvm_lba_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def BundleAdjustment(self, cam_q, cam_t, cam_fixed, pts, edge_cam, edge_pt, edge_obs, edge_w, K, nIterations=5,
                         bRobust=True, abort=None):
        """Optimizer::BundleAdjustment (global BA) for maps of up to max_free_cameras keyframes."""
        delta = float(np.float32(np.sqrt(5.99))) if bRobust else float("inf")
        return self.LocalBundleAdjustment(cam_q, cam_t, cam_fixed, pts, edge_cam, edge_pt, edge_obs, edge_w, K, nIterations,
                                          abort, huber_delta=delta)

    def LocalBundleAdjustment(self, cam_q, cam_t, cam_fixed, pts, edge_cam, edge_pt, edge_obs, edge_w, K, iterations=10,
                              abort=None, huber_delta=None):
        q, t = _c(cam_q, np.float32).copy(), _c(cam_t, np.float32).copy()
        p = _c(pts, np.float32).copy()
        fx = _c(cam_fixed, np.uint8)
        ec, ep = _c(edge_cam, np.int32), _c(edge_pt, np.int32)
        eo, ew = _c(edge_obs, np.float32), _c(edge_w, np.float32)
        ne = len(ec)
        chi2 = np.empty(max(ne, 1), np.float64)   # fully written by the call unless it is a no-op (handled below)
        bad = np.empty(max(ne, 1), np.uint8)
        stats = np.zeros(4, np.float64)
        iters = C.c_int()
        ab = _c([abort], np.uint8) if abort is not None else None   # pbStopFlag: a one-byte bool
        delta = float(np.float32(np.sqrt(5.991))) if huber_delta is None else float(huber_delta)
        check(self.L.dvm_bundle_adjustment(self.h, len(fx), q.ctypes.data, t.ctypes.data, fx.ctypes.data, len(p),
                                           p.ctypes.data, ne, ec.ctypes.data, ep.ctypes.data, eo.ctypes.data,
                                           ew.ctypes.data, _c(K, np.float32).ctypes.data, iterations, C.c_float(delta),
                                           ab.ctypes.data if ab is not None else None, chi2.ctypes.data,
                                           bad.ctypes.data, stats.ctypes.data, C.byref(iters)))
        if iters.value < 0:   # no fixed keyframe / stop flag already set: nothing was computed
            chi2[:] = 0
            bad[:] = 0
        return dict(cam_q=q, cam_t=t, pts=p, chi2=chi2[:ne], bad=bad[:ne], iters=int(stats[0]), trials=int(stats[1]),
                    chi_first=stats[2], chi_last=stats[3], rc=iters.value, kernel_ms=self.kernel_ms())

    def MergeBundleAdjustment(self, cam_q, cam_t, cam_fixed, pts, edge_cam, edge_pt, edge_obs, edge_w, K, abort=None):
        """Optimizer::LocalBundleAdjustment(pMainKF, vpAdjustKF, vpFixedKF, pbStopFlag) -- the welding BA of a map merge
        (O3/src/Optimizer.cc:3257-3675): cam_fixed = 1 for vpFixedKF, 0 for vpAdjustKF.  Returns LocalBundleAdjustment's
        dict plus iters_first (LM iterations of the Huber pass) and excluded (edges moved to level 1)."""
        q, t = _c(cam_q, np.float32).copy(), _c(cam_t, np.float32).copy()
        p = _c(pts, np.float32).copy()
        fx = _c(cam_fixed, np.uint8)
        ec, ep = _c(edge_cam, np.int32), _c(edge_pt, np.int32)
        eo, ew = _c(edge_obs, np.float32), _c(edge_w, np.float32)
        ne = len(ec)
        chi2 = np.empty(max(ne, 1), np.float64)
        bad = np.empty(max(ne, 1), np.uint8)
        stats = np.zeros(6, np.float64)
        iters = C.c_int()
        ab = _c([abort], np.uint8) if abort is not None else None
        check(self.L.dvm_merge_ba(self.h, len(fx), q.ctypes.data, t.ctypes.data, fx.ctypes.data, len(p), p.ctypes.data, ne,
                                  ec.ctypes.data, ep.ctypes.data, eo.ctypes.data, ew.ctypes.data,
                                  _c(K, np.float32).ctypes.data, ab.ctypes.data if ab is not None else None,
                                  chi2.ctypes.data, bad.ctypes.data, stats.ctypes.data, C.byref(iters)))
        if iters.value < 0:
            chi2[:] = 0
            bad[:] = 0
        return dict(cam_q=q, cam_t=t, pts=p, chi2=chi2[:ne], bad=bad[:ne], iters=int(stats[0]), trials=int(stats[1]),
                    chi_first=stats[2], chi_last=stats[3], iters_first=int(stats[4]), excluded=int(stats[5]),
                    rc=iters.value, kernel_ms=self.kernel_ms())

    def set_camera_intrinsics(self, cam_K):
        """cam_K[nc][4] = fx, fy, cx, cy of every camera for the NEXT bundle-adjustment call (each edge is projected with its
        own keyframe's camera, O3/src/Optimizer.cc:1219)."""
        k = _c(cam_K, np.float32)
        self.L.dvm_lba_set_camera_intrinsics.argtypes = [_vp, C.c_int, _vp]
        check(self.L.dvm_lba_set_camera_intrinsics(self.h, len(k), k.ctypes.data))

    def kernel_ms(self) -> float:
        return float(self.L.dvm_lba_last_kernel_ms(self.h))


class Sim3Optimizer:
    """Optimizer::OptimizeSim3 (O3/src/Optimizer.cc:1960-2212) over the C-ABI (dvm_optimize_sim3)."""

    def __init__(self, device: int = 0):
        self.L = lib()
        L = self.L
        L.dvm_sim3_create.argtypes = [C.POINTER(_vp), C.c_int]
        L.dvm_sim3_destroy.argtypes = [_vp]
        L.dvm_sim3_destroy.restype = None
        L.dvm_optimize_sim3.argtypes = [_vp, C.c_int] + [_vp] * 11 + [C.c_float, C.c_int, _vp, _ip, _vp]
        self.h = _vp()
        check(L.dvm_sim3_create(C.byref(self.h), device))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.dvm_sim3_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def OptimizeSim3(self, p1c, p2c, obs1, obs2, w1, w2, K1, K2, q, t, s, th2=10.0, bFixScale=False):
        """Flattened correspondences (see include/dvmslam_b200.h).  Returns dict(q, t, s, inlier, n_in, iters1, iters2,
        trials, n_bad, chi_first, chi_last)."""
        a = [_c(x, np.float32) for x in (p1c, p2c, obs1, obs2, w1, w2, K1, K2)]
        n = len(a[4])
        qq, tt, ss = _c(q, np.float64).copy(), _c(t, np.float64).copy(), np.array([s], np.float64)
        inl = np.zeros(max(n, 1), np.uint8)
        st = np.zeros(6, np.float64)
        n_in = C.c_int()
        check(self.L.dvm_optimize_sim3(self.h, n, *[x.ctypes.data for x in a], qq.ctypes.data, tt.ctypes.data, ss.ctypes.data,
                                       C.c_float(th2), int(bFixScale), inl.ctypes.data, C.byref(n_in), st.ctypes.data))
        return dict(q=qq, t=tt, s=float(ss[0]), inlier=inl[:n], n_in=n_in.value, iters1=int(st[0]), iters2=int(st[1]),
                    trials=int(st[2]), n_bad=int(st[3]), chi_first=st[4], chi_last=st[5])


class EssentialGraphOptimizer:
    """The solve of Optimizer::OptimizeEssentialGraph (O3/src/Optimizer.cc:1389-1651) over the C-ABI
    (dvm_optimize_essential_graph): Sim3 pose graph, numeric Jacobians, LM with lambda_init 1e-16, 20 iterations."""

    def __init__(self, device: int = 0):
        self.L = lib()
        L = self.L
        L.dvm_essential_graph_create.argtypes = [C.POINTER(_vp), C.c_int]
        L.dvm_essential_graph_destroy.argtypes = [_vp]
        L.dvm_essential_graph_destroy.restype = None
        L.dvm_optimize_essential_graph.argtypes = [_vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_double, _vp]
        L.dvm_essential_graph_last_kernel_ms.argtypes = [_vp]
        L.dvm_essential_graph_last_kernel_ms.restype = C.c_float
        self.h = _vp()
        check(L.dvm_essential_graph_create(C.byref(self.h), device))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.dvm_essential_graph_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def OptimizeEssentialGraph(self, sim3, fixed, vi, vj, meas, bFixScale=False, iterations=20, lambda_init=1e-16):
        """sim3 [nv, 8] (q xyzw, t, s) = Scw; edges (vi, vj, Sji [ne, 8]).  Returns dict(sim3, iters, trials, chi_first,
        chi_last, kernel_ms)."""
        S = _c(sim3, np.float64).copy()
        fx = _c(fixed, np.uint8)
        a, b = _c(vi, np.int32), _c(vj, np.int32)
        m = _c(meas, np.float64)
        st = np.zeros(4)
        check(self.L.dvm_optimize_essential_graph(self.h, len(S), S.ctypes.data, fx.ctypes.data, len(a), a.ctypes.data, b.ctypes.data,
                                                  m.ctypes.data, int(bFixScale), int(iterations), float(lambda_init), st.ctypes.data))
        return dict(sim3=S, iters=int(st[0]), trials=int(st[1]), chi_first=st[2], chi_last=st[3],
                    kernel_ms=float(self.L.dvm_essential_graph_last_kernel_ms(self.h)))
