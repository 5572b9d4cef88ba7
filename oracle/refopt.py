"""Bindings of oracle/_ref/libref_opt.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The library holds the reference's own Optimizer::PoseOptimization, LocalBundleAdjustment (local mapping and welding BA),
BundleAdjustment and OptimizeSim3 (bodies cut out of O3/src/Optimizer.cc at build time), its edge types
(O3/src/OptimizableTypes.cpp) and its vendored g2o, compiled unmodified from /root/reference over the mini Eigen of
oracle/g2oshim (oracle/Makefile, target `ref`).  It exists only where /root/reference is present (or where the built .so
travelled); tests that need it skip otherwise.  Call signatures mirror oracle.track / oracle.lba / oracle.sim3 so that the
same inputs go to both."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_opt.so")
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
        L.refopt_pose_optimization.argtypes = [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]
        L.refopt_local_ba.argtypes = [C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp]
        L.refopt_local_ba.restype = None
        L.refopt_bundle_adjustment.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int]
        L.refopt_bundle_adjustment.restype = None
        L.refopt_merge_ba.argtypes = [C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp]
        L.refopt_merge_ba.restype = None
        L.refopt_optimize_sim3.argtypes = [C.c_int] + [_vp] * 9 + [C.c_float, C.c_int, _vp, _vp]
        L.refopt_optimize_essential_graph.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp,
                                                      C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp]
        L.refopt_optimize_essential_graph.restype = None
        _LIB = L
    return _LIB


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def pose_optimization(q, t, K, Xw, kp_xy, inv_sigma2):
    """Optimizer::PoseOptimization.  Returns (n_inliers, q[4] f32, t[3] f32, outlier[n] u8) like oracle.track.pose_optimization."""
    L = _L()
    pose = np.concatenate([_c(q, np.float32), _c(t, np.float32)]).astype(np.float32)
    Xw, xy, w = _c(Xw, np.float32), _c(kp_xy, np.float32), _c(inv_sigma2, np.float32)
    n = len(w)
    outl = np.zeros(max(n, 1), np.uint8)
    r = L.refopt_pose_optimization(n, Xw.ctypes.data, xy.ctypes.data, w.ctypes.data, _c(K, np.float32).ctypes.data,
                                   pose.ctypes.data, outl.ctypes.data)
    return r, pose[:4].copy(), pose[4:].copy(), outl[:n]


def _cam_K(K, nc):
    k = _c(K, np.float32)
    return _c(np.broadcast_to(k, (nc, 4)) if k.ndim == 1 else k, np.float32)


def local_ba(cam_q, cam_t, cam_fixed, pts, edge_cam, edge_pt, edge_obs, edge_w, K, abort=0):
    """Optimizer::LocalBundleAdjustment(pKF, ...).  Returns dict(cam_q, cam_t, pts, bad, stats)."""
    L = _L()
    q, t, p = _c(cam_q, np.float32).copy(), _c(cam_t, np.float32).copy(), _c(pts, np.float32).copy()
    fx = _c(cam_fixed, np.uint8)
    ec, ep, eo, ew = _c(edge_cam, np.int32), _c(edge_pt, np.int32), _c(edge_obs, np.float32), _c(edge_w, np.float32)
    bad = np.zeros(max(len(ec), 1), np.uint8)
    stats = np.zeros(4, np.int32)
    ck = _cam_K(K, len(fx))
    L.refopt_local_ba(len(fx), q.ctypes.data, t.ctypes.data, fx.ctypes.data, ck.ctypes.data, len(p), p.ctypes.data, len(ec),
                      ec.ctypes.data, ep.ctypes.data, eo.ctypes.data, ew.ctypes.data, int(abort), bad.ctypes.data, stats.ctypes.data)
    return dict(cam_q=q, cam_t=t, pts=p, bad=bad[:len(ec)], stats=stats)


def bundle_adjustment(cam_q, cam_t, pts, edge_cam, edge_pt, edge_obs, edge_w, K, iterations=5, robust=True):
    """Optimizer::BundleAdjustment with camera 0 as the map's first (fixed) keyframe.  Returns dict(cam_q, cam_t, pts)."""
    L = _L()
    q, t, p = _c(cam_q, np.float32).copy(), _c(cam_t, np.float32).copy(), _c(pts, np.float32).copy()
    ec, ep, eo, ew = _c(edge_cam, np.int32), _c(edge_pt, np.int32), _c(edge_obs, np.float32), _c(edge_w, np.float32)
    ck = _cam_K(K, len(q))
    L.refopt_bundle_adjustment(len(q), q.ctypes.data, t.ctypes.data, ck.ctypes.data, len(p), p.ctypes.data, len(ec),
                               ec.ctypes.data, ep.ctypes.data, eo.ctypes.data, ew.ctypes.data, int(iterations), int(bool(robust)))
    return dict(cam_q=q, cam_t=t, pts=p)


def merge_ba(cam_q, cam_t, cam_fixed, pts, edge_cam, edge_pt, edge_obs, edge_w, K, abort=0):
    """Optimizer::LocalBundleAdjustment(pMainKF, vpAdjustKF, vpFixedKF, pbStopFlag).  Returns dict(cam_q, cam_t, pts, bad)."""
    L = _L()
    q, t, p = _c(cam_q, np.float32).copy(), _c(cam_t, np.float32).copy(), _c(pts, np.float32).copy()
    fx = _c(cam_fixed, np.uint8)
    ec, ep, eo, ew = _c(edge_cam, np.int32), _c(edge_pt, np.int32), _c(edge_obs, np.float32), _c(edge_w, np.float32)
    bad = np.zeros(max(len(ec), 1), np.uint8)
    ck = _cam_K(K, len(fx))
    L.refopt_merge_ba(len(fx), q.ctypes.data, t.ctypes.data, fx.ctypes.data, ck.ctypes.data, len(p), p.ctypes.data, len(ec),
                      ec.ctypes.data, ep.ctypes.data, eo.ctypes.data, ew.ctypes.data, int(abort), bad.ctypes.data)
    return dict(cam_q=q, cam_t=t, pts=p, bad=bad[:len(ec)])


def optimize_sim3(p1c, p2c, obs1, obs2, w1, w2, K1, K2, q, t, s, th2=10.0, fix_scale=False):
    """Optimizer::OptimizeSim3.  Returns dict(n_in, q, t, s, inlier, hessian)."""
    L = _L()
    a = [_c(x, np.float32) for x in (p1c, p2c, obs1, obs2, w1, w2, K1, K2)]
    n = len(a[4])
    sim3 = np.concatenate([_c(q, np.float64), _c(t, np.float64), [float(s)]]).astype(np.float64)
    inl = np.zeros(max(n, 1), np.uint8)
    H = np.zeros(49, np.float64)
    r = L.refopt_optimize_sim3(n, *[x.ctypes.data for x in a], sim3.ctypes.data, float(th2), int(bool(fix_scale)),
                               inl.ctypes.data, H.ctypes.data)
    return dict(n_in=r, q=sim3[:4].copy(), t=sim3[4:7].copy(), s=float(sim3[7]), inlier=inl[:n], hessian=H.reshape(7, 7))


def optimize_essential_graph(kf_q, kf_t, parent, init_kf, loop_kf, cur_kf, loop_edges, cov, non_corrected, corrected, loop_connections,
                             fix_scale, pts, pt_ref, corrected_by_cur, corrected_ref):
    """Optimizer::OptimizeEssentialGraph on keyframes 0..n-1.  loop_edges [m, 2], cov [m, 3] (i, j, weight), non_corrected /
    corrected: dict keyframe -> sim3[8] (q xyzw, t, s), loop_connections [m, 2].  Returns dict(kf_q, kf_t, pts)."""
    L = _L()
    q, t, p = _c(kf_q, np.float32).copy(), _c(kf_t, np.float32).copy(), _c(pts, np.float32).copy()
    par = _c(parent, np.int32)
    le = _c(np.asarray(loop_edges, np.int32).reshape(-1, 2), np.int32)
    cv = _c(np.asarray(cov, np.int32).reshape(-1, 3), np.int32)
    lc = _c(np.asarray(loop_connections, np.int32).reshape(-1, 2), np.int32)
    nck = _c(sorted(non_corrected), np.int32)
    ncs = _c([non_corrected[k] for k in sorted(non_corrected)], np.float64).reshape(-1, 8)
    ck = _c(sorted(corrected), np.int32)
    cs = _c([corrected[k] for k in sorted(corrected)], np.float64).reshape(-1, 8)
    pr, cb, cr = _c(pt_ref, np.int32), _c(corrected_by_cur, np.uint8), _c(corrected_ref, np.int32)
    L.refopt_optimize_essential_graph(len(q), q.ctypes.data, t.ctypes.data, par.ctypes.data, int(init_kf), int(loop_kf), int(cur_kf),
                                      len(le), le.ctypes.data, len(cv), cv.ctypes.data, len(nck), nck.ctypes.data, ncs.ctypes.data,
                                      len(ck), ck.ctypes.data, cs.ctypes.data, len(lc), lc.ctypes.data, int(bool(fix_scale)), len(p),
                                      p.ctypes.data, pr.ctypes.data, cb.ctypes.data, cr.ctypes.data)
    return dict(kf_q=q, kf_t=t, pts=p)
