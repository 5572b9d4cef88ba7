"""Local-BA oracle bindings (oracle/lba_oracle.cpp) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib

_vp = C.c_void_p


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def set_camera_intrinsics(cam_K):
    """cam_K[nc][4] for the NEXT local_ba / merge_ba call of the oracle (mirrors dvm_lba_set_camera_intrinsics)."""
    L = lib()
    L.lbao_set_camera_intrinsics.argtypes = [C.c_int, _vp]
    L.lbao_set_camera_intrinsics.restype = None
    k = _c(cam_K, np.float32)
    L.lbao_set_camera_intrinsics(len(k), k.ctypes.data)


def local_ba(cam_q, cam_t, cam_fixed, pts, edge_cam, edge_pt, edge_obs, edge_w, K, iterations=10, abort=None,
             huber_delta=None):
    """Returns dict(cam_q, cam_t, pts, chi2, bad, iters, trials, chi_first, chi_last, rc).  huber_delta: None =
    LocalBundleAdjustment's float sqrt(5.991); a float = BundleAdjustment's delta (inf = no robust kernel)."""
    L = lib()
    L.lbao_bundle_adjustment.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int,
                                         C.c_float, _vp, _vp, _vp, _vp]
    delta = float(np.float32(np.sqrt(5.991))) if huber_delta is None else float(np.float32(huber_delta))
    q, t = _c(cam_q, np.float32).copy(), _c(cam_t, np.float32).copy()
    p = _c(pts, np.float32).copy()
    fx = _c(cam_fixed, np.uint8)
    ec, ep = _c(edge_cam, np.int32), _c(edge_pt, np.int32)
    eo, ew = _c(edge_obs, np.float32), _c(edge_w, np.float32)
    ne = len(ec)
    chi2 = np.zeros(max(ne, 1), np.float64)
    bad = np.zeros(max(ne, 1), np.uint8)
    stats = np.zeros(4, np.float64)
    ab = _c([abort], np.int32) if abort is not None else None
    rc = L.lbao_bundle_adjustment(len(fx), q.ctypes.data, t.ctypes.data, fx.ctypes.data, len(p), p.ctypes.data, ne,
                                  ec.ctypes.data, ep.ctypes.data, eo.ctypes.data, ew.ctypes.data,
                                  _c(K, np.float32).ctypes.data, iterations, delta,
                                  ab.ctypes.data if ab is not None else None, chi2.ctypes.data, bad.ctypes.data,
                                  stats.ctypes.data)
    return dict(cam_q=q, cam_t=t, pts=p, chi2=chi2[:ne], bad=bad[:ne], iters=int(stats[0]), trials=int(stats[1]),
                chi_first=stats[2], chi_last=stats[3], rc=rc)


def merge_ba(cam_q, cam_t, cam_fixed, pts, edge_cam, edge_pt, edge_obs, edge_w, K, abort=None):
    """The welding BA (Optimizer.cc:3257-3675).  As local_ba plus iters_first (LM iterations of the Huber pass) and
    excluded (edges moved to level 1 before the second pass)."""
    L = lib()
    L.lbao_merge_ba.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    q, t = _c(cam_q, np.float32).copy(), _c(cam_t, np.float32).copy()
    p = _c(pts, np.float32).copy()
    fx = _c(cam_fixed, np.uint8)
    ec, ep = _c(edge_cam, np.int32), _c(edge_pt, np.int32)
    eo, ew = _c(edge_obs, np.float32), _c(edge_w, np.float32)
    ne = len(ec)
    chi2 = np.zeros(max(ne, 1), np.float64)
    bad = np.zeros(max(ne, 1), np.uint8)
    stats = np.zeros(6, np.float64)
    ab = _c([abort], np.int32) if abort is not None else None
    rc = L.lbao_merge_ba(len(fx), q.ctypes.data, t.ctypes.data, fx.ctypes.data, len(p), p.ctypes.data, ne, ec.ctypes.data,
                         ep.ctypes.data, eo.ctypes.data, ew.ctypes.data, _c(K, np.float32).ctypes.data,
                         ab.ctypes.data if ab is not None else None, chi2.ctypes.data, bad.ctypes.data, stats.ctypes.data)
    return dict(cam_q=q, cam_t=t, pts=p, chi2=chi2[:ne], bad=bad[:ne], iters=int(stats[0]), trials=int(stats[1]),
                chi_first=stats[2], chi_last=stats[3], iters_first=int(stats[4]), excluded=int(stats[5]), rc=rc)


def edge_jacobians(q, t, X, K):
    L = lib()
    L.lbao_edge_jacobians.argtypes = [_vp] * 6
    A, B = np.zeros((2, 3)), np.zeros((2, 6))
    L.lbao_edge_jacobians(_c(q, np.float32).ctypes.data, _c(t, np.float32).ctypes.data, _c(X, np.float32).ctypes.data,
                          _c(K, np.float32).ctypes.data, A.ctypes.data, B.ctypes.data)
    return A, B


def edge_error_perturbed(q, t, X, K, obs, d6, d3):
    L = lib()
    L.lbao_edge_error_perturbed.argtypes = [_vp] * 8
    e = np.zeros(2)
    L.lbao_edge_error_perturbed(_c(q, np.float32).ctypes.data, _c(t, np.float32).ctypes.data,
                                _c(X, np.float32).ctypes.data, _c(K, np.float32).ctypes.data,
                                _c(obs, np.float32).ctypes.data, _c(d6, np.float64).ctypes.data,
                                _c(d3, np.float64).ctypes.data, e.ctypes.data)
    return e
