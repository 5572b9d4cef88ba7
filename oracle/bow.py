"""Bindings of oracle/bow_oracle.cpp and trko_search_for_initialization -- TEST INFRASTRUCTURE, NOT PRODUCT CODE."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib
from .orb import KP_DTYPE

_vp = C.c_void_p


def _L():
    L = lib()
    if getattr(L, "_bow_bound", False):
        return L
    L.bowo_search_by_bow.argtypes = [C.c_int] + [C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp] * 2 + [C.c_float, C.c_int, _vp, _vp]
    L.bowo_hamming_knn.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp]
    L.bowo_hamming_knn.restype = None
    L.trko_search_for_initialization.argtypes = [C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_int, _vp]
    L.bowo_search_for_triangulation.argtypes = [C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp] * 2 + [_vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp]
    L.trko_fuse_search.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_float, _vp, C.c_int] + [_vp] * 6 + [C.c_float, _vp, _vp]
    L.trko_fuse_search.restype = None
    L._bow_bound = True
    return L


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def _csr(fv):
    nodes = sorted(fv)
    node_id = np.array(nodes, np.uint32)
    start = np.zeros(len(nodes) + 1, np.int32)
    for i, k in enumerate(nodes):
        start[i + 1] = start[i] + len(fv[k])
    idx = np.concatenate([np.asarray(fv[k], np.uint32) for k in nodes]) if nodes else np.zeros(0, np.uint32)
    return node_id, start, _c(idx, np.uint32)


def search_by_bow(kf_kf, desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2, nnratio=0.6, check_ori=True):
    """-> (nmatches, match12, match21); fv = {node: [feature indices]}"""
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
        args += [n, d.ctypes.data, a.ctypes.data, v.ctypes.data if v is not None else None, nn, nid.ctypes.data,
                 st.ctypes.data, idx.ctypes.data]
    n = L.bowo_search_by_bow(int(kf_kf), *args, float(nnratio), int(check_ori), m12.ctypes.data, m21.ctypes.data)
    return n, m12[:sides[0][0]], m21[:sides[1][0]]


def hamming_knn(a, b):
    L = _L()
    a, b = _c(a, np.uint8), _c(b, np.uint8)
    out = [np.zeros(max(len(a), 1), np.int32) for _ in range(3)]
    L.bowo_hamming_knn(a.ctypes.data, len(a), b.ctypes.data, len(b), *(o.ctypes.data for o in out))
    return tuple(o[:len(a)] for o in out)


def search_for_initialization(kps1, desc1, frame2_oracle, prev_matched, window=100, nnratio=0.9, check_ori=True):
    """frame2_oracle: oracle.track.FrameOracle of F2 -> (nmatches, matches12, prev_matched)"""
    L = _L()
    k1, d1 = _c(kps1, KP_DTYPE), _c(desc1, np.uint8)
    pm = _c(prev_matched, np.float32).copy()
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = L.trko_search_for_initialization(len(k1), k1.ctypes.data, d1.ctypes.data, frame2_oracle.h, pm.ctypes.data,
                                         int(window), float(nnratio), int(check_ori), m12.ctypes.data)
    return n, m12[:len(k1)], pm


def search_for_triangulation(desc1, kps1, has_mp1, fv1, desc2, kps2, has_mp2, fv2, F12, ep, scale_factors2, level_sigma2_2,
                             coarse=False, check_ori=True):
    """-> (nmatches, matches12)"""
    L = _L()
    args, keep = [], []
    for d, k, v, fv in ((desc1, kps1, has_mp1, fv1), (desc2, kps2, has_mp2, fv2)):
        d, k, v = _c(d, np.uint8), _c(k, KP_DTYPE), _c(v, np.uint8)
        nid, st, idx = _csr(fv)
        keep += [d, k, v, nid, st, idx]
        args += [len(k), d.ctypes.data, k.ctypes.data, v.ctypes.data, len(nid), nid.ctypes.data, st.ctypes.data, idx.ctypes.data]
    n1 = args[0]
    m12 = np.full(max(n1, 1), -1, np.int32)
    F12, ep = _c(F12, np.float32), _c(ep, np.float32)
    sf, s2 = _c(scale_factors2, np.float32), _c(level_sigma2_2, np.float32)
    n = L.bowo_search_for_triangulation(*args, F12.ctypes.data, ep.ctypes.data, sf.ctypes.data, s2.ctypes.data, int(coarse),
                                        int(check_ori), m12.ctypes.data)
    return n, m12[:n1]


def fundamental_from_poses(q1, t1, q2, t2, K1, K2):
    """(F12 row-major float32[9], ep float32[2]) from the two keyframe poses Tcw as stored, in the reference's float32
    arithmetic (Sophus products / inverse, Eigen 3x3 inverse and products; bowo_fundamental_from_poses)."""
    L = _L()
    L.bowo_fundamental_from_poses.argtypes = [_vp] * 8
    L.bowo_fundamental_from_poses.restype = None
    a = [_c(x, np.float32) for x in (q1, t1, q2, t2, K1, K2)]
    F12, ep = np.zeros(9, np.float32), np.zeros(2, np.float32)
    L.bowo_fundamental_from_poses(*(x.ctypes.data for x in a), F12.ctypes.data, ep.ctypes.data)
    return F12, ep


def fuse_search(frame_oracle, q, t, K, log_scale, inv_sigma2, xw, normal, min_dist, max_dist, mp_desc, skip=None, th=3.0):
    """frame_oracle: oracle.track.FrameOracle of the keyframe -> (best_idx, best_dist)"""
    L = _L()
    a = [_c(xw, np.float32), _c(normal, np.float32), _c(min_dist, np.float32), _c(max_dist, np.float32), _c(mp_desc, np.uint8)]
    m = len(a[2])
    sk = _c(skip, np.uint8) if skip is not None else None
    isg = _c(inv_sigma2, np.float32)
    bi, bd = np.full(max(m, 1), -1, np.int32), np.full(max(m, 1), 256, np.int32)
    L.trko_fuse_search(frame_oracle.h, _c(q, np.float32).ctypes.data, _c(t, np.float32).ctypes.data,
                       _c(K, np.float32).ctypes.data, len(isg), float(log_scale), isg.ctypes.data, m,
                       *(x.ctypes.data for x in a), sk.ctypes.data if sk is not None else None, float(th), bi.ctypes.data,
                       bd.ctypes.data)
    return bi[:m], bd[:m]


def _pts(xw, normal, min_dist, max_dist, mp_desc, skip):
    a = [_c(xw, np.float32), _c(normal, np.float32), _c(min_dist, np.float32), _c(max_dist, np.float32), _c(mp_desc, np.uint8)]
    sk = _c(skip, np.uint8) if skip is not None else None
    return a, sk, len(a[2])


def search_by_projection_sim3(frame_oracle, sq, st, K, log_scale, nlevels, xw, normal, min_dist, max_dist, mp_desc, skip, kp_matched,
                              th, ratio_hamming=1.0):
    """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) -> (nmatches, kp_point[kf n])."""
    L = _L()
    L.trko_search_by_projection_sim3.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_int] + [_vp] * 7 + [C.c_int, C.c_float, _vp]
    a, sk, m = _pts(xw, normal, min_dist, max_dist, mp_desc, skip)
    km = _c(kp_matched, np.uint8)
    out = np.full(max(frame_oracle.n, 1), -1, np.int32)
    n = L.trko_search_by_projection_sim3(frame_oracle.h, _c(sq, np.float32).ctypes.data, _c(st, np.float32).ctypes.data,
                                         _c(K, np.float32).ctypes.data, int(nlevels), float(log_scale), m, *(x.ctypes.data for x in a),
                                         sk.ctypes.data if sk is not None else None, km.ctypes.data, int(th), float(ratio_hamming),
                                         out.ctypes.data)
    return n, out[:frame_oracle.n]


def fuse_search_sim3(frame_oracle, sq, st, K, log_scale, nlevels, xw, normal, min_dist, max_dist, mp_desc, skip=None, th=3.0):
    """The search half of Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) -> (best_idx, best_dist)."""
    L = _L()
    L.trko_fuse_search_sim3.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_int] + [_vp] * 6 + [C.c_float, _vp, _vp]
    L.trko_fuse_search_sim3.restype = None
    a, sk, m = _pts(xw, normal, min_dist, max_dist, mp_desc, skip)
    bi, bd = np.full(max(m, 1), -1, np.int32), np.full(max(m, 1), 256, np.int32)
    L.trko_fuse_search_sim3(frame_oracle.h, _c(sq, np.float32).ctypes.data, _c(st, np.float32).ctypes.data, _c(K, np.float32).ctypes.data,
                            int(nlevels), float(log_scale), m, *(x.ctypes.data for x in a), sk.ctypes.data if sk is not None else None,
                            float(th), bi.ctypes.data, bd.ctypes.data)
    return bi[:m], bd[:m]


def search_by_sim3(f1, f2, q1, t1, q2, t2, s12q, s12t, K, log_scale, nlevels, side1, side2, th=7.5):
    """SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) -> (nFound, match12).  side = (skip, xw, min_dist, max_dist, mp_desc)."""
    L = _L()
    L.trko_search_by_sim3.argtypes = [_vp] * 9 + [C.c_int, C.c_float] + [_vp] * 10 + [C.c_float, _vp]
    p = [_c(x, np.float32) for x in (q1, t1, q2, t2, s12q, s12t, K)]
    sides = []
    for sk, xw, mn, mx, d in (side1, side2):
        sides += [_c(sk, np.uint8), _c(xw, np.float32), _c(mn, np.float32), _c(mx, np.float32), _c(d, np.uint8)]
    m12 = np.full(max(f1.n, 1), -1, np.int32)
    n = L.trko_search_by_sim3(f1.h, f2.h, *(x.ctypes.data for x in p), int(nlevels), float(log_scale), *(x.ctypes.data for x in sides),
                              float(th), m12.ctypes.data)
    return n, m12[:f1.n]
