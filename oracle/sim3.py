"""Sim3 oracle bindings (oracle/sim3_oracle.cpp) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib

_vp = C.c_void_p


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def optimize_sim3(p1c, p2c, obs1, obs2, w1, w2, K1, K2, q, t, s, th2=10.0, fix_scale=False):
    """Optimizer::OptimizeSim3 on flattened correspondences.  Returns dict(q, t, s, inlier, n_in, iters1, iters2, trials,
    n_bad, chi_first, chi_last)."""
    L = lib()
    L.sim3o_optimize_sim3.argtypes = [C.c_int] + [_vp] * 11 + [C.c_float, C.c_int, _vp, _vp]
    n = len(_c(w1, np.float32))
    qq, tt, ss = _c(q, np.float64).copy(), _c(t, np.float64).copy(), np.array([s], np.float64)
    inl = np.zeros(max(n, 1), np.uint8)
    st = np.zeros(6, np.float64)
    a = [_c(x, np.float32) for x in (p1c, p2c, obs1, obs2, w1, w2, K1, K2)]
    n_in = L.sim3o_optimize_sim3(n, *[x.ctypes.data for x in a], qq.ctypes.data, tt.ctypes.data, ss.ctypes.data,
                                 C.c_float(th2), int(fix_scale), inl.ctypes.data, st.ctypes.data)
    return dict(q=qq, t=tt, s=float(ss[0]), inlier=inl[:n], n_in=n_in, iters1=int(st[0]), iters2=int(st[1]), trials=int(st[2]),
                n_bad=int(st[3]), chi_first=st[4], chi_last=st[5])


def sim3_exp(u):
    L = lib()
    L.sim3o_exp.argtypes = [_vp, _vp]
    out = np.zeros(8)
    L.sim3o_exp(_c(u, np.float64).ctypes.data, out.ctypes.data)
    return out


def sim3_log(s8):
    L = lib()
    L.sim3o_log.argtypes = [_vp, _vp]
    out = np.zeros(7)
    L.sim3o_log(_c(s8, np.float64).ctypes.data, out.ctypes.data)
    return out


def optimize_essential_graph(sim3, fixed, vi, vj, meas, fix_scale=False, iterations=20, lambda_init=1e-16):
    """The solve of Optimizer::OptimizeEssentialGraph on a flattened pose graph: sim3 [nv, 8] (q xyzw, t, s) = Scw,
    edges (vi, vj, Sji).  Returns dict(sim3, iters, trials, chi_first, chi_last).  Oracle only: no product kernel yet."""
    L = lib()
    L.sim3o_optimize_essential_graph.argtypes = [C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, C.c_double, _vp]
    S = _c(sim3, np.float64).copy()
    fx = _c(fixed, np.uint8)
    a, b = _c(vi, np.int32), _c(vj, np.int32)
    m = _c(meas, np.float64)
    st = np.zeros(4)
    L.sim3o_optimize_essential_graph(len(S), S.ctypes.data, fx.ctypes.data, len(a), a.ctypes.data, b.ctypes.data, m.ctypes.data,
                                     int(fix_scale), iterations, lambda_init, st.ctypes.data)
    return dict(sim3=S, iters=int(st[0]), trials=int(st[1]), chi_first=st[2], chi_last=st[3])
