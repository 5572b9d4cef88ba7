"""Bindings of oracle/dbow_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib

_vp = C.c_void_p


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def transform_features(tree, L, feat, levelsup=4):
    """tree = (child_start, children, desc, weight, word_id) -> (word, weight, node) per feature"""
    Lb = lib()
    Lb.dbowo_transform_features.restype = None
    Lb.dbowo_transform_features.argtypes = [C.c_int] + [_vp] * 5 + [C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp]
    cs, ch, d, w, wid = (_c(a, t) for a, t in zip(tree, (np.int32, np.int32, np.uint8, np.float64, np.int32)))
    f = _c(feat, np.uint8)
    n = len(f)
    word, ww, nid = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64), np.zeros(max(n, 1), np.int32)
    Lb.dbowo_transform_features(len(wid), cs.ctypes.data, ch.ctypes.data, d.ctypes.data, w.ctypes.data, wid.ctypes.data, int(L),
                                f.ctypes.data, n, int(levelsup), word.ctypes.data, ww.ctypes.data, nid.ctypes.data)
    return word[:n], ww[:n], nid[:n]


def transform(tree, L, weighting, scoring, feat, levelsup=4):
    """-> (BowVector dict, FeatureVector dict) in std::map order"""
    Lb = lib()
    Lb.dbowo_transform.restype = None
    Lb.dbowo_transform.argtypes = [C.c_int] + [_vp] * 5 + [C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int] + [_vp] * 6
    cs, ch, d, w, wid = (_c(a, t) for a, t in zip(tree, (np.int32, np.int32, np.uint8, np.float64, np.int32)))
    f = _c(feat, np.uint8)
    n = len(f)
    bw, bv = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.float64)
    fn, fs, fi, cnt = np.zeros(n + 1, np.int32), np.zeros(n + 2, np.int32), np.zeros(n + 1, np.int32), np.zeros(2, np.int32)
    Lb.dbowo_transform(len(wid), cs.ctypes.data, ch.ctypes.data, d.ctypes.data, w.ctypes.data, wid.ctypes.data, int(L),
                       int(weighting), int(scoring), f.ctypes.data, n, int(levelsup), bw.ctypes.data, bv.ctypes.data,
                       fn.ctypes.data, fs.ctypes.data, fi.ctypes.data, cnt.ctypes.data)
    bow = {int(bw[i]): float(bv[i]) for i in range(cnt[0])}
    fv = {int(fn[i]): [int(x) for x in fi[fs[i]:fs[i + 1]]] for i in range(cnt[1])}
    return bow, fv
