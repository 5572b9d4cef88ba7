"""Bindings of oracle/_ref/libref_dbow.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The library holds the reference's own vendored DBoW2 (TemplatedVocabulary<FORB::TDescriptor, FORB> = ORBVocabulary,
O3/Thirdparty/DBoW2/DBoW2/*.{h,cpp} and DUtils), compiled unmodified from /root/reference over the stand-in
<opencv2/core/core.hpp> of oracle/dbowshim (oracle/Makefile, target `ref`).  It exists only where /root/reference is present
(or where the built .so travelled); tests that need it skip otherwise.  Return shapes mirror oracle.dbow."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_dbow.so")
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
        L.refdbow_load_text.argtypes = [C.c_char_p]
        L.refdbow_load_text.restype = _vp
        L.refdbow_free.argtypes = [_vp]
        L.refdbow_free.restype = None
        L.refdbow_info.argtypes = [_vp, _vp]
        L.refdbow_info.restype = None
        L.refdbow_transform_features.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]
        L.refdbow_transform_features.restype = None
        L.refdbow_transform.argtypes = [_vp, _vp, C.c_int, C.c_int] + [_vp] * 6
        L.refdbow_transform.restype = None
        _LIB = L
    return _LIB


class RefVocabulary:
    """The reference's ORBVocabulary loaded with ITS loadFromTextFile (TemplatedVocabulary.h:1211-1290)."""

    def __init__(self, path: str):
        self.h = _L().refdbow_load_text(path.encode())
        if not self.h:
            raise OSError(f"the reference's loadFromTextFile rejected {path}")
        info = np.zeros(5, np.int32)
        _L().refdbow_info(self.h, info.ctypes.data)
        self.k, self.L, self.scoring, self.weighting, self.n_words = (int(x) for x in info)

    def close(self):
        if self.h:
            _L().refdbow_free(self.h)
            self.h = None

    def transform_features(self, feat, levelsup=4):
        f = np.ascontiguousarray(feat, np.uint8)
        n = len(f)
        word, w, nid = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64), np.zeros(max(n, 1), np.int32)
        _L().refdbow_transform_features(self.h, f.ctypes.data, n, int(levelsup), word.ctypes.data, w.ctypes.data, nid.ctypes.data)
        return word[:n], w[:n], nid[:n]

    def transform(self, feat, levelsup=4):
        """-> (BowVector dict, FeatureVector dict) in std::map order (Frame::ComputeBoW, O3/src/Frame.cc:784-789)"""
        f = np.ascontiguousarray(feat, np.uint8)
        n = len(f)
        bw, bv = np.zeros(n + 1, np.int32), np.zeros(n + 1, np.float64)
        fn, fs, fi, cnt = np.zeros(n + 1, np.int32), np.zeros(n + 2, np.int32), np.zeros(n + 1, np.int32), np.zeros(2, np.int32)
        _L().refdbow_transform(self.h, f.ctypes.data, n, int(levelsup), bw.ctypes.data, bv.ctypes.data, fn.ctypes.data, fs.ctypes.data,
                               fi.ctypes.data, cnt.ctypes.data)
        bow = {int(bw[i]): float(bv[i]) for i in range(cnt[0])}
        fv = {int(fn[i]): [int(x) for x in fi[fs[i]:fs[i + 1]]] for i in range(cnt[1])}
        return bow, fv
