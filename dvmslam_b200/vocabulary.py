"""Python mirror of the DBoW2 vocabulary transform over the C-ABI (dvm_vocabulary_*): the reference's
ORBVocabulary (O3/include/ORBVocabulary.h) with loadFromTextFile and transform(features, BowVector, FeatureVector,
levelsup) -- same argument meaning; BowVector / FeatureVector are returned as dicts in key order."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from ._lib import check, lib

_vp = C.c_void_p
TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3
L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = range(6)


def flatten_tree(parent, is_leaf, desc, weight):
    """Per-node arrays in loadFromTextFile order (node ids 1..N as the lines of the file; parent ids) ->
    (child_start, children, desc[N+1,32], weight[N+1], word_id[N+1]) with the root as node 0."""
    parent = np.asarray(parent, np.int64)
    n = len(parent) + 1
    order = np.argsort(parent, kind="stable")            # children of a node in the order they appear
    children = (order + 1).astype(np.int32)
    counts = np.bincount(parent, minlength=n)
    child_start = np.zeros(n + 1, np.int32)
    child_start[1:] = np.cumsum(counts)
    leaf = np.concatenate([[0], np.asarray(is_leaf, np.int64)]) > 0
    word_id = np.full(n, -1, np.int32)
    word_id[leaf] = np.arange(int(leaf.sum()), dtype=np.int32)      # words are numbered in file order
    d = np.zeros((n, 32), np.uint8)
    d[1:] = desc
    w = np.zeros(n, np.float64)
    w[1:] = weight
    return child_start, children, d, w, word_id


def load_text(path):
    """Parses ORBvoc.txt's format (DBoW2/TemplatedVocabulary.h:1211-1287): 'k L scoring weighting', then one line
    per node: parent is_leaf d0 .. d31 weight.  -> (k, L, scoring, weighting, parent, is_leaf, desc, weight)"""
    with open(path) as f:
        k, L, scoring, weighting = (int(x) for x in f.readline().split()[:4])
        try:
            import pandas as pd

            # round_trip: the weights must be the doubles the reference's `ssnode >> weight` (strtod, correctly rounded) reads; the
            # default fast converter is off by an ulp now and then (found by tests/test_ref_dbow.py)
            rows = pd.read_csv(f, sep=r"\s+", header=None, float_precision="round_trip").to_numpy(np.float64)
        except ImportError:
            rows = np.loadtxt(f, dtype=np.float64, ndmin=2)
    return (k, L, scoring, weighting, rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64),
            rows[:, 2:34].astype(np.uint8), rows[:, 34].copy())


class Vocabulary:
    def __init__(self, k, L, scoring, weighting, parent, is_leaf, desc, weight, device: int = 0):
        self.k, self.L, self.scoring, self.weighting = int(k), int(L), int(scoring), int(weighting)
        self.child_start, self.children, self.desc, self.weight, self.word_id = flatten_tree(parent, is_leaf, desc, weight)
        self.Lb = lib()
        self.Lb.dvm_vocabulary_create.argtypes = [C.POINTER(_vp), C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int]
        self.Lb.dvm_vocabulary_destroy.argtypes = [_vp]
        self.Lb.dvm_vocabulary_destroy.restype = None
        self.Lb.dvm_vocabulary_transform.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]
        self.h = _vp()
        check(self.Lb.dvm_vocabulary_create(C.byref(self.h), device, len(self.word_id), self.child_start.ctypes.data,
                                            self.children.ctypes.data, self.desc.ctypes.data, self.weight.ctypes.data,
                                            self.word_id.ctypes.data, self.L))

    @classmethod
    def loadFromTextFile(cls, path, device: int = 0):
        """ORBVocabulary::loadFromTextFile"""
        return cls(*load_text(path), device=device)

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.Lb.dvm_vocabulary_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform_features(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8)
        n = len(d)
        word, w, nid = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64), np.zeros(max(n, 1), np.int32)
        check(self.Lb.dvm_vocabulary_transform(self.h, d.ctypes.data if n else None, n, int(levelsup), word.ctypes.data,
                                               w.ctypes.data, nid.ctypes.data))
        return word[:n], w[:n], nid[:n]

    def transform(self, desc, levelsup=4):
        """-> (BowVector {word: value}, FeatureVector {node: [feature indices]}), both in ascending key order."""
        word, w, nid = self.transform_features(desc, levelsup)
        bow, fv = {}, {}
        accumulate = self.weighting in (TF_IDF, TF)
        for i in range(len(word)):
            if not w[i] > 0:
                continue
            k = int(word[i])
            if accumulate:
                bow[k] = bow.get(k, 0.0) + float(w[i])
            elif k not in bow:
                bow[k] = float(w[i])
            fv.setdefault(int(nid[i]), []).append(i)
        bow = dict(sorted(bow.items()))
        must = self.scoring != DOT_PRODUCT
        if accumulate and bow and not must:
            nd = float(len(bow))
            bow = {k: v / nd for k, v in bow.items()}
        if must:
            norm = 0.0
            if self.scoring == L2_NORM:
                for v in bow.values():
                    norm += v * v
                norm = math.sqrt(norm)
            else:
                for v in bow.values():
                    norm += abs(v)
            if norm > 0.0:
                bow = {k: v / norm for k, v in bow.items()}
        return bow, dict(sorted(fv.items()))
