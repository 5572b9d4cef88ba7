"""Python mirrors of the descriptor matchers that do not project (ORBmatcher::SearchByBoW, both overloads,
ORBmatcher::SearchForInitialization; O3/include/ORBmatcher.h:58-76) and of the exhaustive Hamming search
of the inter-agent exchange, over the C-ABI.  The pointer-graph arguments are flattened as
include/dvmslam_b200.h documents; a DBoW2::FeatureVector is a dict {node id: [feature indices]}."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib
from .extractor import KP_DTYPE
from .tracking import Frame, _bind as _bind_tracking

_vp = C.c_void_p
_ip = C.POINTER(C.c_int)


class _BowFeatures(C.Structure):
    _fields_ = [("n", C.c_int32), ("desc", _vp), ("angle", _vp), ("has_mp", _vp), ("n_nodes", C.c_int32),
                ("node_id", _vp), ("node_start", _vp), ("feat_idx", _vp)]


def _bind(L):
    if getattr(L, "_bow_bound", False):
        return
    _bind_tracking(L)
    L.dvm_match_by_bow.argtypes = [_vp, C.c_int, C.POINTER(_BowFeatures), C.POINTER(_BowFeatures), C.c_float, C.c_int,
                                   _vp, _vp, _ip]
    L.dvm_match_for_initialization.argtypes = [_vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_float, C.c_int, _vp, _ip]
    L.dvm_match_for_triangulation.argtypes = [_vp, C.POINTER(_BowFeatures), _vp, C.POINTER(_BowFeatures), _vp, _vp, _vp, _vp,
                                              _vp, C.c_int, C.c_int, C.c_int, _vp, _ip]
    L.dvm_fuse_search.argtypes = [_vp, _vp, _vp, _vp, C.c_int] + [_vp] * 6 + [C.c_float, _vp, _vp]
    L.dvm_hamming_create.argtypes = [C.POINTER(_vp), C.c_int, _vp]
    L.dvm_hamming_destroy.argtypes = [_vp]
    L.dvm_hamming_destroy.restype = None
    L.dvm_hamming_knn.argtypes = [_vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp]
    L.dvm_hamming_knn_device.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int,
                                         C.c_float]
    L.dvm_hamming_sync.argtypes = [_vp]
    L.dvm_hamming_set_mode.argtypes = [_vp, C.c_int]
    L._bow_bound = True


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def feature_vector_csr(fv):
    """{node id: [feature indices]} (a DBoW2::FeatureVector) -> (node_id u32, node_start i32, feat_idx u32)."""
    nodes = sorted(fv)
    node_id = np.array(nodes, np.uint32)
    start = np.zeros(len(nodes) + 1, np.int32)
    for i, k in enumerate(nodes):
        start[i + 1] = start[i] + len(fv[k])
    idx = np.concatenate([np.asarray(fv[k], np.uint32) for k in nodes]) if nodes else np.zeros(0, np.uint32)
    return node_id, start, _c(idx, np.uint32)


class BowFeatures:
    """One side of SearchByBoW: descriptors, keypoint angles, map-point validity, feature vector."""

    def __init__(self, desc, angle, has_mp, fv):
        self.desc, self.angle = _c(desc, np.uint8), _c(angle, np.float32)
        self.has_mp = None if has_mp is None else _c(has_mp, np.uint8)
        self.node_id, self.node_start, self.feat_idx = feature_vector_csr(fv) if isinstance(fv, dict) else fv
        self.n = len(self.angle)

    def struct(self):
        p = lambda a: a.ctypes.data if a is not None and a.size else None  # noqa: E731
        return _BowFeatures(self.n, p(self.desc), p(self.angle), p(self.has_mp), len(self.node_id), p(self.node_id),
                            self.node_start.ctypes.data, p(self.feat_idx))


class BowMatcher:
    """ORBmatcher's SearchByBoW / SearchForInitialization (same argument meaning; see the header)."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True):
        self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri
        self.L = lib()
        _bind(self.L)

    def _bow(self, ctx: Frame, kf_kf, a: BowFeatures, b: BowFeatures):
        m12 = np.full(max(a.n, 1), -1, np.int32)
        m21 = np.full(max(b.n, 1), -1, np.int32)
        n = C.c_int()
        sa, sb = a.struct(), b.struct()
        check(self.L.dvm_match_by_bow(ctx.h, kf_kf, C.byref(sa), C.byref(sb), float(self.mfNNratio),
                                      int(self.mbCheckOrientation), m12.ctypes.data, m21.ctypes.data, C.byref(n)))
        return n.value, m12[:a.n], m21[:b.n]

    def SearchByBoW_KF_F(self, ctx: Frame, kf: BowFeatures, f: BowFeatures):
        """SearchByBoW(pKF, F, vpMapPointMatches) -> (nmatches, match_f): match_f[i] = the KF feature whose
        map point Frame feature i now holds, or -1."""
        n, _, m21 = self._bow(ctx, 0, kf, f)
        return n, m21

    def SearchByBoW_KF_KF(self, ctx: Frame, kf1: BowFeatures, kf2: BowFeatures):
        """SearchByBoW(pKF1, pKF2, vpMatches12) -> (nmatches, match12): the KF2 feature matched to KF1
        feature i, or -1."""
        n, m12, _ = self._bow(ctx, 1, kf1, kf2)
        return n, m12

    def SearchForInitialization(self, f2: Frame, kps1_un, desc1, prev_matched, windowSize=10):
        """-> (nmatches, vnMatches12, vbPrevMatched updated)."""
        k1, d1 = _c(kps1_un, KP_DTYPE), _c(desc1, np.uint8)
        pm = _c(prev_matched, np.float32).copy()
        m12 = np.full(max(len(k1), 1), -1, np.int32)
        n = C.c_int()
        check(self.L.dvm_match_for_initialization(f2.h, len(k1), k1.ctypes.data, d1.ctypes.data, pm.ctypes.data,
                                                  int(windowSize), float(self.mfNNratio),
                                                  int(self.mbCheckOrientation), m12.ctypes.data, C.byref(n)))
        return n.value, m12[:len(k1)], pm


def fundamental_from_poses(q1, t1, q2, t2, K1, K2):
    """(F12 row-major float32[9], ep float32[2]) as SearchForTriangulation derives them from the two keyframe poses Tcw
    (O3/src/ORBmatcher.cc:841-860) and Pinhole::epipolarConstrain rebuilds F12 = K1^-T [t12]x R12 K2^-1
    (O3/src/CameraModels/Pinhole.cpp:104-110), in the reference's float32 Sophus / Eigen arithmetic
    (dvm_fundamental_from_poses; host arithmetic inside the library)."""
    L = lib()
    L.dvm_fundamental_from_poses.argtypes = [_vp] * 8
    a = [_c(x, np.float32) for x in (q1, t1, q2, t2, K1, K2)]
    F12, ep = np.zeros(9, np.float32), np.zeros(2, np.float32)
    check(L.dvm_fundamental_from_poses(*(x.ctypes.data for x in a), F12.ctypes.data, ep.ctypes.data))
    return F12, ep


def SearchForTriangulation(ctx: Frame, kf1: BowFeatures, kps1, kf2: BowFeatures, kps2, F12, ep, scale_factors2,
                           level_sigma2_2, bCoarse=False, checkOri=True):
    """ORBmatcher::SearchForTriangulation on mono keyframes -> (nmatches, vMatches12); BowFeatures.has_mp marks the
    features that already hold a map point."""
    L = lib()
    _bind(L)
    k1, k2 = _c(kps1, KP_DTYPE), _c(kps2, KP_DTYPE)
    sf, s2 = _c(scale_factors2, np.float32), _c(level_sigma2_2, np.float32)
    m12 = np.full(max(kf1.n, 1), -1, np.int32)
    n = C.c_int()
    sa, sb = kf1.struct(), kf2.struct()
    check(L.dvm_match_for_triangulation(ctx.h, C.byref(sa), k1.ctypes.data, C.byref(sb), k2.ctypes.data,
                                        _c(F12, np.float32).ctypes.data, _c(ep, np.float32).ctypes.data, sf.ctypes.data,
                                        s2.ctypes.data, len(sf), int(bCoarse), int(checkOri), m12.ctypes.data, C.byref(n)))
    return n.value, m12[:kf1.n]


def FuseSearch(kf: Frame, q, t, K, xw, normal, min_dist, max_dist, mp_desc, skip=None, th=3.0):
    """The search half of ORBmatcher::Fuse(pKF, vpMapPoints, th) -> (best_idx, best_dist) per map point."""
    L = lib()
    _bind(L)
    a = [_c(xw, np.float32), _c(normal, np.float32), _c(min_dist, np.float32), _c(max_dist, np.float32), _c(mp_desc, np.uint8)]
    m = len(a[2])
    sk = _c(skip, np.uint8) if skip is not None else None
    bi, bd = np.full(max(m, 1), -1, np.int32), np.full(max(m, 1), 256, np.int32)
    check(L.dvm_fuse_search(kf.h, _c(q, np.float32).ctypes.data, _c(t, np.float32).ctypes.data,
                            _c(K, np.float32).ctypes.data, m, *(x.ctypes.data for x in a),
                            sk.ctypes.data if sk is not None else None, float(th), bi.ctypes.data, bd.ctypes.data))
    return bi[:m], bd[:m]


class HammingKnn:
    """Exhaustive nearest / second-nearest Hamming search on one B200 (dvm_hamming_*)."""
    NO_KEY = 256 << 20

    def __init__(self, device: int = 0, stream=None):
        """stream: None = a private stream; an integer cudaStream_t to launch on (0, CUDA's legacy default
        stream -- what torch.cuda.current_stream().cuda_stream returns by default -- is passed as
        cudaStreamLegacy, since a NULL handle means "private stream" in the C-ABI)."""
        self.L = lib()
        _bind(self.L)
        self.h = _vp()
        check(self.L.dvm_hamming_create(C.byref(self.h), device, None if stream is None else _vp(stream if stream else 1)))

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.dvm_hamming_destroy(self.h)
            self.h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def knn(self, a, b):
        """host arrays [na,32], [nb,32] -> (best_idx, best_dist, second_dist)"""
        a, b = _c(a, np.uint8), _c(b, np.uint8)
        na, nb = len(a), len(b)
        out = [np.zeros(max(na, 1), np.int32) for _ in range(3)]
        check(self.L.dvm_hamming_knn(self.h, a.ctypes.data if na else None, na, b.ctypes.data if nb else None, nb,
                                     *(o.ctypes.data for o in out)))
        return tuple(o[:na] for o in out)

    def knn_device(self, a_ptr, ba, na, b_ptr, bb, nb, key1_ptr, key2_ptr, counts_ptr=0, th_low=50, nnratio=0.75):
        """device pointers (ints); enqueues on the handle's stream"""
        check(self.L.dvm_hamming_knn_device(self.h, _vp(a_ptr), ba, na, _vp(b_ptr), bb, nb, _vp(key1_ptr),
                                            _vp(key2_ptr), _vp(counts_ptr) if counts_ptr else None, int(th_low),
                                            float(nnratio)))

    def sync(self):
        check(self.L.dvm_hamming_sync(self.h))

    def set_mode(self, mode: int):
        """0 = by size, 1 = popcount kernel, 2 = tcgen05 int8 kernel, 3 / 4 = timing probes without outputs (MMA only / + TMEM read-out)."""
        check(self.L.dvm_hamming_set_mode(self.h, int(mode)))


def _points(xw, normal, min_dist, max_dist, mp_desc, skip):
    a = [_c(xw, np.float32), None if normal is None else _c(normal, np.float32), _c(min_dist, np.float32), _c(max_dist, np.float32),
         _c(mp_desc, np.uint8)]
    return a, (_c(skip, np.uint8) if skip is not None else None), len(a[2])


def _ptr(a):
    return a.ctypes.data if a is not None else None


def SearchByProjectionSim3(kf: Frame, sim3_q, sim3_t, K, xw, normal, min_dist, max_dist, mp_desc, skip, kp_matched, th, ratioHamming=1.0):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) -> (nmatches, kp_point[kf n]): the
    candidate index newly written to vpMatched[k], -1 where unchanged."""
    L = lib()
    a, sk, m = _points(xw, normal, min_dist, max_dist, mp_desc, skip)
    km = _c(kp_matched, np.uint8)
    out = np.full(max(kf.n, 1), -1, np.int32)
    n = C.c_int()
    sq, st, Kc = _c(sim3_q, np.float32), _c(sim3_t, np.float32), _c(K, np.float32)
    L.dvm_match_by_projection_sim3.argtypes = [_vp, _vp, _vp, _vp, C.c_int] + [_vp] * 7 + [C.c_int, C.c_float, _vp, _vp]
    check(L.dvm_match_by_projection_sim3(kf.h, _ptr(sq), _ptr(st), _ptr(Kc), m, *(_ptr(x) for x in a), _ptr(sk), _ptr(km), int(th),
                                         float(ratioHamming), _ptr(out), C.addressof(n)))
    return n.value, out[:kf.n]


def FuseSearchSim3(kf: Frame, sim3_q, sim3_t, K, xw, normal, min_dist, max_dist, mp_desc, skip=None, th=3.0):
    """The search half of ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) -> (best_idx, best_dist)."""
    L = lib()
    a, sk, m = _points(xw, normal, min_dist, max_dist, mp_desc, skip)
    bi, bd = np.full(max(m, 1), -1, np.int32), np.full(max(m, 1), 256, np.int32)
    sq, st, Kc = _c(sim3_q, np.float32), _c(sim3_t, np.float32), _c(K, np.float32)
    L.dvm_fuse_search_sim3.argtypes = [_vp, _vp, _vp, _vp, C.c_int] + [_vp] * 6 + [C.c_float, _vp, _vp]
    check(L.dvm_fuse_search_sim3(kf.h, _ptr(sq), _ptr(st), _ptr(Kc), m, *(_ptr(x) for x in a), _ptr(sk), float(th), _ptr(bi), _ptr(bd)))
    return bi[:m], bd[:m]


def SearchBySim3(kf1: Frame, kf2: Frame, q1, t1, q2, t2, s12_q, s12_t, K, side1, side2, th=7.5):
    """ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th) -> (nFound, match12).  side = (skip, xw, min_dist, max_dist,
    mp_desc), one entry per keypoint of the keyframe."""
    L = lib()
    p = [_c(x, np.float32) for x in (q1, t1, q2, t2, s12_q, s12_t, K)]
    sides = []
    for sk, xw, mn, mx, d in (side1, side2):
        sides += [_c(sk, np.uint8), _c(xw, np.float32), _c(mn, np.float32), _c(mx, np.float32), _c(d, np.uint8)]
    m12 = np.full(max(kf1.n, 1), -1, np.int32)
    n = C.c_int()
    L.dvm_match_by_sim3.argtypes = [_vp] * 19 + [C.c_float, _vp, _vp]
    check(L.dvm_match_by_sim3(kf1.h, kf2.h, *(_ptr(x) for x in p), *(_ptr(x) for x in sides), float(th), _ptr(m12), C.addressof(n)))
    return n.value, m12[:kf1.n]
