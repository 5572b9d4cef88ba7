"""Self-checks of the descriptor-matcher oracle (oracle/bow_oracle.cpp, trko_search_for_initialization):
hand-computable cases and a literal pure-Python restatement of the reference loops on small inputs.
The reference holds no fixture for these; the restatement is pinned to the reference's own ORBmatcher.cc
in tests/test_ref_matchers.py and regression-pinned by tests/golden/bow_small.npz."""
import os

import numpy as np
import pytest

from tests import bow_cases

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dist(a, b):
    return int(np.unpackbits(a ^ b).sum())


def _py_search_by_bow(kf_kf, c, nnratio, check_ori):
    """ORBmatcher.cc:214-393 / 709-834 written out in Python (mono)."""
    d1, d2 = c["desc1"], c["desc2"]
    m12 = -np.ones(len(d1), np.int64)
    m21 = -np.ones(len(d2), np.int64)
    hist = [[] for _ in range(30)]
    n = 0
    for node in sorted(set(c["fv1"]) & set(c["fv2"])):
        for r1 in c["fv1"][node]:
            if not c["valid1"][r1]:
                continue
            b1, b2, bi = 256, 256, -1
            for r2 in c["fv2"][node]:
                if m21[r2] >= 0 or (kf_kf and not c["valid2"][r2]):
                    continue
                d = _dist(d1[r1], d2[r2])
                if d < b1:
                    b2, b1, bi = b1, d, r2
                elif d < b2:
                    b2 = d
            if (b1 < 50 if kf_kf else b1 <= 50) and np.float32(b1) < np.float32(nnratio) * np.float32(b2):
                m21[bi], m12[r1] = r1, bi
                n += 1
                if check_ori:
                    rot = np.float32(c["angle1"][r1]) - np.float32(c["angle2"][bi])
                    if rot < 0:
                        rot = np.float32(rot + np.float32(360.0))
                    v = float(np.float32(rot * np.float32(1.0 / 30)))
                    b = int(np.floor(v + 0.5)) if v >= 0 else int(np.ceil(v - 0.5))
                    hist[0 if b == 30 else b].append(r1)
    if check_ori:
        sizes = [len(h) for h in hist]
        order = sorted(range(30), key=lambda i: (-sizes[i], i))
        mx = [order[0] if sizes[order[0]] > 0 else -1, order[1] if sizes[order[1]] > 0 else -1,
              order[2] if sizes[order[2]] > 0 else -1]
        s = [sizes[i] if i >= 0 else 0 for i in mx]
        if s[1] < np.float32(0.1) * np.float32(s[0]):
            mx[1] = mx[2] = -1
        elif s[2] < np.float32(0.1) * np.float32(s[0]):
            mx[2] = -1
        for i in range(30):
            if i in mx:
                continue
            for r1 in hist[i]:
                m21[m12[r1]] = -1
                m12[r1] = -1
                n -= 1
    return n, m12, m21


@pytest.mark.parametrize("kf_kf", [0, 1])
@pytest.mark.parametrize("seed,dup", [(0, False), (1, True), (2, True)])
def test_bow_oracle_matches_python_restatement(kf_kf, seed, dup):
    from oracle.bow import search_by_bow

    c = bow_cases.bow_synthetic(140, 170, seed, dup=dup, nodes=5)
    for ori in (True, False):
        n0, a0, b0 = _py_search_by_bow(kf_kf, c, 0.75, ori)
        n1, a1, b1 = search_by_bow(kf_kf, c["desc1"], c["angle1"], c["valid1"], c["fv1"], c["desc2"], c["angle2"],
                                   c["valid2"], c["fv2"], 0.75, ori)
        assert n0 == n1 and np.array_equal(a0, a1) and np.array_equal(b0, b1)
        assert n1 >= 5


def test_bow_oracle_identity_and_edges():
    from oracle.bow import search_by_bow

    rng = np.random.default_rng(0)
    d = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    ang = np.zeros(64, np.float32)
    fv = {3: list(range(0, 32)), 9: list(range(32, 64))}
    n, m12, m21 = search_by_bow(0, d, ang, None, fv, d, ang, None, fv, 0.6, True)
    assert n == 64 and np.array_equal(m12, np.arange(64)) and np.array_equal(m21, np.arange(64))
    # disjoint node sets: nothing to compare
    n, m12, m21 = search_by_bow(1, d, ang, None, {1: list(range(64))}, d, ang, None, {2: list(range(64))}, 0.6, True)
    assert n == 0 and (m12 < 0).all() and (m21 < 0).all()
    # KF-KF accepts only bestDist1 < 50, KF-F accepts <= 50
    a = np.zeros((1, 32), np.uint8)
    b = np.zeros((1, 32), np.uint8)
    b[0, :6] = 0xFF
    b[0, 6] = 0x03   # distance 50
    for kf_kf, want in ((0, 1), (1, 0)):
        n, _, _ = search_by_bow(kf_kf, a, ang[:1], None, {0: [0]}, b, ang[:1], None, {0: [0]}, 0.6, False)
        assert n == want


def test_hamming_knn_oracle():
    from oracle.bow import hamming_knn

    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (80, 32), dtype=np.uint8)
    b[10] = a[7]
    b[20] = a[7]   # tie: the first wins, the second is the second-nearest at distance 0
    idx, d1, d2 = hamming_knn(a, b)
    D = np.array([[_dist(x, y) for y in b] for x in a])
    assert np.array_equal(idx, D.argmin(1)) and np.array_equal(d1, D.min(1))
    assert np.array_equal(d2, np.sort(D, 1)[:, 1])
    assert idx[7] == 10 and d1[7] == 0 and d2[7] == 0
    idx, d1, d2 = hamming_knn(a, b[:1])
    assert (idx == 0).all() and (d2 == 256).all()
    idx, d1, d2 = hamming_knn(a, b[:0])
    assert (idx == -1).all() and (d1 == 256).all()


def test_search_for_initialization_oracle_properties():
    from oracle.bow import search_for_initialization
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    orc = OrbOracle(1000)
    T = orc.tables()
    c = bow_cases.init_pair(orc.extract)
    F2 = FrameOracle(c["kps2"], c["desc2"], c["bounds"], T["scale"])
    n, m12, pm = search_for_initialization(c["kps1"], c["desc1"], F2, c["prev"], 100, 0.9, True)
    assert n == (m12 >= 0).sum() and n > 50
    used = m12[m12 >= 0]
    assert len(np.unique(used)) == len(used)                       # one-to-one
    assert (c["kps1"]["octave"][m12 >= 0] == 0).all() and (c["kps2"]["octave"][used] == 0).all()
    assert np.array_equal(pm[m12 >= 0], np.stack([c["kps2"]["x"][used], c["kps2"]["y"][used]], 1))
    assert np.array_equal(pm[m12 < 0], c["prev"][m12 < 0])
    # identical frames: every level-0 keypoint whose nearest neighbour is unambiguous matches itself
    F1 = FrameOracle(c["kps1"], c["desc1"], c["bounds"], T["scale"])
    n, m12, _ = search_for_initialization(c["kps1"], c["desc1"], F1, c["prev"], 100, 0.9, False)
    idx = np.nonzero(m12 >= 0)[0]
    assert n > 100 and np.array_equal(m12[idx], idx)


def test_bow_golden():
    from oracle.bow import hamming_knn, search_by_bow, search_for_initialization
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    g = np.load(os.path.join(G, "bow_small.npz"))
    orc = OrbOracle(600)
    T = orc.tables()
    c = bow_cases.bow_pair(orc.extract)
    for kf_kf in (0, 1):
        n, m12, m21 = search_by_bow(kf_kf, c["desc1"], c["angle1"], c["valid1"], c["fv1"], c["desc2"], c["angle2"],
                                    c["valid2"], c["fv2"], 0.7, True)
        assert n == int(g[f"bow{kf_kf}_n"]) and np.array_equal(m12, g[f"bow{kf_kf}_m12"]) and np.array_equal(m21, g[f"bow{kf_kf}_m21"])
    ci = bow_cases.init_pair(orc.extract)
    F2 = FrameOracle(ci["kps2"], ci["desc2"], ci["bounds"], T["scale"])
    n, m12, pm = search_for_initialization(ci["kps1"], ci["desc1"], F2, ci["prev"], 100, 0.9, True)
    assert n == int(g["init_n"]) and np.array_equal(m12, g["init_m12"]) and np.array_equal(pm, g["init_prev"])
    idx, d1, d2 = hamming_knn(c["desc1"], c["desc2"])
    assert np.array_equal(np.stack([idx, d1, d2]), g["knn"])
    from oracle.bow import fuse_search, search_for_triangulation

    ct = bow_cases.triangulation_pair(orc.extract)
    assert np.array_equal(ct["F12"], g["tri_F12"]) and np.array_equal(ct["ep"], g["tri_ep"])
    n, m12 = search_for_triangulation(ct["desc1"], ct["kps1"], ct["has_mp1"], ct["fv1"], ct["desc2"], ct["kps2"], ct["has_mp2"],
                                      ct["fv2"], g["tri_F12"], g["tri_ep"], T["scale"], T["sigma2"])
    assert n == int(g["tri_n"]) and np.array_equal(m12, g["tri_m12"])
    cf = bow_cases.fuse_case(orc.extract)
    Fk = FrameOracle(cf["kps"], cf["desc"], cf["bounds"], T["scale"])
    bi, bd = fuse_search(Fk, cf["q"], cf["t"], cf["K"], float(np.float32(np.log(np.float64(T["scale"][1])))), T["inv_sigma2"],
                         cf["xw"], cf["normal"], cf["min_dist"], cf["max_dist"], cf["mp_desc"], cf["skip"], 3.0)
    assert np.array_equal(bi, g["fuse_idx"]) and np.array_equal(bd, g["fuse_dist"])


def test_triangulation_oracle_properties():
    """Hand-checkable properties of the SearchForTriangulation restatement on two views of a plane with the
    true relative pose: matches only between map-point-free features, inside TH_LOW, on the epipolar line."""
    from oracle.bow import search_for_triangulation
    from oracle.orb import OrbOracle

    orc = OrbOracle(1000)
    T = orc.tables()
    c = bow_cases.triangulation_pair(orc.extract)
    args = (c["desc1"], c["kps1"], c["has_mp1"], c["fv1"], c["desc2"], c["kps2"], c["has_mp2"], c["fv2"], c["F12"], c["ep"],
            T["scale"], T["sigma2"])
    n, m12 = search_for_triangulation(*args)
    idx = np.nonzero(m12 >= 0)[0]
    assert n == len(idx) and n > 30
    assert not c["has_mp1"][idx].any() and not c["has_mp2"][m12[idx]].any()
    F = c["F12"].reshape(3, 3).astype(np.float64)
    for i1 in idx:
        k1, k2 = c["kps1"][i1], c["kps2"][m12[i1]]
        assert _dist(c["desc1"][i1], c["desc2"][m12[i1]]) <= 50
        l = np.array([k1["x"], k1["y"], 1.0]) @ F
        d2 = (l[0] * k2["x"] + l[1] * k2["y"] + l[2]) ** 2 / (l[0] ** 2 + l[1] ** 2)
        assert d2 < 3.84 * T["sigma2"][k2["octave"]] * 1.001
    # bCoarse drops the epipolar gate: at least as many raw matches before the orientation filter
    n_c, _ = search_for_triangulation(*args, coarse=True, check_ori=False)
    n_f, _ = search_for_triangulation(*args, coarse=False, check_ori=False)
    assert n_c >= n_f
    # a wrong geometry (transposed F) rejects most pairs
    bad = np.ascontiguousarray(c["F12"].reshape(3, 3).T.reshape(9))
    n_b, _ = search_for_triangulation(*args[:8], bad, *args[9:], check_ori=False)
    assert n_b < n_f


def test_fuse_oracle_properties():
    from oracle.bow import fuse_search
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    orc = OrbOracle(1000)
    T = orc.tables()
    c = bow_cases.fuse_case(orc.extract)
    F = FrameOracle(c["kps"], c["desc"], c["bounds"], T["scale"])
    bi, bd = fuse_search(F, c["q"], c["t"], c["K"], float(np.log(np.float32(1.2))), T["inv_sigma2"], c["xw"], c["normal"],
                         c["min_dist"], c["max_dist"], c["mp_desc"], c["skip"], 3.0)
    ok = bi >= 0
    assert ok.sum() > 100 and not ok[c["skip"] != 0].any()
    assert (bd[ok] <= 50).all() and (bd[~ok] == 256).all()
    for i in np.nonzero(ok)[0][:200]:
        assert _dist(c["mp_desc"][i], c["desc"][bi[i]]) == bd[i]
    # everything skipped / a camera looking away: nothing fuses
    bi2, _ = fuse_search(F, c["q"], c["t"], c["K"], float(np.log(np.float32(1.2))), T["inv_sigma2"], c["xw"], c["normal"],
                         c["min_dist"], c["max_dist"], c["mp_desc"], np.ones_like(c["skip"]), 3.0)
    assert (bi2 < 0).all()
    bi3, _ = fuse_search(F, c["q"], c["t"], c["K"], float(np.log(np.float32(1.2))), T["inv_sigma2"], c["xw"], -c["normal"],
                         c["min_dist"], c["max_dist"], c["mp_desc"], c["skip"], 3.0)
    assert (bi3 < 0).all()
