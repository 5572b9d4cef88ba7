"""GPU parity of the descriptor matchers that do not project -- SearchByBoW (both overloads),
SearchForInitialization -- and of the exhaustive Hamming search, against the oracle and the golden
fixture.  Integer work: bit-exact match arrays and counts."""
import os

import numpy as np
import pytest

from tests import bow_cases

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    from dvmslam_b200.tracking import Frame

    f = Frame(2048, np.ones(8, np.float32), np.ones(8, np.float32))
    yield f
    f.close()


def _both(ctx, kf_kf, c, nnratio, ori):
    from dvmslam_b200.matching import BowFeatures, BowMatcher
    from oracle.bow import search_by_bow

    n0, a0, b0 = search_by_bow(kf_kf, c["desc1"], c["angle1"], c["valid1"], c["fv1"], c["desc2"], c["angle2"],
                               c["valid2"], c["fv2"], nnratio, ori)
    m = BowMatcher(nnratio, ori)
    n1, a1, b1 = m._bow(ctx, kf_kf, BowFeatures(c["desc1"], c["angle1"], c["valid1"], c["fv1"]),
                        BowFeatures(c["desc2"], c["angle2"], c["valid2"], c["fv2"]))
    assert n0 == n1
    assert np.array_equal(a0, a1) and np.array_equal(b0, b1)
    return n1


@pytest.mark.parametrize("kf_kf", [0, 1])
def test_search_by_bow_frames(ctx, kf_kf):
    from oracle.orb import OrbOracle

    orc = OrbOracle(2000)
    c = bow_cases.bow_pair(orc.extract, w=1280, h=720)
    for nnratio in (0.6, 0.7, 0.9):
        for ori in (True, False):
            n = _both(ctx, kf_kf, c, nnratio, ori)
    assert n > 100
    # every feature in ONE vocabulary node: thousands of candidates per query (the chunked path)
    c1 = bow_cases.bow_pair(orc.extract, w=1280, h=720, one_node=True)
    assert _both(ctx, kf_kf, c1, 0.9, True) > 100


@pytest.mark.parametrize("kf_kf", [0, 1])
@pytest.mark.parametrize("n1,n2,seed,dup,nodes", [(140, 170, 0, False, 5), (900, 1100, 1, True, 9), (3000, 2500, 2, True, 3),
                                                  (40, 2000, 3, True, 1), (2000, 40, 4, False, 2), (1, 1, 5, False, 1)])
def test_search_by_bow_synthetic(ctx, kf_kf, n1, n2, seed, dup, nodes):
    c = bow_cases.bow_synthetic(n1, n2, seed, dup=dup, nodes=nodes)
    for ori in (True, False):
        _both(ctx, kf_kf, c, 0.75, ori)


def test_search_by_bow_edges(ctx):
    from dvmslam_b200 import DvmError
    from dvmslam_b200.matching import BowFeatures, BowMatcher

    c = bow_cases.bow_synthetic(60, 70, 7)
    empty = dict(c, desc1=c["desc1"][:0], angle1=c["angle1"][:0], valid1=c["valid1"][:0], fv1={})
    assert _both(ctx, 0, empty, 0.7, True) == 0
    nomp = dict(c, valid1=np.zeros_like(c["valid1"]))
    assert _both(ctx, 1, nomp, 0.7, True) == 0
    disjoint = dict(c, fv1={1000 + k: v for k, v in c["fv1"].items()})
    assert _both(ctx, 0, disjoint, 0.7, True) == 0
    # a feature listed under two nodes is not a DBoW2 feature vector: rejected, not silently matched
    bad = BowFeatures(c["desc1"], c["angle1"], c["valid1"], {1: [0, 1], 2: [1]})
    ok = BowFeatures(c["desc2"], c["angle2"], c["valid2"], c["fv2"])
    with pytest.raises(DvmError):
        BowMatcher()._bow(ctx, 0, bad, ok)


@pytest.mark.parametrize("nfeat,w,h", [(1000, 640, 480), (2000, 1280, 720), (10000, 1280, 720)])
def test_search_for_initialization(nfeat, w, h):
    """The mono initialiser's matcher (the 5 x nFeatures extractor feeds it, Tracking.cc:575-581)."""
    from dvmslam_b200.matching import BowMatcher
    from dvmslam_b200.tracking import Frame
    from oracle.bow import search_for_initialization
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    orc = OrbOracle(nfeat)
    T = orc.tables()
    c = bow_cases.init_pair(orc.extract, w=w, h=h)
    F0 = FrameOracle(c["kps2"], c["desc2"], c["bounds"], T["scale"])
    F1 = Frame(len(c["kps2"]) + 16, T["scale"], T["inv_sigma2"])
    F1.assign(c["kps2"], c["desc2"], c["bounds"])
    rng = np.random.default_rng(0)
    jitter = c["prev"] + rng.uniform(-30, 30, c["prev"].shape).astype(np.float32)
    for prev, window, nnratio, ori in [(c["prev"], 100, 0.9, True), (c["prev"], 100, 0.9, False), (jitter, 40, 0.75, True),
                                       (c["prev"], 300, 1.0, True)]:
        n0, m0, p0 = search_for_initialization(c["kps1"], c["desc1"], F0, prev, window, nnratio, ori)
        n1, m1, p1 = BowMatcher(nnratio, ori).SearchForInitialization(F1, c["kps1"], c["desc1"], prev, window)
        assert n0 == n1 and np.array_equal(m0, m1) and np.array_equal(p0, p1)
    assert n0 > 50
    F1.close()


def test_search_for_initialization_contention():
    """Many identical descriptors in F2: later queries displace earlier ones only at a strictly smaller
    distance (vMatchedDistance), which is what the fixed-point rounds must reproduce."""
    from dvmslam_b200.matching import BowMatcher
    from dvmslam_b200.tracking import Frame
    from oracle.bow import search_for_initialization
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    orc = OrbOracle(1000)
    T = orc.tables()
    c = bow_cases.init_pair(orc.extract)
    rng = np.random.default_rng(5)
    d2 = c["desc2"].copy()
    lvl0 = np.nonzero(c["kps2"]["octave"] == 0)[0]
    d1 = c["desc1"].copy()
    src = rng.choice(lvl0, 12)
    for j, i1 in enumerate(np.nonzero(c["kps1"]["octave"] == 0)[0]):   # every query is near one of 12 targets
        d1[i1] = d2[src[j % 12]]
        d1[i1, rng.integers(0, 32)] ^= np.uint8(1 << rng.integers(0, 8)) * (j % 3 != 0)
    F0 = FrameOracle(c["kps2"], d2, c["bounds"], T["scale"])
    F1 = Frame(len(c["kps2"]) + 16, T["scale"], T["inv_sigma2"])
    F1.assign(c["kps2"], d2, c["bounds"])
    centre = np.tile(np.array([[320.0, 240.0]], np.float32), (len(d1), 1))
    n0, m0, p0 = search_for_initialization(c["kps1"], d1, F0, centre, 400, 1.1, False)
    n1, m1, p1 = BowMatcher(1.1, False).SearchForInitialization(F1, c["kps1"], d1, centre, 400)
    assert n0 == n1 and np.array_equal(m0, m1) and np.array_equal(p0, p1)
    assert 0 < n0 <= 12
    F1.close()


@pytest.mark.parametrize("na,nb", [(0, 10), (1, 1), (5, 0), (100, 3000), (2000, 2000), (777, 70001)])
def test_hamming_knn_host(na, nb):
    from dvmslam_b200.matching import HammingKnn
    from oracle.bow import hamming_knn

    rng = np.random.default_rng(na + nb)
    a = rng.integers(0, 256, (na, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (nb, 32), dtype=np.uint8)
    if na and nb > 2:
        b[rng.integers(0, nb, na)] = a          # exact matches, some duplicated (ties)
        b[nb // 2] = b[nb // 3]
    h = HammingKnn()
    got = h.knn(a, b)
    want = hamming_knn(a, b)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    h.close()


def test_hamming_knn_batched_device():
    """Config C3's unit: keyframe blocks of one agent against another's; keys, accept counts and the
    size-independent properties (self-match at distance 0; planted pairs found)."""
    import torch

    from dvmslam_b200 import synth
    from dvmslam_b200.matching import HammingKnn
    from oracle.bow import hamming_knn

    KA, KB, N = 3, 4, 500
    A = synth.keyframe_blocks(KA, N, seed=1)
    B = synth.keyframe_blocks(KB, N, seed=2, shared_from=np.concatenate([A, A[:1]]))
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    k1 = torch.empty((KA, KB, N), dtype=torch.int32, device="cuda")
    k2 = torch.empty_like(k1)
    cnt = torch.empty((KA, KB), dtype=torch.int32, device="cuda")
    h = HammingKnn(stream=torch.cuda.current_stream().cuda_stream)
    h.knn_device(a.data_ptr(), KA, N, b.data_ptr(), KB, N, k1.data_ptr(), k2.data_ptr(), cnt.data_ptr(), 50, 0.75)
    h.sync()
    k1, k2, cnt = k1.cpu().numpy().view(np.uint32), k2.cpu().numpy().view(np.uint32), cnt.cpu().numpy()
    for i in range(KA):
        for j in range(KB):
            idx, d1, d2 = hamming_knn(A[i], B[j])
            assert np.array_equal(k1[i, j] >> 20, d1) and np.array_equal(k1[i, j] & 0xFFFFF, idx)
            assert np.array_equal(np.minimum(k2[i, j] >> 20, 256), d2)
            acc = (d1 <= 50) & (d1.astype(np.float32) < np.float32(0.75) * d2.astype(np.float32))
            assert cnt[i, j] == acc.sum()
    assert cnt[0, 0] > 100 and cnt[1, 1] > 100 and cnt[0, 1] == 0 and cnt[0, 3] > 100
    h.close()


def test_bow_golden_gpu(ctx):
    from dvmslam_b200.matching import BowFeatures, BowMatcher, HammingKnn
    from dvmslam_b200.tracking import Frame
    from oracle.orb import OrbOracle

    g = np.load(os.path.join(G, "bow_small.npz"))
    orc = OrbOracle(600)
    T = orc.tables()
    c = bow_cases.bow_pair(orc.extract)
    m = BowMatcher(0.7, True)
    for kf_kf in (0, 1):
        n, m12, m21 = m._bow(ctx, kf_kf, BowFeatures(c["desc1"], c["angle1"], c["valid1"], c["fv1"]),
                             BowFeatures(c["desc2"], c["angle2"], c["valid2"], c["fv2"]))
        assert n == int(g[f"bow{kf_kf}_n"]) and np.array_equal(m12, g[f"bow{kf_kf}_m12"]) and np.array_equal(m21, g[f"bow{kf_kf}_m21"])
    ci = bow_cases.init_pair(orc.extract)
    F = Frame(len(ci["kps2"]) + 16, T["scale"], T["inv_sigma2"])
    F.assign(ci["kps2"], ci["desc2"], ci["bounds"])
    n, m12, pm = BowMatcher(0.9, True).SearchForInitialization(F, ci["kps1"], ci["desc1"], ci["prev"], 100)
    assert n == int(g["init_n"]) and np.array_equal(m12, g["init_m12"]) and np.array_equal(pm, g["init_prev"])
    F.close()
    h = HammingKnn()
    assert np.array_equal(np.stack(h.knn(c["desc1"], c["desc2"])), g["knn"])
    h.close()
    from dvmslam_b200.matching import FuseSearch, SearchForTriangulation

    ct = bow_cases.triangulation_pair(orc.extract)
    n, m12 = SearchForTriangulation(ctx, BowFeatures(ct["desc1"], ct["kps1"]["angle"], ct["has_mp1"], ct["fv1"]), ct["kps1"],
                                    BowFeatures(ct["desc2"], ct["kps2"]["angle"], ct["has_mp2"], ct["fv2"]), ct["kps2"],
                                    g["tri_F12"], g["tri_ep"], T["scale"], T["sigma2"])
    assert n == int(g["tri_n"]) and np.array_equal(m12, g["tri_m12"])
    cf = bow_cases.fuse_case(orc.extract)
    Fk = Frame(len(cf["kps"]) + 16, T["scale"], T["inv_sigma2"])
    Fk.assign(cf["kps"], cf["desc"], cf["bounds"])
    bi, bd = FuseSearch(Fk, cf["q"], cf["t"], cf["K"], cf["xw"], cf["normal"], cf["min_dist"], cf["max_dist"], cf["mp_desc"],
                        cf["skip"], 3.0)
    assert np.array_equal(bi, g["fuse_idx"]) and np.array_equal(bd, g["fuse_dist"])
    Fk.close()


@pytest.mark.parametrize("w,h,nf", [(640, 480, 1000), (1280, 720, 2000)])
def test_search_for_triangulation(ctx, w, h, nf):
    from dvmslam_b200.matching import BowFeatures, SearchForTriangulation
    from oracle.bow import search_for_triangulation
    from oracle.orb import OrbOracle

    orc = OrbOracle(nf)
    T = orc.tables()
    c = bow_cases.triangulation_pair(orc.extract, w=w, h=h)
    one = dict(c, fv1={4: list(range(len(c["kps1"])))}, fv2={4: list(range(len(c["kps2"])))})   # one node: long candidate lists
    for case in (c, one):
        for coarse, ori in ((False, True), (False, False), (True, True)):
            n0, m0 = search_for_triangulation(case["desc1"], case["kps1"], case["has_mp1"], case["fv1"], case["desc2"], case["kps2"],
                                              case["has_mp2"], case["fv2"], case["F12"], case["ep"], T["scale"], T["sigma2"], coarse, ori)
            a = BowFeatures(case["desc1"], case["kps1"]["angle"], case["has_mp1"], case["fv1"])
            b = BowFeatures(case["desc2"], case["kps2"]["angle"], case["has_mp2"], case["fv2"])
            n1, m1 = SearchForTriangulation(ctx, a, case["kps1"], b, case["kps2"], case["F12"], case["ep"], T["scale"], T["sigma2"],
                                            coarse, ori)
            assert n0 == n1 and np.array_equal(m0, m1), (coarse, ori)
    assert n0 > 30


@pytest.mark.parametrize("w,h,nf,th", [(640, 480, 1000, 3.0), (1280, 720, 2000, 3.0), (1280, 720, 2000, 8.0)])
def test_fuse_search(w, h, nf, th):
    from dvmslam_b200.matching import FuseSearch
    from dvmslam_b200.tracking import Frame
    from oracle.bow import fuse_search
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    orc = OrbOracle(nf)
    T = orc.tables()
    c = bow_cases.fuse_case(orc.extract, w=w, h=h, n_points=3000)
    F0 = FrameOracle(c["kps"], c["desc"], c["bounds"], T["scale"])
    F1 = Frame(len(c["kps"]) + 16, T["scale"], T["inv_sigma2"])
    F1.assign(c["kps"], c["desc"], c["bounds"])
    log_scale = float(np.float32(np.log(np.float64(T["scale"][1]))))
    i0, d0 = fuse_search(F0, c["q"], c["t"], c["K"], log_scale, T["inv_sigma2"], c["xw"], c["normal"], c["min_dist"], c["max_dist"],
                         c["mp_desc"], c["skip"], th)
    i1, d1 = FuseSearch(F1, c["q"], c["t"], c["K"], c["xw"], c["normal"], c["min_dist"], c["max_dist"], c["mp_desc"], c["skip"], th)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)
    assert (i1 >= 0).sum() > 100
    F1.close()


@pytest.mark.parametrize("KA,NA,KB,NB", [(1, 128, 1, 128), (2, 300, 3, 257), (1, 1, 1, 1), (3, 2000, 2, 2000), (1, 513, 2, 129),
                                         (2, 256, 1, 3000)])
def test_hamming_tensor_core_matches_popcount_and_oracle(KA, NA, KB, NB):
    """The tcgen05 int8 formulation (a . b = 256 - 2 * distance over +1 / -1 bytes) against the popcount kernel and the
    oracle: keys (distance << 20 | first nearest index), second-smallest keys and accept counts bit-identical, with
    duplicated rows (ties -> first index), all-ones against all-zeros rows (distance 256 never matches), ragged sizes
    (rows and columns past the last full 128-tile) and blocks that are not multiples of the tile."""
    import torch

    from dvmslam_b200.matching import HammingKnn
    from oracle.bow import hamming_knn

    rng = np.random.default_rng(KA * 1000 + NB)
    A = rng.integers(0, 256, (KA, NA, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (KB, NB, 32), dtype=np.uint8)
    if NB > 4:
        for j in range(KB):
            take = rng.integers(0, NA, max(1, NA // 3))
            B[j, rng.integers(0, NB, len(take))] = A[j % KA, take] ^ (rng.random((len(take), 32)) < 0.02).astype(np.uint8)
            B[j, NB - 1] = B[j, NB // 2]          # a tie across tiles: the lower index wins
            B[j, 1] = B[j, 0]
        A[0, 0] = 0xFF
        B[0, 2] = 0x00                            # distance 256
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    out = {}
    h = HammingKnn(stream=torch.cuda.current_stream().cuda_stream)
    for mode in (1, 2):
        k1 = torch.full((KA, KB, NA), -1, dtype=torch.int32, device="cuda")
        k2 = torch.full_like(k1, -1)
        cnt = torch.full((KA, KB), -1, dtype=torch.int32, device="cuda")
        h.set_mode(mode)
        h.knn_device(a.data_ptr(), KA, NA, b.data_ptr(), KB, NB, k1.data_ptr(), k2.data_ptr(), cnt.data_ptr(), 50, 0.75)
        h.sync()
        out[mode] = (k1.cpu().numpy().view(np.uint32), k2.cpu().numpy().view(np.uint32), cnt.cpu().numpy())
    h.close()
    for x, y in zip(out[1], out[2]):
        assert np.array_equal(x, y)
    k1, k2, cnt = out[2]
    for i in range(KA):
        for j in range(KB):
            idx, d1, d2 = hamming_knn(A[i], B[j])
            assert np.array_equal(np.minimum(k1[i, j] >> 20, 256), d1)
            assert np.array_equal(np.where(d1 < 256, k1[i, j] & 0xFFFFF, -1), idx)
            assert np.array_equal(np.minimum(k2[i, j] >> 20, 256), d2)


def _kf(c, T, tag=""):
    from dvmslam_b200.tracking import Frame
    from oracle.track import FrameOracle

    kps, desc = c["kps" + tag], c["desc" + tag]
    F0 = FrameOracle(kps, desc, c["bounds"], T["scale"])
    F1 = Frame(len(kps) + 16, T["scale"], T["inv_sigma2"])
    F1.assign(kps, desc, c["bounds"])
    return F0, F1


@pytest.mark.parametrize("w,h,nf,th,ratio", [(640, 480, 1000, 8, 1.5), (1280, 720, 2000, 4, 1.0), (640, 480, 1000, 30, 1.0)])
def test_search_by_projection_sim3(w, h, nf, th, ratio):
    """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (O3/src/ORBmatcher.cc:395-603): gate pass +
    sequential-greedy resolution on the GPU against the oracle (pinned to the reference in tests/test_ref_matchers.py);
    th = 30 packs many candidates onto each keypoint."""
    from dvmslam_b200.matching import SearchByProjectionSim3
    from oracle.bow import search_by_projection_sim3
    from oracle.orb import OrbOracle

    orc = OrbOracle(nf)
    T = orc.tables()
    c = bow_cases.sim3_projection_case(orc.extract, w=w, h=h, n_points=3000)
    F0, F1 = _kf(c, T)
    pts = (c["xw"], c["normal"], c["min_dist"], c["max_dist"], c["mp_desc"], c["skip"], c["kp_matched"])
    n0, k0 = search_by_projection_sim3(F0, c["sq"], c["st"], c["K"], float(np.log(np.float32(T["scale"][1]))), 8, *pts, th, ratio)
    n1, k1 = SearchByProjectionSim3(F1, c["sq"], c["st"], c["K"], *pts, th, ratio)
    assert n0 == n1 and np.array_equal(k0, k1)
    assert n0 > 100
    F1.close()


@pytest.mark.parametrize("w,h,nf,th", [(640, 480, 1000, 3.0), (1280, 720, 2000, 4.0)])
def test_fuse_search_sim3(w, h, nf, th):
    from dvmslam_b200.matching import FuseSearchSim3
    from oracle.bow import fuse_search_sim3
    from oracle.orb import OrbOracle

    orc = OrbOracle(nf)
    T = orc.tables()
    c = bow_cases.sim3_projection_case(orc.extract, w=w, h=h, n_points=3000, scale=0.7)
    F0, F1 = _kf(c, T)
    pts = (c["xw"], c["normal"], c["min_dist"], c["max_dist"], c["mp_desc"], c["skip"])
    i0, d0 = fuse_search_sim3(F0, c["sq"], c["st"], c["K"], float(np.log(np.float32(T["scale"][1]))), 8, *pts, th)
    i1, d1 = FuseSearchSim3(F1, c["sq"], c["st"], c["K"], *pts, th)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1) and (i1 >= 0).sum() > 100
    F1.close()


@pytest.mark.parametrize("w,h,nf,th", [(640, 480, 1000, 7.5), (1280, 720, 2000, 7.5)])
def test_search_by_sim3(w, h, nf, th):
    from dvmslam_b200.matching import SearchBySim3
    from oracle.bow import search_by_sim3
    from oracle.orb import OrbOracle

    orc = OrbOracle(nf)
    T = orc.tables()
    c = bow_cases.search_by_sim3_case(orc.extract, w=w, h=h)
    (A0, A1), (B0, B1) = _kf(c, T, "1"), _kf(c, T, "2")
    sides = [(c["skip" + s], c["xw" + s], c["min" + s], c["max" + s], c["mpdesc" + s]) for s in "12"]
    poses = (c["q1"], c["t1"], c["q2"], c["t2"], c["s12q"], c["s12t"])
    n0, m0 = search_by_sim3(A0, B0, *poses, c["K"], float(np.log(np.float32(T["scale"][1]))), 8, *sides, th)
    n1, m1 = SearchBySim3(A1, B1, *poses, c["K"], *sides, th)
    assert n0 == n1 and np.array_equal(m0, m1) and n0 > 50
    A1.close()
    B1.close()
