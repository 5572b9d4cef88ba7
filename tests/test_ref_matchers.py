"""The reference's own ORBmatcher.cc -- compiled unmodified from /root/reference against oracle/slamshim (mini Eigen / Sophus
that restate the evaluation order, shim Frame / KeyFrame / MapPoint whose member bodies are extracted from the reference's .cc
files at build time; `make -C oracle ref` -> oracle/_ref/libref_matcher.so) -- must give the same match indices as the oracle
restatements (oracle/track_oracle.cpp, oracle/bow_oracle.cpp) on the inputs of the GPU parity tests.  This pins the matcher
oracle, and through it the CUDA kernels (tests/test_track_gpu.py, tests/test_bow_gpu.py compare kernels and oracle on these very
cases), to the reference's source.  Skipped where neither the built library nor /root/reference is available."""
import numpy as np
import pytest

from tests import bow_cases  # noqa: E402
from dvmslam_b200 import synth
from oracle import refm

pytestmark = pytest.mark.skipif(not refm.available(), reason="oracle/_ref/libref_matcher.so not built and /root/reference absent")


@pytest.fixture(scope="module")
def world():
    from oracle.orb import OrbOracle

    S = synth.PlaneStream(seed=0)
    orc = OrbOracle(2000)
    T = orc.tables()
    cases = {k: synth.tracking_case(S, k, orc.extract) for k in (3, 12)}
    return dict(S=S, T=T, cases=cases, orc=orc)


def _frames(world, case):
    from oracle.track import FrameOracle

    T = world["T"]
    return (FrameOracle(case["cur_kps"], case["cur_desc"], case["bounds"], T["scale"]),
            refm.RefFrame(case["cur_kps"], case["cur_desc"], case["bounds"], T["scale"], case["K"]))


@pytest.mark.parametrize("k", [3, 12])
def test_grid_and_area_queries(world, k):
    """Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (O3/src/Frame.cc:481-506,712-782) and
    KeyFrame::GetFeaturesInArea (O3/src/KeyFrame.cc:750-790)."""
    F0, FR = _frames(world, world["cases"][k])
    rng = np.random.default_rng(k)
    for ix, iy in [(0, 0), (63, 47), (10, 20), (32, 24)] + [tuple(rng.integers(0, [64, 48])) for _ in range(60)]:
        assert np.array_equal(F0.grid_cell(int(ix), int(iy)), FR.grid_cell(int(ix), int(iy))), (ix, iy)
    for _ in range(150):
        x, y = float(rng.uniform(-50, 1330)), float(rng.uniform(-50, 770))
        r = float(rng.choice([5.0, 15.0, 37.3, 120.0]))
        lv = int(rng.integers(0, 8))
        for (a, b) in [(-1, -1), (lv - 1, lv + 1), (lv - 1, lv), (0, lv), (lv, -1)]:
            assert np.array_equal(F0.features_in_area(x, y, r, a, b), FR.features_in_area(x, y, r, a, b)), (x, y, r, a, b)
        assert np.array_equal(F0.features_in_area(x, y, r), FR.kf_features_in_area(x, y, r))


def test_descriptor_distance():
    from oracle.track import descriptor_distance

    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b = rng.integers(0, 256, 32, dtype=np.uint8), rng.integers(0, 256, 32, dtype=np.uint8)
        assert descriptor_distance(a, b) == refm.descriptor_distance(a, b)


@pytest.mark.parametrize("k,th", [(3, 15.0), (3, 30.0), (12, 15.0), (12, 7.0)])
def test_search_by_projection_last(world, k, th):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono), O3/src/ORBmatcher.cc:1553-1748."""
    case = world["cases"][k]
    F0, FR = _frames(world, case)
    lk = case["last_kps"]
    q = case["qcw_prior"]
    tail = (case["has_mp"], case["outlier"], case["last_Xw"], case["last_desc"], case["obs_pos"], lk["octave"], lk["angle"], th)
    for ori in (True, False):
        n0, m0 = F0.search_by_projection_last(q, case["tcw_prior"], case["K"], *tail, check_ori=ori)
        n1, m1 = FR.search_by_projection_last(q, case["tcw_prior"], *tail, check_ori=ori)
        assert n0 == n1 and np.array_equal(m0, m1), ori
        assert n0 > 300


def test_search_by_projection_last_heavy_contention(world):
    case = world["cases"][3]
    F0, FR = _frames(world, case)
    lk = case["last_kps"]
    rng = np.random.default_rng(1)
    Xw = case["last_Xw"].copy()
    Xw[:, :2] = Xw[rng.integers(0, 40, len(Xw)), :2] + rng.normal(0, 0.01, (len(Xw), 2)).astype(np.float32)
    obs = (rng.random(len(Xw)) < 0.7).astype(np.uint8)
    q = case["qcw_prior"]
    tail = (np.ones_like(case["has_mp"]), np.zeros_like(case["outlier"]), Xw, case["last_desc"], obs, lk["octave"], lk["angle"], 30.0)
    n0, m0 = F0.search_by_projection_last(q, case["tcw_prior"], case["K"], *tail)
    n1, m1 = FR.search_by_projection_last(q, case["tcw_prior"], *tail)
    assert n0 == n1 and np.array_equal(m0, m1)


def _map_points(case, rng):
    """Local-map points of the case with normals and distance ranges as MapPoint::UpdateNormalAndDepth leaves them."""
    X = case["map_Xw"].astype(np.float32)
    C0 = np.array([0.0, 0.0, -3.0])
    PO = X.astype(np.float64) - C0
    d = np.linalg.norm(PO, axis=1)
    normal = (PO / d[:, None]).astype(np.float32)
    max_d = (d * 1.2 ** case["map_octave"].astype(np.float64)).astype(np.float32)
    min_d = (max_d / np.float32(1.2 ** 7)).astype(np.float32)
    return X, normal, min_d, max_d


@pytest.mark.parametrize("k", [3, 12])
def test_is_in_frustum(world, k):
    """Frame::isInFrustum + MapPoint::PredictScale (O3/src/Frame.cc:575-636, O3/src/MapPoint.cc:573-587): flags, projections
    (bit-equal floats), predicted levels and viewing cosines."""
    from oracle.track import is_in_frustum

    case = world["cases"][k]
    rng = np.random.default_rng(k)
    X, normal, min_d, max_d = _map_points(case, rng)
    X = np.concatenate([X, X[:200] * np.float32(1.7), -X[:50]])            # some points out of view / behind the camera
    normal = np.concatenate([normal, normal[:200], normal[:50]])
    min_d, max_d = np.concatenate([min_d, min_d[:200], min_d[:50]]), np.concatenate([max_d, max_d[:200], max_d[:50]])
    skip = (rng.random(len(X)) < 0.1).astype(np.uint8)
    q = synth.quat_from_R(np.asarray(case["Rcw_true"], np.float64)).astype(np.float32)
    t = np.asarray(case["tcw_true"], np.float32)
    T = world["T"]
    for cos_limit in (0.5, 0.9):
        a = is_in_frustum(q, t, case["K"], case["bounds"], 8, T["scale"][1], X, normal, min_d, max_d, skip, cos_limit)
        b = refm.is_in_frustum(q, t, case["K"], case["bounds"], 8, T["scale"][1], X, normal, min_d, max_d, skip, cos_limit)
        assert np.array_equal(a[0], b[0])
        v = a[0] != 0
        assert v.sum() > 1000
        for x, y in zip(a[1:], b[1:]):
            assert np.array_equal(x[v].view(np.uint32) if x.dtype == np.float32 else x[v],
                                  y[v].view(np.uint32) if y.dtype == np.float32 else y[v])


@pytest.mark.parametrize("k,th,nnratio", [(3, 1.0, 0.8), (12, 1.0, 0.8), (12, 5.0, 0.8), (3, 3.0, 0.6)])
def test_search_by_projection_map(world, k, th, nnratio):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th), O3/src/ORBmatcher.cc:44-205 (inputs: the mTrack* fields)."""
    case = world["cases"][k]
    F0, FR = _frames(world, case)
    X = case["map_Xw"].astype(np.float32)
    Xc = X @ case["Rcw_true"].T.astype(np.float32) + case["tcw_true"].astype(np.float32)
    K = case["K"]
    u = (K[0] * Xc[:, 0] / Xc[:, 2] + K[2]).astype(np.float32)
    v = (K[1] * Xc[:, 1] / Xc[:, 2] + K[3]).astype(np.float32)
    ok = (Xc[:, 2] > 0) & (u >= 0) & (u < 1280) & (v >= 0) & (v < 720)
    rng = np.random.default_rng(k)
    level = case["map_octave"][ok].astype(np.int32)
    cosv = rng.choice([0.9995, 0.9], ok.sum()).astype(np.float32)
    obs = (rng.random(ok.sum()) < 0.95).astype(np.uint8)
    blocked = (rng.random(len(case["cur_kps"])) < 0.3).astype(np.uint8)
    args = (u[ok], v[ok], level, cosv, case["map_desc"][ok], obs, th, nnratio, blocked)
    n0, m0 = F0.search_by_projection_map(*args)
    n1, m1 = FR.search_by_projection_map(*args)
    assert n0 == n1 and np.array_equal(m0, m1)
    assert n0 > 200


@pytest.mark.parametrize("kf_kf", [0, 1])
def test_search_by_bow(world, kf_kf):
    """SearchByBoW(pKF, F, ...) O3/src/ORBmatcher.cc:214-393 and SearchByBoW(pKF1, pKF2, ...) :709-834."""
    from oracle.bow import search_by_bow
    from oracle.orb import OrbOracle

    orc = OrbOracle(600)
    cases = [bow_cases.bow_pair(orc.extract), bow_cases.bow_pair(orc.extract, one_node=True), bow_cases.bow_synthetic(300, 400, 1),
             bow_cases.bow_synthetic(500, 200, 2, dup=True), bow_cases.bow_synthetic(40, 40, 3, dup=True, nodes=2)]
    for c in cases:
        for nnratio, ori in ((0.7, True), (0.6, False), (0.9, True)):
            a = (c["desc1"], c["angle1"], c["valid1"], c["fv1"], c["desc2"], c["angle2"], c["valid2"] if kf_kf else None, c["fv2"])
            r0 = search_by_bow(kf_kf, *a, nnratio, ori)
            r1 = refm.search_by_bow(kf_kf, *a, nnratio, ori)
            assert r0[0] == r1[0] and np.array_equal(r0[1], r1[1]) and np.array_equal(r0[2], r1[2]), (nnratio, ori)


def test_search_for_initialization(world):
    """SearchForInitialization, O3/src/ORBmatcher.cc:605-707 (vbPrevMatched updated in place)."""
    from oracle.bow import search_for_initialization
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    for nf in (1000, 5000):
        orc = OrbOracle(nf)
        T = orc.tables()
        c = bow_cases.init_pair(orc.extract)
        F0 = FrameOracle(c["kps2"], c["desc2"], c["bounds"], T["scale"])
        FR = refm.RefFrame(c["kps2"], c["desc2"], c["bounds"], T["scale"])
        for window, nnratio, ori in ((100, 0.9, True), (30, 0.9, False), (100, 0.6, True)):
            r0 = search_for_initialization(c["kps1"], c["desc1"], F0, c["prev"], window, nnratio, ori)
            r1 = refm.search_for_initialization(c["kps1"], c["desc1"], FR, c["prev"], window, nnratio, ori)
            assert r0[0] == r1[0] and np.array_equal(r0[1], r1[1]) and np.array_equal(r0[2], r1[2])
            assert r0[0] > 50


@pytest.mark.parametrize("w,h,nf", [(640, 480, 1000), (1280, 720, 2000)])
def test_search_for_triangulation(w, h, nf):
    """SearchForTriangulation (O3/src/ORBmatcher.cc:836-1058) with the epipolar gate of Pinhole::epipolarConstrain
    (O3/src/CameraModels/Pinhole.cpp:104-127): the reference derives R12, t12 from the keyframe poses with Sophus and rebuilds
    F12 = K1^-T [t12]x R12 K2^-1 per candidate; the oracle takes F12 and the epipole from oracle.bow.fundamental_from_poses,
    which restates that arithmetic."""
    from oracle.bow import fundamental_from_poses, search_for_triangulation
    from oracle.orb import OrbOracle

    orc = OrbOracle(nf)
    T = orc.tables()
    c = bow_cases.triangulation_pair(orc.extract, w=w, h=h)
    (q1, t1), (q2, t2) = c["poses"]
    F12, ep = fundamental_from_poses(q1, t1, q2, t2, c["K"], c["K"])
    bounds = (0.0, 0.0, float(w), float(h))
    one = dict(c, fv1={4: list(range(len(c["kps1"])))}, fv2={4: list(range(len(c["kps2"])))})
    for case in (c, one):
        R1 = refm.RefFrame(case["kps1"], case["desc1"], bounds, T["scale"], c["K"])
        R2 = refm.RefFrame(case["kps2"], case["desc2"], bounds, T["scale"], c["K"])
        R1.set_feature_vector(case["fv1"])
        R2.set_feature_vector(case["fv2"])
        for coarse, ori in ((False, True), (False, False), (True, True)):
            n0, m0 = search_for_triangulation(case["desc1"], case["kps1"], case["has_mp1"], case["fv1"], case["desc2"], case["kps2"],
                                              case["has_mp2"], case["fv2"], F12, ep, T["scale"], T["sigma2"], coarse, ori)
            n1, m1 = refm.search_for_triangulation(R1, case["has_mp1"], q1, t1, R2, case["has_mp2"], q2, t2, 0.6, ori, coarse)
            assert n0 == n1 and np.array_equal(m0, m1), (coarse, ori)
        assert n0 > 30


@pytest.mark.parametrize("w,h,nf,th", [(640, 480, 1000, 3.0), (1280, 720, 2000, 3.0), (1280, 720, 2000, 8.0)])
def test_fuse(w, h, nf, th):
    """Fuse(pKF, vpMapPoints, th), O3/src/ORBmatcher.cc:1060-1228: the keypoint every map point is fused with."""
    from oracle.bow import fuse_search
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    orc = OrbOracle(nf)
    T = orc.tables()
    c = bow_cases.fuse_case(orc.extract, w=w, h=h, n_points=3000)
    F0 = FrameOracle(c["kps"], c["desc"], c["bounds"], T["scale"])
    FR = refm.RefFrame(c["kps"], c["desc"], c["bounds"], T["scale"], c["K"])
    log_scale = float(np.log(np.float32(T["scale"][1])))
    i0, d0 = fuse_search(F0, c["q"], c["t"], c["K"], log_scale, T["inv_sigma2"], c["xw"], c["normal"], c["min_dist"], c["max_dist"],
                         c["mp_desc"], c["skip"], th)
    n1, i1 = FR.fuse(c["q"], c["t"], c["xw"], c["normal"], c["min_dist"], c["max_dist"], c["mp_desc"], c["skip"], th)
    assert np.array_equal(i0, i1)
    assert n1 == (i0 >= 0).sum() > 100


def _log_scale(T):
    return float(np.log(np.float32(T["scale"][1])))


@pytest.mark.parametrize("w,h,nf,th,ratio", [(640, 480, 1000, 8, 1.5), (1280, 720, 2000, 4, 1.0), (640, 480, 1000, 30, 1.0)])
def test_search_by_projection_sim3(w, h, nf, th, ratio):
    """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming), O3/src/ORBmatcher.cc:395-494 (and the overload with
    vpPointsKFs, :496-603, whose matching is the same): which keypoint ends up holding which candidate."""
    from oracle.bow import search_by_projection_sim3
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    orc = OrbOracle(nf)
    T = orc.tables()
    c = bow_cases.sim3_projection_case(orc.extract, w=w, h=h, n_points=3000)
    F0 = FrameOracle(c["kps"], c["desc"], c["bounds"], T["scale"])
    FR = refm.RefFrame(c["kps"], c["desc"], c["bounds"], T["scale"], c["K"])
    pts = (c["xw"], c["normal"], c["min_dist"], c["max_dist"], c["mp_desc"], c["skip"], c["kp_matched"])
    n0, k0 = search_by_projection_sim3(F0, c["sq"], c["st"], c["K"], _log_scale(T), 8, *pts, th, ratio)
    n1, k1 = refm.search_by_projection_sim3(FR, c["sq"], c["st"], *pts, th, ratio)
    assert n0 == n1 and np.array_equal(k0, k1)
    assert n0 > 100 and not (k0[c["kp_matched"] != 0] >= 0).any()


@pytest.mark.parametrize("w,h,nf,th", [(640, 480, 1000, 3.0), (1280, 720, 2000, 4.0)])
def test_fuse_sim3(w, h, nf, th):
    """Fuse(pKF, Scw, vpPoints, th, vpReplacePoint), O3/src/ORBmatcher.cc:1236-1345."""
    from oracle.bow import fuse_search_sim3
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    orc = OrbOracle(nf)
    T = orc.tables()
    c = bow_cases.sim3_projection_case(orc.extract, w=w, h=h, n_points=3000, scale=0.7)
    F0 = FrameOracle(c["kps"], c["desc"], c["bounds"], T["scale"])
    FR = refm.RefFrame(c["kps"], c["desc"], c["bounds"], T["scale"], c["K"])
    pts = (c["xw"], c["normal"], c["min_dist"], c["max_dist"], c["mp_desc"], c["skip"])
    i0, d0 = fuse_search_sim3(F0, c["sq"], c["st"], c["K"], _log_scale(T), 8, *pts, th)
    n1, i1 = refm.fuse_sim3(FR, c["sq"], c["st"], *pts, th)
    assert np.array_equal(i0, i1) and n1 == (i0 >= 0).sum() > 100


@pytest.mark.parametrize("w,h,nf,th", [(640, 480, 1000, 7.5), (1280, 720, 2000, 7.5)])
def test_search_by_sim3(w, h, nf, th):
    """SearchBySim3(pKF1, pKF2, vpMatches12, S12, th), O3/src/ORBmatcher.cc:1347-1551."""
    from oracle.bow import search_by_sim3
    from oracle.orb import OrbOracle
    from oracle.track import FrameOracle

    orc = OrbOracle(nf)
    T = orc.tables()
    c = bow_cases.search_by_sim3_case(orc.extract, w=w, h=h)
    F = [FrameOracle(c["kps" + s], c["desc" + s], c["bounds"], T["scale"]) for s in "12"]
    R = [refm.RefFrame(c["kps" + s], c["desc" + s], c["bounds"], T["scale"], c["K"]) for s in "12"]
    sides = [(c["skip" + s], c["xw" + s], c["min" + s], c["max" + s], c["mpdesc" + s]) for s in "12"]
    poses = (c["q1"], c["t1"], c["q2"], c["t2"], c["s12q"], c["s12t"])
    n0, m0 = search_by_sim3(F[0], F[1], *poses, c["K"], _log_scale(T), 8, *sides, th)
    n1, m1 = refm.search_by_sim3(R[0], R[1], *poses, *sides, th)
    assert n0 == n1 and np.array_equal(m0, m1)
    assert n0 > 50
