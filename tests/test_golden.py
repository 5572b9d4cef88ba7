"""Oracle (and, on the GPU box, the CUDA path) against the committed golden vectors of tests/golden/
(made by tests/golden/make_golden.py from cv2 4.13.0 + the restated in-tree logic)."""
import ctypes as C
import os

import numpy as np
import pytest

from dvmslam_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
u8p = C.POINTER(C.c_uint8)


def P(a):
    return a.ctypes.data_as(u8p)


def test_oracle_models_reproduce_cv2_golden():
    from oracle import lib

    L = lib()
    g = np.load(os.path.join(G, "cv2_primitives.npz"))
    img = np.ascontiguousarray(g["img"])
    h, w = img.shape
    dh, dw = g["resized"].shape
    out = np.zeros((dh, dw), np.uint8)
    L.cvm_resize_linear_u8(P(img), w, h, w, P(out), dw, dh, dw)
    assert np.array_equal(out, g["resized"])
    out = np.zeros_like(img)
    L.cvm_gaussian7_u8(P(img), w, h, w, P(out), w)
    assert np.array_equal(out, g["blurred"])

    class KP(C.Structure):
        _fields_ = [("x", C.c_int), ("y", C.c_int), ("r", C.c_int)]

    for th, key in ((20, "fast20"), (7, "fast7")):
        buf = (KP * 100000)()
        n = L.cvm_fast_detect(P(img), w, h, w, th, buf, 100000)
        got = np.array([(buf[i].x, buf[i].y, buf[i].r) for i in range(n)], np.int32).reshape(-1, 3)
        assert np.array_equal(got, g[key])
    L.cvm_fast_atan2.restype = C.c_float
    L.cvm_fast_atan2.argtypes = [C.c_float, C.c_float]
    got = np.array([L.cvm_fast_atan2(float(y), float(x)) for y, x in g["atan_yx"]], np.float32)
    assert np.array_equal(got, g["atan_deg"])


def test_oracle_extractor_reproduces_golden():
    from oracle.orb import OrbOracle

    g = np.load(os.path.join(G, "orb_small.npz"))
    k, d, m = OrbOracle(int(g["nfeatures"])).extract(g["frame"])
    assert m == int(g["mono"]) and np.array_equal(k, g["kps"]) and np.array_equal(d, g["desc"])


def _track_case():
    from oracle.orb import OrbOracle

    S = synth.PlaneStream(640, 480, seed=3, K=(500.0, 500.0, 320.0, 240.0))
    orc = OrbOracle(800)
    return S, orc, orc.tables(), synth.tracking_case(S, 4, orc.extract, n_local=1500)


def test_oracle_tracking_and_ba_reproduce_golden():
    from oracle.lba import local_ba
    from oracle.track import FrameOracle, pose_optimization

    g = np.load(os.path.join(G, "track_small.npz"))
    S, orc, T, case = _track_case()
    F = FrameOracle(case["cur_kps"], case["cur_desc"], case["bounds"], T["scale"])
    lk = case["last_kps"]
    n, cur_mp = F.search_by_projection_last(case["qcw_prior"], case["tcw_prior"], case["K"], case["has_mp"],
                                            case["outlier"], case["last_Xw"], case["last_desc"], case["obs_pos"],
                                            lk["octave"], lk["angle"], 15.0)
    assert n == int(g["nmatches"]) and np.array_equal(cur_mp, g["cur_mp"])
    idx = np.nonzero(cur_mp >= 0)[0]
    ck = case["cur_kps"]
    q = synth.quat_from_R(case["Rcw_prior"].astype(np.float64)).astype(np.float32)
    r, q2, t2, outl, _ = pose_optimization(q, case["tcw_prior"], case["K"], case["last_Xw"][cur_mp[idx]],
                                           np.stack([ck["x"][idx], ck["y"][idx]], 1), T["inv_sigma2"][ck["octave"][idx]])
    assert r == int(g["pose_inliers"]) and np.array_equal(outl, g["pose_outlier"])
    assert np.allclose(q2, g["pose_q"], atol=1e-7) and np.allclose(t2, g["pose_t"], atol=1e-7)
    B = synth.ba_scene(6, 2, 120, seed=5)
    ba = local_ba(B["cam_q"], B["cam_t"], B["cam_fixed"], B["pts"], B["edge_cam"], B["edge_pt"], B["edge_obs"],
                  B["edge_w"], B["K"])
    assert ba["iters"] == int(g["ba_iters"]) and np.array_equal(ba["bad"], g["ba_bad"])
    assert np.allclose(ba["cam_t"], g["ba_cam_t"], atol=1e-6) and np.allclose(ba["pts"], g["ba_pts"], atol=1e-5)


def test_oracle_keyframe_operators_reproduce_golden():
    """Welding BA, OptimizeSim3 and the essential-graph solve of the oracle against their committed outputs (regression
    pins).  The Sim3 paths differentiate numerically with delta 1e-9, so a different libm may move them by ~1e-7."""
    from oracle.lba import merge_ba
    from oracle.sim3 import optimize_essential_graph, optimize_sim3
    from tests.sim3_cases import loop_graph

    g = np.load(os.path.join(G, "keyframe_ops.npz"))
    B = synth.ba_scene(8, 3, 200, seed=11)
    m = merge_ba(B["cam_q"], B["cam_t"], B["cam_fixed"], B["pts"], B["edge_cam"], B["edge_pt"], B["edge_obs"], B["edge_w"], B["K"])
    assert [m["iters"], m["iters_first"], m["excluded"], m["trials"]] == list(g["merge_counts"])
    assert np.array_equal(m["bad"], g["merge_bad"]) and np.allclose([m["chi_first"], m["chi_last"]], g["merge_chi"], rtol=1e-9)
    assert np.allclose(m["cam_t"], g["merge_cam_t"], atol=1e-6) and np.allclose(m["cam_q"], g["merge_cam_q"], atol=1e-7)
    assert np.allclose(m["pts"], g["merge_pts"], atol=1e-5)
    S = synth.sim3_scene(80, seed=11, scale=1.25)
    s3 = optimize_sim3(S["p1c"], S["p2c"], S["obs1"], S["obs2"], S["w1"], S["w2"], S["K"], S["K"], S["q0"], S["t0"], S["s0"], 10.0, False)
    assert [s3["n_in"], s3["iters1"], s3["iters2"], s3["n_bad"]] == list(g["sim3_counts"])
    assert np.array_equal(s3["inlier"], g["sim3_inlier"])
    assert np.allclose(s3["q"], g["sim3_q"], atol=1e-6) and np.allclose(s3["t"], g["sim3_t"], atol=1e-5) and abs(s3["s"] - float(g["sim3_s"])) < 1e-6
    est, fixed, vi, vj, meas, _ = loop_graph(16, seed=2)
    eg = optimize_essential_graph(est, fixed, vi, vj, meas)
    assert eg["iters"] == int(g["eg_counts"][0]) and np.allclose(eg["sim3"], g["eg_sim3"], atol=1e-6)
    assert np.allclose([eg["chi_first"], eg["chi_last"]], g["eg_chi"], rtol=1e-4)
    egf = optimize_essential_graph(est, fixed, vi, vj, meas, fix_scale=True)
    assert egf["iters"] == int(g["egf_counts"][0]) and np.allclose(egf["sim3"], g["egf_sim3"], atol=1e-6)


@pytest.mark.gpu
def test_gpu_extractor_reproduces_golden():
    from dvmslam_b200.extractor import ORBextractor

    g = np.load(os.path.join(G, "orb_small.npz"))
    ext = ORBextractor(int(g["nfeatures"]), max_width=320, max_height=240)
    k, d, m = ext(g["frame"])
    ext.close()
    assert m == int(g["mono"]) and np.array_equal(k, g["kps"]) and np.array_equal(d, g["desc"])


@pytest.mark.gpu
def test_gpu_tracking_and_ba_reproduce_golden():
    from dvmslam_b200.optimizer import LocalBA
    from dvmslam_b200.tracking import Frame, ORBmatcher, PoseOptimization

    g = np.load(os.path.join(G, "track_small.npz"))
    S, orc, T, case = _track_case()
    F = Frame(len(case["cur_kps"]) + 8, T["scale"], T["inv_sigma2"])
    F.assign(case["cur_kps"], case["cur_desc"], case["bounds"])
    lk = case["last_kps"]
    n, cur_mp = ORBmatcher(0.9, True).SearchByProjectionLast(F, case["qcw_prior"], case["tcw_prior"], case["K"],
                                                             case["has_mp"], case["outlier"], case["last_Xw"],
                                                             case["last_desc"], case["obs_pos"], lk["octave"],
                                                             lk["angle"], 15.0)
    assert n == int(g["nmatches"]) and np.array_equal(cur_mp, g["cur_mp"])
    idx = np.nonzero(cur_mp >= 0)[0]
    ck = case["cur_kps"]
    q = synth.quat_from_R(case["Rcw_prior"].astype(np.float64)).astype(np.float32)
    r, q2, t2, outl, _ = PoseOptimization(F, q, case["tcw_prior"], case["K"], case["last_Xw"][cur_mp[idx]],
                                          np.stack([ck["x"][idx], ck["y"][idx]], 1), T["inv_sigma2"][ck["octave"][idx]])
    F.close()
    assert r == int(g["pose_inliers"]) and np.array_equal(outl, g["pose_outlier"])
    assert np.abs(q2 - g["pose_q"]).max() < 1e-6 and np.abs(t2 - g["pose_t"]).max() < 1e-5
    B = synth.ba_scene(6, 2, 120, seed=5)
    s = LocalBA(16)
    ba = s.LocalBundleAdjustment(B["cam_q"], B["cam_t"], B["cam_fixed"], B["pts"], B["edge_cam"], B["edge_pt"],
                                 B["edge_obs"], B["edge_w"], B["K"])
    s.close()
    assert ba["iters"] == int(g["ba_iters"]) and np.array_equal(ba["bad"], g["ba_bad"])
    assert np.abs(ba["cam_t"] - g["ba_cam_t"]).max() < 6e-5 and np.abs(ba["pts"] - g["ba_pts"]).max() < 1e-4
