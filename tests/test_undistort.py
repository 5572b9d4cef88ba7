"""Frame::UndistortKeyPoints / ComputeImageBounds (cv::undistortPoints): the oracle model against the cv2 golden vectors
(CPU) and the CUDA path against both (GPU) -- bit-exact floats."""
import ctypes as C
import os

import numpy as np
import pytest

from dvmslam_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle(pts, K, dist):
    from oracle import lib

    L = lib()
    L.cvm_undistort_points.restype = None
    L.cvm_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    p = np.ascontiguousarray(pts, np.float32)
    out = np.zeros_like(p)
    K = np.ascontiguousarray(K, np.float32)
    d = np.ascontiguousarray(dist, np.float32)
    L.cvm_undistort_points(p.ctypes.data, len(p), K.ctypes.data, d.ctypes.data, K.ctypes.data, out.ctypes.data)
    return out


def test_oracle_model_reproduces_cv2_golden():
    g = np.load(os.path.join(G, "undistort.npz"))
    for d, want in zip(g["dist"], g["out"]):
        assert np.array_equal(_oracle(g["pts"], g["K"], d), want)


def test_oracle_model_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    K = np.array([458.654, 457.296, 367.215, 248.375], np.float32)          # EuRoC cam0
    d = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0], np.float32)
    pts = rng.uniform([0, 0], [752, 480], (4000, 2)).astype(np.float32)
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    want = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, d, None, Km).reshape(-1, 2)
    assert np.array_equal(_oracle(pts, K, d), want)


@pytest.mark.gpu
def test_undistort_gpu_matches_cv2_golden():
    from dvmslam_b200.extractor import KP_DTYPE
    from dvmslam_b200.tracking import Frame

    g = np.load(os.path.join(G, "undistort.npz"))
    F = Frame(4096, np.ones(8, np.float32), np.ones(8, np.float32))
    kps = np.zeros(len(g["pts"]), KP_DTYPE)
    kps["x"], kps["y"] = g["pts"][:, 0], g["pts"][:, 1]
    kps["octave"], kps["angle"] = 3, 45.0
    for d, want in zip(g["dist"], g["out"]):
        un = F.UndistortKeyPoints(kps, g["K"], d)
        assert np.array_equal(np.stack([un["x"], un["y"]], 1), want)
        assert np.array_equal(un["octave"], kps["octave"]) and np.array_equal(un["angle"], kps["angle"])
        b = F.ComputeImageBounds(g["K"], d, 1280, 720)
        c = want[:4]
        assert b == (min(c[0, 0], c[2, 0]), min(c[0, 1], c[1, 1]), max(c[1, 0], c[3, 0]), max(c[2, 1], c[3, 1]))
    same = F.UndistortKeyPoints(kps, g["K"], np.zeros(5, np.float32))       # k1 == 0: mvKeysUn = mvKeys
    assert np.array_equal(same, kps) and F.ComputeImageBounds(g["K"], np.zeros(5), 1280, 720) == (0.0, 0.0, 1280.0, 720.0)
    F.close()


@pytest.mark.gpu
def test_tracker_with_distortion_matches_oracle_chain():
    """The device-resident tracker undistorts every frame between ExtractORB and AssignFeaturesToGrid; the oracle chain
    gets the same keypoints through the cv2-pinned model."""
    from dvmslam_b200.extractor import ORBextractor
    from dvmslam_b200.tracking import Frame, Tracker
    from oracle.orb import OrbOracle
    from oracle.track import TrackerOracle

    S = synth.OrbitStream(seed=0, period=320)
    dist = np.array([-0.05, 0.01, 0.0003, -0.0002, 0.0], np.float32)
    K = np.array(S.K, np.float32)
    ext = ORBextractor(2000, max_width=1280, max_height=720)
    T = ext.tables()
    ctx = Frame(64, T["scale"], T["inv_sigma2"])
    bounds = ctx.ComputeImageBounds(K, dist, 1280, 720)
    orc = OrbOracle(2000)

    def extract_un(img):
        k, d, m = orc.extract(img)
        un = _oracle(np.stack([k["x"], k["y"]], 1), K, dist)
        k = k.copy()
        k["x"], k["y"] = un[:, 0], un[:, 1]
        return k, d, m

    M = synth.plane_map(S, extract_un, list(range(0, 320, 40)), T["scale"], 6000)
    trk = Tracker(ext, K, bounds, M, dist_coef=dist)
    ref = TrackerOracle(extract_un, T, K, bounds, M)
    R, t = S.pose(0)
    q0 = synth.quat_from_R(R).astype(np.float32)
    n1 = trk.bootstrap(S.frame(0), q0, t)
    n0 = ref.bootstrap(S.frame(0), q0, t)
    assert n0 == n1 and n1 > 500
    pq, pt = q0, np.asarray(t, np.float32)
    for k in range(1, 4):
        r = ref.track(S.frame(k), pq, pt)
        q, tt, c = trk.track(S.frame(k), pq, pt)
        assert tuple(int(x) for x in c) == tuple(int(x) for x in r["counts"]), (k, c, r["counts"])
        assert np.abs(tt - r["t"]).max() < 1e-5 and np.abs(q - r["q"]).max() < 1e-6
        pq, pt = r["q"], r["t"]
    trk.close(); ext.close(); ctx.close()
