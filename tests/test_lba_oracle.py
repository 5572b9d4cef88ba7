"""Self-checks of the local-BA oracle (it has no upstream fixtures to pin against):
finite-difference Jacobians, zero-noise convergence, monotone chi2, no-op conditions."""
import numpy as np

from dvmslam_b200 import synth
from oracle.lba import edge_error_perturbed, edge_jacobians, local_ba


def _args(S):
    return (S["cam_q"], S["cam_t"], S["cam_fixed"], S["pts"], S["edge_cam"], S["edge_pt"], S["edge_obs"], S["edge_w"], S["K"])


def test_jacobians_match_finite_differences():
    S = synth.ba_scene(8, 2, 200, seed=1)
    rng = np.random.default_rng(0)
    for e in rng.choice(len(S["edge_cam"]), 20, replace=False):
        q, t, X = S["cam_q"][S["edge_cam"][e]], S["cam_t"][S["edge_cam"][e]], S["pts"][S["edge_pt"][e]]
        A, B = edge_jacobians(q, t, X, S["K"])
        obs = S["edge_obs"][e]
        h = 1e-6
        Bn, An = np.zeros((2, 6)), np.zeros((2, 3))
        for k in range(6):
            d = np.zeros(6); d[k] = h
            Bn[:, k] = (edge_error_perturbed(q, t, X, S["K"], obs, d, np.zeros(3)) -
                        edge_error_perturbed(q, t, X, S["K"], obs, -d, np.zeros(3))) / (2 * h)
        for k in range(3):
            d = np.zeros(3); d[k] = h
            An[:, k] = (edge_error_perturbed(q, t, X, S["K"], obs, np.zeros(6), d) -
                        edge_error_perturbed(q, t, X, S["K"], obs, np.zeros(6), -d)) / (2 * h)
        assert np.abs(B - Bn).max() < 1e-5 * max(1.0, np.abs(B).max())
        assert np.abs(A - An).max() < 1e-5 * max(1.0, np.abs(A).max())


def test_zero_noise_scene_converges():
    S = synth.ba_scene(10, 3, 400, seed=2, pix_sigma=0.0, outlier_frac=0.0)
    r = local_ba(*_args(S))
    assert r["chi_last"] < 1e-6 * r["chi_first"]
    assert np.abs(r["cam_t"] - S["t_true"]).max() < 1e-4
    assert np.median(np.abs(r["pts"] - S["pts_true"]).max(1)) < 1e-4  # (points seen once keep their depth ambiguity)
    assert r["bad"].sum() == 0


def test_noisy_scene_reduces_chi2_and_flags_outliers():
    S = synth.ba_scene(12, 4, 600, seed=3)
    r = local_ba(*_args(S))
    assert r["chi_last"] < r["chi_first"] and r["iters"] >= 2 and r["trials"] >= r["iters"]
    assert 0.02 < r["bad"].mean() < 0.25
    assert np.abs(r["cam_t"] - S["t_true"]).max() < np.abs(S["cam_t"] - S["t_true"]).max()
    fixed = S["cam_fixed"].astype(bool)
    assert np.array_equal(r["cam_q"][fixed], S["cam_q"][fixed]) and np.array_equal(r["cam_t"][fixed], S["cam_t"][fixed])


def test_noop_conditions():
    S = synth.ba_scene(6, 2, 100, seed=4)
    a = list(_args(S))
    a[2] = np.zeros_like(S["cam_fixed"])          # no fixed keyframe -> "LBA aborted"
    r = local_ba(*a)
    assert r["rc"] == -1 and np.array_equal(r["pts"], S["pts"])
    r = local_ba(*_args(S), abort=1)              # pbStopFlag already set
    assert r["rc"] == -1 and np.array_equal(r["cam_t"], S["cam_t"])


def test_bundle_adjustment_delta_variants():
    """The robust-kernel delta is the caller's: sqrt(5.991) (local BA), sqrt(5.99) (global BA), infinity (bRobust =
    false).  Without outliers and without a kernel the problem is plain least squares and converges to the noise floor."""
    from oracle.lba import local_ba

    S = synth.ba_scene(10, 1, 300, seed=4, outlier_frac=0.0)
    a = (S["cam_q"], S["cam_t"], S["cam_fixed"], S["pts"], S["edge_cam"], S["edge_pt"], S["edge_obs"], S["edge_w"], S["K"])
    r_inf = local_ba(*a, iterations=20, huber_delta=float("inf"))
    r_gba = local_ba(*a, iterations=20, huber_delta=np.sqrt(5.99))
    r_lba = local_ba(*a, iterations=20)
    ne = len(S["edge_cam"])
    assert r_inf["chi_last"] < 3.0 * ne and r_inf["chi_last"] < 0.2 * r_inf["chi_first"]
    assert r_gba["chi_last"] <= r_inf["chi_last"] * (1 + 1e-9)             # Huber never exceeds the quadratic cost
    assert abs(r_gba["chi_last"] - r_lba["chi_last"]) < 1e-2 * r_lba["chi_last"]   # 5.99 vs 5.991: nearly the same kernel
    assert r_gba["chi_last"] != r_lba["chi_last"]


def test_merge_ba_two_passes():
    """The welding BA (Optimizer.cc:3257-3675): pass 1 is BundleAdjustment's optimize(5) with delta sqrt(5.99); the edges
    it would flag go to level 1 and keep the error of pass 1; pass 2 runs without a robust kernel on the rest."""
    from oracle.lba import local_ba, merge_ba

    S = synth.ba_scene(12, 3, 600, seed=1)
    a = _args(S)
    r1 = local_ba(*a, iterations=5, huber_delta=np.sqrt(5.99))
    r = merge_ba(*a)
    assert r["iters_first"] == r1["iters"] and r["excluded"] == int(r1["bad"].sum()) and r["excluded"] > 0
    assert r["iters"] > r["iters_first"] and r["chi_first"] == r1["chi_first"]
    lvl1 = r1["bad"].astype(bool)
    assert np.array_equal(r["chi2"][lvl1], r1["chi2"][lvl1])            # level-1 edges: stale error
    assert not np.array_equal(r["chi2"][~lvl1], r1["chi2"][~lvl1])      # level-0 edges moved on
    # the second pass minimises the plain sum of squares of the level-0 edges: it cannot end above where it started
    assert r["chi_last"] <= r1["chi2"][~lvl1].sum() * (1 + 1e-12)
    assert abs(r["chi_last"] - r["chi2"][~lvl1].sum()) < 1e-9 * r["chi_last"]
    # stop flag up: nothing runs
    r = merge_ba(*a, abort=1)
    assert r["rc"] == -1 and np.array_equal(r["pts"], S["pts"])
    # no outliers, no noise: nothing goes to level 1 and the optimum is the true scene
    S = synth.ba_scene(8, 2, 300, seed=3, pix_sigma=0.0, outlier_frac=0.0)
    r = merge_ba(*_args(S))
    assert r["excluded"] == 0 and r["chi_last"] < 1e-6 * r["chi_first"] and not r["bad"].any()


def test_merge_ba_second_pass_reaches_the_least_squares_optimum():
    """Independent check of the two-pass flow: pass 2 is plain least squares over the level-0 edges, so scipy started from
    the oracle's result (own residual code: rotation-vector increments on the free poses, additive points) must not find a
    noticeably lower cost."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation
    from oracle.lba import local_ba, merge_ba

    S = synth.ba_scene(5, 2, 60, seed=13)
    a = _args(S)
    r = merge_ba(*a)
    lvl0 = ~local_ba(*a, iterations=5, huber_delta=np.sqrt(5.99))["bad"].astype(bool)
    assert r["excluded"] == int((~lvl0).sum())
    ec, ep = S["edge_cam"][lvl0], S["edge_pt"][lvl0]
    obs, w = S["edge_obs"][lvl0].astype(np.float64), np.sqrt(S["edge_w"][lvl0].astype(np.float64))
    fx, fy, cx, cy = [float(v) for v in S["K"]]
    free = np.nonzero(S["cam_fixed"] == 0)[0]
    R0 = Rotation.from_quat(r["cam_q"].astype(np.float64)).as_matrix()
    t0 = r["cam_t"].astype(np.float64)
    P0 = r["pts"].astype(np.float64)

    def residuals(x):
        R, t = R0.copy(), t0.copy()
        for k, c in enumerate(free):
            dR = Rotation.from_rotvec(x[6 * k:6 * k + 3]).as_matrix()
            R[c] = dR @ R0[c]
            t[c] = dR @ t0[c] + x[6 * k + 3:6 * k + 6]
        P = P0 + x[6 * len(free):].reshape(-1, 3)
        Xc = np.einsum("eij,ej->ei", R[ec], P[ep]) + t[ec]
        uv = np.stack([fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy], 1)
        return ((obs - uv) * w[:, None]).ravel()

    x0 = np.zeros(6 * len(free) + P0.size)
    c0 = float(np.sum(residuals(x0) ** 2))
    # the outputs are float32: the cost at the rounded result is the oracle's final chi2 up to that rounding
    assert abs(c0 - r["chi_last"]) < 1e-3 * r["chi_last"]
    sol = least_squares(residuals, x0, method="trf", xtol=1e-12, ftol=1e-12, gtol=1e-12, max_nfev=50)
    c1 = float(np.sum(sol.fun ** 2))
    assert c1 <= c0 and c0 - c1 < 2e-3 * c0
