"""Self-checks of the OptimizeSim3 oracle (no upstream fixture exists for it): recovery of a planted
similarity, the fixed-scale variant, outlier removal, the "fewer than 10 survivors" exit."""
import numpy as np

from dvmslam_b200 import synth
from oracle.sim3 import optimize_sim3
from tests.sim3_cases import loop_graph as _loop_graph, sim3_inv as _inv, sim3_mul as _mul


def _args(S):
    return (S["p1c"], S["p2c"], S["obs1"], S["obs2"], S["w1"], S["w2"], S["K"], S["K"], S["q0"], S["t0"], S["s0"])


def _qdiff(a, b):
    return min(np.abs(a - b).max(), np.abs(a + b).max())


def test_recovers_planted_sim3_and_drops_outliers():
    S = synth.sim3_scene(250, seed=0, scale=1.3)
    r = optimize_sim3(*_args(S), th2=10.0, fix_scale=False)
    assert r["iters1"] == 5 and r["iters2"] >= 1 and r["n_bad"] > 0
    assert abs(r["s"] - S["s_true"]) < 5e-3 and np.abs(r["t"] - S["t_true"]).max() < 1e-2 and _qdiff(r["q"], S["q_true"]) < 1e-3
    assert r["n_in"] == int(r["inlier"].sum()) and r["n_in"] + r["n_bad"] <= 250
    # points without a keypoint in KF2 carry a NORMALISED measurement against a pixel projection (Optimizer.cc:2118-2125):
    # the reference always loses them in the first pass, and so does the restatement
    normalised = np.abs(S["obs2"]).max(axis=1) < 3.0
    assert normalised.any() and not r["inlier"][normalised].any()
    assert r["chi_last"] < 1e-2 * r["chi_first"]


def test_fixed_scale_keeps_scale_exactly():
    S = synth.sim3_scene(200, seed=1, scale=1.0, perturb=(np.deg2rad(1.0), 0.05, 0.0))
    r = optimize_sim3(*_args(S), th2=10.0, fix_scale=True)
    assert r["s"] == S["s0"] == 1.0
    assert np.abs(r["t"] - S["t_true"]).max() < 1e-2 and _qdiff(r["q"], S["q_true"]) < 1e-3


def test_too_few_survivors_returns_zero_and_leaves_the_estimate():
    S = synth.sim3_scene(12, seed=2, outlier_frac=0.6, not_in_kf2_frac=0.3)
    r = optimize_sim3(*_args(S), th2=10.0)
    assert 12 - r["n_bad"] < 10 and r["n_in"] == 0 and r["iters2"] == 0
    assert np.array_equal(r["q"], S["q0"]) and np.array_equal(r["t"], S["t0"]) and r["s"] == S["s0"]
    assert int((r["inlier"] == 0).sum()) == r["n_bad"]


# ---- essential graph (Optimizer::OptimizeEssentialGraph's solve): oracle only so far, see DESIGN.md 8f ----
def test_sim3_exp_log_round_trip():
    from oracle.sim3 import sim3_exp, sim3_log
    rng = np.random.default_rng(0)
    for _ in range(50):
        u = np.concatenate([rng.normal(0, 0.6, 3), rng.normal(0, 2.0, 3), [rng.normal(0, 0.3)]])
        assert np.abs(sim3_log(sim3_exp(u)) - u).max() < 1e-9
    for u in ([0, 0, 0, 1, 2, 3, 0], [1e-7, 0, 0, 1, 2, 3, 0.2], [0.3, 0.1, -0.2, 1, 2, 3, 1e-7]):    # the small-angle branches
        assert np.abs(sim3_log(sim3_exp(u)) - np.array(u, float)).max() < 1e-9


def test_essential_graph_distributes_the_loop_error():
    from oracle.sim3 import optimize_essential_graph, sim3_log
    est, fixed, vi, vj, meas, true = _loop_graph()
    r = optimize_essential_graph(est, fixed, vi, vj, meas)
    # (the first Gauss-Newton step does the work; later iterations usually die in g2o's Sim3::log small-rotation branch, see
    # the header of oracle/sim3_oracle.cpp -- with a larger lambda the optimisation goes on and gets lower)
    assert r["iters"] >= 1 and r["chi_last"] < 0.1 * r["chi_first"]
    r_auto = optimize_essential_graph(est, fixed, vi, vj, meas, lambda_init=0)
    assert r_auto["iters"] > r["iters"] and r_auto["chi_last"] < r["chi_last"]
    assert np.array_equal(r["sim3"][0], est[0])                                  # the fixed (initial) keyframe
    # before: all of the error sits on the loop edge; after: no edge carries more than a fraction of it
    def edge_norms(S):
        return np.array([np.linalg.norm(sim3_log(_mul(_mul(m, S[a]), _inv(S[b])))) for a, b, m in zip(vi, vj, meas)])
    e0, e1 = edge_norms(est), edge_norms(r["sim3"])
    assert e0[:-1].max() < 1e-9 and e0[-1] > 1e-2 and e1.max() < 0.4 * e0[-1]
    # the last keyframe moved towards its true pose
    d0 = np.linalg.norm(sim3_log(_mul(est[-1], _inv(true[-1]))))
    d1 = np.linalg.norm(sim3_log(_mul(r["sim3"][-1], _inv(true[-1]))))
    assert d1 < 0.5 * d0
    # consistent measurements: nothing to do
    r2 = optimize_essential_graph(est, fixed, vi[:-1], vj[:-1], meas[:-1])
    assert r2["chi_first"] < 1e-18 and np.abs(r2["sim3"] - est).max() < 1e-9


def test_essential_graph_fixed_scale_and_edge_order():
    from oracle.sim3 import optimize_essential_graph
    est, fixed, vi, vj, meas, _ = _loop_graph(16, seed=1, drift=(0.004, 0.01, 0.0))
    r = optimize_essential_graph(est, fixed, vi, vj, meas, fix_scale=True)
    assert np.array_equal(r["sim3"][:, 7], est[:, 7]) and r["chi_last"] < r["chi_first"]
    p = np.random.default_rng(0).permutation(len(vi))
    rp = optimize_essential_graph(est, fixed, vi[p], vj[p], meas[p], fix_scale=True)
    assert rp["iters"] == r["iters"] and np.abs(rp["sim3"] - r["sim3"]).max() < 1e-6
