"""The optimiser oracles against the REFERENCE's own code.  oracle/_ref/libref_opt.so holds Optimizer::PoseOptimization,
LocalBundleAdjustment (local mapping and the welding BA of a merge), BundleAdjustment, OptimizeSim3 and
OptimizeEssentialGraph -- function bodies cut out of O3/src/Optimizer.cc at build time -- with the reference's edge types
(O3/src/OptimizableTypes.cpp) and its vendored g2o, compiled unmodified where they lie over the mini Eigen of
oracle/g2oshim (Eigen itself is not installed here).  The flat inputs of the oracle restatements (oracle/track_oracle.cpp,
lba_oracle.cpp, sim3_oracle.cpp) are turned into the Frame / KeyFrame / MapPoint / Map objects that make the reference
assemble the same problem; what it writes back must agree with the restatement: poses and points within a few float32
ulps, the same outlier / erased-observation / inlier sets.  Skipped where neither /root/reference nor the built library
is present (the GPU box gets the .so with the snapshot)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from dvmslam_b200 import synth
from oracle import refopt

pytestmark = pytest.mark.skipif(not refopt.available(), reason="reference sources / oracle/_ref/libref_opt.so not present")

K = np.array([994.3, 993.4, 638.0, 372.6], np.float32)


@pytest.mark.parametrize("n,seed", [(400, 0), (400, 1), (300, 3), (1200, 5), (40, 7)])
def test_pose_optimization(n, seed):
    from oracle.track import pose_optimization
    from test_pose_oracle import _scene

    S = _scene(n, seed=seed)
    r0 = pose_optimization(S["q0"], S["t0"], K, S["Xw"], S["uv"], S["w"])
    r1 = refopt.pose_optimization(S["q0"], S["t0"], K, S["Xw"], S["uv"], S["w"])
    assert r0[0] == r1[0] and np.array_equal(r0[3], r1[3])
    assert np.abs(r0[1] - r1[1]).max() <= 2e-7 and np.abs(r0[2] - r1[2]).max() <= 1e-6
    # degenerate sizes: below 3 correspondences nothing is optimised, below 10 edges one round only
    for m in (0, 2, 5, 9):
        a = pose_optimization(S["q0"], S["t0"], K, S["Xw"][:m], S["uv"][:m], S["w"][:m])
        b = refopt.pose_optimization(S["q0"], S["t0"], K, S["Xw"][:m], S["uv"][:m], S["w"][:m])
        assert a[0] == b[0] and np.array_equal(a[3], b[3]), m
        assert np.abs(a[1] - b[1]).max() <= 2e-7 and np.abs(a[2] - b[2]).max() <= 1e-6, m


def _local_only(B):
    """The reference's graph assembly takes the points seen by a LOCAL keyframe; drop the ones only fixed cameras see."""
    free_obs = np.zeros(len(B["pts"]), bool)
    free_obs[B["edge_pt"][B["cam_fixed"][B["edge_cam"]] == 0]] = True
    keep = free_obs[B["edge_pt"]]
    B = dict(B)
    for k in ("edge_cam", "edge_pt", "edge_obs", "edge_w"):
        B[k] = B[k][keep]
    return B, free_obs


def _ba_args(B):
    return (B["cam_q"], B["cam_t"], B["cam_fixed"], B["pts"], B["edge_cam"], B["edge_pt"], B["edge_obs"], B["edge_w"], B["K"])


def _close(a, b, moved):
    assert np.abs(a["cam_q"] - b["cam_q"]).max() <= 3e-7
    assert np.abs(a["cam_t"] - b["cam_t"]).max() <= 2e-6
    assert np.abs(a["pts"] - b["pts"])[moved].max() <= 2e-6


@pytest.mark.parametrize("nf,nx,npts,seed", [(8, 3, 300, 0), (12, 4, 600, 1), (50, 10, 5000, 0)])
def test_bundle_adjustments(nf, nx, npts, seed):
    """LocalBundleAdjustment (one optimize(10), erase chi2 > 5.991 or negative depth), BundleAdjustment (robust and plain,
    camera 0 fixed) and the two-pass welding BA of a merge."""
    from oracle import lba

    B, moved = _local_only(synth.ba_scene(nf, nx, npts, seed=seed))
    r0, r1 = lba.local_ba(*_ba_args(B)), refopt.local_ba(*_ba_args(B))
    _close(r0, r1, moved)
    assert np.array_equal(r0["bad"], r1["bad"]) and r0["bad"].sum() > 0
    assert tuple(r1["stats"]) == (nx, nf, 0, len(B["edge_cam"]))
    seen = np.zeros(len(B["pts"]), bool)
    seen[B["edge_pt"]] = True
    only0 = (np.arange(nf + nx) == 0).astype(np.uint8)
    for robust in (True, False):
        g0 = lba.local_ba(B["cam_q"], B["cam_t"], only0, *_ba_args(B)[3:], iterations=5,
                          huber_delta=np.sqrt(5.99) if robust else np.inf)
        g1 = refopt.bundle_adjustment(B["cam_q"], B["cam_t"], *_ba_args(B)[3:], iterations=5, robust=robust)
        _close(g0, g1, seen)
    m0, m1 = lba.merge_ba(*_ba_args(B)), refopt.merge_ba(*_ba_args(B))
    _close(m0, m1, moved)
    assert np.array_equal(m0["bad"], m1["bad"]) and m0["bad"].sum() > 0


def test_local_ba_without_fixed_keyframes_and_abort():
    """num_fixedKF == 0 -> the reference returns before optimising (Optimizer.cc:1088-1091); *pbStopFlag set -> it returns
    after assembling the graph (:1306-1308)."""
    from oracle import lba

    B, _ = _local_only(synth.ba_scene(6, 0, 200, seed=2))
    r1 = refopt.local_ba(*_ba_args(B))
    r0 = lba.local_ba(*_ba_args(B))
    # (the stand-in SE3f normalises the quaternion it is given, like Sophus: one ulp)
    assert np.abs(r1["cam_q"] - B["cam_q"]).max() < 2e-7 and np.array_equal(r1["pts"], B["pts"]) and r1["bad"].sum() == 0
    assert np.abs(r0["cam_q"] - B["cam_q"]).max() < 2e-7 and np.array_equal(r0["pts"], B["pts"])
    B, _ = _local_only(synth.ba_scene(6, 2, 200, seed=2))
    r1 = refopt.local_ba(*_ba_args(B), abort=1)
    r0 = lba.local_ba(*_ba_args(B), abort=1)
    assert np.abs(r1["cam_q"] - B["cam_q"]).max() < 2e-7 and np.array_equal(r1["pts"], B["pts"]) and r1["bad"].sum() == 0
    assert np.abs(r0["cam_q"] - B["cam_q"]).max() < 2e-7 and np.array_equal(r0["pts"], B["pts"])


def _sim3_args(S):
    return (S["p1c"], S["p2c"], S["obs1"], S["obs2"], S["w1"], S["w2"], S["K"], S["K"], S["q0"], S["t0"], S["s0"])


@pytest.mark.parametrize("n,seed,scale,fix,th2", [(250, 0, 1.3, False, 10.0), (60, 1, 0.8, False, 10.0), (200, 2, 1.0, True, 10.0),
                                                   (1500, 3, 1.1, False, 10.0), (30, 4, 1.0, True, 7.0), (400, 5, 2.0, False, 10.0)])
def test_optimize_sim3(n, seed, scale, fix, th2):
    from oracle.sim3 import optimize_sim3

    S = synth.sim3_scene(n, seed=seed, scale=scale, perturb=(np.deg2rad(1.0), 0.05, 0.0 if fix else 0.03))
    r0 = optimize_sim3(*_sim3_args(S), th2=th2, fix_scale=fix)
    r1 = refopt.optimize_sim3(*_sim3_args(S), th2=th2, fix_scale=fix)
    assert r0["n_in"] == r1["n_in"] and np.array_equal(r0["inlier"], r1["inlier"])
    assert np.abs(r0["q"] - r1["q"]).max() < 1e-8 and np.abs(r0["t"] - r1["t"]).max() < 1e-7 and abs(r0["s"] - r1["s"]) < 1e-8


def _essential_case(n, seed, fix_scale):
    """A drifted circular trajectory closed between its last keyframe and keyframe 0, as LoopClosing::CorrectLoop hands it
    over: the current keyframe and its two covisible neighbours carry corrected Sim3 (and their non-corrected ones), the
    spanning tree is the chain, a few covisibility links above and below the 100-match bar, one older loop edge."""
    from oracle.sim3 import sim3_exp
    from sim3_cases import sim3_inv, sim3_mul

    rng = np.random.default_rng(seed)
    true = [sim3_exp([0, 2 * np.pi * k / n, 0, 3 * np.cos(2 * np.pi * k / n), 0.1 * np.sin(6 * np.pi * k / n), 3 * np.sin(2 * np.pi * k / n), 0])
            for k in range(n)]
    est = [true[0].copy()]
    D = sim3_exp(np.zeros(7))
    for k in range(1, n):
        D = sim3_mul(sim3_exp(np.concatenate([rng.normal(0, 0.004, 3), rng.normal(0, 0.01, 3), [0.0]])), D)
        est.append(sim3_mul(D, true[k]))
    est = np.array(est)
    kf_q, kf_t = est[:, :4].astype(np.float32), est[:, 4:7].astype(np.float32)
    parent = np.arange(-1, n - 1, dtype=np.int32)
    cur, loop = n - 1, 0
    # the loop correction: Sim3 that takes the drifted current keyframe onto its true place, with a scale
    s = 1.0 if fix_scale else 1.04
    def as_sim3(i):
        qq = kf_q[i].astype(np.float64)
        qq = qq / np.sqrt((qq[0] * qq[0] + qq[2] * qq[2]) + (qq[1] * qq[1] + qq[3] * qq[3]))
        return np.concatenate([qq, kf_t[i].astype(np.float64), [1.0]])
    corr_cur = np.concatenate([true[cur][:4], s * true[cur][4:7], [s]])
    non_corrected, corrected = {}, {}
    for i in (cur, cur - 1, cur - 2):
        non_corrected[i] = as_sim3(i)
        rel = sim3_mul(as_sim3(i), sim3_inv(as_sim3(cur)))            # Tic
        corrected[i] = sim3_mul(rel, corr_cur)
    cov = [(k, k - 1, 300) for k in range(1, n)] + [(k, k - 2, int(w)) for k, w in zip(range(2, n), rng.integers(60, 160, n - 2))]
    cov += [(cur, loop, 40), (cur - 1, loop, 120)]
    loop_edges = [(n // 2, 1)]
    loop_connections = [(cur, loop), (cur - 1, loop), (cur - 2, loop + 1)]
    npts = 200
    pt_ref = rng.integers(0, n, npts).astype(np.int32)
    pts = rng.uniform(-4, 4, (npts, 3)).astype(np.float32)
    by_cur = (rng.random(npts) < 0.2).astype(np.uint8)
    corr_ref = rng.integers(0, n, npts).astype(np.int32)
    return dict(kf_q=kf_q, kf_t=kf_t, parent=parent, init=0, loop=loop, cur=cur, loop_edges=loop_edges, cov=cov,
                non_corrected=non_corrected, corrected=corrected, loop_connections=loop_connections, pts=pts, pt_ref=pt_ref,
                by_cur=by_cur, corr_ref=corr_ref, as_sim3=as_sim3)


def _flatten_essential(E):
    """The pose graph Optimizer::OptimizeEssentialGraph assembles (O3/src/Optimizer.cc:1423-1592), for the oracle's solver."""
    from sim3_cases import sim3_inv, sim3_mul

    n = len(E["kf_q"])
    vScw = np.array([E["corrected"].get(i, E["as_sim3"](i)) for i in range(n)])
    weight = {}
    for i, j, w in E["cov"]:
        weight[(i, j)] = weight[(j, i)] = w
    children = {i: set() for i in range(n)}
    for i, p in enumerate(E["parent"]):
        if p >= 0:
            children[int(p)].add(i)
    loop_of = {i: set() for i in range(n)}
    for i, j in E["loop_edges"]:
        loop_of[i].add(j); loop_of[j].add(i)
    vi, vj, meas, inserted = [], [], [], set()
    for i, j in E["loop_connections"]:
        if (i != E["cur"] or j != E["loop"]) and weight.get((i, j), 0) < 100:
            continue
        vi.append(i); vj.append(j); meas.append(sim3_mul(vScw[j], sim3_inv(vScw[i])))
        inserted.add((min(i, j), max(i, j)))
    nc = E["non_corrected"]
    for i in range(n):
        Swi = sim3_inv(nc.get(i, vScw[i]))
        p = int(E["parent"][i])
        if p >= 0:
            vi.append(i); vj.append(p); meas.append(sim3_mul(nc.get(p, vScw[p]), Swi))
        for l in sorted(loop_of[i]):
            if l < i:
                vi.append(i); vj.append(l); meas.append(sim3_mul(nc.get(l, vScw[l]), Swi))
        for k in sorted((k for k in range(n) if weight.get((i, k), 0) >= 100), key=lambda k: (-weight[(i, k)], k)):
            if k != p and k not in children[i] and k < i and (min(i, k), max(i, k)) not in inserted:
                vi.append(i); vj.append(k); meas.append(sim3_mul(nc.get(k, vScw[k]), Swi))
    fixed = (np.arange(n) == E["init"]).astype(np.uint8)
    return vScw, fixed, np.array(vi, np.int32), np.array(vj, np.int32), np.array(meas)


@pytest.mark.parametrize("n,seed,fix_scale", [(24, 0, False), (40, 1, False), (30, 2, True)])
def test_optimize_essential_graph(n, seed, fix_scale):
    from oracle.sim3 import optimize_essential_graph
    from sim3_cases import sim3_inv

    E = _essential_case(n, seed, fix_scale)
    vScw, fixed, vi, vj, meas = _flatten_essential(E)
    r0 = optimize_essential_graph(vScw, fixed, vi, vj, meas, fix_scale=fix_scale)
    r1 = refopt.optimize_essential_graph(E["kf_q"], E["kf_t"], E["parent"], E["init"], E["loop"], E["cur"], E["loop_edges"], E["cov"],
                                         E["non_corrected"], E["corrected"], E["loop_connections"], fix_scale, E["pts"], E["pt_ref"],
                                         E["by_cur"], E["corr_ref"])
    S = r0["sim3"]
    assert r0["iters"] >= 1
    # pKFi->SetPose(SE3f(rotation, translation / s)), Optimizer.cc:1608-1609
    assert np.abs(S[:, :4] - r1["kf_q"]).max() < 2e-6
    assert np.abs(S[:, 4:7] / S[:, 7:8] - r1["kf_t"]).max() < 2e-5
    # map points: correctedSwr * (Srw * p) with r the reference keyframe (or the one recorded by the loop correction)
    def smap(s8, p):
        q = s8[:4]; uv = 2 * np.cross(q[:3], p)
        return s8[7] * (p + q[3] * uv + np.cross(q[:3], uv)) + s8[4:7]
    ref = np.where(E["by_cur"] != 0, E["corr_ref"], E["pt_ref"])
    want = np.array([smap(sim3_inv(S[r]), smap(vScw[r], p.astype(np.float64))) for r, p in zip(ref, E["pts"])])
    assert np.abs(want - r1["pts"]).max() < 5e-5
