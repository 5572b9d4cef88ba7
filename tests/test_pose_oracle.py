"""Self-checks of the pose-only optimisation oracle (Optimizer::PoseOptimization restated in oracle/track_oracle.cpp; no
upstream fixture exists): recovery of a planted pose, outlier rejection, and -- independently of the
restated g2o machinery -- the least-squares optimum over the final inliers computed by scipy."""
import numpy as np

from oracle.track import pose_optimization

K = np.array([994.3, 993.4, 638.0, 372.6], np.float32)


def _scene(n=400, seed=0, sigma=1.0, outlier_frac=0.1):
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(seed)
    R = Rotation.from_rotvec(rng.normal(0, 0.2, 3))
    t = rng.normal(0, 0.3, 3)
    Xc = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(2, 12, n)], 1)
    Xw = (Xc - t) @ R.as_matrix()                       # Xc = R Xw + t
    octave = rng.integers(0, 8, n)
    uv = np.stack([K[0] * Xc[:, 0] / Xc[:, 2] + K[2], K[1] * Xc[:, 1] / Xc[:, 2] + K[3]], 1)
    uv += rng.normal(0, 1, (n, 2)) * (sigma * 1.2 ** octave)[:, None]
    bad = rng.random(n) < outlier_frac
    uv[bad] += rng.uniform(-80, 80, (int(bad.sum()), 2))
    w = (np.float32(1.0) / (np.float32(1.2) ** octave) ** 2).astype(np.float32)
    q0 = (Rotation.from_rotvec(rng.normal(0, 0.01, 3)) * R).as_quat().astype(np.float32)
    t0 = (t + rng.normal(0, 0.03, 3)).astype(np.float32)
    return dict(Xw=Xw.astype(np.float32), uv=uv.astype(np.float32), w=w, q0=q0, t0=t0, R=R, t=t, bad=bad)


def test_recovers_pose_and_rejects_outliers():
    from scipy.spatial.transform import Rotation

    S = _scene()
    n_in, q, t, outl, (its, trials) = pose_optimization(S["q0"], S["t0"], K, S["Xw"], S["uv"], S["w"])
    assert n_in == int((outl == 0).sum()) and its >= 4
    assert np.abs(t - S["t"]).max() < 0.02
    assert (Rotation.from_quat(q.astype(np.float64)) * S["R"].inv()).magnitude() < 2e-3
    # the planted gross errors are flagged; of the clean observations only the chi2 tail (5.991 is the 95 % point)
    assert outl[S["bad"]].mean() > 0.9 and outl[~S["bad"]].mean() < 0.15


def test_final_pose_is_the_least_squares_optimum_over_the_inliers():
    """The last of the four rounds runs without the robust kernel over the edges classified inliers after the third
    (O3/src/Optimizer.cc:957-1012); when that classification equals the final one, the result is the weighted
    least-squares pose of the final inliers."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation

    S = _scene(300, seed=3)
    n_in, q, t, outl, _ = pose_optimization(S["q0"], S["t0"], K, S["Xw"], S["uv"], S["w"])
    keep = outl == 0
    Xw, uv, sw = S["Xw"][keep].astype(np.float64), S["uv"][keep].astype(np.float64), np.sqrt(S["w"][keep].astype(np.float64))
    R0, t0 = Rotation.from_quat(q.astype(np.float64)), t.astype(np.float64)

    def residuals(x):
        R = (Rotation.from_rotvec(x[:3]) * R0).as_matrix()
        Xc = Xw @ R.T + (Rotation.from_rotvec(x[:3]).as_matrix() @ t0 + x[3:])
        p = np.stack([float(K[0]) * Xc[:, 0] / Xc[:, 2] + float(K[2]), float(K[1]) * Xc[:, 1] / Xc[:, 2] + float(K[3])], 1)
        return ((uv - p) * sw[:, None]).ravel()

    sol = least_squares(residuals, np.zeros(6), method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14)
    c0, c1 = float(np.sum(residuals(np.zeros(6)) ** 2)), float(np.sum(sol.fun ** 2))
    assert c1 <= c0 and c0 - c1 < 1e-5 * c0            # float32 outputs: the cost at the rounded pose is within 1e-5 of the optimum
    assert np.abs(sol.x[:3]).max() < 1e-5 and np.abs(sol.x[3:]).max() < 1e-4
