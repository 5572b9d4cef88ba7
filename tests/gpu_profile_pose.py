"""Profiling driver (not a test): a few PoseOptimization calls on synthetic correspondences; meant to be wrapped by ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvmslam_b200 import synth
from dvmslam_b200.tracking import Frame, PoseOptimization

rng = np.random.default_rng(0)
S = synth.OrbitStream(seed=0, period=320)
K = np.array(S.K, np.float32)
Rp, tp = S.pose(4)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
uv = rng.uniform([50, 50], [1230, 670], (n, 2))
X = synth.backproject_to_plane(S, 5, uv).astype(np.float32)
oct_ = rng.integers(0, 8, n)
xy = (uv + rng.normal(0, 1, (n, 2)) * (1.2 ** oct_)[:, None]).astype(np.float32)
out = rng.random(n) < 0.05
xy[out] += rng.normal(0, 40, (int(out.sum()), 2)).astype(np.float32)
w = (1 / (1.2 ** oct_) ** 2).astype(np.float32)
q = synth.quat_from_R(np.asarray(Rp, np.float64)).astype(np.float32)
sc = (1.2 ** np.arange(8)).astype(np.float32)
F = Frame(2048, sc, (1 / sc ** 2).astype(np.float32))
for _ in range(4):
    r = PoseOptimization(F, q, np.asarray(tp, np.float32), K, X, xy, w)
print("inliers", r[0], "stats", r[4])
