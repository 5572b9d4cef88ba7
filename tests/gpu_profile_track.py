"""Profiling driver (not a test): tracks a few frames and runs a few local BAs; meant to be wrapped by ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvmslam_b200 import synth
from dvmslam_b200.extractor import ORBextractor
from dvmslam_b200.tracking import Tracker
from dvmslam_b200.optimizer import LocalBA

nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 6
nlba = int(sys.argv[2]) if len(sys.argv) > 2 else 1
S = synth.OrbitStream(seed=0, period=320)
ext = ORBextractor(2000, max_width=1280, max_height=720)
T = ext.tables()
M = synth.plane_map(S, lambda im: ext(im), list(range(0, 320, 40)), T["scale"], 6000)
trk = Tracker(ext, S.K, (0, 0, 1280, 720), M)
R, t = S.pose(0)
trk.bootstrap(S.frame(0), synth.quat_from_R(R).astype(np.float32), t)
for k in range(1, nframes + 1):
    q, tt, c = trk.track(S.frame(k))
print("counts", c)
if nlba:
    B = synth.ba_scene(50, 10, 5000, seed=0)
    s = LocalBA(64)
    for _ in range(nlba):
        r = s.LocalBundleAdjustment(B["cam_q"], B["cam_t"], B["cam_fixed"], B["pts"], B["edge_cam"], B["edge_pt"], B["edge_obs"], B["edge_w"], B["K"])
    print("lba iters", r["iters"], "trials", r["trials"], "kernel ms", r["kernel_ms"])
