"""Profiling driver (not a test): one welding BA and a few OptimizeSim3 calls; meant to be wrapped by ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvmslam_b200 import synth
from dvmslam_b200.optimizer import LocalBA, Sim3Optimizer

Sm = synth.ba_scene(30, 10, 3000, seed=30)
sol = LocalBA(64)
r = sol.MergeBundleAdjustment(Sm["cam_q"], Sm["cam_t"], Sm["cam_fixed"], Sm["pts"], Sm["edge_cam"], Sm["edge_pt"], Sm["edge_obs"],
                              Sm["edge_w"], Sm["K"])
print("welding BA iters", r["iters"], "first pass", r["iters_first"], "level-1 edges", r["excluded"], "kernel ms", r["kernel_ms"])
Ss = synth.sim3_scene(300, seed=0, scale=1.3)
s3 = Sim3Optimizer()
for _ in range(3):
    r = s3.OptimizeSim3(Ss["p1c"], Ss["p2c"], Ss["obs1"], Ss["obs2"], Ss["w1"], Ss["w2"], Ss["K"], Ss["K"], Ss["q0"], Ss["t0"], Ss["s0"])
print("OptimizeSim3 n_in", r["n_in"], "iters", r["iters1"], r["iters2"], "trials", r["trials"])
