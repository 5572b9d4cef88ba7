"""Probe (not a test): rounds and wall time of the matcher host calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvmslam_b200 import synth
from dvmslam_b200.tracking import Frame, ORBmatcher
from oracle.orb import OrbOracle

S = synth.PlaneStream(seed=0)
orc = OrbOracle(2000)
T = orc.tables()
case = synth.tracking_case(S, 3, orc.extract)
F1 = Frame(len(case["cur_kps"]) + 16, T["scale"], T["inv_sigma2"])
F1.assign(case["cur_kps"], case["cur_desc"], case["bounds"])
lk = case["last_kps"]
mt = ORBmatcher(0.9, True)
for th in (15.0, 30.0):
    args = (case["qcw_prior"], case["tcw_prior"], case["K"], case["has_mp"], case["outlier"], case["last_Xw"],
            case["last_desc"], case["obs_pos"], lk["octave"], lk["angle"], th)
    mt.SearchByProjectionLast(F1, *args)
    t = time.perf_counter()
    for _ in range(20):
        n, m = mt.SearchByProjectionLast(F1, *args)
    print(f"search_last th={th}: matches {n}, rounds {mt.rounds(F1)}, host call {(time.perf_counter()-t)/20*1e6:.0f} us")
