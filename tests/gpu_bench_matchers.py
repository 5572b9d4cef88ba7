"""Timing driver (not a test): host-call latency of the keyframe-rate operators at config-C2 sizes (1280x720, 2000
features): SearchByBoW (both), SearchForInitialization (10000-feature extractor), SearchForTriangulation, Fuse search,
DBoW2 transform (synthetic k = 10, L = 6 tree of the ORBvoc shape).  Wall clock around the synchronous C-ABI call
(H2D + kernels + D2H), median of 20, next to the single-thread oracle on the same inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvmslam_b200 import synth
from dvmslam_b200.extractor import ORBextractor
from dvmslam_b200.matching import BowFeatures, BowMatcher, FuseSearch, SearchForTriangulation
from dvmslam_b200.tracking import Frame
from dvmslam_b200.vocabulary import Vocabulary, flatten_tree
from oracle import bow as ob, dbow as od
from oracle.track import FrameOracle
from tests import bow_cases


def med(fn, reps=20):
    fn(); fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return 1e6 * float(np.median(ts))


ext = ORBextractor(2000, max_width=1280, max_height=720)
T = ext.tables()
gx = lambda im: ext(im)
ctx = Frame(12000, T["scale"], T["inv_sigma2"])
rows = []
c = bow_cases.bow_pair(gx, w=1280, h=720)
a, b = BowFeatures(c["desc1"], c["angle1"], c["valid1"], c["fv1"]), BowFeatures(c["desc2"], c["angle2"], c["valid2"], c["fv2"])
m = BowMatcher(0.7, True)
for kf_kf, name in ((0, "SearchByBoW(KF, Frame)"), (1, "SearchByBoW(KF, KF)")):
    g = med(lambda: m._bow(ctx, kf_kf, a, b))
    o = med(lambda: ob.search_by_bow(kf_kf, c["desc1"], c["angle1"], c["valid1"], c["fv1"], c["desc2"], c["angle2"], c["valid2"], c["fv2"], 0.7, True), 5)
    rows.append((name, g, o))
ct = bow_cases.triangulation_pair(gx, w=1280, h=720)
ta, tb = BowFeatures(ct["desc1"], ct["kps1"]["angle"], ct["has_mp1"], ct["fv1"]), BowFeatures(ct["desc2"], ct["kps2"]["angle"], ct["has_mp2"], ct["fv2"])
g = med(lambda: SearchForTriangulation(ctx, ta, ct["kps1"], tb, ct["kps2"], ct["F12"], ct["ep"], T["scale"], T["sigma2"]))
o = med(lambda: ob.search_for_triangulation(ct["desc1"], ct["kps1"], ct["has_mp1"], ct["fv1"], ct["desc2"], ct["kps2"], ct["has_mp2"], ct["fv2"], ct["F12"], ct["ep"], T["scale"], T["sigma2"]), 5)
rows.append(("SearchForTriangulation", g, o))
cf = bow_cases.fuse_case(gx, w=1280, h=720, n_points=2000)
Fk = Frame(len(cf["kps"]) + 16, T["scale"], T["inv_sigma2"]); Fk.assign(cf["kps"], cf["desc"], cf["bounds"])
F0 = FrameOracle(cf["kps"], cf["desc"], cf["bounds"], T["scale"])
ls = float(np.float32(np.log(np.float64(T["scale"][1]))))
g = med(lambda: FuseSearch(Fk, cf["q"], cf["t"], cf["K"], cf["xw"], cf["normal"], cf["min_dist"], cf["max_dist"], cf["mp_desc"], cf["skip"], 3.0))
o = med(lambda: ob.fuse_search(F0, cf["q"], cf["t"], cf["K"], ls, T["inv_sigma2"], cf["xw"], cf["normal"], cf["min_dist"], cf["max_dist"], cf["mp_desc"], cf["skip"], 3.0), 5)
rows.append(("Fuse (search, 2000 map points)", g, o))
ext5 = ORBextractor(10000, max_width=1280, max_height=720)
ci = bow_cases.init_pair(lambda im: ext5(im), w=1280, h=720)
Fi = Frame(len(ci["kps2"]) + 16, T["scale"], T["inv_sigma2"]); Fi.assign(ci["kps2"], ci["desc2"], ci["bounds"])
Fo = FrameOracle(ci["kps2"], ci["desc2"], ci["bounds"], T["scale"])
mi = BowMatcher(0.9, True)
g = med(lambda: mi.SearchForInitialization(Fi, ci["kps1"], ci["desc1"], ci["prev"], 100))
o = med(lambda: ob.search_for_initialization(ci["kps1"], ci["desc1"], Fo, ci["prev"], 100, 0.9, True), 5)
rows.append((f"SearchForInitialization ({len(ci['kps1'])} features)", g, o))
v = synth.toy_vocabulary(10, 5, seed=1)
voc = Vocabulary(v["k"], v["L"], 0, 0, v["parent"], v["is_leaf"], v["desc"], v["weight"])
tree = flatten_tree(v["parent"], v["is_leaf"], v["desc"], v["weight"])
feat = c["desc1"]
g = med(lambda: voc.transform_features(feat, 4))
o = med(lambda: od.transform_features(tree, v["L"], feat, 4), 5)
rows.append((f"DBoW2 descent ({len(feat)} descriptors, k 10, L 5, {len(tree[4])} nodes)", g, o))
from dvmslam_b200.optimizer import LocalBA
from oracle.lba import local_ba
Sg = synth.ba_scene(96, 1, 3000, seed=96)
ag = (Sg["cam_q"], Sg["cam_t"], Sg["cam_fixed"], Sg["pts"], Sg["edge_cam"], Sg["edge_pt"], Sg["edge_obs"], Sg["edge_w"], Sg["K"])
sol = LocalBA(100)
g = med(lambda: sol.BundleAdjustment(*ag, nIterations=20, bRobust=True), 5)
o = med(lambda: local_ba(*ag, iterations=20, huber_delta=float(np.float32(np.sqrt(5.99)))), 2)
rows.append((f"BundleAdjustment (global: 96 + 1 keyframes, 3000 points, {len(Sg['edge_cam'])} obs, 20 its)", g, o))
from oracle.lba import merge_ba
Sm = synth.ba_scene(30, 10, 3000, seed=30)
am = (Sm["cam_q"], Sm["cam_t"], Sm["cam_fixed"], Sm["pts"], Sm["edge_cam"], Sm["edge_pt"], Sm["edge_obs"], Sm["edge_w"], Sm["K"])
g = med(lambda: sol.MergeBundleAdjustment(*am), 5)
o = med(lambda: merge_ba(*am), 2)
rows.append((f"welding BA (30 adjust + 10 fixed KFs, {len(Sm['pts'])} points, {len(Sm['edge_cam'])} obs, 5 + 10 its)", g, o))
from dvmslam_b200.optimizer import Sim3Optimizer
from oracle.sim3 import optimize_sim3
Ss = synth.sim3_scene(300, seed=0, scale=1.3)
as3 = (Ss["p1c"], Ss["p2c"], Ss["obs1"], Ss["obs2"], Ss["w1"], Ss["w2"], Ss["K"], Ss["K"], Ss["q0"], Ss["t0"], Ss["s0"])
s3 = Sim3Optimizer()
g = med(lambda: s3.OptimizeSim3(*as3, th2=10.0, bFixScale=False))
o = med(lambda: optimize_sim3(*as3, th2=10.0, fix_scale=False), 5)
rows.append(("OptimizeSim3 (300 correspondences, 5 + 10 its, numeric Jacobians)", g, o))
print(f"{'operator':62s} {'B200 call us':>12s} {'oracle 1 thread us':>18s} {'ratio':>7s}")
for name, g, o in rows:
    print(f"{name:62s} {g:12.1f} {o:18.1f} {o / g:7.1f}")
