"""Regenerates the committed golden vectors (run in the build container, where cv2 4.13.0 is importable):

    python tests/golden/make_golden.py

  cv2_primitives.npz  outputs of the REAL OpenCV primitives the reference calls (cv::resize INTER_LINEAR,
                      cv::GaussianBlur 7x7 sigma 2, cv::FAST 9-16 with NMS, cv::fastAtan2) on small seeded
                      inputs -- the oracle's C models must reproduce them bit for bit;
  undistort.npz       cv::undistortPoints(pts, K, dist, None, K) of the real cv2 for three coefficient sets;
  orb_small.npz       ORBextractor::operator() outputs (keypoints, descriptors, monoIndex) on a 320x240 frame,
                      produced by the cv2-backed pipeline (oracle/orb.py: extract_with_cv2), i.e. OpenCV
                      primitives + the restated in-tree logic;
  track_small.npz     matcher / pose-optimisation / local-BA results of the oracle on small seeded cases
                      (no upstream fixture exists for these: they pin the oracle against regressions).
  bow_small.npz       SearchByBoW (both overloads), SearchForInitialization and exhaustive nearest/second-nearest
                      results of the oracle on the seeded cases of tests/bow_cases.py (regression pins, no upstream
                      fixture).
  dbow_orbvoc.npz     the DBoW2 transform of 240 descriptors through the reference's own ORBvoc.txt (k = 10, L = 6,
                      1 082 073 nodes): the visited excerpt of the tree (every child of every node on a visited path) and the
                      oracle's words / nodes / BowVector -- the vocabulary itself does not travel to the GPU box.
  keyframe_ops.npz    welding BA, OptimizeSim3 and essential-graph results of the oracle on small seeded scenes
                      (regression pins, no upstream fixture); `python tests/golden/make_golden.py keyframe` remakes this
                      file alone (it needs neither cv2 nor /root/reference).
The reference repository holds no fixtures for this path; these are ours.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def keyframe_ops():
    from dvmslam_b200 import synth
    from oracle.lba import merge_ba
    from oracle.sim3 import optimize_essential_graph, optimize_sim3
    from tests.sim3_cases import loop_graph

    B = synth.ba_scene(8, 3, 200, seed=11)
    m = merge_ba(B["cam_q"], B["cam_t"], B["cam_fixed"], B["pts"], B["edge_cam"], B["edge_pt"], B["edge_obs"], B["edge_w"], B["K"])
    S = synth.sim3_scene(80, seed=11, scale=1.25)
    s3 = optimize_sim3(S["p1c"], S["p2c"], S["obs1"], S["obs2"], S["w1"], S["w2"], S["K"], S["K"], S["q0"], S["t0"], S["s0"], 10.0, False)
    est, fixed, vi, vj, meas, _ = loop_graph(16, seed=2)
    eg = optimize_essential_graph(est, fixed, vi, vj, meas)
    egf = optimize_essential_graph(est, fixed, vi, vj, meas, fix_scale=True)
    np.savez_compressed(os.path.join(HERE, "keyframe_ops.npz"),
                        merge_cam_q=m["cam_q"], merge_cam_t=m["cam_t"], merge_pts=m["pts"], merge_bad=m["bad"],
                        merge_counts=np.array([m["iters"], m["iters_first"], m["excluded"], m["trials"]], np.int32),
                        merge_chi=np.array([m["chi_first"], m["chi_last"]]),
                        sim3_q=s3["q"], sim3_t=s3["t"], sim3_s=np.float64(s3["s"]), sim3_inlier=s3["inlier"],
                        sim3_counts=np.array([s3["n_in"], s3["iters1"], s3["iters2"], s3["n_bad"]], np.int32),
                        eg_sim3=eg["sim3"], eg_counts=np.array([eg["iters"], eg["trials"]], np.int32),
                        eg_chi=np.array([eg["chi_first"], eg["chi_last"]]),
                        egf_sim3=egf["sim3"], egf_counts=np.array([egf["iters"], egf["trials"]], np.int32))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "keyframe":
        keyframe_ops()
        return
    keyframe_ops()
    import cv2

    from dvmslam_b200 import synth
    from oracle.lba import local_ba
    from oracle.orb import OrbOracle, extract_with_cv2
    from oracle.track import FrameOracle, pose_optimization

    cv2.setNumThreads(1)
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(2026)

    # ---- OpenCV primitives ----
    img = synth.texture(200, 150, seed=7)
    inv = np.float32(1) / np.float32(1.2)
    dw, dh = int(np.rint(np.float32(200) * inv)), int(np.rint(np.float32(150) * inv))
    resized = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    blurred = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    fast = {}
    for th in (20, 7):
        det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        fast[th] = np.array([(k.pt[0], k.pt[1], k.response) for k in det.detect(img)], np.int32)
    yx = rng.integers(-3000000, 3000000, (4000, 2))
    yx[:50, 0] = 0
    yx[50:100, 1] = 0
    yx[100:150, 0] = yx[100:150, 1]
    at = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    np.savez_compressed(os.path.join(HERE, "cv2_primitives.npz"), img=img, resized=resized, blurred=blurred,
                        fast20=fast[20], fast7=fast[7], atan_yx=yx.astype(np.int32), atan_deg=at,
                        cv2_version=np.array(cv2.__version__))

    # ---- cv::undistortPoints (Frame::UndistortKeyPoints / ComputeImageBounds) ----
    Kc = np.array([[994.3, 0, 638.0], [0, 993.4, 372.6], [0, 0, 1]], np.float32)
    und_pts = rng.uniform([0, 0], [1280, 720], (3000, 2)).astype(np.float32)
    und_pts[:4] = [[0, 0], [1280, 0], [0, 720], [1280, 720]]
    und_dist = np.array([[-0.28, 0.07, 0.0002, -0.0001, 0.0], [0.1, -0.2, 0.001, 0.002, 0.05],
                         [-0.35, 0.15, -0.001, 0.0005, -0.03]], np.float32)
    und_out = np.stack([cv2.undistortPoints(und_pts.reshape(-1, 1, 2), Kc, d, None, Kc).reshape(-1, 2) for d in und_dist])
    np.savez_compressed(os.path.join(HERE, "undistort.npz"), K=np.array([994.3, 993.4, 638.0, 372.6], np.float32), pts=und_pts,
                        dist=und_dist, out=und_out, cv2_version=np.array(cv2.__version__))

    # ---- whole extractor on a small frame ----
    frame = synth.frame(320, 240, seed=11)
    kps, desc, mono = extract_with_cv2(frame, 500)
    np.savez_compressed(os.path.join(HERE, "orb_small.npz"), frame=frame, kps=kps, desc=desc, mono=np.int32(mono),
                        nfeatures=np.int32(500))

    # ---- tracking operators and local BA on small cases ----
    S = synth.PlaneStream(640, 480, seed=3, K=(500.0, 500.0, 320.0, 240.0))
    orc = OrbOracle(800)
    T = orc.tables()
    case = synth.tracking_case(S, 4, orc.extract, n_local=1500)
    F = FrameOracle(case["cur_kps"], case["cur_desc"], case["bounds"], T["scale"])
    lk = case["last_kps"]
    n, cur_mp = F.search_by_projection_last(case["qcw_prior"], case["tcw_prior"], case["K"], case["has_mp"],
                                            case["outlier"], case["last_Xw"], case["last_desc"], case["obs_pos"],
                                            lk["octave"], lk["angle"], 15.0)
    idx = np.nonzero(cur_mp >= 0)[0]
    ck = case["cur_kps"]
    q = synth.quat_from_R(case["Rcw_prior"].astype(np.float64)).astype(np.float32)
    r, q2, t2, outl, _ = pose_optimization(q, case["tcw_prior"], case["K"], case["last_Xw"][cur_mp[idx]],
                                           np.stack([ck["x"][idx], ck["y"][idx]], 1), T["inv_sigma2"][ck["octave"][idx]])
    B = synth.ba_scene(6, 2, 120, seed=5)
    ba = local_ba(B["cam_q"], B["cam_t"], B["cam_fixed"], B["pts"], B["edge_cam"], B["edge_pt"], B["edge_obs"],
                  B["edge_w"], B["K"])
    np.savez_compressed(os.path.join(HERE, "track_small.npz"), nmatches=np.int32(n), cur_mp=cur_mp, pose_q=q2,
                        pose_t=t2, pose_inliers=np.int32(r), pose_outlier=outl, ba_cam_q=ba["cam_q"],
                        ba_cam_t=ba["cam_t"], ba_pts=ba["pts"], ba_bad=ba["bad"], ba_iters=np.int32(ba["iters"]),
                        ba_chi=np.array([ba["chi_first"], ba["chi_last"]]))
    # ---- descriptor matchers without projection ----
    from oracle.bow import hamming_knn, search_by_bow, search_for_initialization
    from tests import bow_cases

    orc = OrbOracle(600)
    T = orc.tables()
    c = bow_cases.bow_pair(orc.extract)
    out = {}
    for kf_kf in (0, 1):
        n, m12, m21 = search_by_bow(kf_kf, c["desc1"], c["angle1"], c["valid1"], c["fv1"], c["desc2"], c["angle2"],
                                    c["valid2"], c["fv2"], 0.7, True)
        out[f"bow{kf_kf}_n"], out[f"bow{kf_kf}_m12"], out[f"bow{kf_kf}_m21"] = np.int32(n), m12, m21
    ci = bow_cases.init_pair(orc.extract)
    F2 = FrameOracle(ci["kps2"], ci["desc2"], ci["bounds"], T["scale"])
    n, m12, pm = search_for_initialization(ci["kps1"], ci["desc1"], F2, ci["prev"], 100, 0.9, True)
    out["init_n"], out["init_m12"], out["init_prev"] = np.int32(n), m12, pm
    out["knn"] = np.stack(hamming_knn(c["desc1"], c["desc2"]))
    from oracle.bow import fuse_search, search_for_triangulation

    ct = bow_cases.triangulation_pair(orc.extract)
    n, m12 = search_for_triangulation(ct["desc1"], ct["kps1"], ct["has_mp1"], ct["fv1"], ct["desc2"], ct["kps2"], ct["has_mp2"],
                                      ct["fv2"], ct["F12"], ct["ep"], T["scale"], T["sigma2"])
    out["tri_n"], out["tri_m12"], out["tri_F12"], out["tri_ep"] = np.int32(n), m12, ct["F12"], ct["ep"]
    cf = bow_cases.fuse_case(orc.extract)
    Fk = FrameOracle(cf["kps"], cf["desc"], cf["bounds"], T["scale"])
    bi, bd = fuse_search(Fk, cf["q"], cf["t"], cf["K"], float(np.float32(np.log(np.float64(T["scale"][1])))), T["inv_sigma2"],
                         cf["xw"], cf["normal"], cf["min_dist"], cf["max_dist"], cf["mp_desc"], cf["skip"], 3.0)
    out["fuse_idx"], out["fuse_dist"] = bi, bd
    np.savez_compressed(os.path.join(HERE, "bow_small.npz"), **out)
    # ---- DBoW2 transform on an excerpt of the reference's own vocabulary (ORBvoc.txt, k = 10, L = 6) ----
    voc_tar = "/root/reference/src/slam_system/orb_slam3/Vocabulary/ORBvoc.txt.tar.gz"
    if os.path.exists(voc_tar):
        import tarfile
        import tempfile

        from dvmslam_b200.vocabulary import flatten_tree, load_text
        from oracle.dbow import transform, transform_features

        with tempfile.TemporaryDirectory() as tmp:
            tarfile.open(voc_tar).extractall(tmp)
            k, L, sc, wt, parent, is_leaf, vdesc, vweight = load_text(os.path.join(tmp, "ORBvoc.txt"))
        tree = flatten_tree(parent, is_leaf, vdesc, vweight)
        cs, ch, vd, vw, wid = tree
        feat = extract_with_cv2(frame, 500)[1][:240]
        word, w, nid = transform_features(tree, L, feat, 4)
        bow, fv = transform(tree, L, wt, sc, feat, 4)
        # keep every node whose parent lies on a visited path (a descent only ever reads those)
        keep = np.zeros(len(wid), bool)
        keep[0] = True
        for i in range(len(feat)):
            node = 0
            while cs[node + 1] > cs[node]:
                kids = ch[cs[node]:cs[node + 1]]
                keep[kids] = True
                d = [int(np.unpackbits(feat[i] ^ vd[c]).sum()) for c in kids]
                node = int(kids[int(np.argmin(d))])
            assert wid[node] == word[i]
        new_id = np.cumsum(keep) - 1
        old = np.nonzero(keep)[0]
        counts = np.array([int(keep[ch[cs[o]:cs[o + 1]]].sum()) for o in old])
        p_cs = np.zeros(len(old) + 1, np.int32)
        p_cs[1:] = np.cumsum(counts)
        p_ch = np.concatenate([new_id[ch[cs[o]:cs[o + 1]][keep[ch[cs[o]:cs[o + 1]]]]] for o in old]).astype(np.int32)
        np.savez_compressed(os.path.join(HERE, "dbow_orbvoc.npz"), k=np.int32(k), L=np.int32(L), scoring=np.int32(sc),
                            weighting=np.int32(wt), child_start=p_cs, children=p_ch, desc=vd[old], weight=vw[old],
                            word_id=wid[old], orig_node=old.astype(np.int32), feat=feat, word=word, w=w, nid=nid,
                            bow_word=np.array(list(bow), np.int32), bow_value=np.array(list(bow.values())),
                            n_nodes_full=np.int32(len(wid)), n_words_full=np.int32(int((wid >= 0).sum())))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
