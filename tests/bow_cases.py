"""Seeded inputs of the SearchByBoW / SearchForInitialization / exhaustive-Hamming parity tests (shared by
the CPU oracle tests, the GPU parity tests and tests/golden/make_golden.py)."""
import numpy as np

from dvmslam_b200 import synth


def bow_pair(extract, k0=2, k1=5, seed=0, w=640, h=480, mp_frac=0.7, one_node=False):
    """Two views of the plane stream -> dict(desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2)."""
    S = synth.PlaneStream(w, h, seed=3, K=(500.0, 500.0, w / 2, h / 2))
    rng = np.random.default_rng(seed)
    out = {}
    for tag, k in (("1", k0), ("2", k1)):
        kps, desc, _ = extract(S.frame(k))
        out["desc" + tag], out["angle" + tag] = desc, kps["angle"].copy()
        out["valid" + tag] = (rng.random(len(kps)) < mp_frac).astype(np.uint8)
        out["fv" + tag] = {5: list(range(len(kps)))} if one_node else synth.toy_feature_vector(desc)
    return out


def bow_synthetic(n1, n2, seed, dup=False, nodes=9):
    """Random descriptors with planted noisy copies; optionally many exact duplicates (ties) and shuffled
    per-node index order (push_back order need not be ascending)."""
    rng = np.random.default_rng(seed)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    m = min(n1, n2) // 2
    if m:
        src = rng.choice(n1, m, replace=False)
        dst = rng.choice(n2, m, replace=False)
        d2[dst] = synth.noisy_copy(d1[src], 0.06, rng)
        if dup:   # exact duplicates on both sides: equal distances, first one must win
            d2[rng.choice(n2, m // 2)] = d2[dst[: m // 2]]
            d1[rng.choice(n1, m // 3)] = d1[src[: m // 3]]
    def fv(d):
        f = {}
        for i in rng.permutation(len(d)):
            f.setdefault(int(d[i, 0] % nodes) * 7 + 1, []).append(int(i))
        return f
    return dict(desc1=d1, angle1=rng.uniform(0, 360, n1).astype(np.float32), valid1=(rng.random(n1) < 0.8).astype(np.uint8),
                fv1=fv(d1), desc2=d2, angle2=rng.uniform(0, 360, n2).astype(np.float32),
                valid2=(rng.random(n2) < 0.8).astype(np.uint8), fv2=fv(d2))


def init_pair(extract, k0=1, k1=4, w=640, h=480):
    """Frames for SearchForInitialization: F1 keypoints/descriptors, F2 keypoints/descriptors, bounds."""
    S = synth.PlaneStream(w, h, seed=3, K=(500.0, 500.0, w / 2, h / 2))
    k1_, d1, _ = extract(S.frame(k0))
    k2_, d2, _ = extract(S.frame(k1))
    prev = np.stack([k1_["x"], k1_["y"]], 1).astype(np.float32)
    return dict(kps1=k1_, desc1=d1, kps2=k2_, desc2=d2, prev=prev, bounds=(0.0, 0.0, float(w), float(h)))


def triangulation_pair(extract, k0=2, k1=6, w=640, h=480, seed=0, mp_frac=0.5):
    """Two keyframes of the plane stream with their true relative pose: inputs of SearchForTriangulation."""
    from oracle.bow import fundamental_from_poses

    S = synth.PlaneStream(w, h, seed=3, K=(500.0, 500.0, w / 2, h / 2))
    rng = np.random.default_rng(seed)
    out = {}
    poses = []
    for tag, k in (("1", k0), ("2", k1)):
        kps, desc, _ = extract(S.frame(k))
        out["kps" + tag], out["desc" + tag] = kps, desc
        out["has_mp" + tag] = (rng.random(len(kps)) < mp_frac).astype(np.uint8)
        out["fv" + tag] = synth.toy_feature_vector(desc)
        R, t = S.pose(k)
        poses.append((synth.quat_from_R(np.asarray(R, np.float64)).astype(np.float32), np.asarray(t, np.float32)))
    K = np.array(S.K, np.float32)
    out["F12"], out["ep"] = fundamental_from_poses(poses[0][0], poses[0][1], poses[1][0], poses[1][1], K, K)
    out["K"], out["poses"], out["stream"] = K, poses, S
    return out


def fuse_case(extract, k=4, k_src=1, w=640, h=480, seed=0, n_points=1500):
    """A keyframe (view k) and map points lifted from another view's keypoints: inputs of Fuse."""
    S = synth.PlaneStream(w, h, seed=3, K=(500.0, 500.0, w / 2, h / 2))
    rng = np.random.default_rng(seed)
    kps, desc, _ = extract(S.frame(k))
    mk, md, _ = extract(S.frame(k_src))
    sel = np.sort(rng.permutation(len(mk))[:n_points])
    X = synth.backproject_to_plane(S, k_src, np.stack([mk["x"][sel], mk["y"][sel]], 1).astype(np.float64)).astype(np.float32)
    R0, t0 = S.pose(k_src)
    Ow = -(np.asarray(R0, np.float64).T @ np.asarray(t0, np.float64))
    PO = X.astype(np.float64) - Ow
    d = np.linalg.norm(PO, axis=1)
    normal = (PO / d[:, None]).astype(np.float32)
    lvl = mk["octave"][sel].astype(np.float64)
    # MapPoint::UpdateNormalAndDepth: mfMaxDistance = dist * scale[level], mfMinDistance = mfMaxDistance / scale[nLevels - 1]
    max_d = (d * 1.2 ** lvl).astype(np.float32)
    min_d = (max_d / np.float32(1.2 ** 7)).astype(np.float32)
    R, t = S.pose(k)
    q = synth.quat_from_R(np.asarray(R, np.float64)).astype(np.float32)
    skip = (rng.random(len(sel)) < 0.1).astype(np.uint8)
    return dict(kps=kps, desc=desc, q=q, t=np.asarray(t, np.float32), K=np.array(S.K, np.float32), xw=X, normal=normal,
                min_dist=min_d, max_dist=max_d, mp_desc=np.ascontiguousarray(md[sel]), skip=skip,
                bounds=(0.0, 0.0, float(w), float(h)))


def _lift(S, k_src, mk, sel):
    """Keypoints `sel` of view k_src lifted to the plane, with normals and distance ranges as UpdateNormalAndDepth leaves them."""
    X = synth.backproject_to_plane(S, k_src, np.stack([mk["x"][sel], mk["y"][sel]], 1).astype(np.float64))
    R0, t0 = S.pose(k_src)
    Ow = -(np.asarray(R0, np.float64).T @ np.asarray(t0, np.float64))
    PO = X - Ow
    d = np.linalg.norm(PO, axis=1)
    max_d = d * 1.2 ** mk["octave"][sel].astype(np.float64)
    return X, (PO / d[:, None]), max_d / 1.2 ** 7, max_d


def sim3_quat(R, s):
    """RxSO3 quaternion (x, y, z, w) of scale s: |q|^2 = s (Sophus::Sim3f::quaternion())."""
    return (synth.quat_from_R(np.asarray(R, np.float64)) * np.sqrt(s)).astype(np.float32)


def sim3_projection_case(extract, k=4, k_src=1, w=640, h=480, seed=0, n_points=1500, scale=1.3):
    """A keyframe (view k) and candidate map points of ANOTHER map whose frame differs from the keyframe's world by a
    similarity: inputs of SearchByProjection(pKF, Scw, ...) and Fuse(pKF, Scw, ...).  With Scw = (s, R, s t) the reference's
    Tcw = SE3(Scw.rotationMatrix(), Scw.translation() / Scw.scale()) is the keyframe's pose."""
    S = synth.PlaneStream(w, h, seed=3, K=(500.0, 500.0, w / 2, h / 2))
    rng = np.random.default_rng(seed)
    kps, desc, _ = extract(S.frame(k))
    mk, md, _ = extract(S.frame(k_src))
    sel = np.sort(rng.permutation(len(mk))[:n_points])
    X, normal, min_d, max_d = _lift(S, k_src, mk, sel)
    R, t = S.pose(k)
    sq = sim3_quat(R, scale)
    st = (np.asarray(t, np.float64) * scale).astype(np.float32)
    skip = (rng.random(len(sel)) < 0.1).astype(np.uint8)
    kp_matched = (rng.random(len(kps)) < 0.2).astype(np.uint8)
    return dict(kps=kps, desc=desc, sq=sq, st=st, K=np.array(S.K, np.float32), xw=X.astype(np.float32),
                normal=normal.astype(np.float32), min_dist=min_d.astype(np.float32), max_dist=max_d.astype(np.float32),
                mp_desc=np.ascontiguousarray(md[sel]), skip=skip, kp_matched=kp_matched, bounds=(0.0, 0.0, float(w), float(h)))


def search_by_sim3_case(extract, k1=2, k2=6, w=640, h=480, seed=0, c=1.7, mp_frac=0.8):
    """Two keyframes of two maps (map B = map A scaled by c) with a map point behind most keypoints and the similarity
    S12 (camera 2 -> camera 1): inputs of SearchBySim3."""
    S = synth.PlaneStream(w, h, seed=3, K=(500.0, 500.0, w / 2, h / 2))
    rng = np.random.default_rng(seed)
    out = {"K": np.array(S.K, np.float32), "bounds": (0.0, 0.0, float(w), float(h))}
    Rs, ts = [], []
    for tag, k, unit in (("1", k1, 1.0), ("2", k2, c)):
        kps, desc, _ = extract(S.frame(k))
        sel = np.arange(len(kps))
        X, _, min_d, max_d = _lift(S, k, kps, sel)
        R, t = S.pose(k)
        Rs.append(np.asarray(R, np.float64)); ts.append(np.asarray(t, np.float64))
        out["kps" + tag], out["desc" + tag] = kps, desc
        out["skip" + tag] = (rng.random(len(kps)) > mp_frac).astype(np.uint8)
        out["xw" + tag] = (X * unit).astype(np.float32)
        out["min" + tag], out["max" + tag] = (min_d * unit).astype(np.float32), (max_d * unit).astype(np.float32)
        # the map point's descriptor: the keypoint's own with a few flipped bits (another observation is the representative)
        out["mpdesc" + tag] = synth.noisy_copy(desc, 0.02, rng)
        out["q" + tag] = synth.quat_from_R(np.asarray(R, np.float64)).astype(np.float32)
        out["t" + tag] = (np.asarray(t, np.float64) * unit).astype(np.float32)
    R12 = Rs[0] @ Rs[1].T
    t12 = ts[0] - R12 @ ts[1]
    out["s12q"] = sim3_quat(R12, 1.0 / c)
    out["s12t"] = t12.astype(np.float32)
    return out
