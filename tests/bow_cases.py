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
