"""DBoW2 vocabulary transform (Frame::ComputeBoW): oracle self-checks on CPU, GPU parity against the oracle
(word ids, node ids: bit-exact; BowVector values: identical doubles, same summation order)."""
import numpy as np
import pytest

from dvmslam_b200 import synth


def _dist(a, b):
    return int(np.unpackbits(a ^ b).sum())


def _py_descend(tree, L, f, levelsup):
    cs, ch, d, w, wid = tree
    nid_level = L - levelsup
    nid = 0 if nid_level <= 0 else -1
    node, level = 0, 0
    while cs[node + 1] > cs[node]:
        level += 1
        kids = ch[cs[node]:cs[node + 1]]
        dists = [_dist(f, d[k]) for k in kids]
        node = int(kids[int(np.argmin(dists))])          # argmin = first of equal minima, like the strict '<'
        if level == nid_level:
            nid = node
    return int(wid[node]), float(w[node]), node if nid < 0 else nid


@pytest.mark.parametrize("k,L,ragged,levelsup", [(10, 3, False, 1), (10, 3, False, 4), (7, 4, True, 2), (3, 5, True, 2)])
def test_oracle_descent_matches_python(k, L, ragged, levelsup):
    from dvmslam_b200.vocabulary import flatten_tree
    from oracle.dbow import transform_features

    v = synth.toy_vocabulary(k, L, seed=k + L, ragged=ragged)
    tree = flatten_tree(v["parent"], v["is_leaf"], v["desc"], v["weight"])
    rng = np.random.default_rng(1)
    feat = synth.noisy_copy(v["desc"][rng.integers(0, len(v["desc"]), 300)], 0.05, rng)
    word, w, nid = transform_features(tree, L, feat, levelsup)
    for i in range(0, 300, 7):
        assert (int(word[i]), float(w[i]), int(nid[i])) == _py_descend(tree, L, feat[i], levelsup)
    assert (word >= 0).all() and len(np.unique(word)) > 50


def test_oracle_bow_vector_properties():
    from dvmslam_b200.vocabulary import flatten_tree
    from oracle.dbow import transform, transform_features

    v = synth.toy_vocabulary(10, 3, seed=2)
    tree = flatten_tree(v["parent"], v["is_leaf"], v["desc"], v["weight"])
    rng = np.random.default_rng(2)
    feat = synth.noisy_copy(v["desc"][rng.integers(0, len(v["desc"]), 500)], 0.05, rng)
    word, w, nid = transform_features(tree, 3, feat, 2)
    bow, fv = transform(tree, 3, 0, 0, feat, 2)                     # TF_IDF, L1
    assert abs(sum(bow.values()) - 1.0) < 1e-12
    assert list(bow) == sorted(bow) and list(fv) == sorted(fv)
    live = w > 0
    assert set(bow) == set(word[live].tolist())
    assert sorted(i for l in fv.values() for i in l) == np.nonzero(live)[0].tolist()   # stopped words carry no feature
    for node, idx in fv.items():
        assert idx == sorted(idx) and (nid[idx] == node).all()
    bow_b, _ = transform(tree, 3, 3, 5, feat, 2)                    # BINARY, DOT_PRODUCT: first weight, no norm
    for k_, val in bow_b.items():
        assert val == w[np.nonzero(word == k_)[0][0]]
    bow2, _ = transform(tree, 3, 1, 1, feat, 2)                     # TF, L2
    assert abs(sum(x * x for x in bow2.values()) - 1.0) < 1e-12


def test_text_loader_roundtrip(tmp_path):
    """loadFromTextFile's format: the flat tree built from a written file equals the one built from the arrays."""
    from dvmslam_b200.vocabulary import flatten_tree

    v = synth.toy_vocabulary(5, 3, seed=4, ragged=True)
    p = tmp_path / "voc.txt"
    synth.write_vocabulary_text(str(p), v)
    with open(p) as f:
        k, L, sc, wt = (int(x) for x in f.readline().split()[:4])
        rows = np.loadtxt(f, dtype=np.float64, ndmin=2)
    assert (k, L, sc, wt) == (5, 3, 0, 0)
    a = flatten_tree(v["parent"], v["is_leaf"], v["desc"], v["weight"])
    b = flatten_tree(rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64), rows[:, 2:34].astype(np.uint8), rows[:, 34])
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    cs, ch = a[0], a[1]
    assert cs[0] == 0 and cs[-1] == len(ch) == len(v["parent"])
    for node in range(len(cs) - 1):                                  # children keep file order under every parent
        kids = ch[cs[node]:cs[node + 1]]
        assert (np.diff(kids) > 0).all() and (v["parent"][kids - 1] == node).all()


@pytest.mark.gpu
@pytest.mark.parametrize("k,L,ragged,levelsup,weighting,scoring", [(10, 3, False, 1, 0, 0), (10, 4, False, 2, 0, 0),
                                                                   (7, 4, True, 2, 1, 1), (3, 5, True, 4, 3, 5), (20, 2, True, 1, 2, 2)])
def test_vocabulary_transform_gpu(tmp_path, k, L, ragged, levelsup, weighting, scoring):
    from dvmslam_b200.vocabulary import Vocabulary
    from oracle.dbow import transform, transform_features
    from oracle.orb import OrbOracle

    v = synth.toy_vocabulary(k, L, seed=k * L, ragged=ragged)
    v["weighting"], v["scoring"] = weighting, scoring
    p = tmp_path / "voc.txt"
    synth.write_vocabulary_text(str(p), v)
    voc = Vocabulary.loadFromTextFile(str(p))
    tree = (voc.child_start, voc.children, voc.desc, voc.weight, voc.word_id)
    rng = np.random.default_rng(3)
    _, real, _ = OrbOracle(1000).extract(synth.frame(640, 480, 2))
    planted = synth.noisy_copy(v["desc"][rng.integers(0, len(v["desc"]), 1500)], 0.04, rng)
    ties = np.repeat(v["desc"][rng.integers(0, len(v["desc"]), 50)], 2, axis=0)          # exact node descriptors
    for feat in (np.concatenate([planted, ties, real]), real[:1], real[:0]):
        w0, ww0, n0 = transform_features(tree, L, feat, levelsup)
        w1, ww1, n1 = voc.transform_features(feat, levelsup)
        assert np.array_equal(w0, w1) and np.array_equal(n0, n1) and np.array_equal(ww0, ww1)
        bow0, fv0 = transform(tree, L, weighting, scoring, feat, levelsup)
        bow1, fv1 = voc.transform(feat, levelsup)
        assert list(bow0.items()) == list(bow1.items())              # identical doubles, identical order
        assert fv0 == fv1 and list(fv0) == list(fv1)
    assert len(bow1) == 0 and len(fv1) == 0
    voc.close()


def _orbvoc_excerpt():
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dbow_orbvoc.npz"))
    tree = (g["child_start"], g["children"], g["desc"], g["weight"], g["word_id"])
    return g, tree


def test_oracle_on_orbvoc_excerpt():
    """The reference's own vocabulary (ORBvoc.txt: k 10, L 6, 1 082 073 nodes, 971 814 words), cut down to the nodes
    that the golden descriptors' descents read: the oracle on the excerpt reproduces the full-tree results."""
    from oracle.dbow import transform, transform_features

    g, tree = _orbvoc_excerpt()
    assert (int(g["k"]), int(g["L"]), int(g["n_nodes_full"]), int(g["n_words_full"])) == (10, 6, 1082073, 971814)
    word, w, nid = transform_features(tree, int(g["L"]), g["feat"], 4)
    assert np.array_equal(word, g["word"]) and np.array_equal(w, g["w"]) and np.array_equal(g["orig_node"][nid], g["nid"])
    assert len(np.unique(g["nid"])) <= 100                          # level L - 4 = 2 of a 10-ary tree holds 100 nodes
    bow, _ = transform(tree, int(g["L"]), int(g["weighting"]), int(g["scoring"]), g["feat"], 4)
    assert np.array_equal(np.array(list(bow), np.int32), g["bow_word"]) and np.array_equal(np.array(list(bow.values())), g["bow_value"])


@pytest.mark.gpu
def test_gpu_on_orbvoc_excerpt():
    import ctypes as C

    from dvmslam_b200._lib import check, lib

    g, tree = _orbvoc_excerpt()
    L = lib()
    vp = C.c_void_p
    L.dvm_vocabulary_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_int]
    L.dvm_vocabulary_transform.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.dvm_vocabulary_destroy.argtypes = [vp]
    cs, ch, d, w, wid = (np.ascontiguousarray(a, t) for a, t in zip(tree, (np.int32, np.int32, np.uint8, np.float64, np.int32)))
    h = vp()
    check(L.dvm_vocabulary_create(C.byref(h), 0, len(wid), cs.ctypes.data, ch.ctypes.data, d.ctypes.data, w.ctypes.data,
                                  wid.ctypes.data, int(g["L"])))
    feat = np.ascontiguousarray(g["feat"])
    n = len(feat)
    word, ww, nid = np.zeros(n, np.int32), np.zeros(n, np.float64), np.zeros(n, np.int32)
    check(L.dvm_vocabulary_transform(h, feat.ctypes.data, n, 4, word.ctypes.data, ww.ctypes.data, nid.ctypes.data))
    L.dvm_vocabulary_destroy(h)
    assert np.array_equal(word, g["word"]) and np.array_equal(ww, g["w"]) and np.array_equal(g["orig_node"][nid], g["nid"])
