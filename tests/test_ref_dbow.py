"""Pins the DBoW2 restatement (oracle/dbow_oracle.cpp) and the vocabulary loader (dvmslam_b200/vocabulary.py: load_text,
flatten_tree) to the reference's OWN vendored DBoW2, compiled unmodified from /root/reference (oracle/Makefile `ref`,
oracle/_ref/libref_dbow.so): its loadFromTextFile and the transform Frame::ComputeBoW calls.  Word ids, node ids and
weights per feature, the BowVector (identical doubles: same summation order) and the FeatureVector must be equal.
CPU only; skipped where neither the reference nor the built library is present."""
import os

import numpy as np
import pytest

from dvmslam_b200 import synth

refdbow = pytest.importorskip("oracle.refdbow")
pytestmark = pytest.mark.skipif(not refdbow.available(), reason="no /root/reference and no oracle/_ref/libref_dbow.so")

VOC_TAR = "/root/reference/src/slam_system/orb_slam3/Vocabulary/ORBvoc.txt.tar.gz"


def _both(tmp_path, v, feat, levelsup):
    from dvmslam_b200.vocabulary import flatten_tree, load_text
    from oracle.dbow import transform, transform_features

    p = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(p, v)
    # The reference's loader loops `while (!f.eof())` (TemplatedVocabulary.h:1249): after a final newline it parses one more,
    # empty, line into a node whose parent is 0 and whose descriptor, weight and leaf flag are whatever the failed
    # extractions left behind (uninitialised).  That artefact is undefined behaviour, not an algorithm: the pin feeds both
    # sides a file WITHOUT the final newline.
    with open(p) as f:
        text = f.read().rstrip("\n")
    with open(p, "w") as f:
        f.write(text)
    ref = refdbow.RefVocabulary(p)
    k, L, sc, wt, parent, is_leaf, desc, weight = load_text(p)
    assert (ref.k, ref.L, ref.scoring, ref.weighting) == (k, L, sc, wt)
    tree = flatten_tree(parent, is_leaf, desc, weight)
    assert ref.n_words == int((tree[4] >= 0).sum())
    r = ref.transform_features(feat, levelsup), ref.transform(feat, levelsup)
    o = transform_features(tree, L, feat, levelsup), transform(tree, L, wt, sc, feat, levelsup)
    ref.close()
    # Where a leaf sits ABOVE level L - levelsup (ragged trees; ORBvoc.txt has such leaves too) the reference never writes
    # `nid`: transform(features, v, fv, levelsup) declares `NodeId nid;` inside its loop and hands the uninitialised value to
    # fv.addFeature (TemplatedVocabulary.h:1044-1055) -- undefined behaviour (in practice the previous feature's node).  The
    # restatement and the kernel file such a feature under its leaf node.  `defined` marks the features whose node id the
    # reference does define.
    cs, ch = tree[0], tree[1]
    depth = np.zeros(len(cs) - 1, np.int64)
    for node in range(len(cs) - 1):          # children follow their parent in loadFromTextFile order
        depth[ch[cs[node]:cs[node + 1]]] = depth[node] + 1
    leaf_of_word = np.full(int(tree[4].max()) + 1, -1, np.int64)
    leaf_of_word[tree[4][tree[4] >= 0]] = np.nonzero(tree[4] >= 0)[0]
    defined = (depth[leaf_of_word[o[0][0]]] >= L - levelsup) | (L - levelsup <= 0) if len(feat) else np.zeros(0, bool)
    return r, o, defined


def _node_of_feature(fv, n):
    out = np.full(n, -1, np.int64)
    for node, idx in fv.items():
        out[idx] = node
    return out


@pytest.mark.parametrize("k,L,ragged,levelsup,weighting,scoring",
                         [(10, 3, False, 1, 0, 0), (10, 3, False, 4, 0, 0), (7, 4, True, 2, 0, 0), (3, 5, True, 2, 1, 1),
                          (5, 3, True, 0, 2, 0), (6, 3, False, 3, 3, 5), (4, 4, True, 1, 0, 2)])
def test_transform_matches_the_reference_dbow2(tmp_path, k, L, ragged, levelsup, weighting, scoring):
    v = synth.toy_vocabulary(k, L, seed=k + L, ragged=ragged)
    v["weighting"], v["scoring"] = weighting, scoring
    rng = np.random.default_rng(k * 7 + L)
    feat = synth.noisy_copy(v["desc"][rng.integers(0, len(v["desc"]), 400)], 0.05, rng)
    (rf, (rbow, rfv)), (of, (obow, ofv)), defined = _both(tmp_path, v, feat, levelsup)
    assert np.array_equal(rf[0], of[0]), "word ids"
    assert np.array_equal(rf[1], of[1]), "word weights"
    assert defined.sum() > len(feat) // 2
    assert np.array_equal(rf[2][defined], of[2][defined]), "node ids"
    assert list(rbow) == list(obow) and list(rbow.values()) == list(obow.values()), "BowVector (identical doubles)"
    rn, on = _node_of_feature(rfv, len(feat)), _node_of_feature(ofv, len(feat))
    assert np.array_equal(rn >= 0, on >= 0) and np.array_equal(on >= 0, of[1] > 0), "FeatureVector: the features of live words"
    assert np.array_equal(rn[defined], on[defined]), "FeatureVector: node of every feature"
    assert list(ofv) == sorted(ofv) and all(idx == sorted(idx) for idx in ofv.values())
    if defined.all():
        assert rfv == ofv and list(rfv) == list(ofv), "FeatureVector"


def test_empty_and_single_feature(tmp_path):
    v = synth.toy_vocabulary(5, 3, seed=9)
    (rf, (rbow, rfv)), (of, (obow, ofv)), _ = _both(tmp_path, v, np.zeros((0, 32), np.uint8), 2)
    assert rbow == obow == {} and rfv == ofv == {}
    feat = v["desc"][-1:].copy()
    (rf, (rbow, rfv)), (of, (obow, ofv)), defined = _both(tmp_path, v, feat, 2)
    assert defined.all()
    assert np.array_equal(rf[0], of[0]) and np.array_equal(rf[2], of[2]) and rbow == obow and rfv == ofv


@pytest.mark.skipif(not os.path.exists(VOC_TAR), reason="the reference's ORBvoc.txt is not here")
def test_full_orbvoc_through_the_reference_loader_matches_the_golden(tmp_path):
    """The reference's own vocabulary (k 10, L 6, 1 082 073 nodes) through the reference's own loader and transform: the
    committed golden (made with the restatement, tests/golden/dbow_orbvoc.npz) must be what it returns."""
    import tarfile

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dbow_orbvoc.npz"))
    tarfile.open(VOC_TAR).extractall(tmp_path, filter="data")
    ref = refdbow.RefVocabulary(str(tmp_path / "ORBvoc.txt"))
    assert (ref.k, ref.L, ref.scoring, ref.weighting) == (int(g["k"]), int(g["L"]), int(g["scoring"]), int(g["weighting"]))
    # (ORBvoc.txt ends with a newline: the reference's loader appends its artefact node, see _both)
    assert ref.n_words in (int(g["n_words_full"]), int(g["n_words_full"]) + 1)
    word, w, nid = ref.transform_features(g["feat"], 4)
    assert np.array_equal(word, g["word"]) and np.array_equal(w, g["w"]) and np.array_equal(nid, g["nid"])
    bow, fv = ref.transform(g["feat"], 4)
    assert list(bow) == g["bow_word"].tolist() and list(bow.values()) == g["bow_value"].tolist()
    ref.close()
