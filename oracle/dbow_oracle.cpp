/*
 * oracle/dbow_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the DBoW2 vocabulary transform that Frame::ComputeBoW / KeyFrame::ComputeBoW call
 * (O3/src/Frame.cc:784-789: mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)), following
 * (DB/ = /root/reference/src/slam_system/orb_slam3/Thirdparty/DBoW2/DBoW2/):
 *   TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)   DB/TemplatedVocabulary.h:1106-1146
 *   TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)   DB/TemplatedVocabulary.h:1025-1086
 *   BowVector::addWeight / addIfNotExist / normalize                          DB/BowVector.cpp:30-71
 *   FeatureVector::addFeature                                                 DB/FeatureVector.cpp:27-37
 *   FORB::distance                                                            DB/FORB.cpp:80-97
 *   scoring -> normalisation table                                            DB/ScoringObject.h:76-91
 * The tree is passed flat: children of node i are children[child_start[i] .. child_start[i+1]) in the order
 * loadFromTextFile pushes them (DB/TemplatedVocabulary.h:1248-1285); a node without children is a leaf (word).
 * Convention where the reference leaves a value undefined: a feature that reaches a leaf above the requested
 * level (nid never assigned, :1140-1141) reports that leaf as its node.
 *
 * Parity status: PINNED to the reference source.  The reference's vendored DBoW2 (TemplatedVocabulary<FORB::TDescriptor,
 * FORB>, FORB.cpp, BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp) is compiled unmodified over a stand-in
 * <opencv2/core/core.hpp> into oracle/_ref/libref_dbow.so (oracle/Makefile `ref`, oracle/dbowshim); tests/test_ref_dbow.py
 * requires word ids, weights, node ids, BowVector (identical doubles) and FeatureVector to be equal on toy vocabularies of
 * every weighting and on the reference's own ORBvoc.txt loaded by the reference's own loadFromTextFile.  Excluded: the two
 * undefined behaviours of the reference named in that test (the loader's artefact node after a final newline; the
 * uninitialised node id of a leaf above the requested level, for which the convention above applies).
 */
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

namespace {

int forb_distance(const uint8_t* a, const uint8_t* b)
{
    const int32_t* pa = (const int32_t*)a;
    const int32_t* pb = (const int32_t*)b;
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        unsigned int v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

struct Tree {
    int n_nodes;
    const int32_t* child_start;
    const int32_t* children;
    const uint8_t* desc;
    const double* weight;
    const int32_t* word_id;
    int L;
};

void transform_one(const Tree& t, const uint8_t* f, int levelsup, int& word, double& w, int& nid)
{
    const int nid_level = t.L - levelsup;
    nid = -1;
    if (nid_level <= 0) nid = 0;
    int final_id = 0, current_level = 0;
    do {
        ++current_level;
        const int c0 = t.child_start[final_id], c1 = t.child_start[final_id + 1];
        final_id = t.children[c0];
        double best_d = forb_distance(f, t.desc + (size_t)final_id * 32);
        for (int c = c0 + 1; c < c1; c++) {
            const int id = t.children[c];
            const double d = forb_distance(f, t.desc + (size_t)id * 32);
            if (d < best_d) { best_d = d; final_id = id; }
        }
        if (current_level == nid_level) nid = final_id;
    } while (t.child_start[final_id + 1] > t.child_start[final_id]);
    if (nid < 0) nid = final_id;
    word = t.word_id[final_id];
    w = t.weight[final_id];
}

} // namespace

extern "C" {

/* per feature: word id, word weight, node id at level L - levelsup */
void dbowo_transform_features(int n_nodes, const int32_t* child_start, const int32_t* children, const uint8_t* desc,
                              const double* weight, const int32_t* word_id, int L, const uint8_t* feat, int n, int levelsup,
                              int32_t* word, double* w, int32_t* nid)
{
    const Tree t = { n_nodes, child_start, children, desc, weight, word_id, L };
    for (int i = 0; i < n; i++) transform_one(t, feat + (size_t)i * 32, levelsup, word[i], w[i], nid[i]);
}

/* the whole transform(features, BowVector, FeatureVector, levelsup).  weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY;
 * scoring: 0 L1_NORM .. 5 DOT_PRODUCT.  Outputs in std::map order: bow_word/bow_value [<= n], fv_node [<= n],
 * fv_start [<= n + 1], fv_idx [n]; counts[2] = { BowVector size, FeatureVector size }. */
void dbowo_transform(int n_nodes, const int32_t* child_start, const int32_t* children, const uint8_t* desc,
                     const double* weight, const int32_t* word_id, int L, int weighting, int scoring, const uint8_t* feat,
                     int n, int levelsup, int32_t* bow_word, double* bow_value, int32_t* fv_node, int32_t* fv_start,
                     int32_t* fv_idx, int32_t* counts)
{
    const Tree t = { n_nodes, child_start, children, desc, weight, word_id, L };
    std::map<unsigned, double> v;
    std::map<unsigned, std::vector<unsigned>> fv;
    const bool must = scoring != 5;
    const bool l2 = scoring == 1;
    for (int i = 0; i < n; i++) {
        int id, nid;
        double w;
        transform_one(t, feat + (size_t)i * 32, levelsup, id, w, nid);
        if (!(w > 0)) continue;                       /* stopped word */
        if (weighting == 0 || weighting == 1) v[(unsigned)id] += w;         /* addWeight */
        else v.insert(std::make_pair((unsigned)id, w));                     /* addIfNotExist */
        fv[(unsigned)nid].push_back((unsigned)i);
    }
    if ((weighting == 0 || weighting == 1) && !v.empty() && !must) {
        const double nd = (double)v.size();
        for (auto& kv : v) kv.second /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!l2) for (auto& kv : v) norm += std::fabs(kv.second);
        else { for (auto& kv : v) norm += kv.second * kv.second; norm = std::sqrt(norm); }
        if (norm > 0.0) for (auto& kv : v) kv.second /= norm;
    }
    int k = 0;
    for (auto& kv : v) { bow_word[k] = (int32_t)kv.first; bow_value[k] = kv.second; k++; }
    counts[0] = k;
    k = 0;
    int p = 0;
    fv_start[0] = 0;
    for (auto& kv : fv) {
        fv_node[k] = (int32_t)kv.first;
        for (unsigned i : kv.second) fv_idx[p++] = (int32_t)i;
        fv_start[++k] = p;
    }
    counts[1] = k;
}

} // extern "C"
