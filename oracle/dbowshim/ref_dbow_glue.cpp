// ref_dbow_glue.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// C entry points over the reference's OWN DBoW2 (TemplatedVocabulary<FORB::TDescriptor, FORB>, the ORBVocabulary of
// O3/include/ORBVocabulary.h:28), compiled unmodified from /root/reference by oracle/Makefile (target `ref`): the text
// loader ORB-SLAM3 added (TemplatedVocabulary.h:1211-1290) and the transform Frame::ComputeBoW calls (:1025-1093).
// Used by tests/test_ref_dbow.py to pin oracle/dbow_oracle.cpp and the vocabulary loader of dvmslam_b200/vocabulary.py.
#include "TemplatedVocabulary.h"
#include "FORB.h"
#include <cstdint>

namespace {
typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> Voc;
struct VocX : Voc { using Voc::transform; };   // the per-feature overload with weight and node id is protected

cv::Mat row_of(const uint8_t* d)
{
    cv::Mat m(1, DBoW2::FORB::L, CV_8U);
    std::memcpy(m.data, d, DBoW2::FORB::L);
    return m;
}
}

extern "C" {

void* refdbow_load_text(const char* path)
{
    VocX* v = new VocX();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void refdbow_free(void* h) { delete (VocX*)h; }

void refdbow_info(void* h, int* out5)
{
    const VocX* v = (const VocX*)h;
    out5[0] = v->getBranchingFactor(); out5[1] = v->getDepthLevels(); out5[2] = (int)v->getScoringType();
    out5[3] = (int)v->getWeightingType(); out5[4] = (int)v->size();
}

// per feature: word id, word weight, node id `levelsup` levels above the leaf
void refdbow_transform_features(void* h, const uint8_t* desc, int n, int levelsup, int* word, double* weight, int* node)
{
    const VocX* v = (const VocX*)h;
    for (int i = 0; i < n; i++) {
        DBoW2::WordId id; DBoW2::WordValue w; DBoW2::NodeId nid;
        v->transform(row_of(desc + 32 * i), id, w, &nid, levelsup);
        word[i] = (int)id; weight[i] = w; node[i] = (int)nid;
    }
}

// Frame::ComputeBoW: BowVector and FeatureVector in std::map order; counts[0] = words, counts[1] = nodes
void refdbow_transform(void* h, const uint8_t* desc, int n, int levelsup, int* bow_word, double* bow_val, int* fv_node, int* fv_start,
                       int* fv_idx, int* counts)
{
    const VocX* v = (const VocX*)h;
    std::vector<cv::Mat> feats;
    for (int i = 0; i < n; i++) feats.push_back(row_of(desc + 32 * i));
    DBoW2::BowVector bow;
    DBoW2::FeatureVector fv;
    v->transform(feats, bow, fv, levelsup);
    int k = 0;
    for (auto& e : bow) { bow_word[k] = (int)e.first; bow_val[k] = e.second; k++; }
    counts[0] = k;
    int m = 0, pos = 0;
    for (auto& e : fv) {
        fv_node[m] = (int)e.first;
        fv_start[m] = pos;
        for (unsigned int idx : e.second) fv_idx[pos++] = (int)idx;
        m++;
    }
    fv_start[m] = pos;
    counts[1] = m;
}

}
