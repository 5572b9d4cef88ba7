// bow_kernels.cuh -- device structures and launchers of SearchByBoW, SearchForInitialization and the
// exhaustive Hamming search (bow_kernels.cu, hamming.cu).
#pragma once
#include "track_kernels.cuh"

namespace dvm {

// One side of SearchByBoW: descriptors, keypoint angles, map-point validity and the DBoW2 FeatureVector
// in CSR form (node ids ascending = std::map order; feat_idx in push_back order).  Device pointers.
struct BowSide {
    int n;
    const uint8_t* desc;
    const float* angle;
    const uint8_t* valid;      // may be nullptr: every feature takes part
    int n_nodes;
    const uint32_t* node_id;
    const int* node_start;
    const uint32_t* feat_idx;
};

struct BowArgs {
    BowSide a, b;
    int kf_kf;        // 0: SearchByBoW(KF, Frame)   1: SearchByBoW(KF, KF)
    float nnratio;
    int check_ori;
    int* match12;     // [a.n] partner in b or -1 (preset to -1)
    int* match21;     // [b.n] partner in a or -1 (preset to -1)
    int* histo;       // [kHistoLength] preset to 0
    int* counters;    // [2]: accepted pairs (preset 0), nmatches (out)
};

void launch_bow_match(const BowArgs& g, cudaStream_t stream);

// SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize); F2 is the FrameDev
struct InitMatchArgs {
    int n1;
    const dvm_keypoint* kps1;  // F1.mvKeysUn
    const uint8_t* desc1;      // F1.mDescriptors
    float* prev_matched;       // [n1 * 2] in/out
    int window;
    float nnratio;
    int check_ori;
    int* matches12;            // [n1] out
    int* result;               // [2] out: nmatches, fixed-point rounds
    // scratch
    int* choice_a; int* choice_b; int* cdist_a; int* cdist_b; int* next;  // [n1] each
    int* head;                                                            // [F2 cap]
    unsigned long long* cache;   // [n1 * 8] nearest candidates per query: distance << 48 | traversal order << 24 | index
    int* ncand;                  // [n1] candidates in the window (-1: the query does not take part)
};

void launch_init_match(const FrameDev& f2, const InitMatchArgs& a, cudaStream_t stream);

// SearchForTriangulation(pKF1, pKF2, ...) on mono keyframes: BowSide::valid here means "the feature already
// holds a map point" (such features are skipped on both sides)
struct TriArgs {
    BowSide a, b;
    const dvm_keypoint* kps1;   // pKF1->mvKeysUn
    const dvm_keypoint* kps2;   // pKF2->mvKeysUn
    float F12[9], ep[2];
    float scale2[kTrackMaxLevels], sigma2_2[kTrackMaxLevels];   // pKF2->mvScaleFactors / mvLevelSigma2
    int coarse, check_ori;
    int* matches12;   // [a.n] preset to -1
    int* histo;       // [kHistoLength] preset to 0
    int* counters;    // [2]: accepted (preset 0), nmatches (out)
};
void launch_triangulation_match(const TriArgs& g, cudaStream_t stream);

// Project-and-search, one warp per map point: the common core of Fuse(pKF, vpMapPoints, th) (:1060-1228), Fuse(pKF, Scw,
// ...) (:1236-1345), one direction of SearchBySim3 (:1347-1551) and the gate pass of SearchByProjection(pKF, Scw, ...)
// (:395-603).  The point goes through the SE3 (q, t) -- and then through the Sim3 (sq, st) when chain_sim3 is set -- is
// gated (depth, KeyFrame::IsInImage, scale-invariance range, viewing angle), its level predicted, and the keypoints of the
// window th * scale[level] with octave in [level - 1, level] compared; the first of the nearest wins.
struct FuseArgs {
    float q[4], t[3], K[4];
    int chain_sim3; float sq[4], st[3];   // SearchBySim3: p = S * (T * P)
    int proj_invz;        // SearchBySim3's u = fx * (x * (float)(1.0 / z)) + cx instead of Pinhole::project
    int dist_camera;      // SearchBySim3: the distance is |p| in the target camera frame, not |P - Ow|
    int check_normal;     // viewing-angle gate PO . Pn < 0.5 * dist
    int check_chi;        // Fuse(pKF, vpMapPoints, th) only: reprojection error gate e^2 * invSigma2 > 5.99
    int accept_th;        // TH_LOW (50) or TH_HIGH (100)
    int gate_only;        // SearchByProjection(pKF, Scw, ...): no search; gate_u / gate_v / gate_level receive the projection
    int nlevels; float logScale;
    float inv_sigma2[kTrackMaxLevels];
    int m;
    const float* xw; const float* normal; const float* min_dist; const float* max_dist;
    const uint8_t* mp_desc; const uint8_t* skip;
    float th;
    int* best_idx; int* best_dist;
    float* gate_u; float* gate_v; int* gate_level;   // gate_level[i] = -1 when the point is rejected
};
void launch_fuse_search(const FrameDev& kf, const FuseArgs& a, cudaStream_t stream);

// Exhaustive nearest / second-nearest search, batched over blocks (keyframes):
//   a [ba][na][32], b [bb][nb][32]  ->  key1 / key2 [ba][bb][na]
// key = distance << 20 | index in the b block; key1 = nearest (first of equal distances), key2 = the
// second smallest key (its distance is the reference's bestDist2); 256 << 20 when absent.
// counts (may be nullptr) [ba][bb]: rows with dist1 <= th_low and dist1 < nnratio * dist2.
struct KnnArgs {
    const uint8_t* a; int ba, na;
    const uint8_t* b; int bb, nb;
    uint32_t* key1; uint32_t* key2;
    int* counts;   // preset to 0 by the launcher
    int th_low; float nnratio;
};
struct KnnScratch {
    uint32_t* part[2] = { nullptr, nullptr }; size_t cap = 0;   // per-split partial keys (POPC path)
    uint8_t* expanded = nullptr; size_t exp_cap = 0;            // descriptors as +1 / -1 bytes (tensor-core path)
};
// mode: 0 = by size (tensor cores for batched keyframe blocks, POPC for small calls), 1 = POPC, 2 = tensor cores,
//       3 / 4 = timing probes of the tensor-core kernel that write nothing (3: TMA + MMA only, 4: + TMEM read-out)
int launch_hamming_knn(const KnnArgs& k, KnnScratch& scratch, cudaStream_t stream, int mode = 0);
bool hamming_tc_applicable(const KnnArgs& k);
int launch_hamming_knn_tc(const KnnArgs& k, KnnScratch& scratch, cudaStream_t stream, int epilogue);

} // namespace dvm
