/*
 * oracle/bow_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement, on flat arrays, of the descriptor matchers of the reference that do not project
 * (O3/ = /root/reference/src/slam_system/orb_slam3/):
 *   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)      O3/src/ORBmatcher.cc:214-393
 *   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)   O3/src/ORBmatcher.cc:709-834
 *   ORBmatcher::SearchForTriangulation (mono, epipolar gate of Pinhole.cpp)  O3/src/ORBmatcher.cc:836-1058
 *   ORBmatcher::ComputeThreeMaxima                                      O3/src/ORBmatcher.cc:1862-1896
 *   ORBmatcher::DescriptorDistance                                      O3/src/ORBmatcher.cc:1900-1914
 * and the exhaustive nearest / second-nearest Hamming search that the inter-agent exchange (config C3)
 * uses in place of the vocabulary prefilter.
 *
 * A DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>, DBoW2/FeatureVector.h) is passed in
 * CSR form: node ids ascending (the map's iteration order), node_start[n_nodes + 1], feat_idx in each
 * node's push_back order.  The mono path only (Nleft == -1, no second camera).
 *
 * Parity status: PINNED to the reference source.  The reference has no tests or fixtures for this path, so its own
 * ORBmatcher.cc (with the Frame / KeyFrame / MapPoint / Pinhole bodies it calls and DBoW2's FeatureVector / BowVector) is
 * compiled unmodified into oracle/_ref/libref_matcher.so (oracle/Makefile `ref`, shims in oracle/slamshim + oracle/cvshim)
 * and tests/test_ref_matchers.py requires every matcher's index arrays to be identical to this restatement's.
 */
#include "sophus_order.h"
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

const int TH_LOW = 50, HISTO_LENGTH = 30;

int descriptor_distance(const uint8_t* a, const uint8_t* b)
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

void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

struct Side {
    int n;
    const uint8_t* desc;   /* [n][32] */
    const float* angle;    /* keypoint angle per feature */
    const uint8_t* valid;  /* feature holds a map point that is not bad (may be NULL = all) */
    int n_nodes;
    const uint32_t* node_id;
    const int32_t* node_start;
    const uint32_t* feat_idx;
};

} // namespace

extern "C" {

/* kf_kf = 0: SearchByBoW(pKF = side 1, F = side 2): every feature of F is a candidate until it is
 *            matched; accept bestDist1 <= TH_LOW; the histogram holds F indices.
 * kf_kf = 1: SearchByBoW(pKF1, pKF2): candidates of KF2 need a map point; accept bestDist1 < TH_LOW;
 *            the histogram holds KF1 indices.
 * match12[n1] / match21[n2] receive the partner index or -1.  Returns nmatches. */
int bowo_search_by_bow(int kf_kf, int n1, const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int nn1,
                       const uint32_t* node1, const int32_t* start1, const uint32_t* idx1, int n2,
                       const uint8_t* desc2, const float* angle2, const uint8_t* valid2, int nn2,
                       const uint32_t* node2, const int32_t* start2, const uint32_t* idx2, float nnratio,
                       int checkOri, int* match12, int* match21)
{
    const Side A = { n1, desc1, angle1, valid1, nn1, node1, start1, idx1 };
    const Side B = { n2, desc2, angle2, valid2, nn2, node2, start2, idx2 };
    for (int i = 0; i < n1; i++) match12[i] = -1;
    for (int i = 0; i < n2; i++) match21[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    int ia = 0, ib = 0;
    while (ia < A.n_nodes && ib < B.n_nodes) {
        if (A.node_id[ia] == B.node_id[ib]) {
            for (int p = A.node_start[ia]; p < A.node_start[ia + 1]; p++) {
                const int r1 = (int)A.feat_idx[p];
                if (A.valid && !A.valid[r1]) continue;
                const uint8_t* d1 = A.desc + (size_t)r1 * 32;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int q = B.node_start[ib]; q < B.node_start[ib + 1]; q++) {
                    const int r2 = (int)B.feat_idx[q];
                    if (match21[r2] >= 0) continue;                       /* vpMapPointMatches[r2] / vbMatched2[r2] */
                    if (kf_kf && B.valid && !B.valid[r2]) continue;       /* !pMP2 || pMP2->isBad() */
                    const int dist = descriptor_distance(d1, B.desc + (size_t)r2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = r2; }
                    else if (dist < bestDist2) { bestDist2 = dist; }
                }
                const bool ok = kf_kf ? bestDist1 < TH_LOW : bestDist1 <= TH_LOW;
                if (ok && static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                    match21[bestIdx2] = r1;
                    match12[r1] = bestIdx2;
                    if (checkOri) {
                        float rot = A.angle[r1] - B.angle[bestIdx2];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(kf_kf ? r1 : bestIdx2);
                    }
                    nmatches++;
                }
            }
            ia++; ib++;
        } else if (A.node_id[ia] < B.node_id[ib]) {
            while (ia < A.n_nodes && A.node_id[ia] < B.node_id[ib]) ia++;   /* lower_bound */
        } else {
            while (ib < B.n_nodes && B.node_id[ib] < A.node_id[ia]) ib++;
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) {
                if (kf_kf) { match21[match12[j]] = -1; match12[j] = -1; }
                else { match12[match21[j]] = -1; match21[j] = -1; }
                nmatches--;
            }
        }
    }
    return nmatches;
}

/* Exhaustive search: for every row of A the nearest row of B (first of equal distances), its
 * distance and the second-smallest distance (256 when B has fewer than two rows) -- the
 * bestDist1 / bestDist2 bookkeeping of the matchers above over all of B. */
void bowo_hamming_knn(const uint8_t* A, int na, const uint8_t* B, int nb, int* best_idx, int* best_dist,
                      int* second_dist)
{
    for (int i = 0; i < na; i++) {
        int b1 = 256, b2 = 256, bi = -1;
        for (int j = 0; j < nb; j++) {
            const int d = descriptor_distance(A + (size_t)i * 32, B + (size_t)j * 32);
            if (d < b1) { b2 = b1; b1 = d; bi = j; }
            else if (d < b2) { b2 = d; }
        }
        best_idx[i] = bi; best_dist[i] = b1; second_dist[i] = b2;
    }
}

/* SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo = false, bCoarse)  (O3/src/ORBmatcher.cc:836-1058),
 * mono keyframes.  has_mp[i] = GetMapPoint(i) != NULL (such features are skipped on both sides); kps = mvKeysUn as
 * {x, y, size, angle, response, octave, class_id} records; ep = the projection of camera 1's centre into image 2;
 * F12 (row-major 3x3) = K1^-T [t12]x R12 K2^-1 as Pinhole::epipolarConstrain builds it
 * (O3/src/CameraModels/Pinhole.cpp:104-127) -- computed once by the caller and shared with the CUDA path, since
 * the check 'dsqr < 3.84 * unc' is evaluated in float.  Note that this fork never sets vbMatched2: features of
 * keyframe 2 can be taken several times, and equal distances go to the LAST candidate ('dist > bestDist' skips).
 * matches12[n1] receives vMatches12; returns nmatches. */
struct OKeyPt { float x, y, size, angle, response; int32_t octave, class_id; };

int bowo_search_for_triangulation(int n1, const uint8_t* desc1, const void* kps1_, const uint8_t* has_mp1, int nn1,
                                  const uint32_t* node1, const int32_t* start1, const uint32_t* idx1, int n2,
                                  const uint8_t* desc2, const void* kps2_, const uint8_t* has_mp2, int nn2,
                                  const uint32_t* node2, const int32_t* start2, const uint32_t* idx2, const float* F12,
                                  const float* ep, const float* scaleFactors2, const float* levelSigma2_2, int bCoarse,
                                  int checkOri, int* matches12)
{
    const OKeyPt* kps1 = (const OKeyPt*)kps1_;
    const OKeyPt* kps2 = (const OKeyPt*)kps2_;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    int ia = 0, ib = 0;
    while (ia < nn1 && ib < nn2) {
        if (node1[ia] == node2[ib]) {
            for (int p = start1[ia]; p < start1[ia + 1]; p++) {
                const int i1 = (int)idx1[p];
                if (has_mp1[i1]) continue;
                const OKeyPt& kp1 = kps1[i1];
                const uint8_t* d1 = desc1 + (size_t)i1 * 32;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int q = start2[ib]; q < start2[ib + 1]; q++) {
                    const int i2 = (int)idx2[q];
                    if (has_mp2[i2]) continue;
                    const int dist = descriptor_distance(d1, desc2 + (size_t)i2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const OKeyPt& kp2 = kps2[i2];
                    const float distex = ep[0] - kp2.x, distey = ep[1] - kp2.y;
                    if (distex * distex + distey * distey < 100 * scaleFactors2[kp2.octave]) continue;
                    bool ok = bCoarse != 0;
                    if (!ok) {
                        const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
                        const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
                        const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
                        const float num = a * kp2.x + b * kp2.y + c;
                        const float den = a * a + b * b;
                        if (den != 0) {
                            const float dsqr = num * num / den;
                            ok = dsqr < 3.84 * levelSigma2_2[kp2.octave];   /* double comparison, as in the reference */
                        }
                    }
                    if (ok) { bestIdx2 = i2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    matches12[i1] = bestIdx2;
                    nmatches++;
                    if (checkOri) {
                        float rot = kp1.angle - kps2[bestIdx2].angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(i1);
                    }
                }
            }
            ia++; ib++;
        } else if (node1[ia] < node2[ib]) {
            while (ia < nn1 && node1[ia] < node2[ib]) ia++;
        } else {
            while (ib < nn2 && node2[ib] < node1[ia]) ib++;
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { matches12[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}


/* The relative geometry SearchForTriangulation works with (O3/src/ORBmatcher.cc:841-860, mono keyframes): from the two
 * keyframe poses Tcw as stored (q = x, y, z, w; t), T12 = T1w * Tw2 with Tw2 = T2w.inverse() (Sophus: normalising
 * quaternion product, so3.hpp:325-340), R12 = T12.rotationMatrix(), t12 = T12.translation(); the epipole
 * ep = project(T2w * Cw) with Cw = T1w.inverse().translation(); and the fundamental matrix
 * Pinhole::epipolarConstrain rebuilds for every candidate pair (O3/src/CameraModels/Pinhole.cpp:104-110):
 * F12 = K1.transpose().inverse() * hat(t12) * R12 * K2.inverse(), Eigen's 3x3 inverse and left-to-right coefficient
 * products.  F12 row-major [9], ep [2]. */
void bowo_fundamental_from_poses(const float* q1, const float* t1, const float* q2, const float* t2, const float* K1,
                                 const float* K2, float* F12, float* ep)
{
    float qw2[4], tw2[3], q12[4], t12[3], qw1[4], Cw[3], C2[3];
    so::se3_inverse(q2, t2, qw2, tw2);
    so::se3_mul(q1, t1, qw2, tw2, q12, t12);
    float R12[9];
    so::quat_to_matrix(q12, R12);
    so::se3_inverse(q1, t1, qw1, Cw);
    so::se3_apply(q2, t2, Cw, C2);
    ep[0] = K2[0] * C2[0] / C2[2] + K2[2];
    ep[1] = K2[1] * C2[1] / C2[2] + K2[3];
    const float t12x[9] = { 0.f, -t12[2], t12[1], t12[2], 0.f, -t12[0], -t12[1], t12[0], 0.f };
    const float K1T[9] = { K1[0], 0.f, 0.f, 0.f, K1[1], 0.f, K1[2], K1[3], 1.f };   /* toK_() transposed */
    const float K2m[9] = { K2[0], 0.f, K2[2], 0.f, K2[1], K2[3], 0.f, 0.f, 1.f };
    float K1Ti[9], K2i[9], A[9], B[9];
    so::mat_inverse(K1T, K1Ti);
    so::mat_inverse(K2m, K2i);
    so::mat_mul(K1Ti, t12x, A);
    so::mat_mul(A, R12, B);
    so::mat_mul(B, K2i, F12);
}

} // extern "C"
