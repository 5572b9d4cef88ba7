// orb_matcher_adapter.h -- the bodies of ORB_SLAM3::ORBmatcher's hot methods (O3/include/ORBmatcher.h:40-95)
// over the C-ABI.  Templates over the reference's own Frame / KeyFrame / MapPoint types: they read exactly
// the members the reference's implementation reads, flatten the pointer graph to the flat arrays of
// include/dvmslam_b200.h, call the library and write the result back where the reference writes it.
// In O3/src/ORBmatcher.cc every replaced body becomes one line.
//
// Mono path only (Frame::Nleft == -1, no right camera), like the C-ABI.
#pragma once
#include <cstring>
#include <list>
#include <memory>
#include <utility>
#include <set>
#include <vector>

#include "dvm_host.h"

namespace dvm_host {

// ---------------------------------------------------------------------------------------------------
// A Frame's device twin (keypoints, descriptors, 64x48 grid in HBM).  Created on first use and kept in a
// small per-thread LRU keyed by Frame::mnId, so that the 2-3 matcher calls and the PoseOptimization calls of
// one tracked frame upload it once.  (A maintainer who prefers an explicit member adds
// `std::shared_ptr<dvm_host::DeviceFrame>` to Frame and calls device_frame_create in the mono constructor,
// O3/src/Frame.cc:411-479.)
// ---------------------------------------------------------------------------------------------------
struct DeviceFrame {
    FrameHandle frame;
    long unsigned int id = ~0ul;
    int n = 0;
};

template <class FrameT>
std::shared_ptr<DeviceFrame> device_frame_create(const FrameT& F)
{
    auto d = std::make_shared<DeviceFrame>();
    const int n = static_cast<int>(F.mvKeysUn.size());
    check(dvm_frame_create(&d->frame.h, device_from_env(), nullptr, n > 0 ? n : 1, static_cast<int>(F.mvScaleFactors.size()),
                           F.mvScaleFactors.data(), F.mvInvLevelSigma2.data()),
          "dvm_frame_create");
    // mvKeysUn is what AssignFeaturesToGrid and every matcher read (O3/src/Frame.cc:481-506)
    std::vector<uint8_t> desc(static_cast<size_t>(n) * 32);
    for (int i = 0; i < n; i++) std::memcpy(&desc[static_cast<size_t>(i) * 32], F.mDescriptors.ptr(i), 32);
    check(dvm_frame_assign(d->frame.h, reinterpret_cast<const dvm_keypoint*>(F.mvKeysUn.data()), desc.data(), n,
                           F.mnMinX, F.mnMinY, F.mnMaxX, F.mnMaxY),
          "dvm_frame_assign");
    d->id = F.mnId;
    d->n = n;
    return d;
}

template <class FrameT>
DeviceFrame& device_frame(const FrameT& F)
{
    thread_local std::list<std::shared_ptr<DeviceFrame>> lru;
    for (auto it = lru.begin(); it != lru.end(); ++it)
        if ((*it)->id == F.mnId && (*it)->n == static_cast<int>(F.mvKeysUn.size())) {
            lru.splice(lru.begin(), lru, it);
            return *lru.front();
        }
    lru.push_front(device_frame_create(F));
    if (lru.size() > 4) lru.pop_back();
    return *lru.front();
}

// ---------------------------------------------------------------------------------------------------
// int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
// O3/src/ORBmatcher.cc:1553-1748.  Call sites clear CurrentFrame.mvpMapPoints first
// (O3/src/Tracking.cc:2603, 2618); entries that are already set are kept and block nothing here.
// ---------------------------------------------------------------------------------------------------
template <class FrameT>
int SearchByProjectionLast(FrameT& Cur, const FrameT& Last, float th, bool checkOrientation)
{
    DeviceFrame& dev = device_frame(Cur);
    const auto Tcw = Cur.GetPose();
    // the pose goes over as the SE3f holds it: the library applies Sophus' quaternion action (x3Dc = Tcw * x3Dw, :1577)
    const auto uq = Tcw.unit_quaternion();
    const auto t = Tcw.translation();
    const float qcw[4] = { uq.x(), uq.y(), uq.z(), uq.w() };
    const float tcw[3] = { t(0), t(1), t(2) };
    const float K[4] = { Cur.fx, Cur.fy, Cur.cx, Cur.cy };
    const int n = Last.N;
    std::vector<uint8_t> has_mp(n), outlier(n), obs_pos(n), desc(static_cast<size_t>(n) * 32);
    std::vector<float> Xw(static_cast<size_t>(n) * 3), angle(n);
    std::vector<int32_t> octave(n);
    for (int i = 0; i < n; i++) {
        auto* pMP = Last.mvpMapPoints[i];
        has_mp[i] = pMP != nullptr;
        outlier[i] = Last.mvbOutlier[i];
        octave[i] = Last.mvKeys[i].octave;
        angle[i] = Last.mvKeysUn[i].angle;
        if (!pMP) continue;
        const auto x3Dw = pMP->GetWorldPos();
        for (int k = 0; k < 3; k++) Xw[static_cast<size_t>(i) * 3 + k] = x3Dw(k);
        const auto d = pMP->GetDescriptor();
        std::memcpy(&desc[static_cast<size_t>(i) * 32], d.ptr(0), 32);
        obs_pos[i] = pMP->Observations() > 0;
    }
    std::vector<int32_t> cur_mp(Cur.N > 0 ? Cur.N : 1, -1);
    int nmatches = 0;
    check(dvm_match_by_projection_last(dev.frame.h, qcw, tcw, K, n, has_mp.data(), outlier.data(), Xw.data(), desc.data(),
                                       obs_pos.data(), octave.data(), angle.data(), th, checkOrientation ? 1 : 0,
                                       cur_mp.data(), &nmatches),
          "ORBmatcher::SearchByProjection(Frame&, const Frame&)");
    for (int i = 0; i < Cur.N; i++)
        if (cur_mp[i] >= 0) Cur.mvpMapPoints[i] = Last.mvpMapPoints[cur_mp[i]];
    return nmatches;
}

// ---------------------------------------------------------------------------------------------------
// int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th,
//                                    const bool bFarPoints, const float thFarPoints)
// O3/src/ORBmatcher.cc:44-205 (+ RadiusByViewingCos :207-212, applied on the device)
// ---------------------------------------------------------------------------------------------------
template <class FrameT, class MapPointT>
int SearchByProjectionMap(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, float th, float nnratio,
                          bool bFarPoints = false, float thFarPoints = 50.0f)
{
    DeviceFrame& dev = device_frame(F);
    std::vector<int> sel;            // vpMapPoints indices that take part, in order
    sel.reserve(vpMapPoints.size());
    for (size_t i = 0; i < vpMapPoints.size(); i++) {
        MapPointT* pMP = vpMapPoints[i];
        if (!pMP->mbTrackInView) continue;                             // :52-53 (mono: no mbTrackInViewR)
        if (bFarPoints && pMP->mTrackDepth > thFarPoints) continue;    // :55-56
        if (pMP->isBad()) continue;                                    // :58-59
        sel.push_back(static_cast<int>(i));
    }
    const int m = static_cast<int>(sel.size());
    std::vector<float> px(m), py(m), vc(m);
    std::vector<int32_t> lvl(m);
    std::vector<uint8_t> desc(static_cast<size_t>(m) * 32), obs_pos(m), blocked(F.N > 0 ? F.N : 1);
    for (int k = 0; k < m; k++) {
        MapPointT* pMP = vpMapPoints[sel[k]];
        px[k] = pMP->mTrackProjX; py[k] = pMP->mTrackProjY; lvl[k] = pMP->mnTrackScaleLevel; vc[k] = pMP->mTrackViewCos;
        const auto d = pMP->GetDescriptor();
        std::memcpy(&desc[static_cast<size_t>(k) * 32], d.ptr(0), 32);
        obs_pos[k] = pMP->Observations() > 0;
    }
    for (int i = 0; i < F.N; i++) blocked[i] = F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0;   // :86-88
    std::vector<int32_t> cur_mp(F.N > 0 ? F.N : 1, -1);
    int nmatches = 0;
    check(dvm_match_by_projection_map(dev.frame.h, m, px.data(), py.data(), lvl.data(), vc.data(), desc.data(),
                                      obs_pos.data(), th, nnratio, blocked.data(), cur_mp.data(), &nmatches),
          "ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&)");
    for (int i = 0; i < F.N; i++)
        if (cur_mp[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[sel[cur_mp[i]]];
    return nmatches;
}

// ---------------------------------------------------------------------------------------------------
// SearchByBoW: a DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>) flattened to CSR form
// ---------------------------------------------------------------------------------------------------
struct FlatFeatures {
    std::vector<uint8_t> desc, has_mp;
    std::vector<float> angle;
    std::vector<uint32_t> node_id, feat_idx;
    std::vector<int32_t> node_start;
    dvm_bow_features view() const
    {
        dvm_bow_features v;
        v.n = static_cast<int32_t>(angle.size());
        v.desc = desc.data(); v.angle = angle.data(); v.has_mp = has_mp.empty() ? nullptr : has_mp.data();
        v.n_nodes = static_cast<int32_t>(node_id.size());
        v.node_id = node_id.data(); v.node_start = node_start.data(); v.feat_idx = feat_idx.data();
        return v;
    }
};

template <class KeyPointVec, class MatT, class FeatVecT, class MapPointT>
FlatFeatures flatten_features(const KeyPointVec& keysUn, const MatT& descriptors, const FeatVecT& featVec,
                              const std::vector<MapPointT*>* mapPoints)
{
    FlatFeatures f;
    const size_t n = keysUn.size();
    f.desc.resize(n * 32);
    f.angle.resize(n);
    for (size_t i = 0; i < n; i++) {
        std::memcpy(&f.desc[i * 32], descriptors.ptr(static_cast<int>(i)), 32);
        f.angle[i] = keysUn[i].angle;
    }
    if (mapPoints) {
        f.has_mp.resize(n);
        for (size_t i = 0; i < n; i++) f.has_mp[i] = (*mapPoints)[i] && !(*mapPoints)[i]->isBad();
    }
    f.node_start.push_back(0);
    for (const auto& kv : featVec) {               // std::map iteration = ascending node id
        f.node_id.push_back(static_cast<uint32_t>(kv.first));
        for (unsigned idx : kv.second) f.feat_idx.push_back(idx);
        f.node_start.push_back(static_cast<int32_t>(f.feat_idx.size()));
    }
    return f;
}

// int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)   :214-393
template <class KeyFrameT, class FrameT, class MapPointT>
int SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches, float nnratio, bool checkOrientation)
{
    const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches.assign(static_cast<size_t>(F.N), static_cast<MapPointT*>(nullptr));
    const FlatFeatures a = flatten_features(pKF->mvKeysUn, pKF->mDescriptors, pKF->mFeatVec, &vpMapPointsKF);
    const FlatFeatures b = flatten_features(F.mvKeys, F.mDescriptors, F.mFeatVec, static_cast<const std::vector<MapPointT*>*>(nullptr));
    const dvm_bow_features va = a.view(), vb = b.view();
    std::vector<int32_t> match21(F.N > 0 ? F.N : 1, -1);
    int nmatches = 0;
    check(dvm_match_by_bow(device_frame(F).frame.h, 0, &va, &vb, nnratio, checkOrientation ? 1 : 0, nullptr, match21.data(),
                           &nmatches),
          "ORBmatcher::SearchByBoW(KeyFrame*, Frame&)");
    for (int i = 0; i < F.N; i++)
        if (match21[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[match21[i]];
    return nmatches;
}

// int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)   :709-834
// ctx: any device frame (e.g. device_frame(currentFrame)); it only provides the GPU stream and staging buffers.
template <class KeyFrameT, class MapPointT>
int SearchByBoW(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, float nnratio, bool checkOrientation,
                dvm_frame* ctx)
{
    const std::vector<MapPointT*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
    vpMatches12.assign(mp1.size(), static_cast<MapPointT*>(nullptr));
    const FlatFeatures a = flatten_features(pKF1->mvKeysUn, pKF1->mDescriptors, pKF1->mFeatVec, &mp1);
    const FlatFeatures b = flatten_features(pKF2->mvKeysUn, pKF2->mDescriptors, pKF2->mFeatVec, &mp2);
    const dvm_bow_features va = a.view(), vb = b.view();
    std::vector<int32_t> match12(mp1.empty() ? 1 : mp1.size(), -1);
    int nmatches = 0;
    check(dvm_match_by_bow(ctx, 1, &va, &vb, nnratio, checkOrientation ? 1 : 0, match12.data(), nullptr, &nmatches),
          "ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*)");
    for (size_t i = 0; i < mp1.size(); i++)
        if (match12[i] >= 0) vpMatches12[i] = mp2[match12[i]];
    return nmatches;
}

// int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched,
//                                         vector<int>& vnMatches12, int windowSize)            :605-707
template <class FrameT, class Point2fT>
int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<Point2fT>& vbPrevMatched, std::vector<int>& vnMatches12,
                            int windowSize, float nnratio, bool checkOrientation)
{
    const int n1 = static_cast<int>(F1.mvKeysUn.size());
    vnMatches12.assign(static_cast<size_t>(n1), -1);
    std::vector<uint8_t> desc1(static_cast<size_t>(n1) * 32);
    std::vector<float> prev(static_cast<size_t>(n1) * 2);
    for (int i = 0; i < n1; i++) {
        std::memcpy(&desc1[static_cast<size_t>(i) * 32], F1.mDescriptors.ptr(i), 32);
        prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y;
    }
    int nmatches = 0;
    check(dvm_match_for_initialization(device_frame(F2).frame.h, n1, reinterpret_cast<const dvm_keypoint*>(F1.mvKeysUn.data()),
                                       desc1.data(), prev.data(), windowSize, nnratio, checkOrientation ? 1 : 0,
                                       vnMatches12.data(), &nmatches),
          "ORBmatcher::SearchForInitialization");
    for (int i = 0; i < n1; i++) { vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
    return nmatches;
}

// ---------------------------------------------------------------------------------------------------
// int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, vector<pair<size_t,size_t>>& vMatchedPairs,
//                                        const bool bOnlyStereo, const bool bCoarse)             :836-1058
// Mono keyframes (bOnlyStereo must be false).  The relative pose, the epipole and the fundamental matrix are
// derived by dvm_fundamental_from_poses as the reference does (:841-860; Pinhole::epipolarConstrain,
// O3/src/CameraModels/Pinhole.cpp:104-110) -- same float32 operations in the same order.
// ---------------------------------------------------------------------------------------------------
template <class KeyFrameT>
int SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<std::pair<size_t, size_t>>& vMatchedPairs,
                           bool bOnlyStereo, bool bCoarse, bool checkOrientation, dvm_frame* ctx)
{
    vMatchedPairs.clear();
    if (bOnlyStereo) return 0;   // mono keyframes have no stereo observations: every feature is skipped (:896-898)
    const auto T1w = pKF1->GetPose();
    const auto T2w = pKF2->GetPose();
    // T12 = T1w * Tw2, the epipole and F12 = K1^-T [t12]x R12 K2^-1 in the reference's float32 Sophus / Eigen arithmetic
    const auto q1 = T1w.unit_quaternion(), q2 = T2w.unit_quaternion();
    const auto tr1 = T1w.translation(), tr2 = T2w.translation();
    const float qa[4] = { q1.x(), q1.y(), q1.z(), q1.w() }, ta[3] = { tr1(0), tr1(1), tr1(2) };
    const float qb[4] = { q2.x(), q2.y(), q2.z(), q2.w() }, tb[3] = { tr2(0), tr2(1), tr2(2) };
    const float K1[4] = { pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy }, K2[4] = { pKF2->fx, pKF2->fy, pKF2->cx, pKF2->cy };
    float F12[9], ep[2];
    check(dvm_fundamental_from_poses(qa, ta, qb, tb, K1, K2, F12, ep), "ORBmatcher::SearchForTriangulation (relative pose)");

    auto flatten = [](KeyFrameT* kf) {
        FlatFeatures f = flatten_features(kf->mvKeysUn, kf->mDescriptors, kf->mFeatVec,
                                          static_cast<const std::vector<decltype(kf->GetMapPoint(0))>*>(nullptr));
        f.has_mp.resize(f.angle.size());
        for (size_t i = 0; i < f.angle.size(); i++) f.has_mp[i] = kf->GetMapPoint(i) != nullptr;
        return f;
    };
    const FlatFeatures a = flatten(pKF1), b = flatten(pKF2);
    const dvm_bow_features va = a.view(), vb = b.view();
    std::vector<int32_t> m12(a.angle.empty() ? 1 : a.angle.size(), -1);
    int nmatches = 0;
    check(dvm_match_for_triangulation(ctx, &va, reinterpret_cast<const dvm_keypoint*>(pKF1->mvKeysUn.data()), &vb,
                                      reinterpret_cast<const dvm_keypoint*>(pKF2->mvKeysUn.data()), F12, ep,
                                      pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(),
                                      static_cast<int>(pKF2->mvScaleFactors.size()), bCoarse ? 1 : 0, checkOrientation ? 1 : 0,
                                      m12.data(), &nmatches),
          "ORBmatcher::SearchForTriangulation");
    vMatchedPairs.reserve(nmatches);
    for (size_t i = 0; i < a.angle.size(); i++)
        if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, static_cast<size_t>(m12[i])));                     // :1049-1054
    return nmatches;
}

// ---------------------------------------------------------------------------------------------------
// int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th, const bool bRight)
// :1060-1228, bRight = false.  The search runs on the GPU for all map points at once; Replace / AddObservation /
// AddMapPoint are then applied in vpMapPoints order exactly as the reference's loop does (:1209-1222).
// ---------------------------------------------------------------------------------------------------
template <class KeyFrameT, class MapPointT>
int Fuse(KeyFrameT* pKF, const std::vector<MapPointT*>& vpMapPoints, float th)
{
    const int m = static_cast<int>(vpMapPoints.size());
    std::vector<float> xw(static_cast<size_t>(m) * 3), nrm(static_cast<size_t>(m) * 3), mind(m), maxd(m);
    std::vector<uint8_t> desc(static_cast<size_t>(m) * 32), skip(m > 0 ? m : 1);
    for (int i = 0; i < m; i++) {
        MapPointT* pMP = vpMapPoints[i];
        skip[i] = !pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF);                                                     // :1086-1098
        if (skip[i]) continue;
        const auto p = pMP->GetWorldPos();
        const auto n = pMP->GetNormal();
        for (int k = 0; k < 3; k++) { xw[3 * i + k] = p(k); nrm[3 * i + k] = n(k); }
        // mfMinDistance / mfMaxDistance themselves: the library forms 0.8f * min and 1.2f * max for the range gate as
        // Get{Min,Max}DistanceInvariance do and feeds mfMaxDistance to PredictScale (MapPoint.cc:547-587).  The members
        // are protected in the reference; INTEGRATION.md adds the two one-line getters used here (dividing the
        // invariance values by 0.8f / 1.2f does not give the members back bit for bit).
        mind[i] = pMP->GetMinDistance();
        maxd[i] = pMP->GetMaxDistance();
        const auto d = pMP->GetDescriptor();
        std::memcpy(&desc[static_cast<size_t>(i) * 32], d.ptr(0), 32);
    }
    float q[4], t[3];
    const auto Tcw = pKF->GetPose();
    const auto uq = Tcw.unit_quaternion();
    q[0] = uq.x(); q[1] = uq.y(); q[2] = uq.z(); q[3] = uq.w();
    const auto tr = Tcw.translation();
    for (int k = 0; k < 3; k++) t[k] = tr(k);
    const float K[4] = { pKF->fx, pKF->fy, pKF->cx, pKF->cy };
    std::vector<int32_t> bestIdx(m > 0 ? m : 1, -1), bestDist(m > 0 ? m : 1, 256);
    check(dvm_fuse_search(device_frame(*pKF).frame.h, q, t, K, m, xw.data(), nrm.data(), mind.data(), maxd.data(), desc.data(),
                          skip.data(), th, bestIdx.data(), bestDist.data()),
          "ORBmatcher::Fuse");
    int nFused = 0;
    for (int i = 0; i < m; i++) {
        if (bestIdx[i] < 0) continue;
        MapPointT* pMP = vpMapPoints[i];
        if (pMP->isBad()) continue;             // replaced by an earlier iteration of this loop (:1086-1090 re-evaluated per iteration)
        if (pMP->IsInKeyFrame(pKF)) continue;   // the same map point listed twice: added by its first occurrence
        MapPointT* pMPinKF = pKF->GetMapPoint(bestIdx[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx[i]);
            pKF->AddMapPoint(pMP, bestIdx[i]);
        }
        nFused++;
    }
    return nFused;
}

// ---------------------------------------------------------------------------------------------------
// Sim3-guided matchers of LoopClosing (O3/src/LoopClosing.cc:823-847 and FindMatchesByProjection).  A Sophus::Sim3f goes to
// the library as quaternion().coeffs() (x, y, z, w; squared norm = scale) + translation(), as stored.
// ---------------------------------------------------------------------------------------------------
namespace detail {
template <class Sim3T>
inline void flatten_sim3(const Sim3T& S, float q[4], float t[3])
{
    const auto sq = S.quaternion();
    const auto st = S.translation();
    q[0] = sq.x(); q[1] = sq.y(); q[2] = sq.z(); q[3] = sq.w();
    for (int k = 0; k < 3; k++) t[k] = st(k);
}
template <class SE3T>
inline void flatten_se3(const SE3T& T, float q[4], float t[3])
{
    const auto uq = T.unit_quaternion();
    const auto tr = T.translation();
    q[0] = uq.x(); q[1] = uq.y(); q[2] = uq.z(); q[3] = uq.w();
    for (int k = 0; k < 3; k++) t[k] = tr(k);
}
// world position, normal, mfMinDistance / mfMaxDistance (getters added by INTEGRATION.md) and descriptor of one map point
struct FlatPoints {
    std::vector<float> xw, nrm, mind, maxd;
    std::vector<uint8_t> desc, skip;
    explicit FlatPoints(size_t m) : xw(m * 3), nrm(m * 3), mind(m), maxd(m), desc(m * 32), skip(m ? m : 1) { }
    template <class MapPointT>
    void set(size_t i, MapPointT* pMP)
    {
        const auto p = pMP->GetWorldPos();
        const auto n = pMP->GetNormal();
        for (int k = 0; k < 3; k++) { xw[3 * i + k] = p(k); nrm[3 * i + k] = n(k); }
        mind[i] = pMP->GetMinDistance();
        maxd[i] = pMP->GetMaxDistance();
        const auto d = pMP->GetDescriptor();
        std::memcpy(&desc[i * 32], d.ptr(0), 32);
    }
};
} // namespace detail

// int ORBmatcher::SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints,
//                                    vector<MapPoint*>& vpMatched, int th, float ratioHamming)          :395-494
template <class KeyFrameT, class Sim3T, class MapPointT>
int SearchByProjection(KeyFrameT* pKF, Sim3T& Scw, const std::vector<MapPointT*>& vpPoints, std::vector<MapPointT*>& vpMatched, int th,
                       float ratioHamming = 1.0f)
{
    const size_t m = vpPoints.size();
    std::set<MapPointT*> spAlreadyFound(vpMatched.begin(), vpMatched.end());                  // :406-407
    spAlreadyFound.erase(static_cast<MapPointT*>(nullptr));
    detail::FlatPoints P(m);
    for (size_t i = 0; i < m; i++) {
        MapPointT* pMP = vpPoints[i];
        P.skip[i] = pMP->isBad() || spAlreadyFound.count(pMP);                                // :416-417
        if (!P.skip[i]) P.set(i, pMP);
    }
    std::vector<uint8_t> taken(vpMatched.size() ? vpMatched.size() : 1);
    for (size_t k = 0; k < vpMatched.size(); k++) taken[k] = vpMatched[k] != nullptr;
    float sq[4], st[3];
    detail::flatten_sim3(Scw, sq, st);
    const float K[4] = { pKF->fx, pKF->fy, pKF->cx, pKF->cy };
    std::vector<int32_t> kp_point(vpMatched.size() ? vpMatched.size() : 1, -1);
    int nmatches = 0;
    check(dvm_match_by_projection_sim3(device_frame(*pKF).frame.h, sq, st, K, static_cast<int>(m), P.xw.data(), P.nrm.data(),
                                       P.mind.data(), P.maxd.data(), P.desc.data(), P.skip.data(), taken.data(), th, ratioHamming,
                                       kp_point.data(), &nmatches),
          "ORBmatcher::SearchByProjection(KeyFrame*, Sim3f&, ...)");
    for (size_t k = 0; k < vpMatched.size(); k++)
        if (kp_point[k] >= 0) vpMatched[k] = vpPoints[kp_point[k]];                           // :487-489
    return nmatches;
}

// The overload that also reports the keyframe each matched point came from                             :496-603
template <class KeyFrameT, class Sim3T, class MapPointT>
int SearchByProjection(KeyFrameT* pKF, Sim3T& Scw, const std::vector<MapPointT*>& vpPoints, const std::vector<KeyFrameT*>& vpPointsKFs,
                       std::vector<MapPointT*>& vpMatched, std::vector<KeyFrameT*>& vpMatchedKF, int th, float ratioHamming = 1.0f)
{
    const std::vector<MapPointT*> before = vpMatched;
    const int n = SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming);
    for (size_t k = 0; k < vpMatched.size(); k++) {
        if (vpMatched[k] == before[k]) continue;
        for (size_t i = 0; i < vpPoints.size(); i++)
            if (vpPoints[i] == vpMatched[k]) { vpMatchedKF[k] = vpPointsKFs[i]; break; }     // :596-598 (first occurrence wins)
    }
    return n;
}

// int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const Sophus::Sim3f& S12,
//                              const float th)                                                         :1347-1551
template <class KeyFrameT, class Sim3T, class MapPointT>
int SearchBySim3(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const Sim3T& S12, float th)
{
    const std::vector<MapPointT*> vp1 = pKF1->GetMapPointMatches(), vp2 = pKF2->GetMapPointMatches();
    const int N1 = static_cast<int>(vp1.size()), N2 = static_cast<int>(vp2.size());
    std::vector<uint8_t> matched1(N1 ? N1 : 1, 0), matched2(N2 ? N2 : 1, 0);
    for (int i = 0; i < N1; i++) {                                                           // :1370-1381
        MapPointT* pMP = vpMatches12[i];
        if (!pMP) continue;
        matched1[i] = 1;
        const int idx2 = std::get<0>(pMP->GetIndexInKeyFrame(pKF2));
        if (idx2 >= 0 && idx2 < N2) matched2[idx2] = 1;
    }
    detail::FlatPoints A(N1), B(N2);
    for (int i = 0; i < N1; i++) {
        A.skip[i] = !vp1[i] || matched1[i] || vp1[i]->isBad();                                // :1389-1394
        if (!A.skip[i]) A.set(i, vp1[i]);
    }
    for (int i = 0; i < N2; i++) {
        B.skip[i] = !vp2[i] || matched2[i] || vp2[i]->isBad();                                // :1465-1470
        if (!B.skip[i]) B.set(i, vp2[i]);
    }
    float q1[4], t1[3], q2[4], t2[3], sq[4], st[3];
    detail::flatten_se3(pKF1->GetPose(), q1, t1);
    detail::flatten_se3(pKF2->GetPose(), q2, t2);
    detail::flatten_sim3(S12, sq, st);
    const float K[4] = { pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy };
    std::vector<int32_t> m12(N1 ? N1 : 1, -1);
    int nFound = 0;
    check(dvm_match_by_sim3(device_frame(*pKF1).frame.h, device_frame(*pKF2).frame.h, q1, t1, q2, t2, sq, st, K, A.skip.data(),
                            A.xw.data(), A.mind.data(), A.maxd.data(), A.desc.data(), B.skip.data(), B.xw.data(), B.mind.data(),
                            B.maxd.data(), B.desc.data(), th, m12.data(), &nFound),
          "ORBmatcher::SearchBySim3");
    for (int i = 0; i < N1; i++)
        if (m12[i] >= 0) vpMatches12[i] = vp2[m12[i]];                                        // :1545
    return nFound;
}

// int ORBmatcher::Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints, float th,
//                      vector<MapPoint*>& vpReplacePoint)                                              :1236-1345
template <class KeyFrameT, class Sim3T, class MapPointT>
int Fuse(KeyFrameT* pKF, Sim3T& Scw, const std::vector<MapPointT*>& vpPoints, float th, std::vector<MapPointT*>& vpReplacePoint)
{
    const size_t m = vpPoints.size();
    const std::set<MapPointT*> spAlreadyFound = pKF->GetMapPoints();                          // :1249
    detail::FlatPoints P(m);
    for (size_t i = 0; i < m; i++) {
        MapPointT* pMP = vpPoints[i];
        P.skip[i] = pMP->isBad() || spAlreadyFound.count(pMP);                                // :1260-1261
        if (!P.skip[i]) P.set(i, pMP);
    }
    float sq[4], st[3];
    detail::flatten_sim3(Scw, sq, st);
    const float K[4] = { pKF->fx, pKF->fy, pKF->cx, pKF->cy };
    std::vector<int32_t> bestIdx(m ? m : 1, -1), bestDist(m ? m : 1, 256);
    check(dvm_fuse_search_sim3(device_frame(*pKF).frame.h, sq, st, K, static_cast<int>(m), P.xw.data(), P.nrm.data(), P.mind.data(),
                               P.maxd.data(), P.desc.data(), P.skip.data(), th, bestIdx.data(), bestDist.data()),
          "ORBmatcher::Fuse(KeyFrame*, Sim3f&, ...)");
    int nFused = 0;
    for (size_t i = 0; i < m; i++) {                                                          // :1328-1341 in vpPoints order
        if (bestIdx[i] < 0) continue;
        MapPointT* pMP = vpPoints[i];
        MapPointT* pMPinKF = pKF->GetMapPoint(bestIdx[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, bestIdx[i]);
            pKF->AddMapPoint(pMP, bestIdx[i]);
        }
        nFused++;
    }
    return nFused;
}

} // namespace dvm_host
