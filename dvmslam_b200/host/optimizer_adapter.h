// optimizer_adapter.h -- the bodies of ORB_SLAM3::Optimizer::PoseOptimization and ::LocalBundleAdjustment
// (O3/include/Optimizer.h:56-59) over the C-ABI.  Templates over the reference's own Frame / KeyFrame /
// MapPoint / Map types; the pointer-graph part (which keyframes and map points form the window, what is
// erased and written back afterwards) is the reference's own logic, restated around one GPU call.
// Mono observations only (mvuRight < 0, no second camera), like the C-ABI.
#pragma once
#include <cmath>
#include <cstdio>
#include <limits>
#include <list>
#include <map>
#include <mutex>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <set>
#include <vector>

#include "orb_matcher_adapter.h"

namespace dvm_host {
namespace detail {
// The BA adapters run on worker threads of the reference (LocalMapping, the detached GBA thread of LoopClosing) where nothing
// catches exceptions.  A window beyond the dense solver's capacity (DVM_ERR_CAPACITY: more than 2000 free keyframes) is
// reported and treated like an aborted optimisation -- the map is left untouched -- instead of terminating the process.
inline bool ba_ok(int rc, const char* what)
{
    if (rc == DVM_ERR_CAPACITY) {
        std::fprintf(stderr, "dvmslam_b200: %s skipped: %s\n", what, dvm_last_error());
        return false;
    }
    check(rc, what);
    return true;
}
} // namespace detail

namespace detail {
// e->pCamera = pKFi->mpCamera (O3/src/Optimizer.cc:1219): every keyframe's own intrinsics go to the solver (merged maps
// mix keyframes of agents with different calibrations)
template <class KeyFrameT>
inline void set_camera_intrinsics(dvm_lba* solver, const std::vector<KeyFrameT*>& cams)
{
    std::vector<float> camK(cams.size() * 4);
    for (size_t c = 0; c < cams.size(); c++) {
        camK[4 * c] = cams[c]->fx; camK[4 * c + 1] = cams[c]->fy; camK[4 * c + 2] = cams[c]->cx; camK[4 * c + 3] = cams[c]->cy;
    }
    check(dvm_lba_set_camera_intrinsics(solver, static_cast<int>(cams.size()), camK.data()), "Optimizer (camera intrinsics)");
}
} // namespace detail


// pose <-> (qx, qy, qz, qw, tx, ty, tz) through the accessors Sophus::SE3f / Eigen offer
template <class PoseT>
void pose_to_floats(const PoseT& Tcw, float q[4], float t[3])
{
    const auto uq = Tcw.unit_quaternion();
    q[0] = uq.x(); q[1] = uq.y(); q[2] = uq.z(); q[3] = uq.w();
    const auto tr = Tcw.translation();
    for (int k = 0; k < 3; k++) t[k] = tr(k);
}

template <class PoseT>
PoseT pose_from_floats(const PoseT& like, const float q[4], const float t[3])
{
    auto uq = like.unit_quaternion();
    uq.x() = q[0]; uq.y() = q[1]; uq.z() = q[2]; uq.w() = q[3];
    auto tr = like.translation();
    for (int k = 0; k < 3; k++) tr(k) = t[k];
    return PoseT(uq, tr);
}

// ---------------------------------------------------------------------------------------------------
// int Optimizer::PoseOptimization(Frame* pFrame)                          O3/src/Optimizer.cc:744-1028
// ---------------------------------------------------------------------------------------------------
template <class FrameT, class MapPointT>
int PoseOptimization(FrameT* pFrame)
{
    const int N = pFrame->N;
    std::vector<float> Xw, xy, w;
    std::vector<int> index;         // vnIndexEdgeMono
    Xw.reserve(3 * N); xy.reserve(2 * N); w.reserve(N); index.reserve(N);
    {
        std::unique_lock<std::mutex> lock(MapPointT::mGlobalMutex);                      // :785
        for (int i = 0; i < N; i++) {
            MapPointT* pMP = pFrame->mvpMapPoints[i];
            if (!pMP) continue;
            pFrame->mvbOutlier[i] = false;                                               // :794
            const auto& kpUn = pFrame->mvKeysUn[i];
            const auto p = pMP->GetWorldPos();
            for (int k = 0; k < 3; k++) Xw.push_back(p(k));
            xy.push_back(kpUn.pt.x); xy.push_back(kpUn.pt.y);
            w.push_back(pFrame->mvInvLevelSigma2[kpUn.octave]);                          // :806
            index.push_back(i);
        }
    }
    const int n = static_cast<int>(index.size());   // nInitialCorrespondences
    if (n < 3) return 0;                                                                 // :923-924
    float q[4], t[3];
    const auto Tcw = pFrame->GetPose();
    pose_to_floats(Tcw, q, t);
    const float K[4] = { pFrame->fx, pFrame->fy, pFrame->cx, pFrame->cy };
    std::vector<uint8_t> outlier(n);
    int inliers = 0;
    check(dvm_pose_optimization(device_frame(*pFrame).frame.h, q, t, K, n, Xw.data(), xy.data(), w.data(), outlier.data(),
                                &inliers, nullptr),
          "Optimizer::PoseOptimization");
    for (int k = 0; k < n; k++) pFrame->mvbOutlier[index[k]] = outlier[k] != 0;          // :951-961
    pFrame->SetPose(pose_from_floats(Tcw, q, t));                                        // :1021-1025
    return inliers;
}

// ---------------------------------------------------------------------------------------------------
// void Optimizer::LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap, int& num_fixedKF,
//                                       int& num_OptKF, int& num_MPs, int& num_edges)   :1030-1387
// solver: one dvm_lba context per local-mapping thread (dvm_lba_create).
// ---------------------------------------------------------------------------------------------------
template <class KeyFrameT, class MapPointT, class MapT>
void LocalBundleAdjustment(dvm_lba* solver, KeyFrameT* pKF, bool* pbStopFlag, MapT* pMap, int& num_fixedKF, int& num_OptKF,
                           int& num_MPs, int& num_edges)
{
    // ---- window assembly: the reference's breadth-first search, :1033-1091 ----
    std::list<KeyFrameT*> lLocalKeyFrames;
    lLocalKeyFrames.push_back(pKF);
    pKF->mnBALocalForKF = pKF->mnId;
    auto* pCurrentMap = pKF->GetMap();
    for (KeyFrameT* pKFi : pKF->GetVectorCovisibleKeyFrames()) {
        pKFi->mnBALocalForKF = pKF->mnId;
        if (!pKFi->isBad() && pKFi->GetMap() == pCurrentMap) lLocalKeyFrames.push_back(pKFi);
    }
    num_fixedKF = 0;
    std::list<MapPointT*> lLocalMapPoints;
    for (KeyFrameT* pKFi : lLocalKeyFrames) {
        if (pKFi->mnId == pMap->GetInitKFid()) num_fixedKF = 1;
        for (MapPointT* pMP : pKFi->GetMapPointMatches())
            if (pMP && !pMP->isBad() && pMP->GetMap() == pCurrentMap && pMP->mnBALocalForKF != pKF->mnId) {
                lLocalMapPoints.push_back(pMP);
                pMP->mnBALocalForKF = pKF->mnId;
            }
    }
    std::list<KeyFrameT*> lFixedCameras;
    for (MapPointT* pMP : lLocalMapPoints)
        for (const auto& ob : pMP->GetObservations()) {
            KeyFrameT* pKFi = ob.first;
            if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) {
                pKFi->mnBAFixedForKF = pKF->mnId;
                if (!pKFi->isBad() && pKFi->GetMap() == pCurrentMap) lFixedCameras.push_back(pKFi);
            }
        }
    num_fixedKF += static_cast<int>(lFixedCameras.size());
    if (num_fixedKF == 0) return;                                                        // :1088-1091

    // ---- flatten: cameras (local first, then fixed), points, one edge per mono observation ----
    std::vector<KeyFrameT*> cams(lLocalKeyFrames.begin(), lLocalKeyFrames.end());
    cams.insert(cams.end(), lFixedCameras.begin(), lFixedCameras.end());
    const int nLocal = static_cast<int>(lLocalKeyFrames.size());
    std::unordered_map<KeyFrameT*, int> camIndex;
    std::vector<float> cam_q(cams.size() * 4), cam_t(cams.size() * 3);
    std::vector<uint8_t> cam_fixed(cams.size());
    for (size_t c = 0; c < cams.size(); c++) {
        camIndex[cams[c]] = static_cast<int>(c);
        pose_to_floats(cams[c]->GetPose(), &cam_q[4 * c], &cam_t[3 * c]);
        cam_fixed[c] = static_cast<int>(c) >= nLocal || cams[c]->mnId == pMap->GetInitKFid();   // :1124, :1140
    }
    num_OptKF = nLocal;
    std::vector<MapPointT*> pts(lLocalMapPoints.begin(), lLocalMapPoints.end());
    std::vector<float> xyz(pts.size() * 3), edge_obs, edge_w;
    std::vector<int32_t> edge_cam, edge_pt;
    std::vector<std::pair<KeyFrameT*, MapPointT*>> edge_owner;
    for (size_t j = 0; j < pts.size(); j++) {
        const auto p = pts[j]->GetWorldPos();
        for (int k = 0; k < 3; k++) xyz[3 * j + k] = p(k);
        for (const auto& ob : pts[j]->GetObservations()) {                               // :1196-1232
            KeyFrameT* pKFi = ob.first;
            if (pKFi->isBad() || pKFi->GetMap() != pCurrentMap) continue;
            const int leftIndex = std::get<0>(ob.second);
            if (leftIndex == -1 || !(pKFi->mvuRight[leftIndex] < 0)) continue;           // mono observation only
            const auto it = camIndex.find(pKFi);
            if (it == camIndex.end()) continue;
            const auto& kpUn = pKFi->mvKeysUn[leftIndex];
            edge_cam.push_back(it->second);
            edge_pt.push_back(static_cast<int32_t>(j));
            edge_obs.push_back(kpUn.pt.x); edge_obs.push_back(kpUn.pt.y);
            edge_w.push_back(pKFi->mvInvLevelSigma2[kpUn.octave]);
            edge_owner.emplace_back(pKFi, pts[j]);
        }
    }
    num_MPs = static_cast<int>(pts.size());
    num_edges = static_cast<int>(edge_cam.size());
    if (pbStopFlag && *pbStopFlag) return;                                               // :1306-1308

    // ---- optimizer.initializeOptimization(); optimizer.optimize(10);  :1310-1311 ----
    // pbStopFlag (a bool set by the tracking thread, O3/src/LocalMapping.cc:359) is polled by the library while
    // the solver runs and forwarded to the device, like g2o's forceStopFlag
    static_assert(sizeof(bool) == 1, "pbStopFlag is read as one byte");
    const float K[4] = { pKF->fx, pKF->fy, pKF->cx, pKF->cy };
    detail::set_camera_intrinsics(solver, cams);
    std::vector<uint8_t> edge_bad(edge_cam.size() ? edge_cam.size() : 1);
    int iters = 0;
    if (!detail::ba_ok(dvm_local_ba(solver, static_cast<int>(cams.size()), cam_q.data(), cam_t.data(), cam_fixed.data(),
                       static_cast<int>(pts.size()), xyz.data(), static_cast<int>(edge_cam.size()), edge_cam.data(),
                       edge_pt.data(), edge_obs.data(), edge_w.data(), K, 10,
                       reinterpret_cast<const volatile uint8_t*>(pbStopFlag), nullptr, edge_bad.data(),
                       nullptr, &iters),
          "Optimizer::LocalBundleAdjustment")) return;
    if (iters < 0) return;

    // ---- cull and write back, :1313-1387 ----
    std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
    for (size_t e = 0; e < edge_owner.size(); e++) {
        MapPointT* pMP = edge_owner[e].second;
        if (pMP->isBad() || !edge_bad[e]) continue;
        edge_owner[e].first->EraseMapPointMatch(pMP);
        pMP->EraseObservation(edge_owner[e].first);
    }
    for (int c = 0; c < nLocal; c++) cams[c]->SetPose(pose_from_floats(cams[c]->GetPose(), &cam_q[4 * c], &cam_t[3 * c]));
    for (size_t j = 0; j < pts.size(); j++) {
        auto p = pts[j]->GetWorldPos();
        for (int k = 0; k < 3; k++) p(k) = xyz[3 * j + k];
        pts[j]->SetWorldPos(p);
        pts[j]->UpdateNormalAndDepth();
    }
    pMap->IncreaseChangeIndex();
}

// ---------------------------------------------------------------------------------------------------
// void Optimizer::BundleAdjustment(const vector<KeyFrame*>& vpKFs, const vector<MapPoint*>& vpMP, int nIterations,
//                                  bool* pbStopFlag, const unsigned long nLoopKF, const bool bRobust)   :55-356
// (GlobalBundleAdjustemnt :46-53 passes every keyframe and map point of the map.)  For maps of up to the solver's
// 2000 free keyframes (dense reduced camera system; larger maps are reported and left untouched).  Mono observations only.
// ---------------------------------------------------------------------------------------------------
template <class KeyFrameT, class MapPointT>
void BundleAdjustment(dvm_lba* solver, const std::vector<KeyFrameT*>& vpKFs, const std::vector<MapPointT*>& vpMP, int nIterations,
                      bool* pbStopFlag, unsigned long nLoopKF, bool bRobust)
{
    if (vpKFs.empty()) return;
    auto* pMap = vpKFs[0]->GetMap();
    std::vector<KeyFrameT*> cams;
    std::unordered_map<KeyFrameT*, int> camIndex;
    unsigned long maxKFid = 0;
    for (KeyFrameT* pKF : vpKFs) {
        if (pKF->isBad()) continue;                                                       // :105-106
        camIndex[pKF] = static_cast<int>(cams.size());
        cams.push_back(pKF);
        if (pKF->mnId > maxKFid) maxKFid = pKF->mnId;
    }
    std::vector<float> cam_q(cams.size() * 4), cam_t(cams.size() * 3);
    std::vector<uint8_t> cam_fixed(cams.size());
    for (size_t c = 0; c < cams.size(); c++) {
        pose_to_floats(cams[c]->GetPose(), &cam_q[4 * c], &cam_t[3 * c]);
        cam_fixed[c] = cams[c]->mnId == pMap->GetInitKFid();                              // :116
    }
    std::vector<MapPointT*> pts;
    std::vector<float> xyz, edge_obs, edge_w;
    std::vector<int32_t> edge_cam, edge_pt;
    for (MapPointT* pMP : vpMP) {
        if (pMP->isBad()) continue;                                                       // :126-127
        const size_t first_edge = edge_cam.size();
        for (const auto& ob : pMP->GetObservations()) {
            KeyFrameT* pKF = ob.first;
            if (pKF->isBad() || pKF->mnId > maxKFid) continue;                            // :141-142
            const auto it = camIndex.find(pKF);
            if (it == camIndex.end()) continue;                                           // optimizer.vertex(pKF->mnId) == NULL
            const int leftIndex = std::get<0>(ob.second);
            if (leftIndex == -1 || !(pKF->mvuRight[leftIndex] < 0)) continue;             // mono observation only
            const auto& kpUn = pKF->mvKeysUn[leftIndex];
            edge_cam.push_back(it->second);
            edge_pt.push_back(static_cast<int32_t>(pts.size()));
            edge_obs.push_back(kpUn.pt.x); edge_obs.push_back(kpUn.pt.y);
            edge_w.push_back(pKF->mvInvLevelSigma2[kpUn.octave]);
        }
        if (edge_cam.size() == first_edge) continue;                                      // nEdges == 0: vertex removed, :243-247
        const auto p = pMP->GetWorldPos();
        for (int k = 0; k < 3; k++) xyz.push_back(p(k));
        pts.push_back(pMP);
    }
    if (edge_cam.empty()) return;
    const float K[4] = { cams[0]->fx, cams[0]->fy, cams[0]->cx, cams[0]->cy };
    detail::set_camera_intrinsics(solver, cams);
    // const float thHuber2D = sqrt(5.99), :122 -- not LocalBundleAdjustment's sqrt(5.991)
    const float delta = bRobust ? static_cast<float>(std::sqrt(5.99)) : std::numeric_limits<float>::infinity();
    std::vector<uint8_t> edge_bad(edge_cam.size());
    int iters = 0;
    if (!detail::ba_ok(dvm_bundle_adjustment(solver, static_cast<int>(cams.size()), cam_q.data(), cam_t.data(), cam_fixed.data(),
                                static_cast<int>(pts.size()), xyz.data(), static_cast<int>(edge_cam.size()), edge_cam.data(),
                                edge_pt.data(), edge_obs.data(), edge_w.data(), K, nIterations, delta,
                                reinterpret_cast<const volatile uint8_t*>(pbStopFlag), nullptr, edge_bad.data(), nullptr, &iters),
          "Optimizer::BundleAdjustment")) return;
    if (iters < 0) return;
    // recover optimised data, :250-356: the loop keyframe's own map gets the values directly, otherwise they are parked
    // in mTcwGBA / mPosGBA for the caller's propagation
    const bool direct = nLoopKF == pMap->GetOriginKF()->mnId;
    for (size_t c = 0; c < cams.size(); c++) {
        const auto T = pose_from_floats(cams[c]->GetPose(), &cam_q[4 * c], &cam_t[3 * c]);
        if (direct) cams[c]->SetPose(T);
        else { cams[c]->mTcwGBA = T; cams[c]->mnBAGlobalForKF = nLoopKF; }
    }
    for (size_t j = 0; j < pts.size(); j++) {
        auto p = pts[j]->GetWorldPos();
        for (int k = 0; k < 3; k++) p(k) = xyz[3 * j + k];
        if (direct) { pts[j]->SetWorldPos(p); pts[j]->UpdateNormalAndDepth(); }
        else { pts[j]->mPosGBA = p; pts[j]->mnBAGlobalForKF = nLoopKF; }
    }
}

// ---------------------------------------------------------------------------------------------------
// void Optimizer::LocalBundleAdjustment(KeyFrame* pMainKF, vector<KeyFrame*> vpAdjustKF, vector<KeyFrame*> vpFixedKF,
//                                       bool* pbStopFlag)   :3257-3675
// The welding BA of a map merge (LoopClosing::MergeLocal).  Mono observations only; both optimisation passes run in one
// dvm_merge_ba call.  Map points without any edge stay out of the solver (g2o never activates such a vertex) but are
// still written back, as the reference does.
// ---------------------------------------------------------------------------------------------------
template <class KeyFrameT, class MapPointT>
void LocalBundleAdjustment(dvm_lba* solver, KeyFrameT* pMainKF, std::vector<KeyFrameT*> vpAdjustKF, std::vector<KeyFrameT*> vpFixedKF,
                           bool* pbStopFlag)
{
    auto* pCurrentMap = pMainKF->GetMap();
    std::vector<KeyFrameT*> cams;
    std::vector<uint8_t> cam_fixed;
    std::unordered_map<KeyFrameT*, int> camIndex;
    std::vector<MapPointT*> vpMPs;
    unsigned long maxKFid = 0;
    auto add_keyframe = [&](KeyFrameT* pKFi, bool fixed) {                                // :3283-3348
        pKFi->mnBALocalForMerge = pMainKF->mnId;
        camIndex[pKFi] = static_cast<int>(cams.size());
        cams.push_back(pKFi);
        cam_fixed.push_back(fixed ? 1 : 0);
        if (pKFi->mnId > maxKFid) maxKFid = pKFi->mnId;
        for (MapPointT* pMPi : pKFi->GetMapPoints())
            if (pMPi && !pMPi->isBad() && pMPi->GetMap() == pCurrentMap && pMPi->mnBALocalForMerge != pMainKF->mnId) {
                vpMPs.push_back(pMPi);
                pMPi->mnBALocalForMerge = pMainKF->mnId;
            }
    };
    for (KeyFrameT* pKFi : vpFixedKF)
        if (!pKFi->isBad() && pKFi->GetMap() == pCurrentMap) add_keyframe(pKFi, true);
    const size_t nFixed = cams.size();
    for (KeyFrameT* pKFi : vpAdjustKF)
        if (!pKFi->isBad() && pKFi->GetMap() == pCurrentMap) add_keyframe(pKFi, false);
    std::vector<float> cam_q(cams.size() * 4), cam_t(cams.size() * 3);
    for (size_t c = 0; c < cams.size(); c++) pose_to_floats(cams[c]->GetPose(), &cam_q[4 * c], &cam_t[3 * c]);

    std::vector<MapPointT*> pts;                       // the points that got at least one edge
    std::vector<float> xyz, edge_obs, edge_w;
    std::vector<int32_t> edge_cam, edge_pt;
    std::vector<std::pair<KeyFrameT*, MapPointT*>> edge_owner;
    for (MapPointT* pMPi : vpMPs) {                                                       // :3370-3466
        if (pMPi->isBad()) continue;
        const size_t first_edge = edge_cam.size();
        for (const auto& ob : pMPi->GetObservations()) {
            KeyFrameT* pKF = ob.first;
            const int leftIndex = std::get<0>(ob.second);
            const auto it = camIndex.find(pKF);
            if (pKF->isBad() || pKF->mnId > maxKFid || it == camIndex.end() || !pKF->GetMapPoint(leftIndex)) continue;
            if (!(pKF->mvuRight[leftIndex] < 0)) continue;                                // mono observation only
            const auto& kpUn = pKF->mvKeysUn[leftIndex];
            edge_cam.push_back(it->second);
            edge_pt.push_back(static_cast<int32_t>(pts.size()));
            edge_obs.push_back(kpUn.pt.x); edge_obs.push_back(kpUn.pt.y);
            edge_w.push_back(pKF->mvInvLevelSigma2[kpUn.octave]);
            edge_owner.emplace_back(pKF, pMPi);
        }
        if (edge_cam.size() == first_edge) continue;
        const auto p = pMPi->GetWorldPos();
        for (int k = 0; k < 3; k++) xyz.push_back(p(k));
        pts.push_back(pMPi);
    }
    if (pbStopFlag && *pbStopFlag) return;                                               // :3468-3470
    if (edge_cam.empty() || nFixed == 0) return;       // nothing to solve / no gauge: the device solver needs a fixed keyframe

    static_assert(sizeof(bool) == 1, "pbStopFlag is read as one byte");
    const float K[4] = { pMainKF->fx, pMainKF->fy, pMainKF->cx, pMainKF->cy };
    detail::set_camera_intrinsics(solver, cams);
    std::vector<uint8_t> edge_bad(edge_cam.size());
    int iters = 0;
    if (!detail::ba_ok(dvm_merge_ba(solver, static_cast<int>(cams.size()), cam_q.data(), cam_t.data(), cam_fixed.data(),
                       static_cast<int>(pts.size()), xyz.data(), static_cast<int>(edge_cam.size()), edge_cam.data(),
                       edge_pt.data(), edge_obs.data(), edge_w.data(), K, reinterpret_cast<const volatile uint8_t*>(pbStopFlag),
                       nullptr, edge_bad.data(), nullptr, &iters),
          "Optimizer::LocalBundleAdjustment (welding)")) return;
    if (iters < 0) return;

    // ---- erase the flagged observations and write back, :3523-3674 ----
    std::unique_lock<std::mutex> lock(pMainKF->GetMap()->mMutexMapUpdate);
    for (size_t e = 0; e < edge_owner.size(); e++) {
        MapPointT* pMP = edge_owner[e].second;
        if (pMP->isBad() || !edge_bad[e]) continue;
        edge_owner[e].first->EraseMapPointMatch(pMP);
        pMP->EraseObservation(edge_owner[e].first);
    }
    for (KeyFrameT* pKFi : vpAdjustKF) {
        if (pKFi->isBad()) continue;
        const auto it = camIndex.find(pKFi);
        if (it == camIndex.end()) continue;
        pKFi->SetPose(pose_from_floats(pKFi->GetPose(), &cam_q[4 * it->second], &cam_t[3 * it->second]));
    }
    for (size_t j = 0; j < pts.size(); j++) {
        if (pts[j]->isBad()) continue;
        auto p = pts[j]->GetWorldPos();
        for (int k = 0; k < 3; k++) p(k) = xyz[3 * j + k];
        pts[j]->SetWorldPos(p);
    }
    for (MapPointT* pMPi : vpMPs)
        if (!pMPi->isBad()) pMPi->UpdateNormalAndDepth();
}

// ---------------------------------------------------------------------------------------------------
// int Optimizer::OptimizeSim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches1, g2o::Sim3& g2oS12,
//                             const float th2, const bool bFixScale, Eigen::Matrix<double,7,7>& mAcumHessian,
//                             const bool bAllPoints)   :1960-2212
// solver: one dvm_sim3 context per loop-closing thread (dvm_sim3_create).  Sim3T is g2o::Sim3 (rotation(), translation(),
// scale() with mutable references), Mat77T an Eigen 7x7 (setZero()).  Pinhole cameras (KeyFrame::fx..cy).
// ---------------------------------------------------------------------------------------------------
template <class KeyFrameT, class MapPointT, class Sim3T, class Mat77T>
int OptimizeSim3(dvm_sim3* solver, KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches1, Sim3T& g2oS12,
                 const float th2, const bool bFixScale, Mat77T& mAcumHessian, const bool bAllPoints)
{
    const auto R1w = pKF1->GetRotation();
    const auto t1w = pKF1->GetTranslation();
    const auto R2w = pKF2->GetRotation();
    const auto t2w = pKF2->GetTranslation();
    // Eigen's fixed-size float product R * p + t, coefficient by coefficient
    auto to_camera = [](const decltype(R1w)& R, const decltype(t1w)& t, const decltype(t1w)& p, float out[3]) {
        for (int r = 0; r < 3; r++) out[r] = (R(r, 0) * p(0) + R(r, 1) * p(1) + R(r, 2) * p(2)) + t(r);
    };
    const int N = static_cast<int>(vpMatches1.size());
    const std::vector<MapPointT*> vpMapPoints1 = pKF1->GetMapPointMatches();
    std::vector<float> p1c, p2c, obs1, obs2, w1, w2;
    std::vector<int> index;
    for (int i = 0; i < N; i++) {                                                         // :2008-2146
        if (!vpMatches1[i]) continue;
        MapPointT* pMP1 = vpMapPoints1[i];
        MapPointT* pMP2 = vpMatches1[i];
        const int i2 = std::get<0>(pMP2->GetIndexInKeyFrame(pKF2));
        if (!pMP1 || !pMP2) continue;                   // the reference only adds an unused fixed vertex here (:2045-2060)
        if (pMP1->isBad() || pMP2->isBad()) continue;                                     // nBadMPs
        float P3D1c[3], P3D2c[3];
        to_camera(R1w, t1w, pMP1->GetWorldPos(), P3D1c);
        to_camera(R2w, t2w, pMP2->GetWorldPos(), P3D2c);
        if (i2 < 0 && !bAllPoints) continue;
        if (P3D2c[2] < 0) continue;
        const auto& kpUn1 = pKF1->mvKeysUn[i];
        float o2x, o2y;
        int octave2;
        if (i2 >= 0) {
            const auto& kpUn2 = pKF2->mvKeysUn[i2];
            o2x = kpUn2.pt.x; o2y = kpUn2.pt.y; octave2 = kpUn2.octave;
        } else {   // cv::KeyPoint(cv::Point2f(x, y), pMP2->mnTrackScaleLevel): that argument is the SIZE, the octave stays 0
            const float invz = 1 / P3D2c[2];
            o2x = P3D2c[0] * invz; o2y = P3D2c[1] * invz; octave2 = 0;
        }
        for (int k = 0; k < 3; k++) { p1c.push_back(P3D1c[k]); p2c.push_back(P3D2c[k]); }
        obs1.push_back(kpUn1.pt.x); obs1.push_back(kpUn1.pt.y);
        obs2.push_back(o2x); obs2.push_back(o2y);
        w1.push_back(pKF1->mvInvLevelSigma2[kpUn1.octave]);
        w2.push_back(pKF2->mvInvLevelSigma2[octave2]);
        index.push_back(i);
    }
    const float K1[4] = { pKF1->fx, pKF1->fy, pKF1->cx, pKF1->cy };
    const float K2[4] = { pKF2->fx, pKF2->fy, pKF2->cx, pKF2->cy };
    double q[4] = { g2oS12.rotation().x(), g2oS12.rotation().y(), g2oS12.rotation().z(), g2oS12.rotation().w() };
    double t[3] = { g2oS12.translation()(0), g2oS12.translation()(1), g2oS12.translation()(2) };
    double s = g2oS12.scale();
    std::vector<uint8_t> inlier(index.size() ? index.size() : 1);
    int nIn = 0;
    double stats[6] = { 0, 0, 0, 0, 0, 0 };
    check(dvm_optimize_sim3(solver, static_cast<int>(index.size()), p1c.data(), p2c.data(), obs1.data(), obs2.data(), w1.data(),
                            w2.data(), K1, K2, q, t, &s, th2, bFixScale ? 1 : 0, inlier.data(), &nIn, stats),
          "Optimizer::OptimizeSim3");
    for (size_t e = 0; e < index.size(); e++)
        if (!inlier[e]) vpMatches1[index[e]] = static_cast<MapPointT*>(nullptr);          // :2160, :2202
    if (static_cast<int>(index.size()) - static_cast<int>(stats[3]) < 10) return 0;   // nCorrespondences - nBad < 10: g2oS12 untouched (:2183)
    mAcumHessian.setZero();                                                               // :2191
    g2oS12.rotation().x() = q[0]; g2oS12.rotation().y() = q[1]; g2oS12.rotation().z() = q[2]; g2oS12.rotation().w() = q[3];
    for (int k = 0; k < 3; k++) g2oS12.translation()(k) = t[k];
    g2oS12.scale() = s;
    return nIn;
}

// ---------------------------------------------------------------------------------------------------
// void Optimizer::OptimizeEssentialGraph(Map* pMap, KeyFrame* pLoopKF, KeyFrame* pCurKF,
//        const LoopClosing::KeyFrameAndPose& NonCorrectedSim3, const LoopClosing::KeyFrameAndPose& CorrectedSim3,
//        const map<KeyFrame*, set<KeyFrame*>>& LoopConnections, const bool& bFixScale)                     :1389-1651
// The graph is flattened exactly as the reference assembles it (vertices :1419-1452; loop-connection edges :1460-1488;
// per keyframe the spanning-tree edge :1507-1528, earlier loop edges :1530-1552, covisibility >= 100 edges :1554-1585;
// inertial edges are out of scope with the rest of the IMU path), solved on the device, and written back: keyframe poses
// Sim3 -> SE3 [R, t / s] (:1602-1615) and map points through their reference keyframe (:1619-1647).
// solver: one dvm_essential_graph context per loop-closing thread.  KeyFrameAndPoseT = map<KeyFrame*, g2o::Sim3>.
// ---------------------------------------------------------------------------------------------------
namespace detail {
struct S3 {   // g2o::Sim3 arithmetic in double on the host (g2o/types/sim3.h: operator*, inverse, map)
    double q[4], t[3], s;
    static void rot(const double q[4], const double v[3], double o[3])
    {
        double uv[3] = { q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0] };
        uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
        o[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
        o[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
        o[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
    }
    S3 operator*(const S3& b) const
    {
        S3 r;
        r.q[3] = q[3] * b.q[3] - q[0] * b.q[0] - q[1] * b.q[1] - q[2] * b.q[2];
        r.q[0] = q[3] * b.q[0] + q[0] * b.q[3] + q[1] * b.q[2] - q[2] * b.q[1];
        r.q[1] = q[3] * b.q[1] + q[1] * b.q[3] + q[2] * b.q[0] - q[0] * b.q[2];
        r.q[2] = q[3] * b.q[2] + q[2] * b.q[3] + q[0] * b.q[1] - q[1] * b.q[0];
        double rt[3];
        rot(q, b.t, rt);
        for (int i = 0; i < 3; i++) r.t[i] = s * rt[i] + t[i];
        r.s = s * b.s;
        return r;
    }
    S3 inverse() const
    {
        S3 r;
        r.q[0] = -q[0]; r.q[1] = -q[1]; r.q[2] = -q[2]; r.q[3] = q[3];
        const double v[3] = { (-1. / s) * t[0], (-1. / s) * t[1], (-1. / s) * t[2] };
        rot(r.q, v, r.t);
        r.s = 1. / s;
        return r;
    }
    void map(const double p[3], double o[3]) const
    {
        double rp[3];
        rot(q, p, rp);
        for (int i = 0; i < 3; i++) o[i] = s * rp[i] + t[i];
    }
    template <class Sim3T> static S3 from(const Sim3T& g)
    {
        S3 r;
        Sim3T& m = const_cast<Sim3T&>(g);
        r.q[0] = m.rotation().x(); r.q[1] = m.rotation().y(); r.q[2] = m.rotation().z(); r.q[3] = m.rotation().w();
        for (int i = 0; i < 3; i++) r.t[i] = m.translation()(i);
        r.s = m.scale();
        return r;
    }
};
} // namespace detail

template <class MapT, class KeyFrameT, class KeyFrameAndPoseT>
void OptimizeEssentialGraph(dvm_essential_graph* solver, MapT* pMap, KeyFrameT* pLoopKF, KeyFrameT* pCurKF,
                            const KeyFrameAndPoseT& NonCorrectedSim3, const KeyFrameAndPoseT& CorrectedSim3,
                            const std::map<KeyFrameT*, std::set<KeyFrameT*>>& LoopConnections, const bool& bFixScale)
{
    using detail::S3;
    const std::vector<KeyFrameT*> vpKFs = pMap->GetAllKeyFrames();
    const auto vpMPs = pMap->GetAllMapPoints();
    const unsigned int nMaxKFid = pMap->GetMaxKFid();
    std::vector<S3> vScw(nMaxKFid + 1), vCorrectedSwc(nMaxKFid + 1);
    std::vector<int> vertex_of(nMaxKFid + 1, -1);     // keyframe id -> vertex index of the flattened graph
    const int minFeat = 100;
    std::vector<double> sim3;
    std::vector<uint8_t> fixed;
    std::vector<KeyFrameT*> vertex_kf;
    for (KeyFrameT* pKF : vpKFs) {                                                        // :1419-1452
        if (pKF->isBad()) continue;
        const int nIDi = static_cast<int>(pKF->mnId);
        const auto it = CorrectedSim3.find(pKF);
        S3 Siw;
        if (it != CorrectedSim3.end()) Siw = S3::from(it->second);
        else {   // g2o::Sim3 Siw(Tcw.unit_quaternion(), Tcw.translation(), 1.0) of the pose cast to double
            const auto Tcw = pKF->GetPose();
            const auto uq = Tcw.unit_quaternion();
            const auto tr = Tcw.translation();
            Siw.q[0] = uq.x(); Siw.q[1] = uq.y(); Siw.q[2] = uq.z(); Siw.q[3] = uq.w();
            for (int k = 0; k < 3; k++) Siw.t[k] = tr(k);
            Siw.s = 1.0;
        }
        vScw[nIDi] = Siw;
        vertex_of[nIDi] = static_cast<int>(vertex_kf.size());
        vertex_kf.push_back(pKF);
        for (int k = 0; k < 4; k++) sim3.push_back(Siw.q[k]);
        for (int k = 0; k < 3; k++) sim3.push_back(Siw.t[k]);
        sim3.push_back(Siw.s);
        fixed.push_back(pKF->mnId == pMap->GetInitKFid());
    }
    std::vector<int32_t> vi, vj;
    std::vector<double> meas;
    auto add_edge = [&](long unsigned int idi, long unsigned int idj, const S3& Sji) {     // vertex(0) = i, vertex(1) = j
        if (vertex_of[idi] < 0 || vertex_of[idj] < 0) return;   // optimizer.vertex(id) of a bad keyframe is null in the reference
        vi.push_back(vertex_of[idi]); vj.push_back(vertex_of[idj]);
        for (int k = 0; k < 4; k++) meas.push_back(Sji.q[k]);
        for (int k = 0; k < 3; k++) meas.push_back(Sji.t[k]);
        meas.push_back(Sji.s);
    };
    std::set<std::pair<long unsigned int, long unsigned int>> sInsertedEdges;
    for (const auto& mit : LoopConnections) {                                             // :1460-1488
        KeyFrameT* pKF = mit.first;
        const long unsigned int nIDi = pKF->mnId;
        const S3 Swi = vScw[nIDi].inverse();
        for (KeyFrameT* pKFj : mit.second) {
            const long unsigned int nIDj = pKFj->mnId;
            if ((nIDi != pCurKF->mnId || nIDj != pLoopKF->mnId) && pKF->GetWeight(pKFj) < minFeat) continue;
            add_edge(nIDi, nIDj, vScw[nIDj] * Swi);
            sInsertedEdges.insert(std::make_pair(std::min(nIDi, nIDj), std::max(nIDi, nIDj)));
        }
    }
    auto non_corrected = [&](KeyFrameT* k) {   // Sjw = NonCorrectedSim3[k] if present, else vScw[k]
        const auto it = NonCorrectedSim3.find(k);
        return it != NonCorrectedSim3.end() ? S3::from(it->second) : vScw[k->mnId];
    };
    for (KeyFrameT* pKF : vpKFs) {                                                        // :1491-1592
        const long unsigned int nIDi = pKF->mnId;
        const S3 Swi = non_corrected(pKF).inverse();
        KeyFrameT* pParentKF = pKF->GetParent();
        if (pParentKF) add_edge(nIDi, pParentKF->mnId, non_corrected(pParentKF) * Swi);  // spanning tree
        for (KeyFrameT* pLKF : pKF->GetLoopEdges())                                       // earlier loop edges
            if (pLKF->mnId < pKF->mnId) add_edge(nIDi, pLKF->mnId, non_corrected(pLKF) * Swi);
        for (KeyFrameT* pKFn : pKF->GetCovisiblesByWeight(minFeat)) {                     // covisibility graph
            if (pKFn && pKFn != pParentKF && !pKF->hasChild(pKFn)) {
                if (!pKFn->isBad() && pKFn->mnId < pKF->mnId) {
                    if (sInsertedEdges.count(std::make_pair(std::min(pKF->mnId, pKFn->mnId), std::max(pKF->mnId, pKFn->mnId)))) continue;
                    add_edge(nIDi, pKFn->mnId, non_corrected(pKFn) * Swi);
                }
            }
        }
    }
    double stats[4];
    check(dvm_optimize_essential_graph(solver, static_cast<int>(vertex_kf.size()), sim3.data(), fixed.data(), static_cast<int>(vi.size()),
                                       vi.data(), vj.data(), meas.data(), bFixScale ? 1 : 0, 20, 1e-16, stats),
          "Optimizer::OptimizeEssentialGraph");
    std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
    for (size_t v = 0; v < vertex_kf.size(); v++) {                                       // :1602-1615
        KeyFrameT* pKFi = vertex_kf[v];
        S3 C;
        for (int k = 0; k < 4; k++) C.q[k] = sim3[8 * v + k];
        for (int k = 0; k < 3; k++) C.t[k] = sim3[8 * v + 4 + k];
        C.s = sim3[8 * v + 7];
        vCorrectedSwc[pKFi->mnId] = C.inverse();
        // Sophus::SE3f Tiw(CorrectedSiw.rotation().cast<float>(), CorrectedSiw.translation().cast<float>() / s)
        float qf[4], tf[3];
        for (int k = 0; k < 4; k++) qf[k] = static_cast<float>(C.q[k]);
        for (int k = 0; k < 3; k++) tf[k] = static_cast<float>(C.t[k]) / static_cast<float>(C.s);   // Vector3f / scalar
        pKFi->SetPose(pose_from_floats(pKFi->GetPose(), qf, tf));
    }
    for (auto* pMP : vpMPs) {                                                             // :1619-1647
        if (pMP->isBad()) continue;
        const long unsigned int nIDr = pMP->mnCorrectedByKF == pCurKF->mnId ? pMP->mnCorrectedReference
                                                                               : pMP->GetReferenceKeyFrame()->mnId;
        const auto p = pMP->GetWorldPos();
        const double P[3] = { p(0), p(1), p(2) };
        double a[3], b[3];
        vScw[nIDr].map(P, a);
        vCorrectedSwc[nIDr].map(a, b);
        auto np = p;
        for (int k = 0; k < 3; k++) np(k) = static_cast<float>(b[k]);
        pMP->SetWorldPos(np);
        pMP->UpdateNormalAndDepth();
    }
    pMap->IncreaseChangeIndex();
}

} // namespace dvm_host
