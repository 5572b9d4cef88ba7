/* oracle/g2oshim/ref_opt_glue.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Flat C entry points around the reference's OWN optimisation functions: Optimizer::PoseOptimization,
 * LocalBundleAdjustment (local mapping and the welding BA of a merge), BundleAdjustment, OptimizeSim3 and
 * OptimizeEssentialGraph (bodies cut out of O3/src/Optimizer.cc at build time, oracle/_ref/gen/opt_bodies.inc), running on
 * the reference's vendored g2o compiled where it lies, over the mini Eigen of this directory.  Each entry builds the
 * Frame / KeyFrame / MapPoint / Map objects (oracle/g2oshim/opt_types.h) that make the reference assemble exactly the flat
 * problem the oracle restatements (oracle/lba_oracle.cpp, sim3_oracle.cpp, track_oracle.cpp) take, calls the reference
 * function, and flattens what it wrote back.  tests/test_ref_optimizer.py compares the two. */
#include "opt_types.h"

#include "Thirdparty/g2o/g2o/core/block_solver.h"
#include "Thirdparty/g2o/g2o/core/optimization_algorithm_gauss_newton.h"
#include "Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h"
#include "Thirdparty/g2o/g2o/core/robust_kernel_impl.h"
#include "Thirdparty/g2o/g2o/core/sparse_block_matrix.h"
#include "Thirdparty/g2o/g2o/solvers/linear_solver_dense.h"
#include "Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h"
#include "Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h"
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.h"

#include "OptimizableTypes.h"

#include <algorithm>
#include <cstdint>
#include <memory>

namespace ORB_SLAM3 {
std::mutex MapPoint::mGlobalMutex;
bool sortByVal(const pair<MapPoint*, int>& a, const pair<MapPoint*, int>& b) { return (a.second < b.second); }   /* Optimizer.cc:46 */
} // namespace ORB_SLAM3

#include "../_ref/gen/opt_bodies.inc"

using namespace ORB_SLAM3;

namespace {

Sophus::SE3f pose_of(const float* q, const float* t)
{   /* q = (x, y, z, w) */
    return Sophus::SE3f(Eigen::Quaternionf(q[3], q[0], q[1], q[2]), Eigen::Vector3f(t[0], t[1], t[2]));
}
void pose_out(const Sophus::SE3f& T, float* q, float* t)
{
    const Eigen::Quaternionf& r = T.unit_quaternion();
    q[0] = r.x(); q[1] = r.y(); q[2] = r.z(); q[3] = r.w();
    for (int i = 0; i < 3; i++) t[i] = T.translation()[i];
}

/* a bundle-adjustment scene as objects: one KeyFrame per camera, one MapPoint per point, one keypoint per edge (its
 * octave indexes a per-keyframe mvInvLevelSigma2 table that holds the edge's own weight) */
struct Scene {
    Map map;
    std::vector<std::unique_ptr<KeyFrame>> kfs;
    std::vector<std::unique_ptr<MapPoint>> mps;
    std::vector<std::unique_ptr<Pinhole>> cams;
    std::vector<int> edge_kp;   /* keypoint index of edge e inside its keyframe */
    Scene(int nc, const float* cam_q, const float* cam_t, const float* cam_K, int np, const float* pts, int ne, const int* edge_cam,
          const int* edge_pt, const float* edge_obs, const float* edge_w)
    {
        map.mnInitKFid = 0xfffffffful;   /* no keyframe is the map's first one */
        for (int c = 0; c < nc; c++) {
            kfs.emplace_back(new KeyFrame);
            cams.emplace_back(new Pinhole(cam_K[4 * c], cam_K[4 * c + 1], cam_K[4 * c + 2], cam_K[4 * c + 3]));
            KeyFrame* kf = kfs.back().get();
            kf->mnId = (unsigned long)c + 1;   /* ids from 1: the marks (mnBALocalForKF ...) start at 0 */
            kf->mpMap = &map;
            kf->mpCamera = cams.back().get();
            kf->fx = cam_K[4 * c]; kf->fy = cam_K[4 * c + 1]; kf->cx = cam_K[4 * c + 2]; kf->cy = cam_K[4 * c + 3];
            kf->SetPose(pose_of(cam_q + 4 * c, cam_t + 3 * c));
            map.mvpKeyFrames.push_back(kf);
        }
        map.mnMaxKFid = (unsigned long)nc;
        for (int p = 0; p < np; p++) {
            mps.emplace_back(new MapPoint);
            MapPoint* mp = mps.back().get();
            mp->mnId = (unsigned long)p;
            mp->mpMap = &map;
            mp->SetWorldPos(Eigen::Vector3f(pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]));
            map.mvpMapPoints.push_back(mp);
        }
        edge_kp.resize((size_t)ne);
        for (int e = 0; e < ne; e++) {
            KeyFrame* kf = kfs[(size_t)edge_cam[e]].get();
            MapPoint* mp = mps[(size_t)edge_pt[e]].get();
            const int idx = kf->N++;
            cv::KeyPoint kp;
            kp.pt.x = edge_obs[2 * e]; kp.pt.y = edge_obs[2 * e + 1];
            kp.octave = idx;
            kf->mvKeysUn.push_back(kp);
            kf->mvKeys.push_back(kp);
            kf->mvuRight.push_back(-1.f);
            kf->mvInvLevelSigma2.push_back(edge_w[e]);
            kf->mvpMapPoints.push_back(mp);
            mp->mObservations[kf] = std::tuple<int, int>(idx, -1);
            edge_kp[(size_t)e] = idx;
        }
    }
};

} // namespace

extern "C" {

/* Optimizer::PoseOptimization (O3/src/Optimizer.cc:744-1028): n correspondences; pose = (qx, qy, qz, qw, tx, ty, tz) in / out;
 * returns nInitialCorrespondences - nBad */
int refopt_pose_optimization(int n, const float* Xw, const float* kp_xy, const float* inv_sigma2, const float* K, float* pose,
                             uint8_t* outlier)
{
    Frame F;
    Pinhole cam(K[0], K[1], K[2], K[3]);
    std::vector<MapPoint> mps((size_t)n);
    F.N = n;
    F.mpCamera = &cam;
    F.fx = K[0]; F.fy = K[1]; F.cx = K[2]; F.cy = K[3];
    F.mvKeysUn.resize((size_t)n); F.mvKeys.resize((size_t)n); F.mvuRight.assign((size_t)n, -1.f);
    F.mvpMapPoints.resize((size_t)n); F.mvbOutlier.assign((size_t)n, false);
    F.mvInvLevelSigma2.resize((size_t)n);
    for (int i = 0; i < n; i++) {
        F.mvKeysUn[(size_t)i].pt.x = kp_xy[2 * i]; F.mvKeysUn[(size_t)i].pt.y = kp_xy[2 * i + 1];
        F.mvKeysUn[(size_t)i].octave = i;
        F.mvInvLevelSigma2[(size_t)i] = inv_sigma2[i];
        mps[(size_t)i].SetWorldPos(Eigen::Vector3f(Xw[3 * i], Xw[3 * i + 1], Xw[3 * i + 2]));
        F.mvpMapPoints[(size_t)i] = &mps[(size_t)i];
    }
    F.SetPose(pose_of(pose, pose + 4));
    const int r = Optimizer::PoseOptimization(&F);
    pose_out(F.GetPose(), pose, pose + 4);
    for (int i = 0; i < n; i++) outlier[i] = F.mvbOutlier[(size_t)i] ? 1 : 0;
    return r;
}

/* Optimizer::LocalBundleAdjustment(pKF, pbStopFlag, pMap, ...) (O3/src/Optimizer.cc:1030-1387).  The current keyframe is the
 * first free camera, its covisible keyframes are the other free cameras; the fixed cameras are found by the reference itself
 * through the observations.  bad[e] = the observation was erased; stats = num_fixedKF, num_OptKF, num_MPs, num_edges */
void refopt_local_ba(int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, const float* cam_K, int np, float* pts, int ne,
                     const int* edge_cam, const int* edge_pt, const float* edge_obs, const float* edge_w, int abort_flag, uint8_t* bad,
                     int* stats)
{
    Scene S(nc, cam_q, cam_t, cam_K, np, pts, ne, edge_cam, edge_pt, edge_obs, edge_w);
    KeyFrame* cur = nullptr;
    for (int c = 0; c < nc; c++)
        if (!cam_fixed[c]) {
            if (!cur) cur = S.kfs[(size_t)c].get();
            else cur->mvpOrderedConnectedKeyFrames.push_back(S.kfs[(size_t)c].get());
        }
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    for (int e = 0; e < ne; e++) bad[e] = 0;
    if (!cur) return;
    bool stop = abort_flag != 0;
    Optimizer::LocalBundleAdjustment(cur, &stop, &S.map, stats[0], stats[1], stats[2], stats[3]);
    for (int c = 0; c < nc; c++) pose_out(S.kfs[(size_t)c]->GetPose(), cam_q + 4 * c, cam_t + 3 * c);
    for (int p = 0; p < np; p++) for (int k = 0; k < 3; k++) pts[3 * p + k] = S.mps[(size_t)p]->GetWorldPos()[k];
    for (int e = 0; e < ne; e++)
        bad[e] = S.mps[(size_t)edge_pt[e]]->mObservations.count(S.kfs[(size_t)edge_cam[e]].get()) ? 0 : 1;
}

/* Optimizer::BundleAdjustment (O3/src/Optimizer.cc:55-356) over all cameras / points; camera 0 must be the map's first
 * keyframe (the only fixed vertex, :88).  Results as the reference leaves them for a loop keyframe other than the origin:
 * mTcwGBA / mPosGBA */
void refopt_bundle_adjustment(int nc, float* cam_q, float* cam_t, const float* cam_K, int np, float* pts, int ne, const int* edge_cam,
                              const int* edge_pt, const float* edge_obs, const float* edge_w, int iterations, int robust)
{
    Scene S(nc, cam_q, cam_t, cam_K, np, pts, ne, edge_cam, edge_pt, edge_obs, edge_w);
    S.map.mnInitKFid = 1;
    S.map.mvpKeyFrameOrigins.push_back(S.kfs[0].get());
    std::vector<KeyFrame*> kfs;
    std::vector<MapPoint*> mps;
    for (auto& k : S.kfs) kfs.push_back(k.get());
    for (auto& m : S.mps) mps.push_back(m.get());
    for (auto& k : S.kfs) k->mTcwGBA = k->GetPose();
    for (auto& m : S.mps) m->mPosGBA = m->GetWorldPos();
    Optimizer::BundleAdjustment(kfs, mps, iterations, nullptr, 12345ul, robust != 0);
    for (int c = 0; c < nc; c++) pose_out(S.kfs[(size_t)c]->mTcwGBA, cam_q + 4 * c, cam_t + 3 * c);
    for (int p = 0; p < np; p++) for (int k = 0; k < 3; k++) pts[3 * p + k] = S.mps[(size_t)p]->mPosGBA[k];
}

/* Optimizer::LocalBundleAdjustment(pMainKF, vpAdjustKF, vpFixedKF, pbStopFlag) (O3/src/Optimizer.cc:3257-3675), the welding BA
 * of a map merge: the free cameras are vpAdjustKF (the first one is pMainKF), the fixed ones vpFixedKF */
void refopt_merge_ba(int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, const float* cam_K, int np, float* pts, int ne,
                     const int* edge_cam, const int* edge_pt, const float* edge_obs, const float* edge_w, int abort_flag, uint8_t* bad)
{
    Scene S(nc, cam_q, cam_t, cam_K, np, pts, ne, edge_cam, edge_pt, edge_obs, edge_w);
    std::vector<KeyFrame*> adjust, fixed;
    for (int c = 0; c < nc; c++) (cam_fixed[c] ? fixed : adjust).push_back(S.kfs[(size_t)c].get());
    for (int e = 0; e < ne; e++) bad[e] = 0;
    if (adjust.empty()) return;
    bool stop = abort_flag != 0;
    Optimizer::LocalBundleAdjustment(adjust[0], adjust, fixed, &stop);
    for (int c = 0; c < nc; c++) pose_out(S.kfs[(size_t)c]->GetPose(), cam_q + 4 * c, cam_t + 3 * c);
    for (int p = 0; p < np; p++) for (int k = 0; k < 3; k++) pts[3 * p + k] = S.mps[(size_t)p]->GetWorldPos()[k];
    for (int e = 0; e < ne; e++)
        bad[e] = S.mps[(size_t)edge_pt[e]]->mObservations.count(S.kfs[(size_t)edge_cam[e]].get()) ? 0 : 1;
}

/* Optimizer::OptimizeSim3 (O3/src/Optimizer.cc:1960-2212) for n matched map points given in the two cameras' frames (both
 * keyframes at the identity pose): sim3 = (qx, qy, qz, qw, tx, ty, tz, s) in / out; inlier[i] = the match survived;
 * returns nIn */
int refopt_optimize_sim3(int n, const float* p1c, const float* p2c, const float* obs1, const float* obs2, const float* w1,
                         const float* w2, const float* K1, const float* K2, double* sim3, float th2, int fix_scale, uint8_t* inlier,
                         double* hessian49)
{
    Map map;
    KeyFrame kf1, kf2;
    Pinhole c1(K1[0], K1[1], K1[2], K1[3]), c2(K2[0], K2[1], K2[2], K2[3]);
    kf1.mpCamera = &c1; kf2.mpCamera = &c2;
    kf1.mpMap = kf2.mpMap = &map;
    kf1.mnId = 1; kf2.mnId = 2;
    std::vector<MapPoint> m1((size_t)n), m2((size_t)n);
    std::vector<MapPoint*> matches((size_t)n);
    for (int i = 0; i < n; i++) {
        cv::KeyPoint k1, k2;
        k1.pt.x = obs1[2 * i]; k1.pt.y = obs1[2 * i + 1]; k1.octave = i;
        k2.pt.x = obs2[2 * i]; k2.pt.y = obs2[2 * i + 1]; k2.octave = i;
        kf1.mvKeysUn.push_back(k1); kf2.mvKeysUn.push_back(k2);
        kf1.mvInvLevelSigma2.push_back(w1[i]); kf2.mvInvLevelSigma2.push_back(w2[i]);
        m1[(size_t)i].SetWorldPos(Eigen::Vector3f(p1c[3 * i], p1c[3 * i + 1], p1c[3 * i + 2]));
        m2[(size_t)i].SetWorldPos(Eigen::Vector3f(p2c[3 * i], p2c[3 * i + 1], p2c[3 * i + 2]));
        m1[(size_t)i].mpMap = m2[(size_t)i].mpMap = &map;
        kf1.mvpMapPoints.push_back(&m1[(size_t)i]);
        kf2.mvpMapPoints.push_back(&m2[(size_t)i]);
        m2[(size_t)i].mObservations[&kf2] = std::tuple<int, int>(i, -1);
        matches[(size_t)i] = &m2[(size_t)i];
    }
    kf1.N = kf2.N = n;
    g2o::Sim3 S12(Eigen::Quaterniond(sim3[3], sim3[0], sim3[1], sim3[2]), Eigen::Vector3d(sim3[4], sim3[5], sim3[6]), sim3[7]);
    Eigen::Matrix<double, 7, 7> H;
    H.setZero();
    const int nIn = Optimizer::OptimizeSim3(&kf1, &kf2, matches, S12, th2, fix_scale != 0, H, true);
    const Eigen::Quaterniond& r = S12.rotation();
    sim3[0] = r.x(); sim3[1] = r.y(); sim3[2] = r.z(); sim3[3] = r.w();
    for (int k = 0; k < 3; k++) sim3[4 + k] = S12.translation()[k];
    sim3[7] = S12.scale();
    for (int i = 0; i < n; i++) inlier[i] = matches[(size_t)i] ? 1 : 0;
    if (hessian49) for (int a = 0; a < 7; a++) for (int b = 0; b < 7; b++) hessian49[7 * a + b] = H(a, b);
    return nIn;
}

/* Optimizer::OptimizeEssentialGraph(pMap, pLoopKF, pCurKF, NonCorrectedSim3, CorrectedSim3, LoopConnections, bFixScale)
 * (O3/src/Optimizer.cc:1389-1651).  n keyframes with ids 0..n-1 (poses in / out), spanning tree parent[i] (-1 = none), old loop
 * edges and covisibility weights as (i, j[, w]) lists (symmetric), the two Sim3 maps as (keyframe, qx qy qz qw tx ty tz s) lists,
 * LoopConnections as (i, j) pairs; map points (positions in / out) with their reference keyframe, or -- corrected_by_cur[p] != 0 --
 * the keyframe recorded in mnCorrectedReference */
void refopt_optimize_essential_graph(int n, float* kf_q, float* kf_t, const int* parent, int init_kf, int loop_kf, int cur_kf,
                                     int n_loop_edges, const int* loop_edges, int n_cov, const int* cov, int n_nc, const int* nc_kf,
                                     const double* nc_sim3, int n_c, const int* c_kf, const double* c_sim3, int n_lc, const int* lc,
                                     int fix_scale, int npts, float* pts, const int* pt_ref, const uint8_t* corrected_by_cur,
                                     const int* corrected_ref)
{
    Map map;
    std::vector<std::unique_ptr<KeyFrame>> kfs;
    for (int i = 0; i < n; i++) {
        kfs.emplace_back(new KeyFrame);
        kfs.back()->mnId = (unsigned long)i;
        kfs.back()->mpMap = &map;
        kfs.back()->SetPose(pose_of(kf_q + 4 * i, kf_t + 3 * i));
        map.mvpKeyFrames.push_back(kfs.back().get());
    }
    map.mnInitKFid = (unsigned long)init_kf;
    map.mnMaxKFid = (unsigned long)(n - 1);
    for (int i = 0; i < n; i++)
        if (parent[i] >= 0) { kfs[(size_t)i]->mpParent = kfs[(size_t)parent[i]].get(); kfs[(size_t)parent[i]]->mspChildrens.insert(kfs[(size_t)i].get()); }
    for (int e = 0; e < n_loop_edges; e++) {
        kfs[(size_t)loop_edges[2 * e]]->mspLoopEdges.insert(kfs[(size_t)loop_edges[2 * e + 1]].get());
        kfs[(size_t)loop_edges[2 * e + 1]]->mspLoopEdges.insert(kfs[(size_t)loop_edges[2 * e]].get());
    }
    for (int e = 0; e < n_cov; e++) {
        KeyFrame* a = kfs[(size_t)cov[3 * e]].get();
        KeyFrame* b = kfs[(size_t)cov[3 * e + 1]].get();
        a->mConnectedKeyFrameWeights[b] = cov[3 * e + 2];
        b->mConnectedKeyFrameWeights[a] = cov[3 * e + 2];
    }
    for (auto& k : kfs) {   /* KeyFrame::UpdateBestCovisibles: best first */
        std::vector<std::pair<int, KeyFrame*>> v;
        for (auto& kv : k->mConnectedKeyFrameWeights) v.push_back(std::make_pair(kv.second, kv.first));
        std::sort(v.begin(), v.end(), [](const std::pair<int, KeyFrame*>& x, const std::pair<int, KeyFrame*>& y) {
            return x.first != y.first ? x.first > y.first : x.second->mnId < y.second->mnId; });
        for (auto& x : v) k->mvpOrderedConnectedKeyFrames.push_back(x.second);
    }
    auto sim3_of = [](const double* s) {
        return g2o::Sim3(Eigen::Quaterniond(s[3], s[0], s[1], s[2]), Eigen::Vector3d(s[4], s[5], s[6]), s[7]);
    };
    LoopClosing::KeyFrameAndPose NonCorrected, Corrected;
    for (int e = 0; e < n_nc; e++) NonCorrected[kfs[(size_t)nc_kf[e]].get()] = sim3_of(nc_sim3 + 8 * e);
    for (int e = 0; e < n_c; e++) Corrected[kfs[(size_t)c_kf[e]].get()] = sim3_of(c_sim3 + 8 * e);
    std::map<KeyFrame*, std::set<KeyFrame*>> LoopConnections;
    for (int e = 0; e < n_lc; e++) LoopConnections[kfs[(size_t)lc[2 * e]].get()].insert(kfs[(size_t)lc[2 * e + 1]].get());
    std::vector<std::unique_ptr<MapPoint>> mps;
    for (int p = 0; p < npts; p++) {
        mps.emplace_back(new MapPoint);
        MapPoint* mp = mps.back().get();
        mp->mnId = (unsigned long)p;
        mp->mpMap = &map;
        mp->SetWorldPos(Eigen::Vector3f(pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]));
        mp->mpRefKF = kfs[(size_t)pt_ref[p]].get();
        mp->mnCorrectedByKF = corrected_by_cur[p] ? (unsigned long)cur_kf : 0xfffffffful;
        mp->mnCorrectedReference = (unsigned long)corrected_ref[p];
        map.mvpMapPoints.push_back(mp);
    }
    const bool fs = fix_scale != 0;
    Optimizer::OptimizeEssentialGraph(&map, kfs[(size_t)loop_kf].get(), kfs[(size_t)cur_kf].get(), NonCorrected, Corrected, LoopConnections, fs);
    for (int i = 0; i < n; i++) pose_out(kfs[(size_t)i]->GetPose(), kf_q + 4 * i, kf_t + 3 * i);
    for (int p = 0; p < npts; p++) for (int k = 0; k < 3; k++) pts[3 * p + k] = mps[(size_t)p]->GetWorldPos()[k];
}

} // extern "C"
