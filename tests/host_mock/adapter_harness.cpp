// adapter_harness.cpp -- TEST INFRASTRUCTURE.  Builds mock Frame / KeyFrame / MapPoint graphs from flat
// arrays, drives the C++ adapters of dvmslam_b200/host/ exactly as the reference's call sites would, and
// hands the results back flat, so that tests/test_host_adapters.py can compare them with the direct C-ABI
// calls and the oracle.
#include <cstring>
#include <memory>
#include <string>

#include "mock_slam.h"
#include "optimizer_adapter.h"
#include "orb_matcher_adapter.h"
#include "vocabulary_adapter.h"

namespace mock {
std::mutex MapPoint::mGlobalMutex;
float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY;
float g_K[4];
}
using namespace mock;

namespace {
std::string g_err;
long unsigned int g_next_frame_id = 1;

void fill_frame(Frame& F, const dvm_keypoint* kps, const uint8_t* desc, int n, const float* scale, const float* invsig2, int nlevels)
{
    F.mnId = g_next_frame_id++;
    F.N = n;
    F.mvKeysUn.resize(n);
    if (n) std::memcpy((void*)F.mvKeysUn.data(), kps, sizeof(dvm_keypoint) * n);
    F.mvKeys = F.mvKeysUn;
    F.mvuRight.assign(n, -1.f);
    F.mDescriptors.create(n > 0 ? n : 1, 32, CV_8U);
    for (int i = 0; i < n; i++) std::memcpy(F.mDescriptors.ptr(i), desc + (size_t)i * 32, 32);
    F.mvpMapPoints.assign(n, nullptr);
    F.mvbOutlier.assign(n, false);
    F.mvScaleFactors.assign(scale, scale + nlevels);
    F.mvInvLevelSigma2.assign(invsig2, invsig2 + nlevels);
}

void set_pose(SE3f& T, const float* q, const float* t)
{
    for (int k = 0; k < 4; k++) T.q.q[k] = q[k];
    for (int k = 0; k < 3; k++) T.t.v[k] = t[k];
}

FeatureVector to_fv(int nn, const uint32_t* node, const int32_t* start, const uint32_t* idx)
{
    FeatureVector fv;
    for (int i = 0; i < nn; i++) fv[node[i]] = std::vector<unsigned>(idx + start[i], idx + start[i + 1]);
    return fv;
}

template <class Fn>
int guarded(Fn fn)
{
    try { return fn(); }
    catch (const dvm_host::Error& e) { g_err = e.what(); return -1000 + e.code; }
}
} // namespace

extern "C" {

const char* hm_last_error() { return g_err.c_str(); }

void hm_set_camera(const float* K, const float* bounds)
{
    Frame::fx = K[0]; Frame::fy = K[1]; Frame::cx = K[2]; Frame::cy = K[3];
    for (int k = 0; k < 4; k++) g_K[k] = K[k];
    Frame::mnMinX = bounds[0]; Frame::mnMinY = bounds[1]; Frame::mnMaxX = bounds[2]; Frame::mnMaxY = bounds[3];
}

// TrackWithMotionModel's matcher call: Cur (pose prior set), Last (map points per keypoint)
int hm_search_by_projection_last(const dvm_keypoint* ckps, const uint8_t* cdesc, int nc, const float* scale, const float* invsig2,
                                 int nlevels, const float* q, const float* t, const dvm_keypoint* lkps, int nl,
                                 const uint8_t* has_mp, const uint8_t* outlier, const float* Xw, const uint8_t* mp_desc,
                                 const uint8_t* obs_pos, float th, int checkOri, int* cur_mp)
{
    Frame Cur, Last;
    fill_frame(Cur, ckps, cdesc, nc, scale, invsig2, nlevels);
    std::vector<uint8_t> zero((size_t)(nl > 0 ? nl : 1) * 32);
    fill_frame(Last, lkps, zero.data(), nl, scale, invsig2, nlevels);
    set_pose(Cur.mTcw, q, t);
    std::vector<std::unique_ptr<MapPoint>> mps(nl);
    for (int i = 0; i < nl; i++) {
        Last.mvbOutlier[i] = outlier[i] != 0;
        if (!has_mp[i]) continue;
        mps[i].reset(new MapPoint);
        for (int k = 0; k < 3; k++) mps[i]->pos(k) = Xw[3 * i + k];
        std::memcpy(mps[i]->desc.ptr(0), mp_desc + (size_t)i * 32, 32);
        mps[i]->nObs = obs_pos[i] ? 1 : 0;
        Last.mvpMapPoints[i] = mps[i].get();
    }
    return guarded([&] {
        const int n = dvm_host::SearchByProjectionLast(Cur, Last, th, checkOri != 0);
        for (int i = 0; i < nc; i++) {
            cur_mp[i] = -1;
            if (Cur.mvpMapPoints[i])
                for (int j = 0; j < nl; j++) if (mps[j].get() == Cur.mvpMapPoints[i]) cur_mp[i] = j;
        }
        return n;
    });
}

// SearchLocalPoints' matcher call: vpMapPoints with the fields isInFrustum stored; in_view[m] = mbTrackInView
int hm_search_by_projection_map(const dvm_keypoint* ckps, const uint8_t* cdesc, int nc, const float* scale, const float* invsig2,
                                int nlevels, int m, const uint8_t* in_view, const float* px, const float* py, const int* level,
                                const float* vcos, const uint8_t* mp_desc, const uint8_t* obs_pos, const uint8_t* cur_blocked,
                                float th, float nnratio, int* cur_mp)
{
    Frame F;
    fill_frame(F, ckps, cdesc, nc, scale, invsig2, nlevels);
    std::vector<std::unique_ptr<MapPoint>> mps(m);
    std::vector<MapPoint*> vp(m);
    for (int i = 0; i < m; i++) {
        mps[i].reset(new MapPoint);
        MapPoint& p = *mps[i];
        p.mbTrackInView = in_view[i] != 0; p.mTrackProjX = px[i]; p.mTrackProjY = py[i]; p.mnTrackScaleLevel = level[i];
        p.mTrackViewCos = vcos[i];
        std::memcpy(p.desc.ptr(0), mp_desc + (size_t)i * 32, 32);
        p.nObs = obs_pos[i] ? 1 : 0;
        vp[i] = &p;
    }
    MapPoint held;   // keypoints that already hold a map point with observations
    held.nObs = 1;
    for (int i = 0; i < nc; i++) if (cur_blocked && cur_blocked[i]) F.mvpMapPoints[i] = &held;
    return guarded([&] {
        const int n = dvm_host::SearchByProjectionMap(F, vp, th, nnratio);
        for (int i = 0; i < nc; i++) {
            cur_mp[i] = -1;
            if (F.mvpMapPoints[i] && F.mvpMapPoints[i] != &held)
                for (int j = 0; j < m; j++) if (vp[j] == F.mvpMapPoints[i]) cur_mp[i] = j;
        }
        return n;
    });
}

// Optimizer::PoseOptimization(Frame*): one map point per listed keypoint
int hm_pose_optimization(const dvm_keypoint* kps, int n, const float* scale, const float* invsig2, int nlevels, float* q, float* t,
                         const int* mp_of_kp /* [n]: index into Xw or -1 */, const float* Xw, uint8_t* outlier)
{
    Frame F;
    std::vector<uint8_t> desc((size_t)(n > 0 ? n : 1) * 32);
    fill_frame(F, kps, desc.data(), n, scale, invsig2, nlevels);
    set_pose(F.mTcw, q, t);
    std::vector<std::unique_ptr<MapPoint>> mps(n);
    for (int i = 0; i < n; i++) {
        if (mp_of_kp[i] < 0) continue;
        mps[i].reset(new MapPoint);
        for (int k = 0; k < 3; k++) mps[i]->pos(k) = Xw[3 * mp_of_kp[i] + k];
        F.mvpMapPoints[i] = mps[i].get();
    }
    return guarded([&] {
        const int inl = dvm_host::PoseOptimization<Frame, MapPoint>(&F);
        for (int k = 0; k < 4; k++) q[k] = F.mTcw.q.q[k];
        for (int k = 0; k < 3; k++) t[k] = F.mTcw.t.v[k];
        for (int i = 0; i < n; i++) outlier[i] = F.mvbOutlier[i];
        return inl;
    });
}

// both SearchByBoW overloads; out[i]: for kf_kf == 0 per feature of side 2 the side-1 feature whose map point it
// received, for kf_kf != 0 per feature of side 1 the side-2 feature whose map point it received (-1 none)
int hm_search_by_bow(int kf_kf, int n1, const uint8_t* d1, const float* a1, const uint8_t* v1, int nn1, const uint32_t* node1,
                     const int32_t* st1, const uint32_t* idx1, int n2, const uint8_t* d2, const float* a2, const uint8_t* v2,
                     int nn2, const uint32_t* node2, const int32_t* st2, const uint32_t* idx2, float nnratio, int checkOri, int* out)
{
    const float one[1] = { 1.f };
    auto make_kps = [](int n, const float* ang) {
        std::vector<dvm_keypoint> k(n);
        for (int i = 0; i < n; i++) { k[i] = dvm_keypoint{ float(i % 640), float(i / 640), 31.f, ang[i], 1.f, 0, -1 }; }
        return k;
    };
    KeyFrame K1, K2;
    Frame F;
    std::vector<std::unique_ptr<MapPoint>> m1(n1), m2(n2);
    auto fill_kf = [&](KeyFrame& K, int n, const uint8_t* d, const float* ang, const uint8_t* v, std::vector<std::unique_ptr<MapPoint>>& mp,
                       const FeatureVector& fv) {
        const std::vector<dvm_keypoint> k = make_kps(n, ang);
        K.N = n;
        K.mvKeysUn.resize(n);
        if (n) std::memcpy((void*)K.mvKeysUn.data(), k.data(), sizeof(dvm_keypoint) * n);
        K.mDescriptors.create(n > 0 ? n : 1, 32, CV_8U);
        for (int i = 0; i < n; i++) std::memcpy(K.mDescriptors.ptr(i), d + (size_t)i * 32, 32);
        K.mapPoints.assign(n, nullptr);
        for (int i = 0; i < n; i++)
            if (!v || v[i]) { mp[i].reset(new MapPoint); K.mapPoints[i] = mp[i].get(); }
        K.mFeatVec = fv;
    };
    fill_kf(K1, n1, d1, a1, v1, m1, to_fv(nn1, node1, st1, idx1));
    // a context frame for the GPU call (SearchByBoW itself has no grid to consult)
    const std::vector<dvm_keypoint> k2 = make_kps(n2, a2);
    fill_frame(F, k2.data(), d2, n2, one, one, 1);
    F.mFeatVec = to_fv(nn2, node2, st2, idx2);
    if (kf_kf) fill_kf(K2, n2, d2, a2, v2, m2, F.mFeatVec);
    return guarded([&] {
        std::vector<MapPoint*> res;
        int n;
        if (!kf_kf) {
            n = dvm_host::SearchByBoW(&K1, F, res, nnratio, checkOri != 0);
            for (int i = 0; i < n2; i++) {
                out[i] = -1;
                if (res[i]) for (int j = 0; j < n1; j++) if (m1[j].get() == res[i]) out[i] = j;
            }
        } else {
            n = dvm_host::SearchByBoW(&K1, &K2, res, nnratio, checkOri != 0, dvm_host::device_frame(F).frame.h);
            for (int i = 0; i < n1; i++) {
                out[i] = -1;
                if (res[i]) for (int j = 0; j < n2; j++) if (m2[j].get() == res[i]) out[i] = j;
            }
        }
        return n;
    });
}

int hm_search_for_initialization(const dvm_keypoint* k1, const uint8_t* d1, int n1, const dvm_keypoint* k2, const uint8_t* d2, int n2,
                                 const float* scale, const float* invsig2, int nlevels, float* prev, int window, float nnratio,
                                 int checkOri, int* matches12)
{
    Frame F1, F2;
    fill_frame(F1, k1, d1, n1, scale, invsig2, nlevels);
    fill_frame(F2, k2, d2, n2, scale, invsig2, nlevels);
    std::vector<cv::Point2f> vbPrev(n1);
    for (int i = 0; i < n1; i++) vbPrev[i] = cv::Point2f(prev[2 * i], prev[2 * i + 1]);
    return guarded([&] {
        std::vector<int> m12;
        const int n = dvm_host::SearchForInitialization(F1, F2, vbPrev, m12, window, nnratio, checkOri != 0);
        for (int i = 0; i < n1; i++) { matches12[i] = m12[i]; prev[2 * i] = vbPrev[i].x; prev[2 * i + 1] = vbPrev[i].y; }
        return n;
    });
}

// Optimizer::LocalBundleAdjustment on a flat scene: free cameras become pKF (camera 0) and its covisible keyframes,
// fixed cameras exist only as observers of the local map points.  counts[4] = num_fixedKF, num_OptKF, num_MPs,
// num_edges; erased[ne] = the observation was culled.
int hm_local_ba(int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne, const int* edge_cam,
                const int* edge_pt, const float* edge_obs, const int* edge_octave, const float* invsig2, int nlevels,
                int stop_flag_mode /* 0 none, 1 present and clear, 2 already set */, int* counts, uint8_t* erased)
{
    Map map;
    std::vector<std::unique_ptr<KeyFrame>> kfs(nc);
    std::vector<std::unique_ptr<MapPoint>> mps(np);
    std::vector<int> next_kp(nc, 0);
    for (int c = 0; c < nc; c++) {
        kfs[c].reset(new KeyFrame);
        kfs[c]->mnId = 100 + c;
        kfs[c]->map = &map;
        kfs[c]->fx = g_K[0]; kfs[c]->fy = g_K[1]; kfs[c]->cx = g_K[2]; kfs[c]->cy = g_K[3];
        set_pose(kfs[c]->Tcw, cam_q + 4 * c, cam_t + 3 * c);
        kfs[c]->mvInvLevelSigma2.assign(invsig2, invsig2 + nlevels);
    }
    map.initKFid = 1;   // not in the window: the fixed set is exactly the non-covisible observers
    for (int j = 0; j < np; j++) {
        mps[j].reset(new MapPoint);
        mps[j]->mnId = j;
        mps[j]->map = &map;
        for (int k = 0; k < 3; k++) mps[j]->pos(k) = pts[3 * j + k];
    }
    std::vector<int> edge_kp(ne);
    for (int e = 0; e < ne; e++) {
        KeyFrame& K = *kfs[edge_cam[e]];
        const int kp = next_kp[edge_cam[e]]++;
        edge_kp[e] = kp;
        cv::KeyPoint k(edge_obs[2 * e], edge_obs[2 * e + 1], 31.f, 0.f, 1.f, edge_octave[e]);
        K.mvKeysUn.push_back(k);
        K.mvuRight.push_back(-1.f);
        K.mapPoints.push_back(mps[edge_pt[e]].get());
        K.N++;
        mps[edge_pt[e]]->observations[&K] = std::make_tuple(kp, -1);
    }
    int first_free = -1;
    for (int c = 0; c < nc; c++)
        if (!cam_fixed[c]) {
            if (first_free < 0) first_free = c;
            else kfs[first_free]->covisible.push_back(kfs[c].get());
        }
    if (first_free < 0) return -1;
    dvm_host::LbaHandle solver;
    bool stop = stop_flag_mode == 2;
    return guarded([&] {
        dvm_host::check(dvm_lba_create(&solver.h, dvm_host::device_from_env(), 64), "dvm_lba_create");
        counts[0] = counts[1] = counts[2] = counts[3] = -7;
        dvm_host::LocalBundleAdjustment<KeyFrame, MapPoint, Map>(solver.h, kfs[first_free].get(), stop_flag_mode ? &stop : nullptr, &map,
                                                                 counts[0], counts[1], counts[2], counts[3]);
        for (int c = 0; c < nc; c++) {
            for (int k = 0; k < 4; k++) cam_q[4 * c + k] = kfs[c]->Tcw.q.q[k];
            for (int k = 0; k < 3; k++) cam_t[3 * c + k] = kfs[c]->Tcw.t.v[k];
        }
        for (int j = 0; j < np; j++) for (int k = 0; k < 3; k++) pts[3 * j + k] = mps[j]->pos(k);
        for (int e = 0; e < ne; e++) erased[e] = kfs[edge_cam[e]]->mapPoints[edge_kp[e]] == nullptr;
        return map.changeIndex;
    });
}

} // extern "C"

namespace {
void fill_keyframe(KeyFrame& K, const dvm_keypoint* kps, const uint8_t* desc, int n, const float* scale, const float* sigma2,
                   const float* invsig2, int nlevels, const float* q, const float* t, unsigned long id)
{
    K.mnId = id; K.N = n;
    K.mvKeysUn.resize(n);
    if (n) std::memcpy((void*)K.mvKeysUn.data(), kps, sizeof(dvm_keypoint) * n);
    K.mDescriptors.create(n > 0 ? n : 1, 32, CV_8U);
    for (int i = 0; i < n; i++) std::memcpy(K.mDescriptors.ptr(i), desc + (size_t)i * 32, 32);
    K.mapPoints.assign(n, nullptr);
    K.mvuRight.assign(n, -1.f);
    K.mvScaleFactors.assign(scale, scale + nlevels);
    K.mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
    K.mvInvLevelSigma2.assign(invsig2, invsig2 + nlevels);
    K.fx = g_K[0]; K.fy = g_K[1]; K.cx = g_K[2]; K.cy = g_K[3];
    K.mnMinX = Frame::mnMinX; K.mnMinY = Frame::mnMinY; K.mnMaxX = Frame::mnMaxX; K.mnMaxY = Frame::mnMaxY;
    set_pose(K.Tcw, q, t);
}
} // namespace

extern "C" {

// LocalMapping::CreateNewMapPoints' matcher call.  pairs[2 * cap] receives vMatchedPairs; returns nmatches.
int hm_search_for_triangulation(const dvm_keypoint* k1, const uint8_t* d1, int n1, const uint8_t* has1, int nn1, const uint32_t* node1,
                                const int32_t* st1, const uint32_t* idx1, const float* q1, const float* t1, const dvm_keypoint* k2,
                                const uint8_t* d2, int n2, const uint8_t* has2, int nn2, const uint32_t* node2, const int32_t* st2,
                                const uint32_t* idx2, const float* q2, const float* t2, const float* scale, const float* sigma2,
                                const float* invsig2, int nlevels, int coarse, int checkOri, int* pairs, int cap, int* npairs)
{
    KeyFrame K1, K2;
    fill_keyframe(K1, k1, d1, n1, scale, sigma2, invsig2, nlevels, q1, t1, 1);
    fill_keyframe(K2, k2, d2, n2, scale, sigma2, invsig2, nlevels, q2, t2, 2);
    K1.mFeatVec = to_fv(nn1, node1, st1, idx1);
    K2.mFeatVec = to_fv(nn2, node2, st2, idx2);
    MapPoint some;
    for (int i = 0; i < n1; i++) if (has1[i]) K1.mapPoints[i] = &some;
    for (int i = 0; i < n2; i++) if (has2[i]) K2.mapPoints[i] = &some;
    return guarded([&] {
        std::vector<std::pair<size_t, size_t>> vp;
        const int n = dvm_host::SearchForTriangulation(&K1, &K2, vp, false, coarse != 0, checkOri != 0, dvm_host::device_frame(K1).frame.h);
        *npairs = (int)vp.size();
        for (size_t i = 0; i < vp.size() && (int)i < cap; i++) { pairs[2 * i] = (int)vp[i].first; pairs[2 * i + 1] = (int)vp[i].second; }
        return n;
    });
}

// LocalMapping::SearchInNeighbors' Fuse call.  kf_has_mp[n]: the keyframe keypoint already holds a map point (with
// kf_mp_obs[n] observations).  Per listed map point: action[i] = 0 nothing, 1 added to the keyframe at kp[i],
// 2 it replaced the keyframe's point, 3 it was replaced by the keyframe's point.
int hm_fuse(const dvm_keypoint* kps, const uint8_t* desc, int n, const float* scale, const float* sigma2, const float* invsig2,
            int nlevels, const float* q, const float* t, const uint8_t* kf_has_mp, const int* kf_mp_obs, int m, const uint8_t* skip,
            const float* xw, const float* normal, const float* min_dist, const float* max_dist, const uint8_t* mp_desc,
            const int* mp_obs, float th, int* action, int* kp)
{
    KeyFrame K;
    fill_keyframe(K, kps, desc, n, scale, sigma2, invsig2, nlevels, q, t, 7);
    std::vector<std::unique_ptr<MapPoint>> held(n), mps(m);
    for (int i = 0; i < n; i++)
        if (kf_has_mp[i]) { held[i].reset(new MapPoint); held[i]->nObs = kf_mp_obs[i]; K.mapPoints[i] = held[i].get(); }
    std::vector<MapPoint*> vp(m, nullptr);
    for (int i = 0; i < m; i++) {
        if (skip[i]) continue;    // null entry of vpMapPoints
        mps[i].reset(new MapPoint);
        MapPoint& p = *mps[i];
        for (int k = 0; k < 3; k++) { p.pos(k) = xw[3 * i + k]; p.normal(k) = normal[3 * i + k]; }
        p.minDistInv = 0.8f * min_dist[i]; p.maxDistInv = 1.2f * max_dist[i]; p.minDist = min_dist[i]; p.maxDist = max_dist[i];
        std::memcpy(p.desc.ptr(0), mp_desc + (size_t)i * 32, 32);
        p.nObs = mp_obs[i];
        vp[i] = &p;
    }
    return guarded([&] {
        const int nFused = dvm_host::Fuse(&K, vp, th);
        for (int i = 0; i < m; i++) {
            action[i] = 0; kp[i] = -1;
            if (!vp[i]) continue;
            if (vp[i]->replacedBy) { action[i] = 3; continue; }
            for (int k = 0; k < n; k++) {
                if (K.mapPoints[k] == vp[i]) { action[i] = 1; kp[i] = k; }
                if (held[k] && held[k]->replacedBy == vp[i]) { action[i] = 2; kp[i] = k; }
            }
        }
        return nFused;
    });
}


// ---- Sim3-guided matchers through the adapters (results flattened to candidate / keypoint indices) ----
static void fill_points(std::vector<std::unique_ptr<MapPoint>>& mps, std::vector<MapPoint*>& vp, int m, const uint8_t* bad, const float* xw,
                        const float* normal, const float* min_dist, const float* max_dist, const uint8_t* mp_desc)
{
    mps.resize(m); vp.assign(m, nullptr);
    for (int i = 0; i < m; i++) {
        mps[i].reset(new MapPoint);
        MapPoint& p = *mps[i];
        p.mnId = (unsigned long)i;
        for (int k = 0; k < 3; k++) { p.pos(k) = xw[3 * i + k]; if (normal) p.normal(k) = normal[3 * i + k]; }
        p.minDist = min_dist[i]; p.maxDist = max_dist[i];
        std::memcpy(p.desc.ptr(0), mp_desc + (size_t)i * 32, 32);
        p.bad = bad && bad[i];
        vp[i] = &p;
    }
}

int hm_search_by_projection_sim3(const dvm_keypoint* kps, const uint8_t* desc, int n, const float* scale, const float* sigma2,
                                 const float* invsig2, int nlevels, const float* sq, const float* st, int m, const uint8_t* bad,
                                 const float* xw, const float* normal, const float* min_dist, const float* max_dist,
                                 const uint8_t* mp_desc, const uint8_t* kp_matched, int th, float ratio, int* kp_point)
{
    KeyFrame K;
    const float I[4] = { 0, 0, 0, 1 }, Z[3] = { 0, 0, 0 };
    fill_keyframe(K, kps, desc, n, scale, sigma2, invsig2, nlevels, I, Z, 21);
    std::vector<std::unique_ptr<MapPoint>> mps;
    std::vector<MapPoint*> vp;
    fill_points(mps, vp, m, bad, xw, normal, min_dist, max_dist, mp_desc);
    MapPoint occupied;
    std::vector<MapPoint*> matched(n, nullptr);
    for (int k = 0; k < n; k++) if (kp_matched[k]) matched[k] = &occupied;
    mock::Sim3f S;
    for (int k = 0; k < 4; k++) S.q.q[k] = sq[k];
    for (int k = 0; k < 3; k++) S.t(k) = st[k];
    return guarded([&] {
        const int r = dvm_host::SearchByProjection(&K, S, vp, matched, th, ratio);
        for (int k = 0; k < n; k++) kp_point[k] = (matched[k] && matched[k] != &occupied) ? (int)matched[k]->mnId : -1;
        return r;
    });
}

int hm_fuse_sim3(const dvm_keypoint* kps, const uint8_t* desc, int n, const float* scale, const float* sigma2, const float* invsig2,
                 int nlevels, const float* sq, const float* st, int m, const uint8_t* bad, const float* xw, const float* normal,
                 const float* min_dist, const float* max_dist, const uint8_t* mp_desc, float th, int* best_idx)
{
    KeyFrame K;
    const float I[4] = { 0, 0, 0, 1 }, Z[3] = { 0, 0, 0 };
    fill_keyframe(K, kps, desc, n, scale, sigma2, invsig2, nlevels, I, Z, 22);
    std::vector<std::unique_ptr<MapPoint>> mps;
    std::vector<MapPoint*> vp;
    fill_points(mps, vp, m, bad, xw, normal, min_dist, max_dist, mp_desc);
    std::vector<MapPoint*> repl(m, nullptr);
    mock::Sim3f S;
    for (int k = 0; k < 4; k++) S.q.q[k] = sq[k];
    for (int k = 0; k < 3; k++) S.t(k) = st[k];
    return guarded([&] {
        const int r = dvm_host::Fuse(&K, S, vp, th, repl);
        for (int i = 0; i < m; i++) {
            best_idx[i] = -1;
            const auto it = vp[i]->observations.find(&K);
            if (it != vp[i]->observations.end()) best_idx[i] = std::get<0>(it->second);
            else if (repl[i]) best_idx[i] = std::get<0>(repl[i]->observations.find(&K)->second);
        }
        return r;
    });
}

int hm_search_by_sim3(const dvm_keypoint* kps1, const uint8_t* desc1, int n1, const dvm_keypoint* kps2, const uint8_t* desc2, int n2,
                      const float* scale, const float* sigma2, const float* invsig2, int nlevels, const float* q1, const float* t1,
                      const float* q2, const float* t2, const float* sq, const float* st, const uint8_t* skip1, const float* xw1,
                      const float* min1, const float* max1, const uint8_t* mpd1, const uint8_t* skip2, const float* xw2,
                      const float* min2, const float* max2, const uint8_t* mpd2, float th, int* match12)
{
    KeyFrame A, B;
    fill_keyframe(A, kps1, desc1, n1, scale, sigma2, invsig2, nlevels, q1, t1, 23);
    fill_keyframe(B, kps2, desc2, n2, scale, sigma2, invsig2, nlevels, q2, t2, 24);
    std::vector<std::unique_ptr<MapPoint>> m1, m2;
    std::vector<MapPoint*> v1, v2;
    fill_points(m1, v1, n1, nullptr, xw1, nullptr, min1, max1, mpd1);
    fill_points(m2, v2, n2, nullptr, xw2, nullptr, min2, max2, mpd2);
    for (int i = 0; i < n1; i++) A.mapPoints[i] = skip1[i] ? nullptr : v1[i];
    for (int i = 0; i < n2; i++) B.mapPoints[i] = skip2[i] ? nullptr : v2[i];
    std::vector<MapPoint*> vm(n1, nullptr);
    mock::Sim3f S;
    for (int k = 0; k < 4; k++) S.q.q[k] = sq[k];
    for (int k = 0; k < 3; k++) S.t(k) = st[k];
    return guarded([&] {
        const int r = dvm_host::SearchBySim3(&A, &B, vm, S, th);
        for (int i = 0; i < n1; i++) match12[i] = vm[i] ? (int)vm[i]->mnId : -1;
        return r;
    });
}


// OptimizeEssentialGraph through the adapter on a chain of n keyframes (parent of k = k - 1, keyframe 0 = the map's initial
// keyframe and the loop keyframe, keyframe n - 1 = the current keyframe with its Sim3 correction and a loop connection to 0;
// covis[k] >= 100 adds a covisibility edge k -> k - 2).  poses[n][7] = q, t in/out; points[m][3] in/out, point_ref[m] = id of the
// point's reference keyframe.
int hm_optimize_essential_graph(int n, float* poses, const double* corrected_last, const double* noncorrected_last, const int* covis,
                                int m, float* points, const int* point_ref, int fix_scale)
{
    Map map;
    std::vector<std::unique_ptr<KeyFrame>> kfs(n);
    for (int k = 0; k < n; k++) {
        kfs[k].reset(new KeyFrame);
        kfs[k]->mnId = (unsigned long)k; kfs[k]->map = &map;
        set_pose(kfs[k]->Tcw, poses + 7 * k, poses + 7 * k + 4);
        map.keyframes.push_back(kfs[k].get());
    }
    for (int k = 1; k < n; k++) {
        kfs[k]->parent = kfs[k - 1].get();
        kfs[k - 1]->children.insert(kfs[k].get());
        if (k >= 2 && covis[k] >= 100) { kfs[k]->weights[kfs[k - 2].get()] = covis[k]; kfs[k - 2]->weights[kfs[k].get()] = covis[k]; }
    }
    map.initKFid = 0;
    std::vector<std::unique_ptr<MapPoint>> mps(m);
    for (int i = 0; i < m; i++) {
        mps[i].reset(new MapPoint);
        for (int k = 0; k < 3; k++) mps[i]->pos(k) = points[3 * i + k];
        mps[i]->refKF = kfs[point_ref[i]].get();
        map.mappoints.push_back(mps[i].get());
    }
    auto to_sim3 = [](const double* p) { mock::Sim3 S; for (int k = 0; k < 4; k++) S.r.q[k] = p[k]; for (int k = 0; k < 3; k++) S.t.v[k] = p[4 + k]; S.s = p[7]; return S; };
    std::map<KeyFrame*, mock::Sim3> corrected, noncorrected;
    corrected[kfs[n - 1].get()] = to_sim3(corrected_last);
    noncorrected[kfs[n - 1].get()] = to_sim3(noncorrected_last);
    std::map<KeyFrame*, std::set<KeyFrame*>> loopConnections;
    loopConnections[kfs[n - 1].get()].insert(kfs[0].get());
    dvm_essential_graph* solver = nullptr;
    const int rc = guarded([&] {
        dvm_host::check(dvm_essential_graph_create(&solver, dvm_host::device_from_env()), "dvm_essential_graph_create");
        dvm_host::OptimizeEssentialGraph(solver, &map, kfs[0].get(), kfs[n - 1].get(), noncorrected, corrected, loopConnections, fix_scale != 0);
        for (int k = 0; k < n; k++) {
            for (int i = 0; i < 4; i++) poses[7 * k + i] = kfs[k]->Tcw.q.q[i];
            for (int i = 0; i < 3; i++) poses[7 * k + 4 + i] = kfs[k]->Tcw.t(i);
        }
        for (int i = 0; i < m; i++)
            for (int k = 0; k < 3; k++) points[3 * i + k] = mps[i]->pos(k);
        return map.changeIndex;
    });
    dvm_essential_graph_destroy(solver);
    return rc;
}

} // extern "C"

// DBoW2::BowVector / FeatureVector with the reference's method semantics (DBoW2/BowVector.cpp:30-71,
// DBoW2/FeatureVector.cpp:27-37)
namespace mock {
enum LNorm { L1, L2 };
struct BowVector : std::map<unsigned, double> {
    void addWeight(unsigned id, double v) { auto it = lower_bound(id); if (it != end() && it->first == id) it->second += v; else insert(it, value_type(id, v)); }
    void addIfNotExist(unsigned id, double v) { auto it = lower_bound(id); if (it == end() || it->first != id) insert(it, value_type(id, v)); }
    void normalize(LNorm t)
    {
        double norm = 0.0;
        if (t == L1) for (auto& kv : *this) norm += std::fabs(kv.second);
        else { for (auto& kv : *this) norm += kv.second * kv.second; norm = std::sqrt(norm); }
        if (norm > 0.0) for (auto& kv : *this) kv.second /= norm;
    }
};
struct FeatureVectorMap : std::map<unsigned, std::vector<unsigned>> {
    void addFeature(unsigned id, unsigned i) { (*this)[id].push_back(i); }
};
}

extern "C" {

// Frame::ComputeBoW through the adapter: vocabulary from a text file, descriptors as cv::Mat rows
int hm_compute_bow(const char* voc_path, const uint8_t* desc, int n, int levelsup, int* bow_word, double* bow_value, int* n_bow,
                   int* fv_node, int* fv_start, int* fv_idx, int* n_fv)
{
    return guarded([&] {
        dvm_host::Vocabulary voc;
        if (!voc.loadFromTextFile(voc_path)) return -1;
        std::vector<cv::Mat> feats(n);
        for (int i = 0; i < n; i++) { feats[i].create(1, 32, CV_8U); std::memcpy(feats[i].ptr(0), desc + (size_t)i * 32, 32); }
        mock::BowVector v;
        mock::FeatureVectorMap fv;
        voc.transform(feats, v, fv, levelsup);
        int k = 0;
        for (auto& kv : v) { bow_word[k] = (int)kv.first; bow_value[k] = kv.second; k++; }
        *n_bow = k;
        k = 0;
        int p = 0;
        fv_start[0] = 0;
        for (auto& kv : fv) { fv_node[k] = (int)kv.first; for (unsigned i : kv.second) fv_idx[p++] = (int)i; fv_start[++k] = p; }
        *n_fv = k;
        return voc.k() * 100 + voc.L();
    });
}

} // extern "C"

extern "C" {

// Optimizer::BundleAdjustment on a flat scene (every camera a keyframe of the map; the fixed one is the map's initial
// keyframe).  direct != 0: nLoopKF is the origin keyframe (poses / points written directly), else parked in mTcwGBA / mPosGBA.
int hm_bundle_adjustment(int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne, const int* edge_cam,
                         const int* edge_pt, const float* edge_obs, const int* edge_octave, const float* invsig2, int nlevels,
                         int iterations, int robust, int direct)
{
    Map map;
    std::vector<std::unique_ptr<KeyFrame>> kfs(nc);
    std::vector<std::unique_ptr<MapPoint>> mps(np);
    std::vector<KeyFrame*> vk;
    std::vector<MapPoint*> vm;
    for (int c = 0; c < nc; c++) {
        kfs[c].reset(new KeyFrame);
        kfs[c]->mnId = 10 + c; kfs[c]->map = &map;
        kfs[c]->fx = g_K[0]; kfs[c]->fy = g_K[1]; kfs[c]->cx = g_K[2]; kfs[c]->cy = g_K[3];
        set_pose(kfs[c]->Tcw, cam_q + 4 * c, cam_t + 3 * c);
        kfs[c]->mvInvLevelSigma2.assign(invsig2, invsig2 + nlevels);
        if (cam_fixed[c]) map.initKFid = kfs[c]->mnId;
        vk.push_back(kfs[c].get());
    }
    map.originKF = kfs[0].get();
    for (int j = 0; j < np; j++) {
        mps[j].reset(new MapPoint);
        mps[j]->mnId = j; mps[j]->map = &map;
        for (int k = 0; k < 3; k++) mps[j]->pos(k) = pts[3 * j + k];
        vm.push_back(mps[j].get());
    }
    for (int e = 0; e < ne; e++) {
        KeyFrame& K = *kfs[edge_cam[e]];
        cv::KeyPoint k(edge_obs[2 * e], edge_obs[2 * e + 1], 31.f, 0.f, 1.f, edge_octave[e]);
        K.mvKeysUn.push_back(k); K.mvuRight.push_back(-1.f); K.mapPoints.push_back(mps[edge_pt[e]].get());
        mps[edge_pt[e]]->observations[&K] = std::make_tuple(K.N, -1);
        K.N++;
    }
    dvm_host::LbaHandle solver;
    return guarded([&] {
        dvm_host::check(dvm_lba_create(&solver.h, dvm_host::device_from_env(), 100), "dvm_lba_create");
        dvm_host::BundleAdjustment<KeyFrame, MapPoint>(solver.h, vk, vm, iterations, nullptr, direct ? kfs[0]->mnId : 999999ul, robust != 0);
        int parked = 0;
        for (int c = 0; c < nc; c++) {
            const SE3f& T = direct ? kfs[c]->Tcw : kfs[c]->mTcwGBA;
            parked += kfs[c]->mnBAGlobalForKF == 999999ul;
            for (int k = 0; k < 4; k++) cam_q[4 * c + k] = T.q.q[k];
            for (int k = 0; k < 3; k++) cam_t[3 * c + k] = T.t.v[k];
        }
        for (int j = 0; j < np; j++) for (int k = 0; k < 3; k++) pts[3 * j + k] = direct ? mps[j]->pos(k) : mps[j]->mPosGBA(k);
        return parked;
    });
}

// Optimizer::LocalBundleAdjustment(pMainKF, vpAdjustKF, vpFixedKF, pbStopFlag) -- the welding BA -- on a flat scene: fixed
// cameras become vpFixedKF, the others vpAdjustKF (the first of them is pMainKF).  Returns the number of erased
// observations; *normal_updates = UpdateNormalAndDepth calls.
int hm_merge_ba(int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne, const int* edge_cam,
                const int* edge_pt, const float* edge_obs, const int* edge_octave, const float* invsig2, int nlevels,
                int* normal_updates)
{
    Map map;
    std::vector<std::unique_ptr<KeyFrame>> kfs(nc);
    std::vector<std::unique_ptr<MapPoint>> mps(np);
    std::vector<KeyFrame*> adjust, fixed;
    for (int c = 0; c < nc; c++) {
        kfs[c].reset(new KeyFrame);
        kfs[c]->mnId = 10 + c; kfs[c]->map = &map;
        kfs[c]->fx = g_K[0]; kfs[c]->fy = g_K[1]; kfs[c]->cx = g_K[2]; kfs[c]->cy = g_K[3];
        set_pose(kfs[c]->Tcw, cam_q + 4 * c, cam_t + 3 * c);
        kfs[c]->mvInvLevelSigma2.assign(invsig2, invsig2 + nlevels);
        (cam_fixed[c] ? fixed : adjust).push_back(kfs[c].get());
    }
    for (int j = 0; j < np; j++) {
        mps[j].reset(new MapPoint);
        mps[j]->mnId = j; mps[j]->map = &map;
        for (int k = 0; k < 3; k++) mps[j]->pos(k) = pts[3 * j + k];
    }
    for (int e = 0; e < ne; e++) {
        KeyFrame& K = *kfs[edge_cam[e]];
        cv::KeyPoint k(edge_obs[2 * e], edge_obs[2 * e + 1], 31.f, 0.f, 1.f, edge_octave[e]);
        K.mvKeysUn.push_back(k); K.mvuRight.push_back(-1.f); K.mapPoints.push_back(mps[edge_pt[e]].get());
        mps[edge_pt[e]]->observations[&K] = std::make_tuple(K.N, -1);
        K.N++;
    }
    dvm_host::LbaHandle solver;
    return guarded([&] {
        dvm_host::check(dvm_lba_create(&solver.h, dvm_host::device_from_env(), 64), "dvm_lba_create");
        dvm_host::LocalBundleAdjustment<KeyFrame, MapPoint>(solver.h, adjust[0], adjust, fixed, nullptr);
        int left = 0, updates = 0;
        for (int j = 0; j < np; j++) { left += (int)mps[j]->observations.size(); updates += mps[j]->normalUpdates; }
        for (int c = 0; c < nc; c++) {
            for (int k = 0; k < 4; k++) cam_q[4 * c + k] = kfs[c]->Tcw.q.q[k];
            for (int k = 0; k < 3; k++) cam_t[3 * c + k] = kfs[c]->Tcw.t.v[k];
        }
        for (int j = 0; j < np; j++) for (int k = 0; k < 3; k++) pts[3 * j + k] = mps[j]->pos(k);
        *normal_updates = updates;
        return ne - left;
    });
}

// Optimizer::OptimizeSim3 on flat correspondences: both keyframes sit at the identity of their own map, so a map point's
// world position is its camera-frame position; point i of KF1 is matched to map point i of map 2, which is observed in
// KF2 (keypoint i) iff in_kf2[i].  matched[i] = vpMatches1[i] != NULL afterwards; *hessian_zero = mAcumHessian was zeroed.
int hm_optimize_sim3(int n, const float* p1c, const float* p2c, const float* obs1, const float* obs2, const int* oct1,
                     const int* oct2, const uint8_t* in_kf2, const float* invsig2, int nlevels, double* q, double* t, double* sc,
                     float th2, int fix_scale, int all_points, uint8_t* matched, int* hessian_zero)
{
    Map map1, map2;
    KeyFrame kf1, kf2;
    kf1.map = &map1; kf2.map = &map2; kf1.mnId = 1; kf2.mnId = 2;
    for (KeyFrame* k : { &kf1, &kf2 }) {
        k->fx = g_K[0]; k->fy = g_K[1]; k->cx = g_K[2]; k->cy = g_K[3];
        k->mvInvLevelSigma2.assign(invsig2, invsig2 + nlevels);
    }
    std::vector<std::unique_ptr<MapPoint>> m1(n), m2(n);
    std::vector<MapPoint*> vpMatches1(n);
    for (int i = 0; i < n; i++) {
        m1[i].reset(new MapPoint); m2[i].reset(new MapPoint);
        m1[i]->map = &map1; m2[i]->map = &map2;
        for (int k = 0; k < 3; k++) { m1[i]->pos(k) = p1c[3 * i + k]; m2[i]->pos(k) = p2c[3 * i + k]; }
        kf1.mvKeysUn.push_back(cv::KeyPoint(obs1[2 * i], obs1[2 * i + 1], 31.f, 0.f, 1.f, oct1[i]));
        kf1.mapPoints.push_back(m1[i].get());
        kf2.mvKeysUn.push_back(cv::KeyPoint(obs2[2 * i], obs2[2 * i + 1], 31.f, 0.f, 1.f, oct2[i]));
        kf2.mapPoints.push_back(in_kf2[i] ? m2[i].get() : nullptr);
        if (in_kf2[i]) m2[i]->observations[&kf2] = std::make_tuple(i, -1);
        vpMatches1[i] = m2[i].get();
    }
    kf1.N = kf2.N = n;
    mock::Sim3 S;
    for (int k = 0; k < 4; k++) S.r.q[k] = q[k];
    for (int k = 0; k < 3; k++) S.t.v[k] = t[k];
    S.s = *sc;
    mock::Mat77 H;
    dvm_host::Sim3Handle solver;
    return guarded([&] {
        dvm_host::check(dvm_sim3_create(&solver.h, dvm_host::device_from_env()), "dvm_sim3_create");
        const int nIn = dvm_host::OptimizeSim3<KeyFrame, MapPoint, mock::Sim3, mock::Mat77>(solver.h, &kf1, &kf2, vpMatches1, S, th2,
                                                                                            fix_scale != 0, H, all_points != 0);
        for (int i = 0; i < n; i++) matched[i] = vpMatches1[i] != nullptr;
        for (int k = 0; k < 4; k++) q[k] = S.r.q[k];
        for (int k = 0; k < 3; k++) t[k] = S.t.v[k];
        *sc = S.s;
        *hessian_zero = 1;
        for (double x : H.m) if (x != 0) *hessian_zero = 0;
        return nIn;
    });
}

} // extern "C"
