/* oracle/slamshim/ref_matcher_glue.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Flat-array C entry points around the reference's own ORB_SLAM3::ORBmatcher (O3/src/ORBmatcher.cc compiled unmodified
 * through oracle/_ref/tree) and the reference's Frame / KeyFrame / MapPoint / Pinhole member bodies extracted into
 * oracle/_ref/gen/ref_bodies.inc.  The glue only builds the object graph those functions expect from the arrays and
 * flattens what they wrote; it holds no matching logic.  Used by tests/test_ref_matchers.py to pin oracle/track_oracle.cpp
 * and oracle/bow_oracle.cpp (and through them the CUDA kernels) to the reference's source. */
#include "ORBmatcher.h"
#include <cstring>
#include <memory>

using namespace ORB_SLAM3;

#include "../_ref/gen/ref_bodies.inc"

namespace ORB_SLAM3 {
float Frame::mnMinX = 0, Frame::mnMaxX = 0, Frame::mnMinY = 0, Frame::mnMaxY = 0;
float Frame::mfGridElementWidthInv = 0, Frame::mfGridElementHeightInv = 0;
}

struct RefKp { float x, y, size, angle, response; int32_t octave, class_id; };

namespace {

Sophus::SE3f pose_of(const float* q, const float* t)
{   /* q = x, y, z, w as stored; handed over without renormalisation */
    return Sophus::SE3f(Sophus::SO3f::fromStored(Eigen::Quaternionf(q[3], q[0], q[1], q[2])), Eigen::Vector3f(t[0], t[1], t[2]));
}
Sophus::Sim3f sim3_of(const float* q, const float* t, float s)
{   /* Sim3 of (unit quaternion, translation, scale): RxSO3(scale, SO3) as LoopClosing builds it from g2o::Sim3 */
    return Sophus::Sim3f(Sophus::RxSO3<float>(s, Sophus::SO3f::fromStored(Eigen::Quaternionf(q[3], q[0], q[1], q[2]))),
                         Eigen::Vector3f(t[0], t[1], t[2]));
}
cv::Mat desc_row(const uint8_t* d)
{
    cv::Mat m(1, 32, CV_8U);
    memcpy(m.ptr(0), d, 32);
    return m;
}
void fill_featvec(DBoW2::FeatureVector& fv, int nnodes, const uint32_t* node_id, const int* start, const uint32_t* idx)
{
    for (int k = 0; k < nnodes; k++)
        for (int j = start[k]; j < start[k + 1]; j++) fv.addFeature(node_id[k], idx[j]);
}

/* a Frame with its grid, plus (on demand) the KeyFrame made from it as the KeyFrame(Frame&, ...) constructor does */
struct RefFrame {
    Frame F;
    std::unique_ptr<KeyFrame> KF;
    std::unique_ptr<Pinhole> cam;
    std::vector<std::unique_ptr<MapPoint>> owned;
    float bounds[4];
    void set_statics()
    {   /* Frame.cc:443-456 (first-frame initialisation of the static members) */
        Frame::mnMinX = bounds[0]; Frame::mnMinY = bounds[1]; Frame::mnMaxX = bounds[2]; Frame::mnMaxY = bounds[3];
        Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
        Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    }
    KeyFrame* keyframe()
    {
        if (!KF) {   /* KeyFrame.cc:98-175: the members the matchers read, copied from the Frame */
            set_statics();
            KF.reset(new KeyFrame(FRAME_GRID_COLS, FRAME_GRID_ROWS, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv,
                                  F.mnScaleLevels, F.mfLogScaleFactor, (int)Frame::mnMinX, (int)Frame::mnMinY, (int)Frame::mnMaxX,
                                  (int)Frame::mnMaxY));
            KF->N = F.N;
            KF->mvKeys = F.mvKeys; KF->mvKeysUn = F.mvKeysUn;
            KF->mvuRight.assign(F.N, -1.f);
            KF->mDescriptors = F.mDescriptors.clone();
            KF->mvScaleFactors = F.mvScaleFactors; KF->mvLevelSigma2 = F.mvLevelSigma2; KF->mvInvLevelSigma2 = F.mvInvLevelSigma2;
            KF->mpCamera = F.mpCamera;
            KF->fx = cam->mvParameters[0]; KF->fy = cam->mvParameters[1]; KF->cx = cam->mvParameters[2]; KF->cy = cam->mvParameters[3];
            KF->mvpMapPoints.assign(F.N, nullptr);
            KF->mGrid.resize(FRAME_GRID_COLS);
            for (int i = 0; i < FRAME_GRID_COLS; i++) {
                KF->mGrid[i].resize(FRAME_GRID_ROWS);
                for (int j = 0; j < FRAME_GRID_ROWS; j++) KF->mGrid[i][j] = F.mGrid[i][j];
            }
            KF->mFeatVec = F.mFeatVec;
        }
        return KF.get();
    }
    MapPoint* new_point(int flat)
    {
        owned.emplace_back(new MapPoint);
        owned.back()->mnFlatIndex = flat;
        return owned.back().get();
    }
};

RefFrame* make_frame(const RefKp* kps, const uint8_t* desc, int n, const float* bounds, const float* scaleFactors, int nlevels,
                     const float* K, bool with_grid)
{
    RefFrame* R = new RefFrame;
    Frame& F = R->F;
    memcpy(R->bounds, bounds, sizeof(R->bounds));
    F.N = n;
    F.mvKeys.resize(n);
    for (int i = 0; i < n; i++) F.mvKeys[i] = cv::KeyPoint(kps[i].x, kps[i].y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id);
    F.mvKeysUn = F.mvKeys;
    F.mvuRight.assign(n, -1.f);
    F.mDescriptors = cv::Mat(std::max(n, 1), 32, CV_8U);
    if (n) memcpy(F.mDescriptors.ptr(0), desc, (size_t)n * 32);
    F.mvpMapPoints.assign(n, nullptr);
    F.mvbOutlier.assign(n, false);
    F.mnScaleLevels = nlevels;
    F.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    F.mvLevelSigma2.resize(nlevels); F.mvInvLevelSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {   /* ORBextractor.cc:291-300 */
        F.mvLevelSigma2[i] = scaleFactors[i] * scaleFactors[i];
        F.mvInvLevelSigma2[i] = 1.0f / F.mvLevelSigma2[i];
    }
    F.mfScaleFactor = nlevels > 1 ? scaleFactors[1] : 1.2f;
    F.mfLogScaleFactor = log(F.mfScaleFactor);   /* Frame.cc:399 */
    R->cam.reset(new Pinhole(K[0], K[1], K[2], K[3]));
    F.mpCamera = R->cam.get();
    R->set_statics();
    if (with_grid) F.AssignFeaturesToGrid();
    return R;
}

} // namespace

extern "C" {

void* refm_frame_create(const RefKp* kps, const uint8_t* desc, int n, const float* bounds, const float* scaleFactors, int nlevels,
                        const float* K)
{
    return make_frame(kps, desc, n, bounds, scaleFactors, nlevels, K, true);
}
void refm_frame_destroy(void* f) { delete (RefFrame*)f; }

int refm_grid_cell(void* f, int ix, int iy, int* out, int cap)
{
    const std::vector<size_t>& c = ((RefFrame*)f)->F.mGrid[ix][iy];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) out[i] = (int)c[i];
    return (int)c.size();
}

int refm_features_in_area(void* f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap)
{
    RefFrame* R = (RefFrame*)f;
    R->set_statics();
    const std::vector<size_t> v = R->F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
    return (int)v.size();
}

int refm_kf_features_in_area(void* f, float x, float y, float r, int* out, int cap)
{
    RefFrame* R = (RefFrame*)f;
    const std::vector<size_t> v = R->keyframe()->GetFeaturesInArea(x, y, r);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
    return (int)v.size();
}

int refm_descriptor_distance(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(desc_row(a), desc_row(b)); }

/* ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono = true), O3/src/ORBmatcher.cc:1553-1748 */
int refm_search_by_projection_last(void* fcur, const float* q, const float* t, int lastN, const uint8_t* has_mp,
                                   const uint8_t* outlier, const float* Xw, const uint8_t* mp_desc, const uint8_t* mp_obs_pos,
                                   const int* last_octave, const float* last_angle, float th, int checkOri, int* cur_mp)
{
    RefFrame* R = (RefFrame*)fcur;
    R->set_statics();
    Frame& C = R->F;
    C.SetPose(pose_of(q, t));
    std::fill(C.mvpMapPoints.begin(), C.mvpMapPoints.end(), nullptr);
    Frame L;
    L.N = lastN;
    L.mvKeys.resize(lastN); L.mvpMapPoints.assign(lastN, nullptr); L.mvbOutlier.assign(lastN, false);
    std::vector<std::unique_ptr<MapPoint>> pts;
    for (int i = 0; i < lastN; i++) {
        L.mvKeys[i].octave = last_octave[i];
        L.mvKeys[i].angle = last_angle[i];
        L.mvbOutlier[i] = outlier[i] != 0;
        if (!has_mp[i]) continue;
        pts.emplace_back(new MapPoint);
        MapPoint* p = pts.back().get();
        p->mnFlatIndex = i;
        p->mWorldPos = Eigen::Vector3f(Xw[3 * i], Xw[3 * i + 1], Xw[3 * i + 2]);
        p->mDescriptor = desc_row(mp_desc + (size_t)i * 32);
        p->nObs = mp_obs_pos[i] ? 1 : 0;
        L.mvpMapPoints[i] = p;
    }
    L.mvKeysUn = L.mvKeys;
    const float I[4] = { 0, 0, 0, 1 }, Z[3] = { 0, 0, 0 };
    L.mTcw = pose_of(I, Z);
    ORBmatcher matcher(0.9f, checkOri != 0);
    const int n = matcher.SearchByProjection(C, L, th, true);
    for (int i = 0; i < C.N; i++) cur_mp[i] = C.mvpMapPoints[i] ? C.mvpMapPoints[i]->mnFlatIndex : -1;
    std::fill(C.mvpMapPoints.begin(), C.mvpMapPoints.end(), nullptr);
    return n;
}

/* Frame::isInFrustum (O3/src/Frame.cc:575-636) over flat map points; min_dist / max_dist are mfMinDistance / mfMaxDistance */
void refm_is_in_frustum(const float* q, const float* t, const float* K, const float* bounds, int nlevels, float scaleFactor, int m,
                        const float* xw, const float* normal, const float* min_dist, const float* max_dist, const uint8_t* skip,
                        float cos_limit, uint8_t* in_view, float* projX, float* projY, int* level, float* view_cos)
{
    std::vector<float> sf(nlevels, 1.f);
    for (int i = 1; i < nlevels; i++) sf[i] = sf[i - 1] * scaleFactor;
    RefKp dummy = {};
    uint8_t d[32] = {};
    RefFrame* R = make_frame(&dummy, d, 0, bounds, sf.data(), nlevels, K, false);
    R->F.mfScaleFactor = scaleFactor;
    R->F.mfLogScaleFactor = log(scaleFactor);
    R->F.SetPose(pose_of(q, t));
    for (int i = 0; i < m; i++) {
        in_view[i] = 0; projX[i] = projY[i] = view_cos[i] = 0; level[i] = 0;
        if (skip && skip[i]) continue;
        MapPoint p;
        p.mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        p.mNormalVector = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        p.mfMinDistance = min_dist[i]; p.mfMaxDistance = max_dist[i];
        if (R->F.isInFrustum(&p, cos_limit)) {
            in_view[i] = 1; projX[i] = p.mTrackProjX; projY[i] = p.mTrackProjY; level[i] = p.mnTrackScaleLevel; view_cos[i] = p.mTrackViewCos;
        }
    }
    delete R;
}

/* ORBmatcher::SearchByProjection(F, vpMapPoints, th, bFarPoints = false), O3/src/ORBmatcher.cc:44-205 */
int refm_search_by_projection_map(void* fcur, int M, const float* projX, const float* projY, const int* level, const float* viewCos,
                                  const uint8_t* mp_desc, const uint8_t* mp_obs_pos, float th, float nnratio,
                                  const uint8_t* cur_blocked, int* cur_mp)
{
    RefFrame* R = (RefFrame*)fcur;
    R->set_statics();
    Frame& C = R->F;
    MapPoint blocker;
    blocker.nObs = 1;
    for (int i = 0; i < C.N; i++) C.mvpMapPoints[i] = (cur_blocked && cur_blocked[i]) ? &blocker : nullptr;
    std::vector<std::unique_ptr<MapPoint>> pts(M);
    std::vector<MapPoint*> v(M);
    for (int i = 0; i < M; i++) {
        pts[i].reset(new MapPoint);
        MapPoint* p = v[i] = pts[i].get();
        p->mnFlatIndex = i;
        p->mbTrackInView = true;
        p->mTrackProjX = projX[i]; p->mTrackProjY = projY[i]; p->mnTrackScaleLevel = level[i]; p->mTrackViewCos = viewCos[i];
        p->mDescriptor = desc_row(mp_desc + (size_t)i * 32);
        p->nObs = mp_obs_pos[i] ? 1 : 0;
    }
    ORBmatcher matcher(nnratio);
    const int n = matcher.SearchByProjection(C, v, th, false, 50.0f);
    for (int i = 0; i < C.N; i++) cur_mp[i] = (C.mvpMapPoints[i] && C.mvpMapPoints[i] != &blocker) ? C.mvpMapPoints[i]->mnFlatIndex : -1;
    std::fill(C.mvpMapPoints.begin(), C.mvpMapPoints.end(), nullptr);
    return n;
}

/* SearchByBoW(pKF, F, vpMapPointMatches) (kf_kf = 0, :214-393) and SearchByBoW(pKF1, pKF2, vpMatches12) (kf_kf = 1, :709-834).
 * Sides as oracle.bow.search_by_bow passes them; valid == NULL means every feature holds a map point. */
int refm_search_by_bow(int kf_kf, int n1, const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int nn1,
                       const uint32_t* node1, const int* start1, const uint32_t* idx1, int n2, const uint8_t* desc2,
                       const float* angle2, const uint8_t* valid2, int nn2, const uint32_t* node2, const int* start2,
                       const uint32_t* idx2, float nnratio, int checkOri, int* m12, int* m21)
{
    const float bounds[4] = { 0, 0, 640, 480 }, sf[1] = { 1.f }, K[4] = { 500, 500, 320, 240 };
    auto side = [&](int n, const uint8_t* desc, const float* angle, const uint8_t* valid, int nn, const uint32_t* node, const int* start,
                    const uint32_t* idx) {
        std::vector<RefKp> kps(std::max(n, 1));
        for (int i = 0; i < n; i++) { kps[i] = RefKp{}; kps[i].angle = angle[i]; }
        RefFrame* R = make_frame(kps.data(), desc, n, bounds, sf, 1, K, false);
        fill_featvec(R->F.mFeatVec, nn, node, start, idx);
        KeyFrame* kf = R->keyframe();
        for (int i = 0; i < n; i++)
            if (!valid || valid[i]) { MapPoint* p = R->new_point(i); p->nObs = 1; kf->mvpMapPoints[i] = p; }
        return R;
    };
    std::unique_ptr<RefFrame> A(side(n1, desc1, angle1, valid1, nn1, node1, start1, idx1));
    std::unique_ptr<RefFrame> B(side(n2, desc2, angle2, valid2, nn2, node2, start2, idx2));
    for (int i = 0; i < n1; i++) m12[i] = -1;
    for (int i = 0; i < n2; i++) m21[i] = -1;
    ORBmatcher matcher(nnratio, checkOri != 0);
    int n;
    if (!kf_kf) {
        std::vector<MapPoint*> vm;
        n = matcher.SearchByBoW(A->keyframe(), B->F, vm);
        for (int i = 0; i < n2 && i < (int)vm.size(); i++)
            if (vm[i]) { m21[i] = vm[i]->mnFlatIndex; m12[vm[i]->mnFlatIndex] = i; }
    } else {
        std::vector<MapPoint*> vm;
        n = matcher.SearchByBoW(A->keyframe(), B->keyframe(), vm);
        for (int i = 0; i < n1 && i < (int)vm.size(); i++)
            if (vm[i]) { m12[i] = vm[i]->mnFlatIndex; m21[vm[i]->mnFlatIndex] = i; }
    }
    return n;
}

/* SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), :605-707 */
int refm_search_for_initialization(int n1, const RefKp* kps1, const uint8_t* desc1, void* f2, float* prev_matched, int window,
                                   float nnratio, int checkOri, int* m12)
{
    RefFrame* R2 = (RefFrame*)f2;
    R2->set_statics();
    std::unique_ptr<RefFrame> R1(make_frame(kps1, desc1, n1, R2->bounds, R2->F.mvScaleFactors.data(), R2->F.mnScaleLevels,
                                            R2->cam->mvParameters.data(), false));
    std::vector<cv::Point2f> prev(n1);
    for (int i = 0; i < n1; i++) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
    std::vector<int> v12;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchForInitialization(R1->F, R2->F, prev, v12, window);
    for (int i = 0; i < n1; i++) { m12[i] = v12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
    return n;
}

/* SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo = false, bCoarse), :836-1058.  Poses as stored (q, t). */
int refm_search_for_triangulation(void* f1, const uint8_t* has_mp1, const float* q1, const float* t1, void* f2,
                                  const uint8_t* has_mp2, const float* q2, const float* t2, float nnratio, int checkOri, int coarse,
                                  int* m12)
{
    RefFrame *R1 = (RefFrame*)f1, *R2 = (RefFrame*)f2;
    KeyFrame *k1 = R1->keyframe(), *k2 = R2->keyframe();
    k1->mFeatVec = R1->F.mFeatVec; k2->mFeatVec = R2->F.mFeatVec;
    k1->SetPose(pose_of(q1, t1)); k2->SetPose(pose_of(q2, t2));
    for (int i = 0; i < k1->N; i++) k1->mvpMapPoints[i] = has_mp1[i] ? R1->new_point(i) : nullptr;
    for (int i = 0; i < k2->N; i++) k2->mvpMapPoints[i] = has_mp2[i] ? R2->new_point(i) : nullptr;
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchForTriangulation(k1, k2, pairs, false, coarse != 0);
    for (int i = 0; i < k1->N; i++) m12[i] = -1;
    for (auto& p : pairs) m12[p.first] = (int)p.second;
    return n;
}

void refm_frame_set_featvec(void* f, int nnodes, const uint32_t* node_id, const int* start, const uint32_t* idx)
{
    RefFrame* R = (RefFrame*)f;
    R->F.mFeatVec.clear();
    fill_featvec(R->F.mFeatVec, nnodes, node_id, start, idx);
    if (R->KF) R->KF->mFeatVec = R->F.mFeatVec;
}

/* Fuse(pKF, vpMapPoints, th, bRight = false), :1060-1228.  The keyframe starts without map points, every candidate has one
 * observation; best_idx[i] = keyframe keypoint map point i was fused with (read back from the side effects: AddObservation
 * on first use of a keypoint, Replace by the earlier point afterwards), -1 if none.  min_dist / max_dist are mfMinDistance /
 * mfMaxDistance; skip[i] = the point is already observed in the keyframe (IsInKeyFrame). */
int refm_fuse(void* fkf, const float* q, const float* t, int m, const float* xw, const float* normal, const float* min_dist,
              const float* max_dist, const uint8_t* mp_desc, const uint8_t* skip, float th, int* best_idx)
{
    RefFrame* R = (RefFrame*)fkf;
    KeyFrame* kf = R->keyframe();
    kf->SetPose(pose_of(q, t));
    std::fill(kf->mvpMapPoints.begin(), kf->mvpMapPoints.end(), nullptr);
    std::vector<std::unique_ptr<MapPoint>> pts(m);
    std::vector<MapPoint*> v(m);
    for (int i = 0; i < m; i++) {
        pts[i].reset(new MapPoint);
        MapPoint* p = v[i] = pts[i].get();
        p->mnFlatIndex = i;
        p->mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        p->mNormalVector = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        p->mfMinDistance = min_dist[i]; p->mfMaxDistance = max_dist[i];
        p->mDescriptor = desc_row(mp_desc + (size_t)i * 32);
        p->nObs = 1;
        if (skip && skip[i]) p->mObservations[kf] = std::tuple<int, int>(-2, -1);
    }
    ORBmatcher matcher;
    const int n = matcher.Fuse(kf, v, th, false);
    std::vector<int> first_at(m, -1);   /* keypoint index a point was ADDED to */
    for (int i = 0; i < m; i++) {
        const auto it = pts[i]->mObservations.find(kf);
        if (it != pts[i]->mObservations.end() && std::get<0>(it->second) >= 0) first_at[i] = std::get<0>(it->second);
    }
    for (int i = 0; i < m; i++) {
        best_idx[i] = first_at[i];
        if (best_idx[i] < 0 && pts[i]->mpReplaced) best_idx[i] = first_at[pts[i]->mpReplaced->mnFlatIndex];
    }
    /* a point that replaced the keypoint's earlier holder: the earlier holder is marked replaced by it */
    for (int i = 0; i < m; i++)
        if (pts[i]->mpReplaced && first_at[i] >= 0 && best_idx[pts[i]->mpReplaced->mnFlatIndex] < 0)
            best_idx[pts[i]->mpReplaced->mnFlatIndex] = first_at[i];
    return n;
}


/* SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming), :395-494.  sq / st: Scw.quaternion() / translation().
 * kp_matched[kf n]: vpMatched[k] != NULL on entry; kp_point[kf n] receives the candidate index newly written to vpMatched[k]. */
int refm_search_by_projection_sim3(void* fkf, const float* sq, const float* st, int m, const float* xw, const float* normal,
                                   const float* min_dist, const float* max_dist, const uint8_t* mp_desc, const uint8_t* skip,
                                   const uint8_t* kp_matched, int th, float ratioHamming, int* kp_point)
{
    RefFrame* R = (RefFrame*)fkf;
    KeyFrame* kf = R->keyframe();
    Sophus::Sim3f Scw(Eigen::Quaternionf(sq[3], sq[0], sq[1], sq[2]), Eigen::Vector3f(st[0], st[1], st[2]));
    MapPoint occupied;
    std::vector<std::unique_ptr<MapPoint>> pts(m);
    std::vector<MapPoint*> v(m);
    for (int i = 0; i < m; i++) {
        pts[i].reset(new MapPoint);
        MapPoint* p = v[i] = pts[i].get();
        p->mnFlatIndex = i;
        p->mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        p->mNormalVector = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        p->mfMinDistance = min_dist[i]; p->mfMaxDistance = max_dist[i];
        p->mDescriptor = desc_row(mp_desc + (size_t)i * 32);
        p->mbBad = skip && skip[i];
    }
    std::vector<MapPoint*> matched(kf->N, nullptr);
    for (int k = 0; k < kf->N; k++) if (kp_matched[k]) matched[k] = &occupied;
    ORBmatcher matcher(0.75f, true);
    const int n = matcher.SearchByProjection(kf, Scw, v, matched, th, ratioHamming);
    for (int k = 0; k < kf->N; k++) kp_point[k] = (matched[k] && matched[k] != &occupied) ? matched[k]->mnFlatIndex : -1;
    return n;
}

/* Fuse(pKF, Scw, vpPoints, th, vpReplacePoint), :1236-1345.  The keyframe holds no map points on entry, so every fused
 * candidate is added (AddObservation / AddMapPoint) or, when an earlier candidate took the keypoint, reported through
 * vpReplacePoint; best_idx[i] = the keypoint either way. */
int refm_fuse_sim3(void* fkf, const float* sq, const float* st, int m, const float* xw, const float* normal, const float* min_dist,
                   const float* max_dist, const uint8_t* mp_desc, const uint8_t* skip, float th, int* best_idx)
{
    RefFrame* R = (RefFrame*)fkf;
    KeyFrame* kf = R->keyframe();
    std::fill(kf->mvpMapPoints.begin(), kf->mvpMapPoints.end(), nullptr);
    Sophus::Sim3f Scw(Eigen::Quaternionf(sq[3], sq[0], sq[1], sq[2]), Eigen::Vector3f(st[0], st[1], st[2]));
    std::vector<std::unique_ptr<MapPoint>> pts(m);
    std::vector<MapPoint*> v(m), repl(m, nullptr);
    for (int i = 0; i < m; i++) {
        pts[i].reset(new MapPoint);
        MapPoint* p = v[i] = pts[i].get();
        p->mnFlatIndex = i;
        p->mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
        p->mNormalVector = Eigen::Vector3f(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        p->mfMinDistance = min_dist[i]; p->mfMaxDistance = max_dist[i];
        p->mDescriptor = desc_row(mp_desc + (size_t)i * 32);
        p->nObs = 1;
        p->mbBad = skip && skip[i];
    }
    ORBmatcher matcher;
    const int n = matcher.Fuse(kf, Scw, v, th, repl);
    std::vector<int> first_at(m, -1);
    for (int i = 0; i < m; i++) {
        const auto it = pts[i]->mObservations.find(kf);
        if (it != pts[i]->mObservations.end()) first_at[i] = std::get<0>(it->second);
    }
    for (int i = 0; i < m; i++) best_idx[i] = first_at[i] >= 0 ? first_at[i] : (repl[i] ? first_at[repl[i]->mnFlatIndex] : -1);
    return n;
}

/* SearchBySim3(pKF1, pKF2, vpMatches12, S12, th), :1347-1551.  has_s[i] = keypoint i of keyframe s holds a (good) map point. */
int refm_search_by_sim3(void* f1, void* f2, const float* q1, const float* t1, const float* q2, const float* t2, const float* s12q,
                        const float* s12t, const uint8_t* skip1, const float* xw1, const float* min1, const float* max1,
                        const uint8_t* desc1, const uint8_t* skip2, const float* xw2, const float* min2, const float* max2,
                        const uint8_t* desc2, float th, int* match12)
{
    RefFrame *R1 = (RefFrame*)f1, *R2 = (RefFrame*)f2;
    KeyFrame *k1 = R1->keyframe(), *k2 = R2->keyframe();
    k1->SetPose(pose_of(q1, t1)); k2->SetPose(pose_of(q2, t2));
    auto fill = [](RefFrame* R, KeyFrame* kf, const uint8_t* skip, const float* xw, const float* mind, const float* maxd, const uint8_t* desc) {
        for (int i = 0; i < kf->N; i++) {
            kf->mvpMapPoints[i] = nullptr;
            if (skip[i]) continue;
            MapPoint* p = R->new_point(i);
            p->mWorldPos = Eigen::Vector3f(xw[3 * i], xw[3 * i + 1], xw[3 * i + 2]);
            p->mfMinDistance = mind[i]; p->mfMaxDistance = maxd[i];
            p->mDescriptor = desc_row(desc + (size_t)i * 32);
            kf->mvpMapPoints[i] = p;
        }
    };
    fill(R1, k1, skip1, xw1, min1, max1, desc1);
    fill(R2, k2, skip2, xw2, min2, max2, desc2);
    Sophus::Sim3f S12(Eigen::Quaternionf(s12q[3], s12q[0], s12q[1], s12q[2]), Eigen::Vector3f(s12t[0], s12t[1], s12t[2]));
    std::vector<MapPoint*> vm(k1->N, nullptr);
    ORBmatcher matcher(0.75f, true);
    const int n = matcher.SearchBySim3(k1, k2, vm, S12, th);
    for (int i = 0; i < k1->N; i++) match12[i] = vm[i] ? vm[i]->mnFlatIndex : -1;
    return n;
}

} // extern "C"
