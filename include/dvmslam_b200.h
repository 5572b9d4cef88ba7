/*
 * dvmslam_b200.h -- C-ABI of libdvmslam_b200.so: the B200-native (sm_100a) ORB front end and
 * local back end of DVM-SLAM behind the reference's own operator surface.
 *
 * The reference (proroklab/DVM-SLAM, /root/reference) has no FFI: its hot path sits behind three
 * C++ classes compiled into libORB_SLAM3 -- ORBextractor, ORBmatcher, Optimizer.  Each entry point
 * below names the reference interface it replaces (O3/ = src/slam_system/orb_slam3/).  The C++
 * drop-in adapters that give back the reference's class signatures on top of these calls live in
 * dvmslam_b200/host/ ; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions: plain pointers and sizes only; every function returns DVM_OK (0) or a negative
 * dvm_status; no exceptions cross the boundary; a handle is bound to one GPU and one CUDA stream,
 * is not thread-safe, and different handles are independent.  "host" entry points are synchronous
 * (results are in host memory on return, like the reference's calls); "_device" entry points
 * enqueue on the handle's stream and leave results in HBM for the next stage.
 */
#ifndef DVMSLAM_B200_H
#define DVMSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVM_API __attribute__((visibility("default")))

typedef enum dvm_status {
    DVM_OK = 0,
    DVM_ERR_INVALID = -1,   /* bad argument (null pointer, size out of range, wrong state) */
    DVM_ERR_CUDA = -2,      /* a CUDA runtime call failed; see dvm_last_error() */
    DVM_ERR_CAPACITY = -3,  /* an internal or caller-provided buffer was too small */
    DVM_ERR_NO_DEVICE = -4, /* no usable sm_100 GPU: there is no CPU fallback */
    DVM_ERR_NUMERIC = -5    /* linear solve failed / non-finite values */
} dvm_status;

/* Text of the last error raised on the calling thread ("" if none). */
DVM_API const char* dvm_last_error(void);
/* Library version / build info, and the number of kernels this library has launched so far in the
 * process (used by bench.py for "gpu_launches"). */
DVM_API const char* dvm_version(void);
DVM_API uint64_t dvm_kernel_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * ORB extractor   -- replaces ORB_SLAM3::ORBextractor (O3/include/ORBextractor.h:44-96)
 * ---------------------------------------------------------------------------------------------- */

/* Same memory layout as cv::KeyPoint (28 bytes), which is what ORBextractor::operator() fills. */
typedef struct dvm_keypoint {
    float x, y;      /* pt, in level-0 pixel units (already multiplied by the level's scale) */
    float size;      /* (int)(31 * scale[octave]) */
    float angle;     /* degrees in [0,360), cv::fastAtan2 of the intensity centroid */
    float response;  /* cv::FAST score */
    int32_t octave;
    int32_t class_id; /* always -1 */
} dvm_keypoint;

typedef struct dvm_orb dvm_orb;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (O3/src/ORBextractor.cc:282-339).  max_width/max_height bound the images later passed to
 * dvm_orb_extract (device buffers are sized once here).  device = CUDA ordinal. */
DVM_API int dvm_orb_create(dvm_orb** out, int device, int nfeatures, float scale_factor, int nlevels,
                           int ini_th_fast, int min_th_fast, int max_width, int max_height);
DVM_API void dvm_orb_destroy(dvm_orb* h);
/* A second extractor with the same parameters, own buffers and own stream (an agent keeps several: the
 * reference constructs two, O3/src/Tracking.cc:575-581; dvm_tracker alternates two so that consecutive frames
 * are extracted concurrently). */
DVM_API int dvm_orb_clone(const dvm_orb* src, dvm_orb** out);

/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (O3/include/ORBextractor.h:57-67) plus the per-level feature quota
 * (mnFeaturesPerLevel).  Any pointer may be NULL.  Arrays hold nlevels entries. */
DVM_API int dvm_orb_tables(const dvm_orb* h, int* nlevels, float* scale, float* inv_scale, float* sigma2,
                           float* inv_sigma2, int* features_per_level);
/* Upper bound on the number of keypoints one extract call can return (size kps/desc with it). */
DVM_API int dvm_orb_max_keypoints(const dvm_orb* h);
/* The tighter bound for the image size of the last extract call (the octree can return at most
 * quota + 4 keypoints per level); equals dvm_orb_max_keypoints() before the first extract. */
DVM_API int dvm_orb_max_keypoints_current(const dvm_orb* h);

/* int ORBextractor::operator()(image, mask (ignored), keypoints, descriptors, vLappingArea)
 * (O3/src/ORBextractor.cc:876-955).  gray: CV_8UC1 host image, `stride` bytes per row.
 * lap0/lap1 = vLappingArea[0..1] (the mono caller passes {0,1000}, O3/src/Frame.cc:411).
 * kps[cap] / desc[cap*32] receive *n_out entries in the reference's output order; *mono_index
 * receives the reference's return value (-1 for an empty image, with *n_out = 0). */
DVM_API int dvm_orb_extract(dvm_orb* h, const uint8_t* gray, int width, int height, int stride, int lap0, int lap1,
                            dvm_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_index);

/* Same computation with the image already in HBM and the results left in HBM (handle-owned
 * buffers, valid until the next extract on this handle).  Enqueues on the handle's stream and
 * returns without synchronising; dvm_orb_sync() waits.  This is the path the per-frame tracker and
 * bench.py's device-resident `value` use. */
DVM_API int dvm_orb_extract_device(dvm_orb* h, const uint8_t* gray_dev, int width, int height, int stride, int lap0,
                                   int lap1);
/* Same, leaving the result in caller-provided HBM: kps_dev[dvm_orb_max_keypoints()], desc_dev[max*32],
 * counts_dev[2] = {n, mono_index} (all three or none; none = the handle's own buffers).  Lets a Frame own
 * its keypoints without a copy while the handle already extracts the next image. */
DVM_API int dvm_orb_extract_device_to(dvm_orb* h, const uint8_t* gray_dev, int width, int height, int stride, int lap0,
                                      int lap1, dvm_keypoint* kps_dev, uint8_t* desc_dev, int32_t* counts_dev);
DVM_API int dvm_orb_sync(dvm_orb* h);
/* Device pointers to the last result: kps (dvm_keypoint[cap]), desc (u8[cap*32]),
 * counts (int32[2] = {n, mono_index}). */
DVM_API int dvm_orb_result_device(const dvm_orb* h, const dvm_keypoint** kps_dev, const uint8_t** desc_dev,
                                  const int32_t** counts_dev);
/* The CUDA stream (cudaStream_t) the handle launches on, for CUDA-event timing by the caller. */
DVM_API void* dvm_orb_stream(const dvm_orb* h);

/* Per-stage device timing for bench.py's roofline: with profiling enabled every extract call records
 * CUDA events on the handle's stream around its four stages {pyramid resize chain, per-cell FAST,
 * octree + output ordering, orientation + blur + rBRIEF} and synchronises; dvm_orb_get_profile returns
 * the number of profiled frames and the summed milliseconds per stage (and resets both). */
DVM_API int dvm_orb_set_profiling(dvm_orb* h, int enable);
DVM_API int dvm_orb_get_profile(dvm_orb* h, int* n_frames, float* stage_ms_sum /* [4] */);

/* Stage read-back for parity tests (valid after an extract; synchronises).
 *  dvm_orb_debug_level_image: mvImagePyramid[level] (O3/include/ORBextractor.h:69), or with
 *    blurred=1 the GaussianBlur'ed working copy the descriptors are sampled from
 *    (O3/src/ORBextractor.cc:919-920).  out holds w*h bytes, tightly packed.
 *  dvm_orb_debug_level_keypoints: which=0 -> vToDistributeKeys (all per-cell FAST keypoints of the
 *    level, order unspecified); which=1 -> the octree's selection in list order
 *    (O3/src/ORBextractor.cc:697-698).  x,y are level pixels relative to the image origin. */
DVM_API int dvm_orb_debug_level_size(const dvm_orb* h, int level, int* w, int* hgt);
DVM_API int dvm_orb_debug_level_image(dvm_orb* h, int level, int blurred, uint8_t* out);
DVM_API int dvm_orb_debug_level_keypoints(dvm_orb* h, int level, int which, int* xs, int* ys, int* responses,
                                          int cap, int* n_out);

/* ------------------------------------------------------------------------------------------------
 * Frame (mono)   -- replaces the parts of ORB_SLAM3::Frame the matchers need
 *                   (O3/src/Frame.cc:371-506 mono constructor, :712-782 grid queries)
 * ---------------------------------------------------------------------------------------------- */
typedef struct dvm_frame dvm_frame;

/* scale_factors / inv_level_sigma2: the extractor's GetScaleFactors() / GetInverseScaleSigmaSquares()
 * that Frame copies (O3/src/Frame.cc:399-405).  cuda_stream: a cudaStream_t to launch on (e.g.
 * dvm_orb_stream() so that frame building is stream-ordered after extraction), or NULL for a private
 * stream. */
DVM_API int dvm_frame_create(dvm_frame** out, int device, void* cuda_stream, int max_keypoints, int nlevels,
                             const float* scale_factors, const float* inv_level_sigma2);
DVM_API void dvm_frame_destroy(dvm_frame* f);
/* UndistortKeyPoints (identity: the caller passes undistorted keypoints; with k1 == 0 they are the
 * extractor's output, O3/src/Frame.cc:791-795) + ComputeImageBounds result (minX..maxY) +
 * AssignFeaturesToGrid.  Host arrays; synchronous. */
DVM_API int dvm_frame_assign(dvm_frame* f, const dvm_keypoint* kps_un, const uint8_t* desc, int n, float min_x,
                             float min_y, float max_x, float max_y);
/* Same, taking the keypoints/descriptors of the extractor's last result straight from HBM
 * (device to device, enqueued on the frame's stream; no synchronisation). */
DVM_API int dvm_frame_assign_from_orb(dvm_frame* f, const dvm_orb* orb, float min_x, float min_y, float max_x,
                                      float max_y);
/* The mono Frame constructor on the device (O3/src/Frame.cc:371-479 with zero distortion): ExtractORB(0, im,
 * 0, 1000) straight into the frame's own buffers, then AssignFeaturesToGrid -- all enqueued on the
 * EXTRACTOR's stream, no synchronisation.  The frame's own stream must be ordered after it by the caller
 * (cudaStreamWaitEvent; dvm_tracker does this). */
DVM_API int dvm_frame_construct_device(dvm_frame* f, dvm_orb* orb, const uint8_t* gray_dev, int width, int height,
                                       int stride, float min_x, float min_y, float max_x, float max_y);
/* Frame::UndistortKeyPoints (O3/src/Frame.cc:791-818): cv::undistortPoints(pts, pts, K, mDistCoef, Mat(), mK) with
 * OpenCV's default five iterations, bit-exact with cv2 4.13; dist5 = k1, k2, p1, p2, k3; the identity when k1 == 0,
 * as in the reference.  In place on the x, y of host keypoints (synchronous). */
DVM_API int dvm_undistort_keypoints(dvm_frame* ctx, dvm_keypoint* kps, int n, const float* K, const float* dist5);
/* The same applied inside dvm_frame_construct_device between ExtractORB and AssignFeaturesToGrid (the frame then
 * holds mvKeysUn); K = NULL or k1 == 0 switches it off. */
DVM_API int dvm_frame_set_distortion(dvm_frame* f, const float* K, const float* dist5);
/* Frame::ComputeImageBounds (O3/src/Frame.cc:820-848): bounds = mnMinX, mnMinY, mnMaxX, mnMaxY. */
DVM_API int dvm_image_bounds(dvm_frame* ctx, const float* K, const float* dist5, int width, int height, float* bounds);
/* Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) (O3/src/Frame.cc:712-770); indices in the
 * reference's order.  *n_out may exceed cap (then only cap entries were written). */
DVM_API int dvm_frame_features_in_area(dvm_frame* f, float x, float y, float r, int min_level, int max_level,
                                       int32_t* out, int cap, int* n_out);
/* mGrid[ix][iy] (O3/include/Frame.h:290) for parity tests. */
DVM_API int dvm_frame_grid_cell(dvm_frame* f, int ix, int iy, int32_t* out, int cap, int* n_out);

/* ------------------------------------------------------------------------------------------------
 * ORBmatcher   -- replaces the projection matchers of ORB_SLAM3::ORBmatcher
 *                 (O3/include/ORBmatcher.h:40-95; TH_HIGH = 100, HISTO_LENGTH = 30)
 * Map points are passed as flat arrays; results are indices into those arrays (-1 = no map point).
 * ---------------------------------------------------------------------------------------------- */

/* int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, float th, bool bMono=true)
 * (O3/src/ORBmatcher.cc:1553-1748).  One entry per last-frame keypoint i: has_mp (mvpMapPoints[i] !=
 * NULL), outlier (mvbOutlier[i]), Xw (GetWorldPos), mp_desc (GetDescriptor, 32 B), mp_obs_pos
 * (Observations() > 0), last_octave (mvKeys[i].octave), last_angle (mvKeysUn[i].angle).
 * qcw (x, y, z, w), tcw: CurrentFrame.GetPose() as the SE3f holds it -- unit_quaternion().coeffs() and translation(),
 * used as they are (no renormalisation): the projection is Sophus' quaternion action x3Dc = Tcw * x3Dw (:1577;
 * O3/Thirdparty/Sophus/sophus/so3.hpp:358-367), not a rotation-matrix product.  K = fx, fy, cx, cy.
 * cur_mp[cur n] receives, per current keypoint, the last-frame index whose map point it now holds
 * (CurrentFrame.mvpMapPoints, which the caller cleared before the call); *nmatches = return value. */
DVM_API int dvm_match_by_projection_last(dvm_frame* cur, const float* qcw, const float* tcw, const float* K,
                                         int last_n, const uint8_t* has_mp, const uint8_t* outlier, const float* Xw,
                                         const uint8_t* mp_desc, const uint8_t* mp_obs_pos,
                                         const int32_t* last_octave, const float* last_angle, float th,
                                         int check_orientation, int32_t* cur_mp, int* nmatches);

/* int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, float th, ...)
 * (O3/src/ORBmatcher.cc:44-205; bFarPoints = false).  The caller lists, in vpMapPoints order, the map
 * points with mbTrackInView set and !isBad(): proj_x/proj_y (mTrackProjX/Y), level
 * (mnTrackScaleLevel), view_cos (mTrackViewCos), descriptor, Observations() > 0.  cur_blocked[i] != 0:
 * keypoint i already holds a map point with Observations() > 0 (NULL = none).  nnratio = mfNNratio.
 * cur_mp[cur n]: index of the map point assigned by THIS call, or -1. */
DVM_API int dvm_match_by_projection_map(dvm_frame* cur, int m, const float* proj_x, const float* proj_y,
                                        const int32_t* level, const float* view_cos, const uint8_t* mp_desc,
                                        const uint8_t* mp_obs_pos, float th, float nnratio,
                                        const uint8_t* cur_blocked, int32_t* cur_mp, int* nmatches);
/* Number of Jacobi rounds the last matcher call on this frame needed (diagnostic). */
DVM_API int dvm_match_last_rounds(dvm_frame* cur);

/* ------------------------------------------------------------------------------------------------
 * ORBmatcher, descriptor matchers without projection
 *   SearchByBoW(KeyFrame*, Frame&, ...)      O3/src/ORBmatcher.cc:214-393  (TrackReferenceKeyFrame, Relocalization)
 *   SearchByBoW(KeyFrame*, KeyFrame*, ...)   O3/src/ORBmatcher.cc:709-834  (loop closing / map merging, the
 *                                            reference's inter-agent descriptor comparison)
 *   SearchForInitialization(F1, F2, ...)     O3/src/ORBmatcher.cc:605-707  (monocular initialisation)
 * ---------------------------------------------------------------------------------------------- */

/* One side of SearchByBoW (host arrays).  The DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>,
 * O3/Thirdparty/DBoW2/DBoW2/FeatureVector.h) is flattened to CSR form in the map's iteration order:
 * node_id ascending, node_start[n_nodes + 1], feat_idx = each node's indices in push_back order (every
 * feature index appears at most once, as DBoW2 guarantees). */
typedef struct dvm_bow_features {
    int32_t n;                 /* number of features (keypoints) */
    const uint8_t* desc;       /* [n * 32]  mDescriptors */
    const float* angle;        /* [n]       mvKeysUn[i].angle (== mvKeys[i].angle) */
    const uint8_t* has_mp;     /* [n]       map point present && !isBad(); NULL = every feature */
    int32_t n_nodes;           /* mFeatVec.size() */
    const uint32_t* node_id;   /* [n_nodes] */
    const int32_t* node_start; /* [n_nodes + 1] */
    const uint32_t* feat_idx;  /* [node_start[n_nodes]] */
} dvm_bow_features;

/* kf_kf == 0: SearchByBoW(pKF = a, F = b): a.has_mp = vpMapPointsKF[i] && !isBad(); b.has_mp is ignored
 *             (every unmatched feature of F is a candidate); a match needs bestDist1 <= TH_LOW.
 * kf_kf != 0: SearchByBoW(pKF1 = a, pKF2 = b): both sides need a map point; bestDist1 < TH_LOW.
 * match12[a.n] / match21[b.n] (either may be NULL) receive the partner's index or -1:
 * vpMapPointMatches[i] of the Frame form is "the map point of a's feature match21[i]", vpMatches12[i] of
 * the KeyFrame form is "the map point of b's feature match12[i]".  ctx: any frame handle (device and
 * stream to run on, staging buffers). */
DVM_API int dvm_match_by_bow(dvm_frame* ctx, int kf_kf, const dvm_bow_features* a, const dvm_bow_features* b,
                             float nnratio, int check_orientation, int32_t* match12, int32_t* match21,
                             int* nmatches);

/* int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched,
 * vector<int>& vnMatches12, int windowSize).  f2 = F2 (assigned frame: grid queries); F1 by its
 * undistorted keypoints and descriptors; prev_matched[n1 * 2] is vbPrevMatched, updated in place;
 * matches12[n1] receives vnMatches12; *nmatches the return value. */
DVM_API int dvm_match_for_initialization(dvm_frame* f2, int n1, const dvm_keypoint* kps1_un, const uint8_t* desc1,
                                         float* prev_matched, int window_size, float nnratio,
                                         int check_orientation, int32_t* matches12, int* nmatches);

/* ---- Sim3-guided matchers of loop closing / map merging (LoopClosing::DetectCommonRegionsFromBoW / ...FromLastKF,
 * O3/src/LoopClosing.cc:823-847 and callers of FindMatchesByProjection).  A Sophus::Sim3f goes over as sim3_q =
 * quaternion().coeffs() (x, y, z, w; squared norm = scale) and sim3_t = translation(), used as stored.  Map points as in
 * dvm_fuse_search: world position, GetNormal(), mfMinDistance / mfMaxDistance (the members), descriptor. */

/* int ORBmatcher::SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints,
 * vector<MapPoint*>& vpMatched, int th, float ratioHamming)  (O3/src/ORBmatcher.cc:395-494) and its overload with
 * vpPointsKFs / vpMatchedKF (:496-603, same matching; the caller copies the keyframe pointers by index).  skip[i] = isBad()
 * or the point is already in vpMatched; kp_matched[kf n] = vpMatched[k] != NULL on entry.  kp_point[kf n] receives the index
 * i of the candidate the call writes into vpMatched[k] (-1: unchanged); *nmatches the return value.  Sequential semantics
 * kept: a keypoint taken by an earlier candidate is not available to a later one. */
DVM_API int dvm_match_by_projection_sim3(dvm_frame* kf, const float* sim3_q, const float* sim3_t, const float* K, int m,
                                         const float* xw, const float* normal, const float* min_dist, const float* max_dist,
                                         const uint8_t* mp_desc, const uint8_t* skip, const uint8_t* kp_matched, int th,
                                         float ratio_hamming, int32_t* kp_point, int* nmatches);

/* int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const Sophus::Sim3f& S12,
 * const float th)  (:1347-1551).  kf1 / kf2 = the keyframes' device twins; q / t = GetPose() of each; s12_q / s12_t = S12;
 * K = pKF1's fx, fy, cx, cy (the reference projects both directions with them, :1349-1352).  Side s, one entry per
 * keypoint: skip (no map point, isBad(), or already matched: vbAlreadyMatched, :1370-1381), the map point's world position,
 * mfMinDistance / mfMaxDistance and descriptor.  match12[kf1 n] receives the keypoint of keyframe 2 whose map point
 * the call writes to vpMatches12[i1] (-1: unchanged); *nfound the return value. */
DVM_API int dvm_match_by_sim3(dvm_frame* kf1, dvm_frame* kf2, const float* q1, const float* t1, const float* q2,
                              const float* t2, const float* s12_q, const float* s12_t, const float* K, const uint8_t* skip1,
                              const float* xw1, const float* min_dist1, const float* max_dist1, const uint8_t* mp_desc1,
                              const uint8_t* skip2, const float* xw2, const float* min_dist2, const float* max_dist2,
                              const uint8_t* mp_desc2, float th, int32_t* match12, int* nfound);

/* The search half of int ORBmatcher::Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints, float th,
 * vector<MapPoint*>& vpReplacePoint)  (:1236-1345; caller LoopClosing::SearchAndFuse).  skip[i] = isBad() or the point is
 * among pKF->GetMapPoints().  best_idx[m] / best_dist[m]: the keypoint each candidate is fused with (bestDist <= TH_LOW) or
 * -1 / 256; the caller then sets vpReplacePoint[i] = pKF->GetMapPoint(best_idx) or adds the observation (:1330-1340). */
DVM_API int dvm_fuse_search_sim3(dvm_frame* kf, const float* sim3_q, const float* sim3_t, const float* K, int m,
                                 const float* xw, const float* normal, const float* min_dist, const float* max_dist,
                                 const uint8_t* mp_desc, const uint8_t* skip, float th, int32_t* best_idx, int32_t* best_dist);

/* The relative geometry SearchForTriangulation derives from the two keyframe poses before it matches
 * (O3/src/ORBmatcher.cc:841-860, mono keyframes), in the reference's own float32 arithmetic: T12 = T1w * T2w.inverse()
 * (Sophus' normalising quaternion product and inverse, O3/Thirdparty/Sophus/sophus/{so3,se3}.hpp), R12 / t12 of it, the
 * fundamental matrix Pinhole::epipolarConstrain rebuilds for every candidate pair, F12 = K1^-T [t12]x R12 K2^-1 with
 * Eigen's 3x3 inverse and left-to-right products (O3/src/CameraModels/Pinhole.cpp:104-110), and the epipole ep =
 * project(T2w * Cw).  q1 / q2 = (x, y, z, w) and t1 / t2 of pKF1->GetPose() / pKF2->GetPose() as stored; K = fx, fy, cx,
 * cy.  Host arithmetic only (no GPU): F12[9] row-major and ep[2] are the arguments of dvm_match_for_triangulation. */
DVM_API int dvm_fundamental_from_poses(const float* q1, const float* t1, const float* q2, const float* t2, const float* K1,
                                       const float* K2, float* F12, float* ep);

/* int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, vector<pair<size_t,size_t>>& vMatchedPairs,
 * const bool bOnlyStereo = false, const bool bCoarse)  (O3/src/ORBmatcher.cc:836-1058; caller
 * LocalMapping::CreateNewMapPoints, O3/src/LocalMapping.cc:519), mono keyframes.  In kf1 / kf2 `has_mp[i]` means
 * GetMapPoint(i) != NULL (such features are skipped; required here).  kps1 / kps2 = mvKeysUn.  F12 (row-major) =
 * K1^-T [t12]x R12 K2^-1 and ep = camera 1's centre projected into image 2, both as the reference computes them
 * from the two poses (:841-849, O3/src/CameraModels/Pinhole.cpp:104-110; the host adapter does this);
 * scale_factors2 / level_sigma2_2 = pKF2->mvScaleFactors / mvLevelSigma2 (nlevels entries).
 * matches12[kf1->n] receives vMatches12 (the caller builds vMatchedPairs from its entries >= 0). */
DVM_API int dvm_match_for_triangulation(dvm_frame* ctx, const dvm_bow_features* kf1, const dvm_keypoint* kps1,
                                        const dvm_bow_features* kf2, const dvm_keypoint* kps2, const float* F12,
                                        const float* ep, const float* scale_factors2, const float* level_sigma2_2,
                                        int nlevels, int coarse, int check_orientation, int32_t* matches12,
                                        int* nmatches);

/* The search half of int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th,
 * const bool bRight = false)  (O3/src/ORBmatcher.cc:1060-1228; caller LocalMapping::SearchInNeighbors,
 * O3/src/LocalMapping.cc:818-849).  kf = the keyframe's device twin (mvKeysUn, mDescriptors, grid: an assigned
 * dvm_frame).  pose_q / pose_t = pKF->GetPose(); K = fx, fy, cx, cy; per map point: world position, GetNormal(),
 * mfMinDistance / mfMaxDistance (the 0.8 / 1.2 invariance factors are applied inside), descriptor, and
 * skip[i] = !pMP || isBad() || IsInKeyFrame(pKF).  best_idx[m] / best_dist[m] receive the keypoint each map point
 * would be fused with (bestDist <= TH_LOW), or -1 / 256.  The search does not read the keyframe's map points, so
 * the caller applies Replace / AddObservation / AddMapPoint (:1209-1222) afterwards in vpMapPoints order. */
DVM_API int dvm_fuse_search(dvm_frame* kf, const float* pose_q, const float* pose_t, const float* K, int m,
                            const float* xw, const float* normal, const float* min_dist, const float* max_dist,
                            const uint8_t* mp_desc, const uint8_t* skip, float th, int32_t* best_idx,
                            int32_t* best_dist);

/* ------------------------------------------------------------------------------------------------
 * DBoW2 vocabulary transform   -- Frame::ComputeBoW / KeyFrame::ComputeBoW (O3/src/Frame.cc:784-789):
 * ORBVocabulary::transform(features, mBowVec, mFeatVec, 4)
 * (O3/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1025-1146)
 * ---------------------------------------------------------------------------------------------- */
typedef struct dvm_vocabulary dvm_vocabulary;
/* The k-ary tree, flat (what loadFromTextFile builds, TemplatedVocabulary.h:1211-1287): node 0 is the root; the
 * children of node i are children[child_start[i] .. child_start[i+1]) in push_back order; a node without children
 * is a word with word_id[i] >= 0 and weight[i]; desc[n_nodes * 32] holds the node descriptors; L = m_L. */
DVM_API int dvm_vocabulary_create(dvm_vocabulary** out, int device, int n_nodes, const int32_t* child_start,
                                  const int32_t* children, const uint8_t* desc, const double* weight,
                                  const int32_t* word_id, int L);
DVM_API void dvm_vocabulary_destroy(dvm_vocabulary* v);
/* transform(feature, word_id, weight, &nid, levelsup) for n descriptors (host arrays, synchronous).  The caller
 * assembles BowVector (addWeight / addIfNotExist per feature in index order, then normalize) and FeatureVector
 * (addFeature(nid, i)) from the three outputs, skipping features with weight <= 0 (stopped words). */
DVM_API int dvm_vocabulary_transform(dvm_vocabulary* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id,
                                     double* weight, int32_t* node_id);

/* ------------------------------------------------------------------------------------------------
 * Exhaustive nearest / second-nearest Hamming search (DescriptorDistance, O3/src/ORBmatcher.cc:1900-1914)
 * -- the inter-agent loop-closure exchange of config C3: every received keyframe's descriptors against
 * every local keyframe's, without the vocabulary prefilter (see csrc/hamming.cu).
 * ---------------------------------------------------------------------------------------------- */
typedef struct dvm_hamming dvm_hamming;
/* cuda_stream: a cudaStream_t to launch on, or NULL for a private stream. */
DVM_API int dvm_hamming_create(dvm_hamming** out, int device, void* cuda_stream);
DVM_API void dvm_hamming_destroy(dvm_hamming* h);
/* Host arrays a[na * 32], b[nb * 32].  Per row of a: best_idx (first of the nearest rows of b, -1 if
 * none is closer than 256), best_dist, second_dist (256 when absent) -- the matchers' bestIdx /
 * bestDist1 / bestDist2 over all of b.  Synchronous. */
DVM_API int dvm_hamming_knn(dvm_hamming* h, const uint8_t* a, int na, const uint8_t* b, int nb, int32_t* best_idx,
                            int32_t* best_dist, int32_t* second_dist);
/* Batched, device-resident: a_dev [ba][na][32] against b_dev [bb][nb][32] (nb < 2^20).  key1_dev /
 * key2_dev [ba][bb][na] receive distance << 20 | index of the nearest row and the second-smallest
 * such key ((256 << 20) when absent); counts_dev [ba][bb] (may be NULL) the number of rows with
 * best_dist <= th_low && best_dist < nnratio * second_dist -- the score by which the exchange ranks
 * keyframe pairs.  Enqueues on the handle's stream; no synchronisation. */
DVM_API int dvm_hamming_knn_device(dvm_hamming* h, const uint8_t* a_dev, int ba, int na, const uint8_t* b_dev, int bb,
                                   int nb, uint32_t* key1_dev, uint32_t* key2_dev, int32_t* counts_dev, int th_low,
                                   float nnratio);
DVM_API int dvm_hamming_sync(dvm_hamming* h);
/* Kernel selection of dvm_hamming_knn[_device]: 0 (default) = by size -- batched keyframe blocks go to the tcgen05 int8
 * kernel (csrc/hamming_tc.cu: a . b = 256 - 2 * distance over +1 / -1 bytes, exact in int32), small calls to the
 * popcount kernel; 1 = popcount; 2 = tensor cores; 3 / 4 = timing probes that write no outputs (3: TMA + MMA pipeline
 * alone, 4: plus the TMEM read-out without the key bookkeeping).  Results of modes 0, 1 and 2 are bit-identical. */
DVM_API int dvm_hamming_set_mode(dvm_hamming* h, int mode);

/* ------------------------------------------------------------------------------------------------
 * Inter-agent loop-closure exchange (BASELINE.json config C3), one agent per rank / GPU.  Replaces the keyframe push of
 * OrbSlam3Wrapper::sendNewKeyFrameBows / receiveNewKeyFrameBows (src/slam_system/src/orb_slam3_wrapper.cpp:457-618):
 * every agent sends the descriptor blocks of its not-yet-sent keyframes (at least MIN_BOW_SHARE_SIZE of them, :37) to
 * the rank that owns the agent pair -- the lower id, the reference's lead-node rule (isLeadNodeInGroup, :1238-1243), or
 * a balanced assignment -- over grouped ncclSend / ncclRecv (NVLink / NVSwitch), and the owner matches every received
 * keyframe against its whole database with the exhaustive Hamming search (csrc/exchange.cu, csrc/hamming_tc.cu).
 * ---------------------------------------------------------------------------------------------- */
typedef struct dvm_exchange dvm_exchange;
/* Test hooks (host logic without GPUs): with hooks the database lives in host memory, `alltoall` carries both exchanges
 * (send: one contiguous buffer, send_bytes[world] / recv_bytes[world] per peer, recv: contiguous in rank order) and
 * `match_counts` fills counts[ka][kb] = accepted descriptor matches of block a[i] against db[j].  Return 0 on success.
 * The library has no CPU matcher of its own. */
typedef struct dvm_exchange_hooks {
    void* ctx;
    int (*alltoall)(void* ctx, const void* send, const size_t* send_bytes, void* recv, const size_t* recv_bytes);
    int (*match_counts)(void* ctx, const uint8_t* a, int ka, const uint8_t* db, int kb, int n_feat, int th_low, float nnratio,
                        int32_t* counts);
} dvm_exchange_hooks;
/* ncclGetUniqueId into id128[128]: rank 0 calls it and hands the bytes to its peers over its own channel. */
DVM_API int dvm_exchange_unique_id(uint8_t* id128);
/* One handle per agent.  world > 1 creates the NCCL communicator from id128 (collective: every rank calls it).
 * balanced_ownership = 0: the lower agent id of a pair matches (the reference's rule); 1: pairs are spread over the ranks.
 * cuda_stream: stream to work on, or NULL for a private one.  hooks: NULL in production. */
DVM_API int dvm_exchange_create(dvm_exchange** out, int device, int rank, int world, const uint8_t* id128, int n_feat,
                                int max_keyframes, int balanced_ownership, void* cuda_stream, const dvm_exchange_hooks* hooks);
DVM_API void dvm_exchange_destroy(dvm_exchange* x);
/* th_low / nnratio: a descriptor match is accepted when best <= th_low and best < nnratio * second best (defaults 50,
 * 0.75); min_matches: accepted matches a keyframe pair needs to be reported (default 20 = nBoWMatches,
 * O3/src/LoopClosing.cc:647,751); min_share: fewer new keyframes than this are held back (default 5). */
DVM_API int dvm_exchange_set_policy(dvm_exchange* x, int th_low, float nnratio, int min_matches, int min_share);
/* Appends n_kf keyframes (u8[n_kf][n_feat][32], host or device memory) to this agent's database; *first_id = id of the first. */
DVM_API int dvm_exchange_add_keyframes(dvm_exchange* x, const uint8_t* desc, int n_kf, int desc_is_device, int* first_id);
/* Empties the database and the sent / received bookkeeping; the communicator stays (every rank must do the same). */
DVM_API int dvm_exchange_reset(dvm_exchange* x);
DVM_API int dvm_exchange_keyframes(const dvm_exchange* x);
DVM_API const uint8_t* dvm_exchange_database(const dvm_exchange* x);
/* counts[ka][keyframes()] (host) = accepted descriptor matches of the ka blocks at `a` (device memory; host with hooks)
 * against every keyframe of the database. */
DVM_API int dvm_exchange_match_counts(dvm_exchange* x, const uint8_t* a, int ka, int32_t* counts);
/* One exchange + matching round (collective over the ranks).  candidates[cap][4] receives (peer rank, peer keyframe id,
 * own keyframe id, accepted matches), best first, for the pairs this rank owns; *n_candidates the total found. */
DVM_API int dvm_exchange_round(dvm_exchange* x, int32_t* candidates, int cap, int* n_candidates);
DVM_API size_t dvm_exchange_last_bytes_sent(const dvm_exchange* x);
DVM_API float dvm_exchange_last_match_ms(const dvm_exchange* x);

/* ------------------------------------------------------------------------------------------------
 * Optimizer::PoseOptimization   (O3/src/Optimizer.cc:744-1028, mono observations)
 * ---------------------------------------------------------------------------------------------- */
/* pose_q (x,y,z,w of Tcw.unit_quaternion()) and pose_t (Tcw.translation()) are in/out (float, like
 * Frame::GetPose/SetPose); K = fx,fy,cx,cy; per correspondence the map point's world position, the
 * undistorted keypoint and mvInvLevelSigma2[octave].  outlier[n] receives mvbOutlier; *n_inliers the
 * return value (nInitialCorrespondences - nBad; 0 when n < 3).  stats (may be NULL): [0] LM
 * iterations, [1] LM trials summed over the 4 rounds. */
DVM_API int dvm_pose_optimization(dvm_frame* ctx, float* pose_q, float* pose_t, const float* K, int n,
                                  const float* Xw, const float* kp_xy, const float* inv_sigma2, uint8_t* outlier,
                                  int* n_inliers, int* stats);

/* ------------------------------------------------------------------------------------------------
 * Per-frame tracker: the call order of Tracking::TrackWithMotionModel + TrackLocalMap
 * (O3/src/Tracking.cc:2584-2666, 2668-2768, 3041-3106) with every operator above chained on the GPU.
 * The pointer-graph state machine itself (keyframe decisions, relocalisation, map management) stays
 * on the host, as in the reference; this object only removes the host round trips between
 * ExtractORB -> Frame -> SearchByProjection(last) -> PoseOptimization -> isInFrustum /
 * SearchByProjection(local map) -> PoseOptimization for a map snapshot resident in HBM.
 * ---------------------------------------------------------------------------------------------- */
typedef struct dvm_tracker dvm_tracker;

/* map snapshot (the local map, flattened): world position, representative descriptor, mean viewing
 * direction (GetNormal) and mfMinDistance / mfMaxDistance of each map point.  K = fx, fy, cx, cy;
 * bounds = mnMinX, mnMinY, mnMaxX, mnMaxY.  The tracker launches on the extractor's stream. */
DVM_API int dvm_tracker_create(dvm_tracker** out, dvm_orb* orb, const float* K, const float* bounds, int map_n,
                               const float* map_xw, const uint8_t* map_desc, const float* map_normal,
                               const float* map_min_dist, const float* map_max_dist);
DVM_API void dvm_tracker_destroy(dvm_tracker* t);
/* mDistCoef of the agent's camera (k1, k2, p1, p2, k3): every frame the tracker builds is undistorted on the device
 * (call before dvm_tracker_bootstrap; `bounds` given at creation must then come from dvm_image_bounds). */
DVM_API int dvm_tracker_set_distortion(dvm_tracker* t, const float* dist5);
/* Bootstrap: extract `gray` (host image), set the frame's pose (q = x,y,z,w; t) and associate its
 * keypoints with map points by projection + descriptor search around the given pose (what the
 * reference's initialisation / relocalisation leaves behind: a last frame with map points). */
DVM_API int dvm_tracker_bootstrap(dvm_tracker* t, const uint8_t* gray, int width, int height, int stride,
                                  const float* pose_q, const float* pose_t, int* n_matched);
/* Track one frame.  gray_is_device != 0: the image is already in HBM.  prior_q/prior_t: the pose
 * prior mVelocity * mLastFrame.GetPose() computed by the caller, or NULL to let the tracker apply the
 * same constant-velocity model on the device from its last two poses.  With sync == 0 the call only
 * enqueues (results are read later with dvm_tracker_result); with sync != 0 it returns the pose
 * (pose_out[7] = qx,qy,qz,qw,tx,ty,tz) and counts[4] = {keypoints, matches after
 * TrackWithMotionModel, inliers of the first PoseOptimization, mnMatchesInliers}. */
DVM_API int dvm_tracker_track(dvm_tracker* t, const uint8_t* gray, int gray_is_device, int width, int height,
                              int stride, const float* prior_q, const float* prior_t, int sync, float* pose_out,
                              int32_t* counts);
DVM_API int dvm_tracker_result(dvm_tracker* t, float* pose_out, int32_t* counts);
/* The result of the frame `lag` frames before the last one handed to dvm_tracker_track (lag 0 = dvm_tracker_result;
 * lag <= 2).  It waits for THAT frame's chain only, so a caller that enqueues frame k + 1 first and then reads frame k
 * (lag = 1) gets every frame's pose without ever leaving the GPU idle between two frames: the chain's last kernel
 * writes pose and counts straight into a pinned, mapped ring (the role of mCurrentFrame.GetPose() /
 * mnMatchesInliers after Tracking::Track, O3/src/Tracking.cc:2057-2150). */
DVM_API int dvm_tracker_result_lag(dvm_tracker* t, int lag, float* pose_out, int32_t* counts);
/* Frame pipelining: ExtractORB does not depend on the previous frame's pose, so the tracker runs it on
 * the extractor's stream while the tracking chain of the previous frame runs on its own stream.
 * dvm_tracker_prefetch enqueues the upload + extraction of the NEXT frame not yet handed over (call it after
 * dvm_tracker_track(..., sync = 0) of the current frame and before dvm_tracker_result); at most two frames
 * may be pending (the tracker alternates two extractors on two streams, so consecutive frames are extracted
 * concurrently).  A dvm_tracker_track call with pending frames ignores its image arguments (gray may be NULL)
 * and uses the oldest pending frame.  With sync == 0 and no prefetch the same overlap happens whenever the host
 * runs ahead. */
DVM_API int dvm_tracker_prefetch(dvm_tracker* t, const uint8_t* gray, int gray_is_device, int width, int height,
                                 int stride);
/* The CUDA stream (cudaStream_t) of the tracking chain, for CUDA-event timing by the caller. */
DVM_API void* dvm_tracker_stream(const dvm_tracker* t);
/* Device time of the tracked-frame chain by segment (a measurement aid: with profiling on, every dvm_tracker_track
 * call records CUDA events between the operators and waits for the frame).  segment_ms[5] accumulates, in the chain's
 * order: constant-velocity prior + per-frame reset | SearchByProjection(cur, last) incl. the 2*th retry |
 * PoseOptimization | SearchLocalPoints (isInFrustum + SearchByProjection(F, map points)) | PoseOptimization. */
DVM_API int dvm_tracker_set_profiling(dvm_tracker* t, int on);
DVM_API int dvm_tracker_get_profile(const dvm_tracker* t, double* segment_ms, long long* frames);
/* Current frame's association after the last track call, for parity tests: cur_map[n] = map point
 * index per keypoint (-1 none), outlier[n] = mvbOutlier. */
DVM_API int dvm_tracker_debug_matches(dvm_tracker* t, int32_t* cur_map, uint8_t* outlier, int cap, int* n_out);

/* Frame::isInFrustum (O3/src/Frame.cc:575-636) + MapPoint::PredictScale (O3/src/MapPoint.cc:573-587)
 * for a batch of map points, through the frame's device context.  pose q,t = the frame's Tcw (float).
 * skip[m] != 0 marks points to leave out (mnLastFrameSeen == frame id, isBad).  Outputs per point:
 * in_view, proj_x, proj_y, level, view_cos (as the reference stores them in the MapPoint). */
DVM_API int dvm_frame_is_in_frustum(dvm_frame* f, const float* pose_q, const float* pose_t, const float* K, int m,
                                    const float* xw, const float* normal, const float* min_dist,
                                    const float* max_dist, const uint8_t* skip, float viewing_cos_limit,
                                    uint8_t* in_view, float* proj_x, float* proj_y, int32_t* level, float* view_cos);

/* ------------------------------------------------------------------------------------------------
 * Optimizer::LocalBundleAdjustment   (O3/src/Optimizer.cc:1030-1387, mono observations)
 * ---------------------------------------------------------------------------------------------- */
typedef struct dvm_lba dvm_lba;

/* One solver context per agent/GPU; buffers grow on demand.  max_free_cameras (1..2000) is a SIZING HINT for the
 * one-cluster solve of the dense reduced camera system (6 unknowns per free keyframe), not a limit: windows of up to 100 free
 * keyframes that fit the hint are factored by one 8-CTA cluster from shared memory (local BA: lowest latency), anything
 * larger -- global BA over hundreds of keyframes -- by a blocked Cholesky over the whole GPU.  Calls with more than 2000
 * free keyframes return DVM_ERR_CAPACITY (a 12000 x 12000 dense system; the reference's sparse solver has no such bound), and
 * so do maps whose landmarks have 2^31 or more observation pairs in total (sum over the landmarks of k (k + 1) / 2 for k
 * observations: the Schur complement's work lists hold one 16-byte entry per pair). */
DVM_API int dvm_lba_create(dvm_lba** out, int device, int max_free_cameras);
DVM_API void dvm_lba_destroy(dvm_lba* h);

/* Intrinsics per camera for the NEXT dvm_local_ba / dvm_bundle_adjustment / dvm_merge_ba call on this handle (consumed by
 * it): cam_K[nc][4] = fx, fy, cx, cy of every camera, as the reference projects each edge with its own keyframe's camera
 * (e->pCamera = pKFi->mpCamera, O3/src/Optimizer.cc:1219) -- merged maps mix keyframes of agents with different
 * calibrations.  Without it (or when nc does not match) every camera uses the call's K. */
DVM_API int dvm_lba_set_camera_intrinsics(dvm_lba* h, int nc, const float* cam_K);

/* The flattened local window that LocalBundleAdjustment assembles (O3/src/Optimizer.cc:1033-1304):
 *   cam_q[nc*4] (x,y,z,w) / cam_t[nc*3]: keyframe poses Tcw as float (KeyFrame::GetPose), in/out;
 *   cam_fixed[nc]: 1 for lFixedCameras and for the map's initial keyframe (vSE3->setFixed);
 *   pts[np*3]: MapPoint::GetWorldPos as float, in/out;
 *   one mono observation per edge: edge_cam, edge_pt (indices into the arrays above), edge_obs
 *   (kpUn.pt), edge_inv_sigma2 (mvInvLevelSigma2[octave]);  K = fx, fy, cx, cy.
 * Runs optimizer.optimize(iterations) (the reference passes 10) of g2o's Levenberg-Marquardt with the
 * Schur complement, entirely on the GPU.  abort_flag (may be NULL) is the reference's pbStopFlag (a
 * one-byte `bool` written by the tracking thread): it is checked before starting and polled between LM
 * iterations and trials.
 * Outputs: updated cam_q/cam_t (free cameras) and pts as float; edge_chi2[ne] (may be NULL);
 * edge_bad[ne] = chi2 > 5.991 || depth <= 0, i.e. the observations the caller must erase
 * (:1313-1329); stats[4] (may be NULL) = {LM iterations, LM trials, initial robust chi2, final robust
 * chi2}.  *iters_done = -1 when the call was a no-op (no fixed camera, abort already set). */
DVM_API int dvm_local_ba(dvm_lba* h, int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts,
                         int ne, const int32_t* edge_cam, const int32_t* edge_pt, const float* edge_obs,
                         const float* edge_inv_sigma2, const float* K, int iterations, const volatile uint8_t* abort_flag,
                         double* edge_chi2, uint8_t* edge_bad, double* stats, int* iters_done);
/* void Optimizer::BundleAdjustment(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust)
 * (O3/src/Optimizer.cc:55-356; GlobalBundleAdjustemnt :46-53 calls it with every keyframe and map point of the map)
 * for maps of up to 2000 free keyframes: the same solver with the caller's Huber delta -- the
 * reference uses (float)sqrt(5.99) here, not LocalBundleAdjustment's sqrt(5.991) -- or INFINITY for bRobust == false.
 * cam_fixed marks the map's initial keyframe (:116); map points without observation are left out by the caller
 * (:243-247).  Outputs as dvm_local_ba (the caller writes mTcwGBA / mPosGBA and ignores edge_bad). */
DVM_API int dvm_bundle_adjustment(dvm_lba* h, int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np,
                                  float* pts, int ne, const int32_t* edge_cam, const int32_t* edge_pt,
                                  const float* edge_obs, const float* edge_inv_sigma2, const float* K, int iterations,
                                  float huber_delta, const volatile uint8_t* abort_flag, double* edge_chi2,
                                  uint8_t* edge_bad, double* stats, int* iters_done);
/* void Optimizer::LocalBundleAdjustment(KeyFrame* pMainKF, vector<KeyFrame*> vpAdjustKF, vector<KeyFrame*> vpFixedKF,
 * bool* pbStopFlag)   (O3/src/Optimizer.cc:3257-3675, mono observations) -- the welding BA of a map merge
 * (LoopClosing::MergeLocal).  cam_fixed = 1 for vpFixedKF, 0 for vpAdjustKF; at least one fixed keyframe.  Both passes run
 * in ONE kernel with the estimates kept in double between them, as the reference's optimizer object does:
 *   pass 1: optimize(5), Huber delta (float)sqrt(5.99) (:3362,3476);
 *   unless *abort_flag is up: edges with chi2 > 5.991 or depth <= 0 go to level 1, every edge loses its robust kernel,
 *   pass 2: optimize(10) over the level-0 edges (:3481-3521).
 * Outputs as dvm_local_ba; edge_chi2 of a level-1 edge is its chi2 after pass 1 (g2o keeps the edge's last error) while
 * the depth in edge_bad is taken at the final estimate, exactly what the reference's final test reads (:3530-3546).
 * stats[6] = {LM iterations of both passes, LM trials, initial robust chi2, final chi2 of the last pass, LM iterations of
 * pass 1, edges moved to level 1}. */
DVM_API int dvm_merge_ba(dvm_lba* h, int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts,
                         int ne, const int32_t* edge_cam, const int32_t* edge_pt, const float* edge_obs,
                         const float* edge_inv_sigma2, const float* K, const volatile uint8_t* abort_flag,
                         double* edge_chi2, uint8_t* edge_bad, double* stats, int* iters_done);
/* ------------------------------------------------------------------------------------------------
 * int Optimizer::OptimizeSim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches1, g2o::Sim3& g2oS12,
 *                             const float th2, const bool bFixScale, Eigen::Matrix<double,7,7>& mAcumHessian,
 *                             const bool bAllPoints)            (O3/src/Optimizer.cc:1960-2212)
 * called by LoopClosing::DetectCommonRegionsFromBoW / DetectAndReffineSim3FromLastKF for loop and merge candidates.
 * The caller flattens the correspondences that receive an edge pair (:2008-2146):
 *   p1c[n*3], p2c[n*3]: P3D1c = R1w*P1w + t1w and P3D2c = R2w*P2w + t2w as float;
 *   obs1[n*2]: kpUn1.pt;  obs2[n*2]: kpUn2.pt, or (x/z, y/z) of P3D2c when the point has no keypoint in KF2 (:2118-2125);
 *   inv_sigma2_1/2[n]: mvInvLevelSigma2[octave] (octave 0 for the projected case);  K1/K2: fx, fy, cx, cy of pCamera1/2;
 *   s12_q (x,y,z,w) / s12_t / s12_s: g2oS12 in/out as double;  th2 and fix_scale as the reference.
 * Runs optimize(5) with Huber delta (float)sqrt(th2), drops the pairs with chi2 > th2 on either edge and the robust
 * kernels of the rest, optimize(10 or 5), and the final inlier test -- g2o's Levenberg-Marquardt with the numeric
 * (central-difference, delta 1e-9) Jacobians the reference's edges get -- in one kernel launch.
 * Outputs: inlier[n] = 0 where the reference resets vpMatches1[idx]; *n_in = the return value (0 with s12 untouched when
 * fewer than 10 correspondences survive the first pass); stats[6] (may be NULL) = {LM iterations pass 1, pass 2, LM
 * trials, pairs dropped after pass 1, initial robust chi2, final chi2}.  mAcumHessian is zero in the reference (:2191). */
typedef struct dvm_sim3 dvm_sim3;
DVM_API int dvm_sim3_create(dvm_sim3** out, int device);
DVM_API void dvm_sim3_destroy(dvm_sim3* h);
DVM_API int dvm_optimize_sim3(dvm_sim3* h, int n, const float* p1c, const float* p2c, const float* obs1, const float* obs2,
                              const float* inv_sigma2_1, const float* inv_sigma2_2, const float* K1, const float* K2,
                              double* s12_q, double* s12_t, double* s12_s, float th2, int fix_scale, uint8_t* inlier,
                              int* n_in, double* stats);

/* Device time of the last dvm_local_ba kernel in milliseconds (CUDA events on the solver's stream). */
DVM_API float dvm_lba_last_kernel_ms(const dvm_lba* h);

/* ------------------------------------------------------------------------------------------------
 * The solve of Optimizer::OptimizeEssentialGraph (O3/src/Optimizer.cc:1389-1651; the 4-DoF variant of :1653 differs only
 * in its vertex / edge types and is not covered) on a flattened pose graph: Sim3 vertices, EdgeSim3 constraints with
 * identity information, g2o's numeric Jacobians, Levenberg-Marquardt with setUserLambdaInit(lambda_init = 1e-16) for
 * `iterations` (20) iterations -- one cooperative kernel, the reduced system factored densely over the whole GPU
 * (csrc/essential_graph.cu).  sim3[nv][8] = Scw of every keyframe as (q x,y,z,w; t; s), in/out; fixed[nv] (the initial /
 * loop keyframe: vSim3->setFixed(true), :1448-1449); edges (vi, vj, meas[ne][8]): vertex(0) = vi, vertex(1) = vj,
 * measurement Sji (:1479-1592).  fix_scale = bFixScale (true for mono after inertial initialisation; false here).
 * stats[4] = LM iterations, trials, initial chi2, final chi2.  The caller writes the poses back and corrects the map
 * points (:1602-1650).
 * ---------------------------------------------------------------------------------------------- */
typedef struct dvm_essential_graph dvm_essential_graph;
DVM_API int dvm_essential_graph_create(dvm_essential_graph** out, int device);
DVM_API void dvm_essential_graph_destroy(dvm_essential_graph* h);
DVM_API int dvm_optimize_essential_graph(dvm_essential_graph* h, int nv, double* sim3, const uint8_t* fixed, int ne,
                                         const int32_t* vi, const int32_t* vj, const double* meas, int fix_scale,
                                         int iterations, double lambda_init, double* stats);
DVM_API float dvm_essential_graph_last_kernel_ms(const dvm_essential_graph* h);

#ifdef __cplusplus
}
#endif
#endif /* DVMSLAM_B200_H */
