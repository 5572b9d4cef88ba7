// track_kernels.cuh -- device structures and launchers of the per-frame tracking operators
// (Frame grid, projection matchers, pose-only optimisation).
#pragma once
#include "sophus_f32.cuh"
#include "common.cuh"

namespace dvm {

constexpr int kGridCols = 64, kGridRows = 48;   // FRAME_GRID_COLS / ROWS, O3/include/Frame.h:44-45
constexpr int kGridCells = kGridCols * kGridRows;
constexpr int kThHigh = 100, kThLow = 50, kHistoLength = 30; // O3/src/ORBmatcher.cc:36-38
constexpr int kTrackMaxLevels = 12;

// Device view of a mono Frame: undistorted keypoints, descriptors, bounds, 64x48 grid in CSR form
// (cell = ix * 48 + iy, items ascending = the reference's push_back order).
struct FrameDev {
    const dvm_keypoint* kps;
    const uint8_t* desc;
    const int* n;          // device int: number of keypoints
    int cap;
    float minX, minY, maxX, maxY, gwInv, ghInv;
    int* cell_start;       // [kGridCells + 1] (allocated kGridCells + 4)
    int* cell_items;       // [cap]
    int4* cell_rec;        // [cap] per grid position j (the order of cell_items): {x bits, y bits, octave, keypoint index}
    int nlevels;
    float scale[kTrackMaxLevels];       // mvScaleFactors
    float inv_sigma2[kTrackMaxLevels];  // mvInvLevelSigma2
};

// SearchByProjection(CurrentFrame, LastFrame): flat last-frame inputs (device pointers)
struct MatchLastArgs {
    float q[4], t[3], K[4];   // Tcw as the frame's SE3f holds it: unit quaternion (x, y, z, w), translation
    int last_n;
    const uint8_t* has_mp;
    const uint8_t* outlier;
    const float* Xw;         // [last_n * 3]
    const uint8_t* mp_desc;  // [last_n * 32]
    const uint8_t* obs_pos;  // Observations() > 0
    const int* octave;       // LastFrame.mvKeys[i].octave
    const float* angle;      // LastFrame.mvKeysUn[i].angle
    float th;
    int check_ori;
    // ---- optional device-resident forms (dvm_tracker): all may be NULL ----
    const int* n_ptr;              // last_n read from the device
    const int* mp_index;           // per last keypoint: index into Xw / mp_desc (map arrays), -1 = no map point
    const dvm_keypoint* last_kps;  // octave / angle source instead of the two arrays
    const float* pose;             // qx,qy,qz,qw,tx,ty,tz on the device instead of q,t
    const int* guard;              // skip the whole search if *guard >= 20 (the 2*th retry as a second launch)
    float retry_th;                // > 0: fewer than 20 matches -> the resolution kernel itself searches again with this th
                                   // (Tracking.cc:2614-2621) and its second result replaces the first
    int* map_out;                  // [cur cap] with mp_index: map point now held by each current keypoint (-1 none)
};

// Frame::isInFrustum (mono) + MapPoint::PredictScale over a flat map snapshot (device pointers)
struct FrustumArgs {
    const float* pose;  // device: qx,qy,qz,qw,tx,ty,tz
    float K[4], bounds[4];
    int nlevels;
    float logScale, cosLimit;
    int m;
    const float* xw; const float* normal; const float* min_dist; const float* max_dist;
    const uint8_t* skip;
    uint8_t* in_view; float* px; float* py; int* level; float* view_cos; // outputs of frustum_kernel
};

// SearchByProjection(F, vpMapPoints): flat in-view map points (device pointers)
struct MatchMapArgs {
    int m;
    const float* projX;
    const float* projY;
    const int* level;
    const float* view_cos;
    const uint8_t* mp_desc;
    const uint8_t* obs_pos;
    float th, nnratio;
    const uint8_t* cur_blocked; // [cur n] or nullptr
    // ---- optional device-resident forms ----
    const int* m_ptr;           // m read from the device
    const int* q_index;         // per query: index into mp_desc (map array)
    const int* cur_map;         // blocked = cur_map[k] >= 0 (instead of cur_blocked)
    // ---- fused SearchLocalPoints (dvm_tracker): queries are ALL map points of the snapshot; each one is
    // run through isInFrustum first (fr.m = m; fr.in_view.. unused) and takes part only if visible.  The
    // query index is then the map index itself, which keeps vpMapPoints order as the greedy priority. ----
    int use_frustum;
    FrustumArgs fr;
    int* merge_into;            // [cur cap] or nullptr: keypoints matched by this call receive their map index here
    // ---- SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (O3/src/ORBmatcher.cc:395-603): level[i] < 0
    // = candidate rejected by the projection gates; radius th * scale[level] (no viewing-cosine factor); the nearest free
    // keypoint is accepted when (float)bestDist <= accept_limit (TH_LOW * ratioHamming), no second-best test ----
    int sim3_mode;
    float accept_limit;
};

struct MatchScratch {
    float* pu;      // [cap_q] projected u
    float* pv;      // [cap_q]
    float* pr;      // [cap_q] radius
    int* plevels;   // [cap_q] (unused by the projection matchers since the compact list carries the level window)
    int* choice;    // [cap_q] chosen keypoint per compact slot (slots beyond the register-resident ones)
    int* claim_a;   // [cap_kp]
    int* claim_b;   // [cap_kp]
    int* iters;     // [1] fixed-point rounds used (diagnostic)
    unsigned long long* cache; // [cap_q * kMatchCacheK] best candidates of each compact slot, sorted by (distance, walk order)
    int* ncand;     // [cap_q] (unused by the projection matchers)
    unsigned* cache8; // [cap_q * kMatchCacheK] the same candidates in 32 bits each (distance:9 | octave:7 | keypoint:16)
    int4* qmeta;    // [cap_q] compact list of the queries that take part, in the order the walk's warps append them:
                    // {query index, level window, candidates seen, 0}; cache rows are in the same compact order
    int* qcount;    // [1] length of the compact list: bumped by the walk, reset to 0 by the resolution kernel
    unsigned long long* prof; // [16] or nullptr: %globaltimer stamps of the kernel phases (DVM_MATCH_PROFILE, diagnostics)
};
constexpr int kMatchCacheK = 8;

struct PoseOptArgs {
    int n;                   // correspondences (<= cap)
    const float* Xw;         // [n*3]
    const float* kp_xy;      // [n*2]
    const float* inv_sigma2; // [n]
    const uint8_t* valid;    // [n] or nullptr: only entries with valid != 0 are edges
    float K[4];
    float* pose;             // [7] in/out: qx,qy,qz,qw,tx,ty,tz
    uint8_t* outlier;        // [n] out
    int* result;             // [4] out: n_inliers, n_edges, lm iterations, lm trials
    double* err;             // [n*2] scratch (the edges' _error)
    // ---- optional device-resident forms: one potential edge per keypoint of a frame ----
    const int* n_ptr;                // number of keypoints on the device
    const int* map_index;            // [n] map point of keypoint k (-1 none); Xw then indexes the map array
    const dvm_keypoint* kps;         // observation and octave source
    float inv_sigma2_table[kTrackMaxLevels]; // mvInvLevelSigma2 (with kps)
    // ---- optional fused tails (dvm_tracker) ----
    // "discard outliers" (Tracking.cc:2634-2654): every matched map point is marked seen; outlier matches
    // are dropped from map_index_rw (= map_index) and their outlier flag cleared
    uint8_t* seen;
    int* map_index_rw;
    // frame hand-over after TrackLocalMap: pose history for the constant-velocity prior and the result block
    float* pose_last; float* pose_prev; float* out_pose; int* out_counts; const int* nm_last; const int* res_first;
    // ... and the start of the NEXT frame, so that its chain begins with the projection search: the constant-velocity prior
    // mVelocity * mLastFrame.GetPose() (Tracking.cc:1990-1991,2598) written over `pose`, the "seen in this frame" marks cleared
    float* next_prior; uint8_t* seen_reset; int seen_n;
    unsigned long long* prof; // [4] or nullptr: ns spent in {edge pass, reduction, solve + update, passes} (DVM_POSE_PROFILE)
};

// cv::undistortPoints (O3/src/Frame.cc:791-818) applied in place to the x, y of n keypoints (n read from the device)
struct UndistortArgs { double k[5]; double fx, fy, cx, cy; };
void launch_undistort(dvm_keypoint* kps, const int* n_ptr, int cap, const UndistortArgs& a, cudaStream_t stream);
void launch_grid_build(const FrameDev& f, cudaStream_t stream);
// copies an extractor result (keypoints, descriptors, count) into the frame's buffers and builds the grid, one kernel
void launch_frame_assign(const FrameDev& f, const dvm_keypoint* src_kps, const uint8_t* src_desc, const int* src_n,
                         cudaStream_t stream);
void launch_frustum(const FrustumArgs& a, cudaStream_t stream);
void launch_match_last(const FrameDev& cur, const MatchLastArgs& a, const MatchScratch& s, int* cur_mp, int* nmatches,
                       cudaStream_t stream);
void launch_match_map(const FrameDev& cur, const MatchMapArgs& a, const MatchScratch& s, int* cur_mp, int* nmatches,
                      cudaStream_t stream);
int launch_pose_opt(const PoseOptArgs& a, cudaStream_t stream);
// Frame::isInFrustum (mono branch, O3/src/Frame.cc:575-636) + MapPoint::PredictScale (O3/src/MapPoint.cc:573-587) for map
// point k.  pose = qx,qy,qz,qw,tx,ty,tz of the frame's SE3f as stored; R = toRotationMatrix(q) and Ow = translation of
// Tcw.inverse() are what Frame::UpdatePoseMatrices (:553-559) caches (FrustumPose, computed once per thread).
struct FrustumPose { float R[9], t[3], Ow[3]; };
__device__ inline void frustum_pose(const float* pose, FrustumPose& fp)
{
    const float q[4] = { pose[0], pose[1], pose[2], pose[3] };
    fp.t[0] = pose[4]; fp.t[1] = pose[5]; fp.t[2] = pose[6];
    so::quat_to_matrix(q, fp.R);
    float qi[4];
    so::se3_inverse(q, fp.t, qi, fp.Ow);
}
__device__ inline bool frustum_eval(const FrustumArgs& a, const FrustumPose& fp, int k, float& u, float& v, int& lvl, float& vc)
{
    u = -1.f; v = -1.f; vc = 0.f; lvl = -1;
    // everything the gates may need is loaded up front (one round trip to L2 instead of one per gate passed)
    const bool skipped = a.skip && a.skip[k];
    const float P[3] = { a.xw[3 * k], a.xw[3 * k + 1], a.xw[3 * k + 2] };
    const float max_dist_k = a.max_dist[k], min_dist_k = a.min_dist[k];
    const float Pn[3] = { a.normal[3 * k], a.normal[3 * k + 1], a.normal[3 * k + 2] };
    if (skipped) return false;
    float Pc[3];
    so::mat_vec(fp.R, P, Pc);                                   // Pc = mRcw * P + mtcw
    const float xc = __fadd_rn(Pc[0], fp.t[0]), yc = __fadd_rn(Pc[1], fp.t[1]), zc = __fadd_rn(Pc[2], fp.t[2]);
    if (zc < 0.0f) return false;
    const float pu = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[0], xc), zc), a.K[2]);
    const float pv = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[1], yc), zc), a.K[3]);
    if ((pu < a.bounds[0] || pu > a.bounds[2]) || (pv < a.bounds[1] || pv > a.bounds[3])) return false;
    u = pu; v = pv;
    const float maxD = __fmul_rn(1.2f, max_dist_k), minD = __fmul_rn(0.8f, min_dist_k);
    const float PO[3] = { __fsub_rn(P[0], fp.Ow[0]), __fsub_rn(P[1], fp.Ow[1]), __fsub_rn(P[2], fp.Ow[2]) };
    const float dist = so::norm3(PO);
    if (dist < minD || dist > maxD) return false;
    const float c = __fdiv_rn(so::dot3(PO, Pn), dist);
    if (c < a.cosLimit) return false;
    lvl = so::predict_scale(max_dist_k, dist, a.logScale, a.nlevels);
    vc = c;
    return true;
}

void launch_features_in_area(const FrameDev& f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap,
                             int* n_out, cudaStream_t stream);

} // namespace dvm
