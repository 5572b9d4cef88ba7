// track_kernels.cuh -- device structures and launchers of the per-frame tracking operators
// (Frame grid, projection matchers, pose-only optimisation).
#pragma once
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
    int* kxyo;             // [cap * 3] compact {x bits, y bits, octave} records written by the grid build
    int nlevels;
    float scale[kTrackMaxLevels];       // mvScaleFactors
    float inv_sigma2[kTrackMaxLevels];  // mvInvLevelSigma2
};

// SearchByProjection(CurrentFrame, LastFrame): flat last-frame inputs (device pointers)
struct MatchLastArgs {
    float R[9], t[3], K[4];
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
    const float* pose;             // qx,qy,qz,qw,tx,ty,tz on the device instead of R,t
    const int* guard;              // skip the whole search if *guard >= 20 (the 2*th retry, Tracking.cc:2614)
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
};

struct MatchScratch {
    float* pu;      // [cap_q] projected u
    float* pv;      // [cap_q]
    float* pr;      // [cap_q] radius
    int* plevels;   // [cap_q] minLevel << 16 | (maxLevel & 0xffff), or -1 when the query is skipped
    int* choice;    // [cap_q] chosen keypoint per query
    int* claim_a;   // [cap_kp]
    int* claim_b;   // [cap_kp]
    int* iters;     // [1] fixed-point rounds used (diagnostic)
    unsigned long long* cache; // [cap_q * kMatchCacheK] best candidates of each query, sorted by (distance, walk order)
    int* ncand;     // [cap_q] candidates seen by the window walk
    int* qlist;     // [cap_q] ordered list of the queries that take part
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
// quaternion (x,y,z,w, float) -> row-major rotation matrix, Eigen's toRotationMatrix in float32 after
// normalisation: the convention shared with oracle/track_oracle.cpp (trko_is_in_frustum)
__device__ inline void quat_to_R_f32(const float* q_in, float R[9])
{
    const float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q_in[0], q_in[0]), __fmul_rn(q_in[1], q_in[1])),
                                                   __fmul_rn(q_in[2], q_in[2])), __fmul_rn(q_in[3], q_in[3])));
    const float x = __fdiv_rn(q_in[0], n), y = __fdiv_rn(q_in[1], n), z = __fdiv_rn(q_in[2], n), w = __fdiv_rn(q_in[3], n);
    const float tx = __fmul_rn(2.f, x), ty = __fmul_rn(2.f, y), tz = __fmul_rn(2.f, z);
    const float twx = __fmul_rn(tx, w), twy = __fmul_rn(ty, w), twz = __fmul_rn(tz, w);
    const float txx = __fmul_rn(tx, x), txy = __fmul_rn(ty, x), txz = __fmul_rn(tz, x);
    const float tyy = __fmul_rn(ty, y), tyz = __fmul_rn(tz, y), tzz = __fmul_rn(tz, z);
    R[0] = __fsub_rn(1.f, __fadd_rn(tyy, tzz)); R[1] = __fsub_rn(txy, twz); R[2] = __fadd_rn(txz, twy);
    R[3] = __fadd_rn(txy, twz); R[4] = __fsub_rn(1.f, __fadd_rn(txx, tzz)); R[5] = __fsub_rn(tyz, twx);
    R[6] = __fsub_rn(txz, twy); R[7] = __fadd_rn(tyz, twx); R[8] = __fsub_rn(1.f, __fadd_rn(txx, tyy));
}

// Frame::isInFrustum (mono branch, O3/src/Frame.cc:575-636) + MapPoint::PredictScale (O3/src/MapPoint.cc:573-587)
// for map point k with rotation R (row-major) and translation pose[4..6]; returns visibility
__device__ inline bool frustum_eval(const FrustumArgs& a, const float R[9], int k, float& u, float& v, int& lvl, float& vc)
{
    u = -1.f; v = -1.f; vc = 0.f; lvl = -1;
    if (a.skip && a.skip[k]) return false;
    const float t0 = a.pose[4], t1 = a.pose[5], t2 = a.pose[6];
    const float X = a.xw[3 * k], Y = a.xw[3 * k + 1], Z = a.xw[3 * k + 2];
    const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], X), __fmul_rn(R[1], Y)), __fmul_rn(R[2], Z)), t0);
    const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], X), __fmul_rn(R[4], Y)), __fmul_rn(R[5], Z)), t1);
    const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], X), __fmul_rn(R[7], Y)), __fmul_rn(R[8], Z)), t2);
    if (zc < 0.0f) return false;
    const float pu = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[0], xc), zc), a.K[2]);
    const float pv = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[1], yc), zc), a.K[3]);
    if ((pu < a.bounds[0] || pu > a.bounds[2]) || (pv < a.bounds[1] || pv > a.bounds[3])) return false;
    u = pu; v = pv;
    const float maxD = __fmul_rn(1.2f, a.max_dist[k]), minD = __fmul_rn(0.8f, a.min_dist[k]);
    float Ow[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
        Ow[i] = __fadd_rn(__fadd_rn(__fmul_rn(R[i], -t0), __fmul_rn(R[3 + i], -t1)), __fmul_rn(R[6 + i], -t2));
    const float p0 = __fsub_rn(X, Ow[0]), p1 = __fsub_rn(Y, Ow[1]), p2 = __fsub_rn(Z, Ow[2]);
    const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(p0, p0), __fmul_rn(p1, p1)), __fmul_rn(p2, p2)));
    if (dist < minD || dist > maxD) return false;
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(p0, a.normal[3 * k]), __fmul_rn(p1, a.normal[3 * k + 1])),
                                __fmul_rn(p2, a.normal[3 * k + 2]));
    const float c = __fdiv_rn(dot, dist);
    if (c < a.cosLimit) return false;
    const float ratio = __fdiv_rn(a.max_dist[k], dist);
    int nScale = (int)ceilf(__fdiv_rn((float)log((double)ratio), a.logScale));
    if (nScale < 0) nScale = 0;
    else if (nScale >= a.nlevels) nScale = a.nlevels - 1;
    lvl = nScale; vc = c;
    return true;
}

void launch_features_in_area(const FrameDev& f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap,
                             int* n_out, cudaStream_t stream);

} // namespace dvm
