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
    int* cell_start;       // [kGridCells + 1]
    int* cell_items;       // [cap]
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
};

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
};

void launch_grid_build(const FrameDev& f, cudaStream_t stream);
void launch_match_last(const FrameDev& cur, const MatchLastArgs& a, const MatchScratch& s, int* cur_mp, int* nmatches,
                       cudaStream_t stream);
void launch_match_map(const FrameDev& cur, const MatchMapArgs& a, const MatchScratch& s, int* cur_mp, int* nmatches,
                      cudaStream_t stream);
void launch_pose_opt(const PoseOptArgs& a, cudaStream_t stream);
void launch_features_in_area(const FrameDev& f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap,
                             int* n_out, cudaStream_t stream);

} // namespace dvm
