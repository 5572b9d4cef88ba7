// orb_kernels.cuh -- device-side configuration and kernel entry points of the ORB extractor.
#pragma once
#include "common.cuh"
#include <cuda.h>   // CUtensorMap (the descriptor is encoded through cudaGetDriverEntryPoint: no libcuda link dependency)

namespace dvm {

constexpr int kMaxLevels = 12;
constexpr int kBorder = 16;        // EDGE_THRESHOLD - 3   (O3/src/ORBextractor.cc:618)
constexpr int kEdge = 19;          // EDGE_THRESHOLD
constexpr int kHalfPatch = 15;     // HALF_PATCH_SIZE
constexpr int kCellTilePitch = 112; // smem pitch of one FAST cell tile (sub-image <= 88 px wide, up to 15 px of TMA alignment slack)
constexpr int kCellMaxDim = 88;

struct ResizeX { int sx0, sx1; short a0, a1; };       // per destination column
struct ResizeY { int sy0, sy1; short b0, b1; };       // per destination row
// Four destination columns 4g .. 4g + 3 at once (the vectorised pyramid): the taps of the four pixels lie within 12 source
// bytes from the 4-byte aligned word `base`; sel holds two PRMT selectors (pixels 0-1 in the low half, 2-3 in the high half),
// each picking the bytes (tap0, tap1, tap0, tap1) of its two pixels out of the word pair (w0, w1) or -- window bit set --
// (w1, w2); c[p] = a0 | a1 << 16 feeds IDP.2A with those bytes.
struct ResizeX4 { uint32_t sel, base /* word index | winA << 16 | winB << 17 */, c[4], pad[2]; };

struct OrbLevel {
    const uint8_t* img;   // level image (level 0 may alias the caller's device image)
    int w, h, pitch;
    int width, height;    // maxBorder - minBorder = (w - 32, h - 32)
    int nCols, nRows, wCell, hCell;
    int cell_base;        // index of this level's first cell in the flattened cell grid
    int quota;            // mnFeaturesPerLevel[level]
    int nIni;             // root nodes of the octree
    float hX;
    int cand_off, cand_cap; // slice of the candidate array
    int node_cap;           // capacity of the node list == of the per-level selection
    int sel_off;            // slice of the selection array
    float scale;            // mvScaleFactor[level]
    float size;             // (int)(31 * scale)
    int xtab_off, ytab_off; // resize tables (levels >= 1)
    int x4_off;             // ResizeX4 table of this level (groups of four columns)
};

// The whole resize chain in ONE launch (pyramid_kernel): a CTA owns a tile of the LAST level and computes, level by level in
// shared memory, the region of every level that tile depends on -- the chain dependency of ComputePyramid (level l from
// level l-1, O3/src/ORBextractor.cc:967) stays inside the CTA -- and writes the part of each level it owns.  Per tile
// column / row and level: first pixel of the region, last pixel it owns, last pixel it needs (host-built from the
// resize tables).
constexpr int kPyrTileW = 32, kPyrTileH = 16;  // 12 x 13 = 156 tiles at 720p: ONE wave (312 tiles of 32 x 8 were two waves of ~10 us, each bounded by its seven dependent levels)
struct OrbPyrPlan {
    int ntx, nty;                 // tiles of the last level
    int soff[kMaxLevels];         // shared-memory offset of each level's region buffer
    int spitch[kMaxLevels];       // and its pitch (max region width over the tiles)
    int smem_bytes;
    int vec_ok;                   // the four-pixels-per-thread kernel applies (tap windows fit, regions 4-aligned)
};

struct OrbCfg {
    int nlevels;
    int total_cells;
    int ini_th, min_th;
    int max_kp;           // sum of node_cap
    OrbLevel lv[kMaxLevels];
    OrbPyrPlan pyr;
};

// packed keypoint: response << 24 | y << 12 | x   (x, y relative to the 16-px border origin)
__host__ __device__ inline uint32_t pack_kp(int x, int y, int resp) { return ((uint32_t)resp << 24) | ((uint32_t)y << 12) | (uint32_t)x; }
__host__ __device__ inline int kp_x(uint32_t p) { return p & 0xfff; }
__host__ __device__ inline int kp_y(uint32_t p) { return (p >> 12) & 0xfff; }
__host__ __device__ inline int kp_resp(uint32_t p) { return p >> 24; }

struct OrbBuffers {
    uint32_t* cand;        // [sum cand_cap] packed candidates, unordered within a level
    int* cand_count;       // [2*kMaxLevels]: live counters, then a copy of the last frame's
    uint16_t* pnode;       // [sum cand_cap] node id of each candidate (octree scratch)
    uint32_t* sel;         // [max_kp] packed selection per level, octree list order
    int* sel_count;        // [kMaxLevels]
    uint32_t* work_kp;     // [max_kp] level-major sequence: packed kp
    uint32_t* work_meta;   // [max_kp] level << 24 | destination row
    int* counts;           // [2] = {n keypoints, monoIndex}
    int* status;           // [1] sticky error bits (1 = candidate overflow, 2 = node overflow)
    unsigned int* ticket;  // [1]
    dvm_keypoint* out_kps; // [max_kp]
    uint8_t* out_desc;     // [max_kp * 32]
    const ResizeX* xtab;
    const ResizeY* ytab;
    const ResizeX4* x4tab; // per level, per group of four columns
    const int* pyr_col;    // [ntx][kMaxLevels][3] = {first, last owned, last needed} column of each level's region
    const int* pyr_row;    // [nty][kMaxLevels][3]
    const int8_t* pattern; // [256*4] device copy of the rBRIEF pattern
};

// TMA descriptors of the pyramid levels for the FAST cell tiles: a 2-D u8 tensor (width x height, row pitch) per level
// with a box of kCellTilePitch x (hCell + 6) bytes; a level whose base or pitch is not 16-byte aligned (a caller's
// image read in place) falls back to ordinary loads.
struct alignas(64) OrbTmaps {
    CUtensorMap map[kMaxLevels];
    int use[kMaxLevels];
    int box_h[kMaxLevels];
};

// launches (all on `stream`)
void launch_resize_level(const OrbCfg& cfg, const OrbBuffers& b, int level, uint8_t* dst, cudaStream_t stream);
void launch_pyramid(const OrbCfg& cfg, const OrbBuffers& b, cudaStream_t stream);   // levels 1 .. nlevels-1 in one launch
void prepare_pyramid_kernel(int smem_bytes);
void launch_fast_cells(const OrbCfg& cfg, const OrbBuffers& b, const OrbTmaps& tm, cudaStream_t stream);
// encodes tm.map[level] for the level image currently in cfg (returns false and clears tm.use[level] when TMA cannot
// address it)
bool encode_level_tmap(const OrbCfg& cfg, int level, OrbTmaps& tm);
int octree_smem_bytes(const OrbCfg& cfg);
int prepare_octree_kernel(int smem_bytes);
void launch_octree(const OrbCfg& cfg, const OrbBuffers& b, int lap0, int lap1, int smem_bytes, cudaStream_t stream);
void launch_describe(const OrbCfg& cfg, const OrbBuffers& b, cudaStream_t stream);
void launch_blur_level_debug(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, cudaStream_t stream);

} // namespace dvm
