// track_device.cuh -- device helpers shared by the matcher kernels (track_kernels.cu, bow_kernels.cu):
// descriptor distance, the view through which a kernel reads a Frame, Frame::GetFeaturesInArea as a window walk (one
// thread or one warp per query), the rotation histogram.  (Staging the whole frame in every CTA's shared memory by TMA
// bulk copies was measured and dropped: DESIGN.md, negative results.)
#pragma once
#include "track_kernels.cuh"

namespace dvm {

// --------------------------------------------------------------------------------------- helpers
__device__ inline int hamming256(const uint32_t a[8], const uint8_t* b)
{
    // ORBmatcher::DescriptorDistance: popcount of the 256-bit XOR
    const uint4* p = reinterpret_cast<const uint4*>(b);
    const uint4 v0 = p[0], v1 = p[1];
    return __popc(a[0] ^ v0.x) + __popc(a[1] ^ v0.y) + __popc(a[2] ^ v0.z) + __popc(a[3] ^ v0.w) +
           __popc(a[4] ^ v1.x) + __popc(a[5] ^ v1.y) + __popc(a[6] ^ v1.z) + __popc(a[7] ^ v1.w);
}

__device__ inline void load_desc(uint32_t a[8], const uint8_t* d)
{
    const uint4* p = reinterpret_cast<const uint4*>(d);
    const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
    a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
    a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
}

// How the matchers read the current frame (global memory through L1/L2; the frame is ~200 KB).
struct FrameLook {
    const int* cell_start;
    const int* cell_items;
    const int4* cell_rec;   // per grid position j: {x bits, y bits, octave, keypoint index} (grid_build_kernel)
    const char* kbase;   // keypoint records: x at +0, y at +4, octave at +oct_off
    int kstride, oct_off;
    const uint8_t* desc;
    float minX, minY, gwInv, ghInv;
    __device__ float x(int i) const { return *reinterpret_cast<const float*>(kbase + (size_t)i * kstride); }
    __device__ float y(int i) const { return *reinterpret_cast<const float*>(kbase + (size_t)i * kstride + 4); }
    __device__ int oct(int i) const { return *reinterpret_cast<const int*>(kbase + (size_t)i * kstride + oct_off); }
};

__device__ inline FrameLook look_global(const FrameDev& f)
{
    FrameLook v;
    v.cell_start = f.cell_start; v.cell_items = f.cell_items; v.cell_rec = f.cell_rec;
    v.kbase = reinterpret_cast<const char*>(f.kps); v.kstride = (int)sizeof(dvm_keypoint); v.oct_off = 20;
    v.desc = f.desc;
    v.minX = f.minX; v.minY = f.minY; v.gwInv = f.gwInv; v.ghInv = f.ghInv;
    return v;
}

// Frame::GetFeaturesInArea: calls fn(idx, octave) for every keypoint of the window, in the
// reference's traversal order (ix outer, iy inner, insertion order inside a cell).
template <class Fn>
__device__ inline void walk_area(const FrameLook& f, float x, float y, float r, int minLevel, int maxLevel, Fn fn)
{
    const float dxm = __fsub_rn(x, f.minX), dym = __fsub_rn(y, f.minY);
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(dxm, r), f.gwInv)));
    if (nMinCellX >= kGridCols) return;
    const int nMaxCellX = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(dxm, r), f.gwInv)));
    if (nMaxCellX < 0) return;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(dym, r), f.ghInv)));
    if (nMinCellY >= kGridRows) return;
    const int nMaxCellY = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(dym, r), f.ghInv)));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        // cells of one grid column are contiguous: one range covers iy = nMinCellY .. nMaxCellY
        const int j0 = f.cell_start[ix * kGridRows + nMinCellY], j1 = f.cell_start[ix * kGridRows + nMaxCellY + 1];
        for (int j = j0; j < j1; j++) {
            const int4 rec = f.cell_rec[j];
            const int idx = rec.w, oct = rec.z;
            if (bCheckLevels) {
                if (oct < minLevel) continue;
                if (maxLevel >= 0 && oct > maxLevel) continue;
            }
            const float distx = __fsub_rn(__int_as_float(rec.x), x), disty = __fsub_rn(__int_as_float(rec.y), y);
            if (fabsf(distx) < r && fabsf(disty) < r) fn(idx, oct);
        }
    }
}

// ORBmatcher::ComputeThreeMaxima
__device__ inline void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

__device__ inline int rot_bin(float last_angle, float cur_angle)
{
    float rot = __fsub_rn(last_angle, cur_angle);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / kHistoLength)); // factor = 1/30: the reference's bug, kept
    if (bin == kHistoLength) bin = 0;
    return bin;
}

// ---- one WARP per query: the lanes share the window's candidates, every lane keeps its own sorted
// top-K, and the K best of the warp are merged by K rounds of a 64-bit warp minimum.  The order key of a
// candidate is its position in the window's traversal (cells in the reference's order, unfiltered), which
// is monotone in the reference's visiting order, so (distance, key) ranks candidates exactly as the
// sequential scan does.
// The grid columns of the window are NOT walked one after the other (each would be a dependent chain of loads for a
// handful of keypoints): lane c fetches the item range of column c, a warp scan concatenates the ranges, and the lanes
// then stride over the concatenation -- one pass of cell_start -> cell_rec -> descriptor for the whole window. ----
template <class Fn>
__device__ inline void walk_area_warp(const FrameLook& f, float x, float y, float r, int minLevel, int maxLevel, int lane, Fn fn)
{
    const float dxm = __fsub_rn(x, f.minX), dym = __fsub_rn(y, f.minY);
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(dxm, r), f.gwInv)));
    if (nMinCellX >= kGridCols) return;
    const int nMaxCellX = min(kGridCols - 1, (int)ceilf(__fmul_rn(__fadd_rn(dxm, r), f.gwInv)));
    if (nMaxCellX < 0) return;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(dym, r), f.ghInv)));
    if (nMinCellY >= kGridRows) return;
    const int nMaxCellY = min(kGridRows - 1, (int)ceilf(__fmul_rn(__fadd_rn(dym, r), f.ghInv)));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    const int ncols = nMaxCellX - nMinCellX + 1;
    int base = 0;
    for (int c0 = 0; c0 < ncols; c0 += 32) {
        int j0 = 0, cnt = 0;
        if (c0 + lane < ncols) { // cells of one grid column are contiguous: one range covers iy = nMinCellY .. nMaxCellY
            const int ix = nMinCellX + c0 + lane;
            j0 = f.cell_start[ix * kGridRows + nMinCellY];
            cnt = f.cell_start[ix * kGridRows + nMaxCellY + 1] - j0;
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - cnt;
        for (int tb = 0; tb < total; tb += 32) {
            const int t = tb + lane;
            int col = 0; // number of columns that end at or before position t = the column position t lies in
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
                const int probe = __shfl_sync(0xffffffffu, incl, col + sft - 1);
                if (probe <= t) col += sft;
            }
            col = min(col, 31);
            const int cj0 = __shfl_sync(0xffffffffu, j0, col), cex = __shfl_sync(0xffffffffu, excl, col);
            if (t >= total) continue;
            const int4 rec = f.cell_rec[cj0 + (t - cex)];
            const int idx = rec.w, oct = rec.z;
            if (bCheckLevels) {
                if (oct < minLevel) continue;
                if (maxLevel >= 0 && oct > maxLevel) continue;
            }
            const float distx = __fsub_rn(__int_as_float(rec.x), x), disty = __fsub_rn(__int_as_float(rec.y), y);
            if (fabsf(distx) < r && fabsf(disty) < r) fn(idx, oct, base + t);
        }
        base += total;
    }
}

__device__ inline unsigned long long warp_min_u64(unsigned long long v)
{
    const unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)(v >> 32));
    const unsigned lo = __reduce_min_sync(0xffffffffu, (unsigned)(v >> 32) == hi ? (unsigned)v : 0xffffffffu);
    return ((unsigned long long)hi << 32) | lo;
}

} // namespace dvm
