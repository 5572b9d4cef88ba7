// track_kernels.cu -- sm_100a kernels of the per-frame tracking operators.
//
//   grid_build_kernel     Frame::AssignFeaturesToGrid / PosInGrid          (O3/src/Frame.cc:481-506,772-782)
//   match_last_kernel     ORBmatcher::SearchByProjection(cur, last, th)    (O3/src/ORBmatcher.cc:1553-1748)
//   match_map_kernel      ORBmatcher::SearchByProjection(F, mapPoints, th) (O3/src/ORBmatcher.cc:44-212)
//   pose_opt_kernel       Optimizer::PoseOptimization + g2o LM             (O3/src/Optimizer.cc:744-1028)
//
// The reference's matchers are sequential and greedy: a map point skips keypoints that an EARLIER
// map point (with Observations() > 0) already took in the same loop.  Here every query walks its
// window in parallel, and the greedy outcome is reached as the fixed point of
//     choice[i] = best candidate of i not claimed by any j < i
// iterated from "nothing claimed" (Jacobi rounds inside one CTA).  At a fixed point the rule holds for
// i = 0, 1, 2, ... in turn, so it is exactly the sequential result; typical frames need 3-5 rounds.
#include "track_kernels.cuh"
#include "orb_math.cuh"

namespace dvm {

// --------------------------------------------------------------------------------------- helpers
__device__ inline int hamming256(const uint32_t a[8], const uint8_t* b)
{
    // ORBmatcher::DescriptorDistance: popcount of the 256-bit XOR
    const uint4* p = reinterpret_cast<const uint4*>(b);
    const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
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

// Frame::GetFeaturesInArea: calls fn(idx, octave) for every keypoint of the window, in the
// reference's traversal order (ix outer, iy inner, insertion order inside a cell).
template <class Fn>
__device__ inline void walk_area(const FrameDev& f, float x, float y, float r, int minLevel, int maxLevel, Fn fn)
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
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int c = ix * kGridRows + iy;
            const int j0 = f.cell_start[c], j1 = f.cell_start[c + 1];
            for (int j = j0; j < j1; j++) {
                const int idx = f.cell_items[j];
                const dvm_keypoint* kp = f.kps + idx;
                const int oct = kp->octave;
                if (bCheckLevels) {
                    if (oct < minLevel) continue;
                    if (maxLevel >= 0 && oct > maxLevel) continue;
                }
                const float distx = __fsub_rn(kp->x, x), disty = __fsub_rn(kp->y, y);
                if (fabsf(distx) < r && fabsf(disty) < r) fn(idx, oct);
            }
        }
}

// -------------------------------------------------------------------------------------- grid build
__global__ void __launch_bounds__(1024) grid_build_kernel(FrameDev f)
{
    __shared__ int counts[kGridCells];
    __shared__ int warp_sums[33];
    const int tid = threadIdx.x;
    const int n = min(*f.n, f.cap);
    for (int c = tid; c < kGridCells; c += 1024) counts[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const dvm_keypoint* kp = f.kps + i;
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kp->x, f.minX), f.gwInv));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(kp->y, f.minY), f.ghInv));
        if (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) continue;
        atomicAdd(&counts[px * kGridRows + py], 1);
    }
    __syncthreads();
    // exclusive scan of 3072 counts: 3 consecutive cells per thread
    const int c0 = tid * 3;
    const int a = counts[c0], b = counts[c0 + 1], c = counts[c0 + 2];
    int total;
    const int off = block_exclusive_scan(a + b + c, warp_sums, &total);
    f.cell_start[c0] = off;
    f.cell_start[c0 + 1] = off + a;
    f.cell_start[c0 + 2] = off + a + b;
    if (tid == 0) f.cell_start[kGridCells] = total;
    counts[c0] = off; counts[c0 + 1] = off + a; counts[c0 + 2] = off + a + b; // running fill cursors
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const dvm_keypoint* kp = f.kps + i;
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kp->x, f.minX), f.gwInv));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(kp->y, f.minY), f.ghInv));
        if (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) continue;
        f.cell_items[atomicAdd(&counts[px * kGridRows + py], 1)] = i;
    }
    __syncthreads();
    // the reference pushes indices in increasing order: sort every (short) cell list
    for (int cidx = tid; cidx < kGridCells; cidx += 1024) {
        const int j0 = f.cell_start[cidx], j1 = counts[cidx];
        for (int j = j0 + 1; j < j1; j++) {
            const int v = f.cell_items[j];
            int k = j - 1;
            while (k >= j0 && f.cell_items[k] > v) { f.cell_items[k + 1] = f.cell_items[k]; k--; }
            f.cell_items[k + 1] = v;
        }
    }
}

void launch_grid_build(const FrameDev& f, cudaStream_t stream) { DVM_LAUNCH(grid_build_kernel, 1, 1024, 0, stream, f); }

__global__ void features_in_area_kernel(FrameDev f, float x, float y, float r, int minLevel, int maxLevel, int* out,
                                        int cap, int* n_out)
{
    int n = 0;
    walk_area(f, x, y, r, minLevel, maxLevel, [&](int idx, int) {
        if (n < cap) out[n] = idx;
        n++;
    });
    *n_out = n;
}

void launch_features_in_area(const FrameDev& f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap,
                             int* n_out, cudaStream_t stream)
{
    DVM_LAUNCH(features_in_area_kernel, 1, 1, 0, stream, f, x, y, r, minLevel, maxLevel, out, cap, n_out);
}

// ----------------------------------------------------------------- SearchByProjection(cur, last)
constexpr int kMatchThreads = 1024;
constexpr int kNoClaim = 0x7fffffff;

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

__global__ void __launch_bounds__(kMatchThreads, 1)
match_last_kernel(FrameDev cur, MatchLastArgs a, MatchScratch s, int* __restrict__ cur_mp, int* __restrict__ nmatches)
{
    __shared__ int histo[kHistoLength];
    __shared__ int s_ind[3], s_events, s_bad;
    const int tid = threadIdx.x;
    if (a.guard && *a.guard >= 20) return; // enough matches at th: the wider retry is not run
    const int ncur = min(*cur.n, cur.cap);
    const int nq = a.n_ptr ? *a.n_ptr : a.last_n;
    if (a.pose) { // pose prior held on the device
        float Rm[9];
        quat_to_R_f32(a.pose, Rm);
        for (int i = 0; i < 9; i++) a.R[i] = Rm[i];
        a.t[0] = a.pose[4]; a.t[1] = a.pose[5]; a.t[2] = a.pose[6];
    }
    auto mp_of = [&](int i) { return a.mp_index ? a.mp_index[i] : (a.has_mp[i] ? i : -1); };
    auto desc_of = [&](int i) { return a.mp_desc + (size_t)(a.mp_index ? a.mp_index[i] : i) * 32; };
    auto obs_of = [&](int i) { return a.obs_pos ? a.obs_pos[i] != 0 : true; };
    auto angle_of = [&](int i) { return a.last_kps ? a.last_kps[i].angle : a.angle[i]; };

    // project every last-frame map point with the current pose prior
    for (int i = tid; i < nq; i += kMatchThreads) {
        int lv = -1;
        const int mi = mp_of(i);
        if (mi >= 0 && !a.outlier[i]) {
            const float X = a.Xw[3 * mi], Y = a.Xw[3 * mi + 1], Z = a.Xw[3 * mi + 2];
            const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.R[0], X), __fmul_rn(a.R[1], Y)), __fmul_rn(a.R[2], Z)), a.t[0]);
            const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.R[3], X), __fmul_rn(a.R[4], Y)), __fmul_rn(a.R[5], Z)), a.t[1]);
            const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a.R[6], X), __fmul_rn(a.R[7], Y)), __fmul_rn(a.R[8], Z)), a.t[2]);
            const float invzc = (float)(1.0 / (double)zc);
            if (!(invzc < 0)) {
                const float u = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[0], xc), zc), a.K[2]);
                const float v = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[1], yc), zc), a.K[3]);
                if (!(u < cur.minX || u > cur.maxX) && !(v < cur.minY || v > cur.maxY)) {
                    const int oct = a.last_kps ? a.last_kps[i].octave : a.octave[i];
                    s.pu[i] = u; s.pv[i] = v;
                    s.pr[i] = __fmul_rn(a.th, cur.scale[oct]);
                    lv = ((oct - 1) << 16) | ((oct + 1) & 0xffff);
                }
            }
        }
        s.plevels[i] = lv;
        s.choice[i] = -1;
    }
    int* claim_prev = s.claim_a;
    int* claim_next = s.claim_b;
    for (int k = tid; k < ncur; k += kMatchThreads) claim_prev[k] = kNoClaim;
    __syncthreads();

    int rounds = 0;
    while (true) {
        int changed = 0;
        for (int i = tid; i < nq; i += kMatchThreads) {
            const int lv = s.plevels[i];
            if (lv == -1) continue;
            uint32_t d[8];
            load_desc(d, desc_of(i));
            int bestDist = 256, bestIdx = -1;
            walk_area(cur, s.pu[i], s.pv[i], s.pr[i], lv >> 16, (int)(short)(lv & 0xffff), [&](int idx, int) {
                if (claim_prev[idx] < i) return; // taken by an earlier map point with observations
                const int dist = hamming256(d, cur.desc + (size_t)idx * 32);
                if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
            });
            const int pick = bestDist <= kThHigh ? bestIdx : -1;
            if (pick != s.choice[i]) { s.choice[i] = pick; changed = 1; }
        }
        rounds++;
        if (!__syncthreads_or(changed)) break;
        for (int k = tid; k < ncur; k += kMatchThreads) claim_next[k] = kNoClaim;
        __syncthreads();
        for (int i = tid; i < nq; i += kMatchThreads) {
            const int k = s.choice[i];
            if (k >= 0 && obs_of(i)) atomicMin(&claim_next[k], i);
        }
        __syncthreads();
        int* t = claim_prev; claim_prev = claim_next; claim_next = t;
    }

    // every accepted query is one write "CurrentFrame.mvpMapPoints[k] = pMP" (+ one histogram entry)
    if (tid < kHistoLength) histo[tid] = 0;
    if (tid == 0) { s_events = 0; s_bad = 0; }
    int* owner = claim_next;   // last writer of each keypoint
    int* nulled = claim_prev;  // keypoint cleared by the rotation check
    __syncthreads();
    for (int k = tid; k < ncur; k += kMatchThreads) { owner[k] = -1; nulled[k] = 0; }
    __syncthreads();
    for (int i = tid; i < nq; i += kMatchThreads) {
        const int k = s.choice[i];
        if (k < 0) continue;
        atomicMax(&owner[k], i);
        atomicAdd(&s_events, 1);
        if (a.check_ori) atomicAdd(&histo[rot_bin(angle_of(i), cur.kps[k].angle)], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int i1 = -1, i2 = -1, i3 = -1;
        if (a.check_ori) three_maxima(histo, kHistoLength, i1, i2, i3);
        s_ind[0] = i1; s_ind[1] = i2; s_ind[2] = i3;
    }
    __syncthreads();
    if (a.check_ori) {
        for (int i = tid; i < nq; i += kMatchThreads) {
            const int k = s.choice[i];
            if (k < 0) continue;
            const int bin = rot_bin(angle_of(i), cur.kps[k].angle);
            if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) {
                nulled[k] = 1;
                atomicAdd(&s_bad, 1);
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < ncur; k += kMatchThreads) cur_mp[k] = nulled[k] ? -1 : owner[k];
    if (tid == 0) { *nmatches = s_events - s_bad; *s.iters = rounds; }
}

void launch_match_last(const FrameDev& cur, const MatchLastArgs& a, const MatchScratch& s, int* cur_mp, int* nmatches,
                       cudaStream_t stream)
{
    DVM_LAUNCH(match_last_kernel, 1, kMatchThreads, 0, stream, cur, a, s, cur_mp, nmatches);
}

// --------------------------------------------------------------- SearchByProjection(F, mapPoints)
__global__ void __launch_bounds__(kMatchThreads, 1)
match_map_kernel(FrameDev cur, MatchMapArgs a, MatchScratch s, int* __restrict__ cur_mp, int* __restrict__ nmatches)
{
    __shared__ int s_events;
    const int tid = threadIdx.x;
    const int ncur = min(*cur.n, cur.cap);
    const int nq = a.m_ptr ? *a.m_ptr : a.m;
    const bool bFactor = a.th != 1.0f;
    auto desc_of = [&](int i) { return a.mp_desc + (size_t)(a.q_index ? a.q_index[i] : i) * 32; };
    auto obs_of = [&](int i) { return a.obs_pos ? a.obs_pos[i] != 0 : true; };
    auto blocked = [&](int k) { return a.cur_map ? a.cur_map[k] >= 0 : (a.cur_blocked && a.cur_blocked[k]); };
    for (int i = tid; i < nq; i += kMatchThreads) {
        const int lvl = a.level[i];
        float r = a.view_cos[i] > 0.998f ? 2.5f : 4.0f; // RadiusByViewingCos
        if (bFactor) r = __fmul_rn(r, a.th);
        s.pr[i] = __fmul_rn(r, cur.scale[lvl]);
        s.choice[i] = -1;
    }
    int* claim_prev = s.claim_a;
    int* claim_next = s.claim_b;
    for (int k = tid; k < ncur; k += kMatchThreads) claim_prev[k] = blocked(k) ? -1 : kNoClaim;
    if (tid == 0) s_events = 0;
    __syncthreads();

    int rounds = 0;
    while (true) {
        int changed = 0;
        for (int i = tid; i < nq; i += kMatchThreads) {
            const int lvl = a.level[i];
            uint32_t d[8];
            load_desc(d, desc_of(i));
            int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
            walk_area(cur, a.projX[i], a.projY[i], s.pr[i], lvl - 1, lvl, [&](int idx, int oct) {
                if (claim_prev[idx] < i) return; // held by a map point with observations
                const int dist = hamming256(d, cur.desc + (size_t)idx * 32);
                if (dist < bestDist) {
                    bestDist2 = bestDist; bestDist = dist;
                    bestLevel2 = bestLevel; bestLevel = oct;
                    bestIdx = idx;
                } else if (dist < bestDist2) {
                    bestLevel2 = oct;
                    bestDist2 = dist;
                }
            });
            int pick = -1;
            if (bestDist <= kThHigh) {
                const float lim = __fmul_rn(a.nnratio, (float)bestDist2);
                const bool reject = (bestLevel == bestLevel2) && ((float)bestDist > lim);
                if (!reject && (bestLevel != bestLevel2 || (float)bestDist <= lim)) pick = bestIdx;
            }
            if (pick != s.choice[i]) { s.choice[i] = pick; changed = 1; }
        }
        rounds++;
        if (!__syncthreads_or(changed)) break;
        for (int k = tid; k < ncur; k += kMatchThreads) claim_next[k] = blocked(k) ? -1 : kNoClaim;
        __syncthreads();
        for (int i = tid; i < nq; i += kMatchThreads) {
            const int k = s.choice[i];
            if (k >= 0 && obs_of(i)) atomicMin(&claim_next[k], i);
        }
        __syncthreads();
        int* t = claim_prev; claim_prev = claim_next; claim_next = t;
    }
    int* owner = claim_next;
    __syncthreads();
    for (int k = tid; k < ncur; k += kMatchThreads) owner[k] = -1;
    __syncthreads();
    for (int i = tid; i < nq; i += kMatchThreads) {
        const int k = s.choice[i];
        if (k < 0) continue;
        atomicMax(&owner[k], i);
        atomicAdd(&s_events, 1);
    }
    __syncthreads();
    for (int k = tid; k < ncur; k += kMatchThreads) cur_mp[k] = owner[k];
    if (tid == 0) { *nmatches = s_events; *s.iters = rounds; }
}

void launch_match_map(const FrameDev& cur, const MatchMapArgs& a, const MatchScratch& s, int* cur_mp, int* nmatches,
                      cudaStream_t stream)
{
    DVM_LAUNCH(match_map_kernel, 1, kMatchThreads, 0, stream, cur, a, s, cur_mp, nmatches);
}

// ----------------------------------------------------------------------------- PoseOptimization
constexpr int kPoseThreads = 512;
constexpr int kPoseWarps = kPoseThreads / 32;

struct Quat { double x, y, z, w; };
struct SE3d { Quat r; double t[3]; };

__device__ inline void quat_normalize(Quat& q)
{
    if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}
__device__ inline Quat quat_mul(const Quat& a, const Quat& b)
{
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
__device__ inline void quat_rotate(const Quat& q, const double v[3], double out[3])
{
    double uv[3] = { q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0] };
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
    out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
    out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}
__device__ inline Quat quat_from_matrix(const double R[9])
{
    Quat q;
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[7] - R[5]) * t;
        q.y = (R[2] - R[6]) * t;
        q.z = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (R[k * 3 + j] - R[j * 3 + k]) * t;
        v[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        v[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}
// SE3Quat::exp (g2o/types/se3quat.h:212-240)
__device__ inline SE3d se3_exp(const double u[6])
{
    const double w0 = u[0], w1 = u[1], w2 = u[2];
    const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    const double O[9] = { 0, -w2, w1, w2, 0, -w0, -w1, w0, 0 };
    double O2[9], R[9], V[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    if (theta < 0.00001) {
#pragma unroll
        for (int i = 0; i < 9; i++) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double sa = sin(theta) / theta, sb = (1 - cos(theta)) / (theta * theta);
        const double sc = (theta - sin(theta)) / pow(theta, 3.0);
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const double I = (i % 4 == 0) ? 1.0 : 0.0;
            R[i] = I + sa * O[i] + sb * O2[i];
            V[i] = I + sb * O[i] + sc * O2[i];
        }
    }
    SE3d T;
    T.r = quat_from_matrix(R);
    quat_normalize(T.r);
#pragma unroll
    for (int i = 0; i < 3; i++) T.t[i] = V[i * 3] * u[3] + V[i * 3 + 1] * u[4] + V[i * 3 + 2] * u[5];
    return T;
}
__device__ inline SE3d se3_mul(const SE3d& a, const SE3d& b)
{
    SE3d r = a;
    double rt[3];
    quat_rotate(a.r, b.t, rt);
    r.t[0] += rt[0]; r.t[1] += rt[1]; r.t[2] += rt[2];
    r.r = quat_mul(a.r, b.r);
    quat_normalize(r.r);
    return r;
}

// unpivoted LDL^T of a symmetric 6x6 system (full storage); false if a pivot is not positive
__device__ inline bool ldlt6_solve(const double* A, const double* b, double* x)
{
    double L[36], D[6], y[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[j * 6 + j];
        for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
        if (!(d > 0)) return false;
        D[j] = d;
        for (int i = j + 1; i < 6; i++) {
            double s = A[i * 6 + j];
            for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k] * D[k];
            L[i * 6 + j] = s / d;
        }
    }
    for (int i = 0; i < 6; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L[i * 6 + k] * y[k];
        y[i] = s;
    }
    for (int i = 5; i >= 0; i--) {
        double s = y[i] / D[i];
        for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * x[k];
        x[i] = s;
    }
    return true;
}

// deterministic block sum of NV doubles per thread; result readable by every thread from `out`
template <int NV>
__device__ inline void block_sum(double (&v)[NV], double* warp_buf /*[kPoseWarps*NV]*/, double* out /*[NV]*/)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) warp_buf[wid * NV + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < kPoseWarps; w++) s += warp_buf[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

struct PoseCam { double fx, fy, cx, cy, delta, dsqr; };

__device__ inline void pose_edge_error(const PoseCam& c, const SE3d& T, const float* Xw, const float* obs, double e[2], double xc[3])
{
    const double X[3] = { (double)Xw[0], (double)Xw[1], (double)Xw[2] };
    quat_rotate(T.r, X, xc);
    xc[0] += T.t[0]; xc[1] += T.t[1]; xc[2] += T.t[2];
    e[0] = (double)obs[0] - (c.fx * xc[0] / xc[2] + c.cx);
    e[1] = (double)obs[1] - (c.fy * xc[1] / xc[2] + c.cy);
}
__device__ inline double pose_chi2(const double e[2], double info) { return e[0] * (info * e[0]) + e[1] * (info * e[1]); }
__device__ inline double huber_rho0(const PoseCam& c, double e) { return e <= c.dsqr ? e : 2 * sqrt(e) * c.delta - c.dsqr; }
__device__ inline double huber_rho1(const PoseCam& c, double e) { return e <= c.dsqr ? 1.0 : c.delta / sqrt(e); }

// One CTA runs the whole of Optimizer::PoseOptimization: 4 rounds x optimize(10) of g2o's LM on one
// SE3 vertex, outlier re-classification between rounds.  Every thread carries the pose and the
// 6x6 system redundantly (identical arithmetic), so no broadcast is needed; edges are strided over
// the threads and reduced in a fixed order (deterministic).
__global__ void __launch_bounds__(kPoseThreads, 1) pose_opt_kernel(PoseOptArgs a)
{
    __shared__ double warp_buf[kPoseWarps * 28];
    __shared__ double red[28];
    const int tid = threadIdx.x;
    const int n = a.n_ptr ? min(*a.n_ptr, a.n) : a.n;
    auto edge_valid = [&](int k) { return a.map_index ? a.map_index[k] >= 0 : (a.valid ? a.valid[k] != 0 : true); };
    auto edge_X = [&](int k) { return a.Xw + 3 * (size_t)(a.map_index ? a.map_index[k] : k); };
    auto edge_info = [&](int k) { return (double)(a.kps ? a.inv_sigma2_table[a.kps[k].octave] : a.inv_sigma2[k]); };
    auto edge_err = [&](const PoseCam& c, const SE3d& T, int k, double e[2], double xc[3]) {
        float o[2];
        if (a.kps) { o[0] = a.kps[k].x; o[1] = a.kps[k].y; }
        else { o[0] = a.kp_xy[2 * k]; o[1] = a.kp_xy[2 * k + 1]; }
        pose_edge_error(c, T, edge_X(k), o, e, xc);
    };
    PoseCam cam;
    cam.fx = a.K[0]; cam.fy = a.K[1]; cam.cx = a.K[2]; cam.cy = a.K[3];
    cam.delta = (double)(float)sqrt(5.991);
    cam.dsqr = cam.delta * cam.delta;

    // edge states (per edge, owned by thread k % kPoseThreads): bit0 excluded (level 1), bit1 robust kernel removed,
    // kept in a.outlier's byte until the end
    int nedges_local = 0;
    for (int k = tid; k < n; k += kPoseThreads) {
        const bool valid = edge_valid(k);
        a.outlier[k] = valid ? 0 : 4; // 4 = not an edge
        nedges_local += valid;
    }
    double cnt[1] = { (double)nedges_local };
    block_sum<1>(cnt, warp_buf, red);
    const int nedges = (int)red[0];
    __syncthreads();

    SE3d T0;
    T0.r.x = a.pose[0]; T0.r.y = a.pose[1]; T0.r.z = a.pose[2]; T0.r.w = a.pose[3];
    T0.t[0] = a.pose[4]; T0.t[1] = a.pose[5]; T0.t[2] = a.pose[6];
    quat_normalize(T0.r);
    SE3d T = T0;
    int nBadEdges = 0, total_iters = 0, total_trials = 0;

    if (nedges >= 3) {
        for (int round = 0; round < 4; round++) {
            T = T0; // the frame's pose is only written back at the end (O3/src/Optimizer.cc:935-936)
            // ---- optimize(10) ----
            double lambda = -1, ni = 2;
            int nBadIter = 0;
            for (int it = 0; it < 10; it++) {
                // computeActiveErrors + activeRobustChi2 + buildSystem at the current estimate
                double acc[28];
#pragma unroll
                for (int i = 0; i < 28; i++) acc[i] = 0;
                int nact = 0;
                for (int k = tid; k < n; k += kPoseThreads) {
                    const int st = a.outlier[k];
                    if (st & 5) continue;
                    nact++;
                    double e[2], xc[3];
                    edge_err(cam, T, k, e, xc);
                    a.err[2 * k] = e[0]; a.err[2 * k + 1] = e[1];
                    const double om = edge_info(k);
                    const double chi = pose_chi2(e, om);
                    const bool robust = !(st & 2);
                    acc[27] += robust ? huber_rho0(cam, chi) : chi;
                    const double w = robust ? huber_rho1(cam, chi) : 1.0;
                    const double x = xc[0], y = xc[1], z = xc[2];
                    const double pj[6] = { cam.fx / z, 0, -cam.fx * x / (z * z), 0, cam.fy / z, -cam.fy * y / (z * z) };
                    const double D[18] = { 0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1 };
                    double J[12];
#pragma unroll
                    for (int r = 0; r < 2; r++)
#pragma unroll
                        for (int c = 0; c < 6; c++)
                            J[r * 6 + c] = -(pj[r * 3] * D[c] + pj[r * 3 + 1] * D[6 + c] + pj[r * 3 + 2] * D[12 + c]);
                    int idx = 0;
#pragma unroll
                    for (int c = 0; c < 6; c++) {
                        acc[21 + c] -= w * (J[c] * (om * e[0]) + J[6 + c] * (om * e[1]));
#pragma unroll
                        for (int d = c; d < 6; d++) acc[idx++] += J[c] * (w * om) * J[d] + J[6 + c] * (w * om) * J[6 + d];
                    }
                }
                // the active count rides along in a 29th slot via a second tiny reduction
                block_sum<28>(acc, warp_buf, red);
                double H[36], b[6];
                {
                    int idx = 0;
                    for (int c = 0; c < 6; c++)
                        for (int d = c; d < 6; d++) { H[c * 6 + d] = red[idx]; H[d * 6 + c] = red[idx]; idx++; }
                    for (int c = 0; c < 6; c++) b[c] = red[21 + c];
                }
                double currentChi = red[27];
                const double iniChi = currentChi;
                __syncthreads();
                double cn[1] = { (double)nact };
                block_sum<1>(cn, warp_buf, red);
                const int nactive = (int)red[0];
                __syncthreads();
                if (nactive == 0) break; // optimize() returns without touching anything
                if (it == 0) {
                    double mx = 0;
                    for (int j = 0; j < 6; j++) mx = fmax(fabs(H[j * 6 + j]), mx);
                    lambda = 1e-5 * mx; // computeLambdaInit, tau = 1e-5
                    ni = 2;
                    nBadIter = 0;
                }
                double rho = 0;
                int qmax = 0;
                do {
                    const SE3d backup = T;
                    double Hl[36], x[6];
                    for (int j = 0; j < 36; j++) Hl[j] = H[j];
                    for (int j = 0; j < 6; j++) Hl[j * 6 + j] += lambda;
                    const bool ok2 = ldlt6_solve(Hl, b, x);
                    if (!ok2) for (int j = 0; j < 6; j++) x[j] = 0;
                    T = se3_mul(se3_exp(x), T);
                    double chi[1] = { 0 };
                    for (int k = tid; k < n; k += kPoseThreads) {
                        const int st = a.outlier[k];
                        if (st & 5) continue;
                        double e[2], xc[3];
                        edge_err(cam, T, k, e, xc);
                        a.err[2 * k] = e[0]; a.err[2 * k + 1] = e[1];
                        const double c2 = pose_chi2(e, edge_info(k));
                        chi[0] += (st & 2) ? c2 : huber_rho0(cam, c2);
                    }
                    block_sum<1>(chi, warp_buf, red);
                    double tempChi = red[0];
                    __syncthreads();
                    if (!ok2) tempChi = 1.7976931348623157e308;
                    rho = currentChi - tempChi;
                    double scale = 0;
                    for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
                    scale += 1e-3;
                    rho /= scale;
                    if (rho > 0 && isfinite(tempChi)) {
                        double alpha = 1. - pow((2 * rho - 1), 3.0);
                        alpha = fmin(alpha, 2. / 3.);
                        const double sf = fmax(1. / 3., alpha);
                        lambda *= sf;
                        ni = 2;
                        currentChi = tempChi;
                    } else {
                        lambda *= ni;
                        ni *= 2;
                        T = backup;
                    }
                    qmax++;
                    total_trials++;
                } while (rho < 0 && qmax < 10);
                total_iters++;
                if (qmax == 10 || rho == 0) break;
                if ((iniChi - currentChi) * 1e3 < iniChi) nBadIter++;
                else nBadIter = 0;
                if (nBadIter >= 3) break;
            }
            // ---- re-classify every edge (O3/src/Optimizer.cc:941-965) ----
            double bad[1] = { 0 };
            for (int k = tid; k < n; k += kPoseThreads) {
                int st = a.outlier[k];
                if (st & 4) continue;
                double e[2] = { a.err[2 * k], a.err[2 * k + 1] };
                if (st & 1) { double xc[3]; edge_err(cam, T, k, e, xc); a.err[2 * k] = e[0]; a.err[2 * k + 1] = e[1]; }
                const float chi2 = (float)pose_chi2(e, edge_info(k));
                if (chi2 > 5.991f) { st |= 1; bad[0] += 1; }
                else st &= ~1;
                if (round == 2) st |= 2;
                a.outlier[k] = (uint8_t)st;
            }
            block_sum<1>(bad, warp_buf, red);
            nBadEdges = (int)red[0];
            __syncthreads();
            if (nedges < 10) break;
        }
    }
    for (int k = tid; k < n; k += kPoseThreads) a.outlier[k] = (a.outlier[k] & 4) ? 0 : (a.outlier[k] & 1);
    if (tid == 0) {
        if (nedges >= 3) {
            a.pose[0] = (float)T.r.x; a.pose[1] = (float)T.r.y; a.pose[2] = (float)T.r.z; a.pose[3] = (float)T.r.w;
            a.pose[4] = (float)T.t[0]; a.pose[5] = (float)T.t[1]; a.pose[6] = (float)T.t[2];
        }
        a.result[0] = nedges >= 3 ? nedges - nBadEdges : 0;
        a.result[1] = nedges;
        a.result[2] = total_iters;
        a.result[3] = total_trials;
    }
}

void launch_pose_opt(const PoseOptArgs& a, cudaStream_t stream) { DVM_LAUNCH(pose_opt_kernel, 1, kPoseThreads, 0, stream, a); }

} // namespace dvm
