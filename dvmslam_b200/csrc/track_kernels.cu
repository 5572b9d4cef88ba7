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
#include "track_device.cuh"

namespace dvm {

// -------------------------------------------------------------------------------------- grid build
// src_kps != nullptr: first copy an extractor result (keypoints, descriptors, count) into the frame's buffers
__global__ void __launch_bounds__(1024) grid_build_kernel(FrameDev f, const dvm_keypoint* __restrict__ src_kps,
                                                          const uint8_t* __restrict__ src_desc, const int* __restrict__ src_n)
{
    __shared__ int counts[kGridCells];
    __shared__ int warp_sums[33];
    const int tid = threadIdx.x;
    if (src_kps) {
        const int ns = min(*src_n, f.cap);
        const int* sk = reinterpret_cast<const int*>(src_kps);
        int* dk = reinterpret_cast<int*>(const_cast<dvm_keypoint*>(f.kps));
        for (int i = tid; i < ns * 7; i += 1024) dk[i] = sk[i];
        const uint4* sd = reinterpret_cast<const uint4*>(src_desc);
        uint4* dd = reinterpret_cast<uint4*>(const_cast<uint8_t*>(f.desc));
        for (int i = tid; i < ns * 2; i += 1024) dd[i] = sd[i];
        if (tid == 0) *const_cast<int*>(f.n) = *src_n;
        __syncthreads();
    }
    const int n = min(src_kps ? *src_n : *f.n, f.cap);
    for (int c = tid; c < kGridCells; c += 1024) counts[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const dvm_keypoint* kp = f.kps + i;
        f.kxyo[3 * i] = __float_as_int(kp->x); f.kxyo[3 * i + 1] = __float_as_int(kp->y); f.kxyo[3 * i + 2] = kp->octave;
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

// Frame::UndistortKeyPoints: cv::undistortPoints(pts, pts, K, distCoef, Mat(), K) with its default five iterations, in
// double, operation by operation as OpenCV evaluates it (this file is compiled with -fmad=false; the zero-coefficient
// terms k4..k6, s1..s4 are kept because they take part in the rounding) -- bit-exact with cv2 4.13 (oracle/cvmodels.c).
__global__ void __launch_bounds__(256) undistort_kernel(dvm_keypoint* __restrict__ kps, const int* __restrict__ n_ptr, int cap,
                                                        UndistortArgs a)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= min(*n_ptr, cap)) return;
    const double ifx = 1. / a.fx, ify = 1. / a.fy;
    const double u = (double)kps[i].x, v = (double)kps[i].y;
    double x = (u - a.cx) * ifx, y = (v - a.cy) * ify;
    const double x0 = x, y0 = y;
    const double k5 = 0., k6 = 0., k7 = 0., k8 = 0., k9 = 0., k10 = 0., k11 = 0.;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k7 * r2 + k6) * r2 + k5) * r2) / (1 + ((a.k[4] * r2 + a.k[1]) * r2 + a.k[0]) * r2);
        if (icdist < 0) { x = (u - a.cx) * ifx; y = (v - a.cy) * ify; break; }
        const double deltaX = 2 * a.k[2] * x * y + a.k[3] * (r2 + 2 * x * x) + k8 * r2 + k9 * r2 * r2;
        const double deltaY = a.k[2] * (r2 + 2 * y * y) + 2 * a.k[3] * x * y + k10 * r2 + k11 * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    const double xx = a.fx * x + 0. * y + a.cx, yy = 0. * x + a.fy * y + a.cy, ww = 1. / (0. * x + 0. * y + 1.);
    kps[i].x = (float)(xx * ww);
    kps[i].y = (float)(yy * ww);
}

void launch_undistort(dvm_keypoint* kps, const int* n_ptr, int cap, const UndistortArgs& a, cudaStream_t stream)
{
    DVM_LAUNCH(undistort_kernel, div_up(max(cap, 1), 256), 256, 0, stream, kps, n_ptr, cap, a);
}

void launch_grid_build(const FrameDev& f, cudaStream_t stream)
{
    DVM_LAUNCH(grid_build_kernel, 1, 1024, 0, stream, f, (const dvm_keypoint*)nullptr, (const uint8_t*)nullptr, (const int*)nullptr);
}
void launch_frame_assign(const FrameDev& f, const dvm_keypoint* src_kps, const uint8_t* src_desc, const int* src_n,
                         cudaStream_t stream)
{
    DVM_LAUNCH(grid_build_kernel, 1, 1024, 0, stream, f, src_kps, src_desc, src_n);
}

// Frame::isInFrustum for a batch of map points, one thread per point
__global__ void __launch_bounds__(256) frustum_kernel(FrustumArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.m) return;
    FrustumPose fp;
    frustum_pose(a.pose, fp);
    float u, v, vc;
    int lvl;
    const bool vis = frustum_eval(a, fp, k, u, v, lvl, vc);
    a.in_view[k] = vis ? 1 : 0; a.px[k] = u; a.py[k] = v; a.level[k] = lvl; a.view_cos[k] = vc;
}
void launch_frustum(const FrustumArgs& a, cudaStream_t stream) { DVM_LAUNCH(frustum_kernel, div_up(a.m, 256), 256, 0, stream, a); }

__global__ void features_in_area_kernel(FrameDev f, float x, float y, float r, int minLevel, int maxLevel, int* out,
                                        int cap, int* n_out)
{
    int n = 0;
    const FrameLook fl = look_global(f);
    walk_area(fl, x, y, r, minLevel, maxLevel, [&](int idx, int) {
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
constexpr int kRegQ = 2;   // queries per thread kept in registers by the resolution kernels
constexpr int kNoClaim = 0x7fffffff;

// ---- candidate cache: the window of every query is walked ONCE; its best kMatchCacheK candidates are
// kept sorted by (Hamming distance, walk order) -- the reference's preference order: strict '<' keeps
// the first of equal distances.  The greedy rounds then only consult the cache; a query whose cached
// candidates are all taken while it had more falls back to a full walk.
__device__ inline unsigned long long pack_cand(int dist, int ord, int oct, int idx)
{
    return ((unsigned long long)(((unsigned)dist << 16) | (unsigned)min(ord, 0xffff)) << 32) | ((unsigned)oct << 16) | (unsigned)idx;
}
__device__ inline int cand_idx(unsigned long long e) { return (int)(e & 0xffffu); }
__device__ inline int cand_oct(unsigned long long e) { return (int)((e >> 16) & 0xffu); }
__device__ inline int cand_dist(unsigned long long e) { return (int)(e >> 48); }

__device__ inline void topk_insert(unsigned long long (&top)[kMatchCacheK], unsigned long long v)
{
#pragma unroll
    for (int p = 0; p < kMatchCacheK; p++)
        if (v < top[p]) { const unsigned long long t = top[p]; top[p] = v; v = t; }
}


// merges the lanes' sorted top-K lists; lane p < K returns the p-th best of the warp (~0 when absent)
__device__ inline unsigned long long warp_topk_merge(unsigned long long (&top)[kMatchCacheK], int lane)
{
    unsigned long long mine = ~0ull;
#pragma unroll
    for (int p = 0; p < kMatchCacheK; p++) {
        const unsigned long long m = warp_min_u64(top[0]);
        if (lane == p) mine = m;
        if (top[0] == m && m != ~0ull) { // keys are unique: exactly one lane pops
#pragma unroll
            for (int q = 0; q + 1 < kMatchCacheK; q++) top[q] = top[q + 1];
            top[kMatchCacheK - 1] = ~0ull;
        }
    }
    return mine;
}

constexpr int kWalkThreads = 128;

// ordered list of the queries that take part (plevels != -1) in shared memory; returns their number
__device__ inline int compact_active(const int* __restrict__ plevels, int nq, int* qlist, int* warp_sums)
{
    const int ipt = div_up(max(nq, 1), kMatchThreads);
    const int i0 = min((int)threadIdx.x * ipt, nq), i1 = min(i0 + ipt, nq);
    int c = 0;
    for (int i = i0; i < i1; i++) c += (plevels[i] != -1);
    int total;
    int off = block_exclusive_scan(c, warp_sums, &total);
    for (int i = i0; i < i1; i++)
        if (plevels[i] != -1) qlist[off++] = i;
    __syncthreads();
    return total;
}

// ---- SearchByProjection(cur, last), phase 1: one WARP per last-frame keypoint projects its map point
// with the pose prior and walks its window once (many CTAs; the frame is read through L1/L2) ----
__global__ void __launch_bounds__(kWalkThreads) match_last_walk_kernel(FrameDev cur, MatchLastArgs a, MatchScratch s)
{
    pdl_trigger(); pdl_wait();
    if (a.guard && *a.guard >= 20) return; // enough matches at th: the wider retry is not run
    const int nq = a.n_ptr ? *a.n_ptr : a.last_n;
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * kWalkThreads + threadIdx.x) >> 5;
    if (i >= nq) return;
    const FrameLook fl = look_global(cur);
    if (a.pose) { // pose prior held on the device
        a.q[0] = a.pose[0]; a.q[1] = a.pose[1]; a.q[2] = a.pose[2]; a.q[3] = a.pose[3];
        a.t[0] = a.pose[4]; a.t[1] = a.pose[5]; a.t[2] = a.pose[6];
    }
    const int mi = a.mp_index ? a.mp_index[i] : (a.has_mp[i] ? i : -1);
    int lv = -1, nc = 0;
    unsigned long long top[kMatchCacheK];
#pragma unroll
    for (int p = 0; p < kMatchCacheK; p++) top[p] = ~0ull;
    if (mi >= 0 && !a.outlier[i]) {
        // x3Dc = Tcw * x3Dw (O3/src/ORBmatcher.cc:1577): Sophus' quaternion action, not a matrix product
        const float Pw[3] = { a.Xw[3 * mi], a.Xw[3 * mi + 1], a.Xw[3 * mi + 2] };
        float Pc[3];
        so::se3_apply(a.q, a.t, Pw, Pc);
        const float xc = Pc[0], yc = Pc[1], zc = Pc[2];
        const float invzc = (float)(1.0 / (double)zc);
        if (!(invzc < 0)) {
            const float u = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[0], xc), zc), a.K[2]);
            const float v = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[1], yc), zc), a.K[3]);
            if (!(u < cur.minX || u > cur.maxX) && !(v < cur.minY || v > cur.maxY)) {
                const int oct = a.last_kps ? a.last_kps[i].octave : a.octave[i];
                const float r = __fmul_rn(a.th, cur.scale[oct]);
                if (lane == 0) { s.pu[i] = u; s.pv[i] = v; s.pr[i] = r; }
                lv = ((oct - 1) << 16) | ((oct + 1) & 0xffff);
                uint32_t d[8];
                load_desc(d, a.mp_desc + (size_t)mi * 32);
                walk_area_warp(fl, u, v, r, oct - 1, oct + 1, lane, [&](int idx, int o, int ord) {
                    const int dist = hamming256(d, fl.desc + (size_t)idx * 32);
                    topk_insert(top, pack_cand(dist, ord, o, idx));
                    nc++;
                });
            }
        }
    }
    // the branch conditions above depend on the query only: the whole warp is here together
    if (lane == 0) s.plevels[i] = lv;
    if (lv != -1) {
        nc = __reduce_add_sync(0xffffffffu, nc);
        const unsigned long long e = warp_topk_merge(top, lane);
        if (lane < kMatchCacheK) s.cache[(size_t)i * kMatchCacheK + lane] = e;
        if (lane == 0) s.ncand[i] = nc;
    }
}

// ---- phase 2: the sequential-greedy outcome as a Jacobi fixed point on ONE CTA (claims in shared
// memory), then the rotation-consistency check ----
__global__ void __launch_bounds__(kMatchThreads, 1)
match_last_kernel(FrameDev cur, MatchLastArgs a, MatchScratch s, int* __restrict__ cur_mp, int* __restrict__ nmatches,
                  int use_smem)
{
    pdl_trigger(); pdl_wait();
    extern __shared__ __align__(16) unsigned char claim_smem[];
    __shared__ int histo[kHistoLength];
    __shared__ int s_ind[3], s_events, s_bad;
    const int tid = threadIdx.x;
    if (a.guard && *a.guard >= 20) return; // enough matches at th: the wider retry is not run
    const int ncur = min(*cur.n, cur.cap);
    const int nq = a.n_ptr ? *a.n_ptr : a.last_n;
    const FrameLook fl = look_global(cur);
    auto desc_of = [&](int i) { return a.mp_desc + (size_t)(a.mp_index ? a.mp_index[i] : i) * 32; };
    auto obs_of = [&](int i) { return a.obs_pos ? a.obs_pos[i] != 0 : true; };
    auto angle_of = [&](int i) { return a.last_kps ? a.last_kps[i].angle : a.angle[i]; };

    __shared__ int warp_sums[33];
    int* claim_prev = use_smem ? reinterpret_cast<int*>(claim_smem) : s.claim_a;
    int* claim_next = use_smem ? reinterpret_cast<int*>(claim_smem) + cur.cap : s.claim_b;
    int* qlist = s.qlist;
    for (int i = tid; i < nq; i += kMatchThreads) s.choice[i] = -1;
    for (int k = tid; k < ncur; k += kMatchThreads) claim_prev[k] = kNoClaim;
    const int nact = compact_active(s.plevels, nq, qlist, warp_sums);

    // the first kRegQ queries of every thread stay in registers across the rounds (index, candidate count, the
    // four best cache entries, current choice): a round then only touches the claims in shared memory
    int r_i[kRegQ], r_nc[kRegQ], r_lv[kRegQ], r_choice[kRegQ];
    ulonglong2 r_e01[kRegQ], r_e23[kRegQ];
#pragma unroll
    for (int q = 0; q < kRegQ; q++) {
        const int j = tid + q * kMatchThreads;
        r_i[q] = -1; r_nc[q] = 0; r_lv[q] = 0; r_choice[q] = -1;
        r_e01[q] = make_ulonglong2(~0ull, ~0ull); r_e23[q] = r_e01[q];
        if (j < nact) {
            const int i = qlist[j];
            r_i[q] = i; r_lv[q] = s.plevels[i]; r_nc[q] = s.ncand[i];
            const unsigned long long* e = s.cache + (size_t)i * kMatchCacheK;
            r_e01[q] = *reinterpret_cast<const ulonglong2*>(e); r_e23[q] = *reinterpret_cast<const ulonglong2*>(e + 2);
        }
    }
    auto evaluate = [&](int i, int lv, int nc, const ulonglong2& e01, const ulonglong2& e23, const int* claim) -> int {
        int bestDist = 256, bestIdx = -1;
        bool found = false;
        const unsigned long long* e = s.cache + (size_t)i * kMatchCacheK;
        for (int p = 0; p < kMatchCacheK && p < nc; p++) {
            const unsigned long long ev = p == 0 ? e01.x : p == 1 ? e01.y : p == 2 ? e23.x : p == 3 ? e23.y : e[p];
            const int idx = cand_idx(ev);
            if (claim[idx] < i) continue; // taken by an earlier map point with observations
            bestDist = cand_dist(ev); bestIdx = idx; found = true;
            break;
        }
        if (!found && nc > kMatchCacheK) { // cache exhausted: full walk against the current claims
            uint32_t d[8];
            load_desc(d, desc_of(i));
            walk_area(fl, s.pu[i], s.pv[i], s.pr[i], lv >> 16, (int)(short)(lv & 0xffff), [&](int idx, int) {
                if (claim[idx] < i) return;
                const int dist = hamming256(d, fl.desc + (size_t)idx * 32);
                if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
            });
        }
        return bestDist <= kThHigh ? bestIdx : -1;
    };

    int rounds = 0;
    while (true) {
        int changed = 0;
#pragma unroll
        for (int q = 0; q < kRegQ; q++) {
            if (r_i[q] < 0) continue;
            const int pick = evaluate(r_i[q], r_lv[q], r_nc[q], r_e01[q], r_e23[q], claim_prev);
            if (pick != r_choice[q]) { r_choice[q] = pick; s.choice[r_i[q]] = pick; changed = 1; }
        }
        for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) {
            const int i = qlist[j];
            const unsigned long long* e = s.cache + (size_t)i * kMatchCacheK;
            const int pick = evaluate(i, s.plevels[i], s.ncand[i], *reinterpret_cast<const ulonglong2*>(e),
                                      *reinterpret_cast<const ulonglong2*>(e + 2), claim_prev);
            if (pick != s.choice[i]) { s.choice[i] = pick; changed = 1; }
        }
        rounds++;
        if (!__syncthreads_or(changed)) break;
        for (int k = tid; k < ncur; k += kMatchThreads) claim_next[k] = kNoClaim;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kRegQ; q++)
            if (r_i[q] >= 0 && r_choice[q] >= 0 && obs_of(r_i[q])) atomicMin(&claim_next[r_choice[q]], r_i[q]);
        for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) {
            const int i = qlist[j];
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
    for (int j = tid; j < nact; j += kMatchThreads) {
        const int i = qlist[j];
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
        for (int j = tid; j < nact; j += kMatchThreads) {
            const int i = qlist[j];
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
    if (a.map_out) // CurrentFrame.mvpMapPoints as map indices (device-resident tracker)
        for (int k = tid; k < cur.cap; k += kMatchThreads)
            a.map_out[k] = (k < ncur && !nulled[k] && owner[k] >= 0) ? a.mp_index[owner[k]] : -1;
    if (tid == 0) { *nmatches = s_events - s_bad; *s.iters = rounds; }
}

constexpr size_t kMatchSmemLimit = 200 * 1024;
static bool prepare_match_kernels(); // raises the dynamic shared memory limit of both matcher kernels once
__host__ inline size_t claim_smem_bytes(int cap) { return (size_t)cap * 8; }

void launch_match_last(const FrameDev& cur, const MatchLastArgs& a, const MatchScratch& s, int* cur_mp, int* nmatches,
                       cudaStream_t stream)
{
    const size_t smem = claim_smem_bytes(cur.cap);
    const int use_smem = smem <= kMatchSmemLimit ? 1 : 0;
    prepare_match_kernels();
    DVM_LAUNCH_PDL(match_last_walk_kernel, div_up(max(a.last_n, 1), kWalkThreads / 32), kWalkThreads, 0, stream, cur, a, s);
    DVM_LAUNCH_PDL(match_last_kernel, 1, kMatchThreads, use_smem ? smem : 0, stream, cur, a, s, cur_mp, nmatches, use_smem);
}

// --------------------------------------------------------------- SearchByProjection(F, mapPoints)
// phase 1: one warp per in-view map point walks its window once
__global__ void __launch_bounds__(kWalkThreads) match_map_walk_kernel(FrameDev cur, MatchMapArgs a, MatchScratch s)
{
    pdl_trigger(); pdl_wait();
    const int nq = a.m_ptr ? *a.m_ptr : a.m;
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * kWalkThreads + threadIdx.x) >> 5;
    if (i >= nq) return;
    const FrameLook fl = look_global(cur);
    int lvl;
    float px, py, vcos;
    if (a.use_frustum) { // SearchLocalPoints: isInFrustum decides whether map point i takes part at all
        FrustumPose fp;
        frustum_pose(a.fr.pose, fp);
        if (!frustum_eval(a.fr, fp, i, px, py, lvl, vcos)) {
            if (lane == 0) { s.plevels[i] = -1; s.ncand[i] = 0; }
            return;
        }
    } else {
        lvl = a.level[i]; px = a.projX[i]; py = a.projY[i]; vcos = a.sim3_mode ? 0.f : a.view_cos[i];
        if (a.sim3_mode && lvl < 0) { // rejected by the projection gates
            if (lane == 0) { s.plevels[i] = -1; s.ncand[i] = 0; }
            return;
        }
    }
    float r;
    if (a.sim3_mode) r = __fmul_rn(a.th, cur.scale[lvl]);   // const float radius = th * pKF->mvScaleFactors[nPredictedLevel]
    else {
        r = vcos > 0.998f ? 2.5f : 4.0f; // RadiusByViewingCos
        if (a.th != 1.0f) r = __fmul_rn(r, a.th);
        r = __fmul_rn(r, cur.scale[lvl]);
    }
    if (lane == 0) { s.pu[i] = px; s.pv[i] = py; s.pr[i] = r; s.plevels[i] = lvl; }
    uint32_t d[8];
    load_desc(d, a.mp_desc + (size_t)(a.q_index ? a.q_index[i] : i) * 32);
    unsigned long long top[kMatchCacheK];
#pragma unroll
    for (int p = 0; p < kMatchCacheK; p++) top[p] = ~0ull;
    int nc = 0;
    walk_area_warp(fl, px, py, r, lvl - 1, lvl, lane, [&](int idx, int oct, int ord) {
        if (a.cur_map ? a.cur_map[idx] >= 0 : (a.cur_blocked && a.cur_blocked[idx])) return; // blocked from the start
        const int dist = hamming256(d, fl.desc + (size_t)idx * 32);
        topk_insert(top, pack_cand(dist, ord, oct, idx));
        nc++;
    });
    nc = __reduce_add_sync(0xffffffffu, nc);
    const unsigned long long e = warp_topk_merge(top, lane);
    if (lane < kMatchCacheK) s.cache[(size_t)i * kMatchCacheK + lane] = e;
    if (lane == 0) s.ncand[i] = nc;
}

__global__ void __launch_bounds__(kMatchThreads, 1)
match_map_kernel(FrameDev cur, MatchMapArgs a, MatchScratch s, int* __restrict__ cur_mp, int* __restrict__ nmatches,
                 int use_smem)
{
    pdl_trigger(); pdl_wait();
    extern __shared__ __align__(16) unsigned char claim_smem[];
    __shared__ int s_events;
    const int tid = threadIdx.x;
    const int ncur = min(*cur.n, cur.cap);
    const int nq = a.m_ptr ? *a.m_ptr : a.m;
    const FrameLook fl = look_global(cur);
    auto desc_of = [&](int i) { return a.mp_desc + (size_t)(a.q_index ? a.q_index[i] : i) * 32; };
    auto obs_of = [&](int i) { return a.obs_pos ? a.obs_pos[i] != 0 : true; };
    auto blocked = [&](int k) { return a.cur_map ? a.cur_map[k] >= 0 : (a.cur_blocked && a.cur_blocked[k]); };
    __shared__ int warp_sums[33];
    int* claim_prev = use_smem ? reinterpret_cast<int*>(claim_smem) : s.claim_a;
    int* claim_next = use_smem ? reinterpret_cast<int*>(claim_smem) + cur.cap : s.claim_b;
    int* qlist = s.qlist;
    for (int i = tid; i < nq; i += kMatchThreads) s.choice[i] = -1;
    for (int k = tid; k < ncur; k += kMatchThreads) claim_prev[k] = blocked(k) ? -1 : kNoClaim;
    if (tid == 0) s_events = 0;
    const int nact = compact_active(s.plevels, nq, qlist, warp_sums); // queries in the frustum, vpMapPoints order

    int r_i[kRegQ], r_nc[kRegQ], r_lv[kRegQ], r_choice[kRegQ];
    ulonglong2 r_e01[kRegQ], r_e23[kRegQ];
#pragma unroll
    for (int q = 0; q < kRegQ; q++) {
        const int j = tid + q * kMatchThreads;
        r_i[q] = -1; r_nc[q] = 0; r_lv[q] = 0; r_choice[q] = -1;
        r_e01[q] = make_ulonglong2(~0ull, ~0ull); r_e23[q] = r_e01[q];
        if (j < nact) {
            const int i = qlist[j];
            r_i[q] = i; r_lv[q] = s.plevels[i]; r_nc[q] = s.ncand[i];
            const unsigned long long* e = s.cache + (size_t)i * kMatchCacheK;
            r_e01[q] = *reinterpret_cast<const ulonglong2*>(e); r_e23[q] = *reinterpret_cast<const ulonglong2*>(e + 2);
        }
    }
    auto evaluate = [&](int i, int lvl, int nc, const ulonglong2& e01, const ulonglong2& e23, const int* claim) -> int {
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        int found = 0;
        // sorted by (distance, walk order): the first two free entries are the reference's best and
        // second best (its scan keeps the earliest of equal distances)
        const unsigned long long* e = s.cache + (size_t)i * kMatchCacheK;
        for (int p = 0; p < kMatchCacheK && p < nc && found < 2; p++) {
            const unsigned long long ev = p == 0 ? e01.x : p == 1 ? e01.y : p == 2 ? e23.x : p == 3 ? e23.y : e[p];
            const int idx = cand_idx(ev);
            if (claim[idx] < i) continue; // held by a map point with observations
            if (found == 0) { bestDist = cand_dist(ev); bestLevel = cand_oct(ev); bestIdx = idx; }
            else { bestDist2 = cand_dist(ev); bestLevel2 = cand_oct(ev); }
            found++;
        }
        if (found < 2 && nc > kMatchCacheK) { // cache exhausted: full walk against the current claims
            uint32_t d[8];
            load_desc(d, desc_of(i));
            bestDist = 256; bestLevel = -1; bestDist2 = 256; bestLevel2 = -1; bestIdx = -1;
            walk_area(fl, s.pu[i], s.pv[i], s.pr[i], lvl - 1, lvl, [&](int idx, int oct) {
                if (claim[idx] < i) return;
                const int dist = hamming256(d, fl.desc + (size_t)idx * 32);
                if (dist < bestDist) {
                    bestDist2 = bestDist; bestDist = dist;
                    bestLevel2 = bestLevel; bestLevel = oct;
                    bestIdx = idx;
                } else if (dist < bestDist2) {
                    bestLevel2 = oct;
                    bestDist2 = dist;
                }
            });
        }
        int pick = -1;
        if (a.sim3_mode) return ((float)bestDist <= a.accept_limit && bestDist < 256) ? bestIdx : -1;
        if (bestDist <= kThHigh) {
            const float lim = __fmul_rn(a.nnratio, (float)bestDist2);
            const bool reject = (bestLevel == bestLevel2) && ((float)bestDist > lim);
            if (!reject && (bestLevel != bestLevel2 || (float)bestDist <= lim)) pick = bestIdx;
        }
        return pick;
    };

    int rounds = 0;
    while (true) {
        int changed = 0;
#pragma unroll
        for (int q = 0; q < kRegQ; q++) {
            if (r_i[q] < 0) continue;
            const int pick = evaluate(r_i[q], r_lv[q], r_nc[q], r_e01[q], r_e23[q], claim_prev);
            if (pick != r_choice[q]) { r_choice[q] = pick; s.choice[r_i[q]] = pick; changed = 1; }
        }
        for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) {
            const int i = qlist[j];
            const unsigned long long* e = s.cache + (size_t)i * kMatchCacheK;
            const int pick = evaluate(i, s.plevels[i], s.ncand[i], *reinterpret_cast<const ulonglong2*>(e),
                                      *reinterpret_cast<const ulonglong2*>(e + 2), claim_prev);
            if (pick != s.choice[i]) { s.choice[i] = pick; changed = 1; }
        }
        rounds++;
        if (!__syncthreads_or(changed)) break;
        for (int k = tid; k < ncur; k += kMatchThreads) claim_next[k] = blocked(k) ? -1 : kNoClaim;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kRegQ; q++)
            if (r_i[q] >= 0 && r_choice[q] >= 0 && obs_of(r_i[q])) atomicMin(&claim_next[r_choice[q]], r_i[q]);
        for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) {
            const int i = qlist[j];
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
    for (int j = tid; j < nact; j += kMatchThreads) {
        const int i = qlist[j];
        const int k = s.choice[i];
        if (k < 0) continue;
        atomicMax(&owner[k], i);
        atomicAdd(&s_events, 1);
    }
    __syncthreads();
    for (int k = tid; k < ncur; k += kMatchThreads) {
        cur_mp[k] = owner[k];
        if (a.merge_into && owner[k] >= 0) a.merge_into[k] = a.q_index ? a.q_index[owner[k]] : owner[k];
    }
    if (tid == 0) { *nmatches = s_events; *s.iters = rounds; }
}

void launch_match_map(const FrameDev& cur, const MatchMapArgs& a, const MatchScratch& s, int* cur_mp, int* nmatches,
                      cudaStream_t stream)
{
    const size_t smem = claim_smem_bytes(cur.cap);
    const int use_smem = smem <= kMatchSmemLimit ? 1 : 0;
    prepare_match_kernels();
    DVM_LAUNCH_PDL(match_map_walk_kernel, div_up(max(a.m, 1), kWalkThreads / 32), kWalkThreads, 0, stream, cur, a, s);
    DVM_LAUNCH_PDL(match_map_kernel, 1, kMatchThreads, use_smem ? smem : 0, stream, cur, a, s, cur_mp, nmatches, use_smem);
}


static bool prepare_match_kernels()
{
    static std::atomic<bool> done[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || done[dev].load()) return true;
    cudaFuncSetAttribute(match_last_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmemLimit);
    cudaFuncSetAttribute(match_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmemLimit);
    done[dev].store(true);
    return true;
}

} // namespace dvm
