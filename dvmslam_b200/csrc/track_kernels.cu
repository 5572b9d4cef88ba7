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
    __syncthreads();
    // what a window walk needs of the keypoint at grid position j, in one 16-byte record
    for (int j = tid; j < total; j += 1024) {
        const int i = f.cell_items[j];
        const dvm_keypoint* kp = f.kps + i;
        f.cell_rec[j] = make_int4(__float_as_int(kp->x), __float_as_int(kp->y), kp->octave, i);
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
constexpr int kMatchThreads = 1024;  // one CTA of 32 warps: a round is bound by the serial path of one warp, so few queries per thread
constexpr int kRegQ = 4;   // queries per thread kept in registers by the resolution kernels
constexpr int kMoreWords = 7 * kRegQ * kMatchThreads;   // shared-memory words for their candidates two to eight

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
// the same candidate in 32 bits for the registers of the resolution kernels: distance:9 | octave:7 | keypoint:16
__device__ inline unsigned cand_short(unsigned long long e)
{
    return e == ~0ull ? 0xffffffffu : ((unsigned)cand_dist(e) << 23) | ((unsigned)(cand_oct(e) & 0x7f) << 16) | (unsigned)cand_idx(e);
}
__device__ inline int short_idx(unsigned c) { return (int)(c & 0xffffu); }
__device__ inline int short_oct(unsigned c) { return (int)((c >> 16) & 0x7fu); }
__device__ inline int short_dist(unsigned c) { return (int)(c >> 23); }

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
        if (m == ~0ull) break; // nothing left in any lane (warp-uniform)
        if (top[0] == m) { // keys are unique: exactly one lane pops
#pragma unroll
            for (int q = 0; q + 1 < kMatchCacheK; q++) top[q] = top[q + 1];
            top[kMatchCacheK - 1] = ~0ull;
        }
    }
    return mine;
}

constexpr int kWalkThreads = 128;

// phase stamps for DVM_MATCH_PROFILE (thread 0 of a one-CTA kernel; min / max over the CTAs of a walk)
__device__ inline unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ inline void prof_stamp(unsigned long long* prof, int slot) { if (prof && threadIdx.x == 0) prof[slot] = gtimer(); }
__device__ inline void prof_walk_begin(unsigned long long* prof) { if (prof && threadIdx.x == 0) atomicMin(&prof[0], gtimer()); }
__device__ inline void prof_walk_end(unsigned long long* prof) { if (prof && threadIdx.x == 0) atomicMax(&prof[1], gtimer()); }

// ---- claims of the greedy rounds.  claim[k] = the smallest index among the queries (with observations) that chose keypoint
// k, 0xffffffff while nobody has.  Shared-memory atomics are what a round costs (two cycles per lane on the one SM this
// runs on), so a query claims only when its choice CHANGES, in place, while the others read: the choices move down the
// candidate lists and the holders' indices only fall, the process is monotone and ends in the unique state in which
// every query holds its best candidate not held by an earlier one -- the sequential outcome -- whatever the interleaving.
// The one non-monotone event is a query that gives up a keypoint it holds (the second-best test of
// SearchByProjection(F, map points) can turn against a candidate when the runner-up changes): the claim array is then
// rebuilt from the current choices before the next round.  A round that changes no choice ends the iteration.
// The common case -- the answer is among a query's first candidates -- is a few dozen instructions on registers and shared
// memory; everything else (the full walk, queries beyond the register-resident ones) sits behind one call that is not
// inlined. ----
constexpr unsigned kNoClaim = 0xffffffffu;
__device__ inline bool held_before(const unsigned* claim, int idx, int i) { return claim[idx] < (unsigned)i; }
constexpr int kIdxBits = 20;                       // query index | candidate count << 20 in one register
constexpr int kIdxMask = (1 << kIdxBits) - 1;
constexpr unsigned kWatchNone = 0xffffu, kWatchAlways = 0xfffeu;   // watched-keypoint markers (keypoint indices stay below 65532)

// what the slow path needs, gathered once per kernel (lives in local memory: its address is passed to a call)
struct SlowCtx {
    FrameLook fl;
    const unsigned long long* cache;
    const float* pu; const float* pv; const float* pr;
    const uint8_t* q_desc;      // descriptors of the queries ...
    const int* q_desc_index;    // ... indexed through this table when it is not null
    const int* cur_map;         // SearchByProjection(F, map points): keypoints blocked from the start
    const uint8_t* cur_blocked;
    unsigned long long* prof;
};

// One query of SearchByProjection(cur, last) (O3/src/ORBmatcher.cc:1570-1640): projects the map point of last-frame
// keypoint i with the pose prior, walks its window once (whole warp) and appends the query -- its index, level window,
// candidate count, angle and its kMatchCacheK best candidates -- to the compact list the resolution kernel reads.
__device__ inline void last_walk_query(const FrameDev& cur, const MatchLastArgs& a, const MatchScratch& s, const FrameLook& fl,
                                       const float* q, const float* t, float th, int i, int lane)
{
    const int mi = a.mp_index ? a.mp_index[i] : (a.has_mp[i] ? i : -1);
    const bool is_outlier = a.outlier[i] != 0;
    const int oct = a.last_kps ? a.last_kps[i].octave : a.octave[i];        // (loaded with the first round trip)
    const float ang = a.last_kps ? a.last_kps[i].angle : a.angle[i];
    if (mi < 0 || is_outlier) return;
    uint32_t d[8];
    load_desc(d, a.mp_desc + (size_t)mi * 32);                              // (with the second, next to the position)
    // x3Dc = Tcw * x3Dw (O3/src/ORBmatcher.cc:1577): Sophus' quaternion action, not a matrix product
    const float Pw[3] = { a.Xw[3 * mi], a.Xw[3 * mi + 1], a.Xw[3 * mi + 2] };
    float Pc[3];
    so::se3_apply(q, t, Pw, Pc);
    const float xc = Pc[0], yc = Pc[1], zc = Pc[2];
    const float invzc = (float)(1.0 / (double)zc);
    if (invzc < 0) return;
    const float u = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[0], xc), zc), a.K[2]);
    const float v = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[1], yc), zc), a.K[3]);
    if ((u < cur.minX || u > cur.maxX) || (v < cur.minY || v > cur.maxY)) return;
    // the conditions above depend on the query only: the whole warp is here together
    int slot = 0;
    if (lane == 0) slot = atomicAdd(s.qcount, 1);
    const float r = __fmul_rn(th, cur.scale[oct]);
    if (lane == 0) { s.pu[i] = u; s.pv[i] = v; s.pr[i] = r; }
    unsigned long long top[kMatchCacheK];
#pragma unroll
    for (int p = 0; p < kMatchCacheK; p++) top[p] = ~0ull;
    int nc = 0;
    walk_area_warp(fl, u, v, r, oct - 1, oct + 1, lane, [&](int idx, int o, int ord) {
        const int dist = hamming256(d, fl.desc + (size_t)idx * 32);
        topk_insert(top, pack_cand(dist, ord, o, idx));
        nc++;
    });
    nc = __reduce_add_sync(0xffffffffu, nc);
    const unsigned long long e = warp_topk_merge(top, lane);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (lane < kMatchCacheK) s.cache[(size_t)slot * kMatchCacheK + lane] = e;
    if (lane < kMatchCacheK) s.cache8[(size_t)slot * kMatchCacheK + lane] = cand_short(e);
    if (lane == 0) s.qmeta[slot] = make_int4(i, ((oct - 1) << 16) | ((oct + 1) & 0xffff), nc, __float_as_int(ang));
}

// ---- SearchByProjection(cur, last), phase 1: one WARP per last-frame keypoint (many CTAs; the frame is read through
// L1/L2) ----
__global__ void __launch_bounds__(kWalkThreads) match_last_walk_kernel(FrameDev cur, MatchLastArgs a, MatchScratch s)
{
    pdl_trigger(); pdl_wait();
    if (a.guard && *a.guard >= 20) return; // enough matches at th: the wider retry is not run
    prof_walk_begin(s.prof);
    const int nq = a.n_ptr ? *a.n_ptr : a.last_n;
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * kWalkThreads + threadIdx.x) >> 5;
    if (i < nq) {
        const FrameLook fl = look_global(cur);
        if (a.pose) { // pose prior held on the device
            a.q[0] = a.pose[0]; a.q[1] = a.pose[1]; a.q[2] = a.pose[2]; a.q[3] = a.pose[3];
            a.t[0] = a.pose[4]; a.t[1] = a.pose[5]; a.t[2] = a.pose[6];
        }
        last_walk_query(cur, a, s, fl, a.q, a.t, a.th, i, lane);
    }
    prof_walk_end(s.prof);
}

// slow path of one query of SearchByProjection(cur, last): all cached candidates, then the full walk
__device__ __noinline__ int slow_pick_last(const SlowCtx* c, int j, int i, int lv, int nc, const unsigned* claim)
{
    if (c->prof) atomicAdd(&c->prof[10], 1ull);
    const unsigned long long* e = c->cache + (size_t)j * kMatchCacheK;
    for (int p = 0; p < kMatchCacheK && p < nc; p++) {
        const unsigned long long ev = e[p];
        if (held_before(claim, cand_idx(ev), i)) continue; // taken by an earlier map point with observations
        return cand_dist(ev) <= kThHigh ? cand_idx(ev) : -1;
    }
    if (nc <= kMatchCacheK) return -1;
    if (c->prof) atomicAdd(&c->prof[8], 1ull);
    int bestDist = 256, bestIdx = -1; // cache exhausted: full walk against the current claims
    uint32_t d[8];
    load_desc(d, c->q_desc + (size_t)(c->q_desc_index ? c->q_desc_index[i] : i) * 32);
    walk_area(c->fl, c->pu[i], c->pv[i], c->pr[i], lv >> 16, (int)(short)(lv & 0xffff), [&](int idx, int) {
        if (held_before(claim, idx, i)) return;
        const int dist = hamming256(d, c->fl.desc + (size_t)idx * 32);
        if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    });
    return bestDist <= kThHigh ? bestIdx : -1;
}

// ---- phase 2: the sequential-greedy outcome as a Jacobi fixed point on ONE CTA (claims in shared
// memory), then the rotation-consistency check.  With a.retry_th > 0 the kernel also holds Tracking's retry
// (Tracking.cc:2614-2621): fewer than 20 matches -> the CTA walks every query again with the wider window (a rare,
// nearly-lost frame: one CTA is enough) and resolves once more; the outputs of the second pass replace the first's. ----
template <bool kClaimsInSmem>
__global__ void __launch_bounds__(kMatchThreads, 1)
match_last_kernel(FrameDev cur, MatchLastArgs a, MatchScratch s, int* __restrict__ cur_mp, int* __restrict__ nmatches)
{
    pdl_trigger(); pdl_wait();
    extern __shared__ __align__(16) unsigned char claim_smem[];
    __shared__ int histo[kHistoLength];
    __shared__ int s_events, s_bad;
    const int tid = threadIdx.x, lane = tid & 31;
    if (a.guard && *a.guard >= 20) return; // enough matches at th: the wider retry is not run
    prof_stamp(s.prof, 2);
    const int ncur = min(*cur.n, cur.cap);
    SlowCtx sc;
    sc.fl = look_global(cur);
    sc.cache = s.cache; sc.pu = s.pu; sc.pv = s.pv; sc.pr = s.pr;
    sc.q_desc = a.mp_desc; sc.q_desc_index = a.mp_index; sc.cur_map = nullptr; sc.cur_blocked = nullptr; sc.prof = s.prof;
    const bool all_obs = a.obs_pos == nullptr;   // every query blocks later ones (the tracker): a keypoint's chooser is unique
    auto obs_of = [&](int i) { return a.obs_pos ? a.obs_pos[i] != 0 : true; };
    // shared memory: candidates two to eight of the register-resident queries, the claim array, the owner array
    unsigned* more = reinterpret_cast<unsigned*>(claim_smem);
    unsigned* claim = kClaimsInSmem ? more + kMoreWords : reinterpret_cast<unsigned*>(s.claim_a);
    int* owner = kClaimsInSmem ? reinterpret_cast<int*>(more + kMoreWords + cur.cap) : s.claim_b;

    for (int attempt = 0; attempt < 2; attempt++) {
        const int nact = *reinterpret_cast<volatile int*>(s.qcount);
        // the first kRegQ queries of every thread stay in registers across the rounds: index and candidate count in one
        // word, the best candidate, the current choice and the WATCHED keypoint -- the first candidate no earlier query
        // holds.  Keypoints held by earlier queries stay held (the process is monotone), so between two evaluations the only
        // thing that can change a query's answer is its watched keypoint being taken: a round costs one shared-memory load
        // and a compare per query, and the (few) queries that lost their keypoint are re-evaluated by one copy of the scan.
        int r_in[kRegQ], r_choice[kRegQ];   // r_in: index | min(candidates, 255) << 20, or -1
        unsigned r_c0[kRegQ], r_watch[kRegQ];
#pragma unroll
        for (int q = 0; q < kRegQ; q++) {
            const int j = tid + q * kMatchThreads;
            r_in[q] = -1; r_c0[q] = ~0u; r_choice[q] = -1; r_watch[q] = kWatchNone;
            if (j < nact) {
                const int4 m = s.qmeta[j];
                const uint4 c = reinterpret_cast<const uint4*>(s.cache8)[2 * j], c2 = reinterpret_cast<const uint4*>(s.cache8)[2 * j + 1];
                r_in[q] = m.x | (min(m.z, 255) << kIdxBits);
                r_c0[q] = c.x;
                unsigned* mo = more + 7 * j;
                mo[0] = c.y; mo[1] = c.z; mo[2] = c.w; mo[3] = c2.x; mo[4] = c2.y; mo[5] = c2.z; mo[6] = c2.w;
                // first evaluation, nothing claimed yet: the best candidate
                if (m.z > 0) {
                    r_watch[q] = (unsigned)short_idx(c.x);
                    r_choice[q] = short_dist(c.x) <= kThHigh ? short_idx(c.x) : -1;
                }
            }
        }
        for (int k = tid; k < ncur; k += kMatchThreads) claim[k] = kNoClaim;
        for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) s.choice[j] = -1;
        if (tid < kHistoLength) histo[tid] = 0;
        if (tid == 0) { s_events = 0; s_bad = 0; }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < kRegQ; q++)
            if (r_choice[q] >= 0 && obs_of(r_in[q] & kIdxMask)) atomicMin(&claim[r_choice[q]], (unsigned)(r_in[q] & kIdxMask));
        __syncthreads();
        prof_stamp(s.prof, 4);
        if (s.prof && tid == 0) s.prof[9] = (unsigned long long)nact;

        int rounds = 1;
        while (true) {
            int changed = 0;
            unsigned redo = 0u;
#pragma unroll
            for (int q = 0; q < kRegQ; q++) {
                // (inactive slots watch nothing; kWatchAlways = the answer came from the full walk: evaluate every round)
                const unsigned w = r_watch[q];
                const bool stale = w == kWatchAlways || (w != kWatchNone && claim[w] < (unsigned)(r_in[q] & kIdxMask));
                if (stale) redo |= 1u << q;
            }
#pragma unroll 1
            while (redo) { // the watched keypoint was taken by an earlier query: scan the candidates again
                const int q = __ffs(redo) - 1;
                redo &= redo - 1u;
                const int j = tid + q * kMatchThreads;
                int rin = r_in[0];
                unsigned c0 = r_c0[0];
#pragma unroll
                for (int qq = 1; qq < kRegQ; qq++) { rin = qq == q ? r_in[qq] : rin; c0 = qq == q ? r_c0[qq] : c0; }
                const int i = rin & kIdxMask, nc = rin >> kIdxBits;
                int pick = -1;
                unsigned watch = kWatchNone;
                bool decided = false;
                for (int p = 0; p < kMatchCacheK && p < nc; p++) {
                    const unsigned c = p == 0 ? c0 : more[7 * j + p - 1];
                    if (held_before(claim, short_idx(c), i)) continue;
                    pick = short_dist(c) <= kThHigh ? short_idx(c) : -1;
                    watch = (unsigned)short_idx(c); decided = true;
                    break;
                }
                if (!decided && nc > kMatchCacheK) {
                    pick = slow_pick_last(&sc, j, i, s.qmeta[j].y, s.qmeta[j].z, claim);
                    watch = kWatchAlways;
                }
                int prev = r_choice[0];
#pragma unroll
                for (int qq = 1; qq < kRegQ; qq++) prev = qq == q ? r_choice[qq] : prev;
#pragma unroll
                for (int qq = 0; qq < kRegQ; qq++) { r_choice[qq] = qq == q ? pick : r_choice[qq]; r_watch[qq] = qq == q ? watch : r_watch[qq]; }
                if (pick != prev) {
                    changed = 1;
                    if (pick >= 0 && obs_of(i)) atomicMin(&claim[pick], (unsigned)i);
                }
            }
#pragma unroll 1
            for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) {
                const int4 m = s.qmeta[j];
                const int pick = slow_pick_last(&sc, j, m.x, m.y, m.z, claim);
                if (pick != s.choice[j]) {
                    s.choice[j] = pick; changed = 1;
                    if (pick >= 0 && obs_of(m.x)) atomicMin(&claim[pick], (unsigned)m.x);
                }
            }
            rounds++;
            if (s.prof && tid == 0 && rounds <= 12) s.prof[16 + rounds] = gtimer();
            // (a query leaves a keypoint only for one an earlier query took from it: no claim is ever given up here)
            if (!__syncthreads_or(changed)) break;
            if (s.prof && tid == 0 && rounds <= 12) s.prof[32 + rounds] = gtimer();
            if (rounds > nact + 8) break; // (the fixed point is reached within one round per query; never seen, kept as a fuse)
        }
        prof_stamp(s.prof, 5);

        // every accepted query is one write "CurrentFrame.mvpMapPoints[k] = pMP" (+ one histogram entry); the last writer of a
        // keypoint owns it.  With observations everywhere the chooser of a keypoint is unique and the claim array IS the owner.
        if (!all_obs) {
            for (int k = tid; k < ncur; k += kMatchThreads) owner[k] = -1;
            __syncthreads();
        }
        int r_bin[kRegQ];
        int events = 0;
#pragma unroll
        for (int q = 0; q < kRegQ; q++) {
            r_bin[q] = -1;
            const bool has = r_in[q] >= 0 && r_choice[q] >= 0;
            if (has) {
                events++;
                if (!all_obs) atomicMax(&owner[r_choice[q]], r_in[q] & kIdxMask);
                if (a.check_ori) r_bin[q] = rot_bin(__int_as_float(s.qmeta[tid + q * kMatchThreads].w), cur.kps[r_choice[q]].angle);
            }
            if (a.check_ori) { // one shared-memory atomic per distinct bin of the warp
                const unsigned peers = __match_any_sync(0xffffffffu, r_bin[q]);
                if (r_bin[q] >= 0 && lane == __ffs(peers) - 1) atomicAdd(&histo[r_bin[q]], __popc(peers));
            }
        }
#pragma unroll 1
        for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) {
            const int k = s.choice[j];
            if (k < 0) continue;
            const int4 m = s.qmeta[j];
            events++;
            if (!all_obs) atomicMax(&owner[k], m.x);
            if (a.check_ori) atomicAdd(&histo[rot_bin(__int_as_float(m.w), cur.kps[k].angle)], 1);
        }
        events = __reduce_add_sync(0xffffffffu, events);
        if (lane == 0 && events) atomicAdd(&s_events, events);
        __syncthreads();
        prof_stamp(s.prof, 6);
        int* own = all_obs ? reinterpret_cast<int*>(claim) : owner;   // kNoClaim reads as -1
        if (a.check_ori) {
            // ComputeThreeMaxima by every warp for itself: the sequential scan with its strict comparisons keeps the three
            // largest positive bins, equal counts in index order -- three warp maxima of count << 5 | (31 - bin)
            unsigned key = (lane < kHistoLength && histo[lane] > 0) ? ((unsigned)histo[lane] << 5) | (unsigned)(31 - lane) : 0u;
            int ind[3], mx[3];
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const unsigned best = __reduce_max_sync(0xffffffffu, key);
                ind[m] = best ? 31 - (int)(best & 31u) : -1;
                mx[m] = (int)(best >> 5);
                if (key == best) key = 0u;
            }
            if ((float)mx[1] < __fmul_rn(0.1f, (float)mx[0])) { ind[1] = -1; ind[2] = -1; }
            else if ((float)mx[2] < __fmul_rn(0.1f, (float)mx[0])) ind[2] = -1;
            // a keypoint whose match falls outside the three bins is cleared, whoever wrote it last (ORBmatcher.cc:1733-1741)
            int bad = 0;
            auto drop = [&](int k, int bin) {
                if (bin != ind[0] && bin != ind[1] && bin != ind[2]) { bad++; atomicOr(reinterpret_cast<unsigned*>(&own[k]), 0x40000000u); }
            };
#pragma unroll
            for (int q = 0; q < kRegQ; q++)
                if (r_bin[q] >= 0) drop(r_choice[q], r_bin[q]);
#pragma unroll 1
            for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) {
                const int k = s.choice[j];
                if (k >= 0) drop(k, rot_bin(__int_as_float(s.qmeta[j].w), cur.kps[k].angle));
            }
            bad = __reduce_add_sync(0xffffffffu, bad);
            if (lane == 0 && bad) atomicAdd(&s_bad, bad);
            __syncthreads();
        }
        // own[k]: -1 nobody, index | 0x40000000 cleared by the rotation check, index otherwise
        for (int k = tid; k < cur.cap; k += kMatchThreads) {
            const int o = k < ncur ? own[k] : -1;
            const int keep = (o < 0 || (o & 0x40000000)) ? -1 : o;
            if (k < ncur) cur_mp[k] = keep;
            if (a.map_out) a.map_out[k] = keep >= 0 ? a.mp_index[keep] : -1; // CurrentFrame.mvpMapPoints as map indices
        }
        const int found = s_events - s_bad;
        if (tid == 0) { *nmatches = found; *s.iters = rounds; *s.qcount = 0; }
        prof_stamp(s.prof, 7);
        if (attempt == 1 || !(a.retry_th > 0.f) || found >= 20) break;
        // ---- nmatches < 20: SearchByProjection(cur, last, 2 * th) (Tracking.cc:2614-2621), walked by this CTA ----
        __syncthreads(); // own / s_events were read by everybody; the reset of qcount is ordered before the appends
        {
            const int nq = a.n_ptr ? *a.n_ptr : a.last_n;
            float q4[4], t3[3];
            for (int c = 0; c < 4; c++) q4[c] = a.pose ? a.pose[c] : a.q[c];
            for (int c = 0; c < 3; c++) t3[c] = a.pose ? a.pose[4 + c] : a.t[c];
            for (int i = tid >> 5; i < nq; i += kMatchThreads / 32) last_walk_query(cur, a, s, sc.fl, q4, t3, a.retry_th, i, lane);
        }
        __threadfence_block();
        __syncthreads();
    }
}

constexpr size_t kMatchSmemLimit = 200 * 1024;
static bool prepare_match_kernels(); // raises the dynamic shared memory limit of both matcher kernels once
__host__ inline size_t claim_smem_bytes(int cap) { return (size_t)cap * 8; }
__host__ inline size_t match_smem_bytes(int use_smem, int cap) { return kMoreWords * 4 + (use_smem ? claim_smem_bytes(cap) : 0); }

// DVM_MATCH_PROFILE: prints the phase times of every matcher launch (synchronises: diagnostics only)
struct MatchProfile {
    unsigned long long* d = nullptr;
    bool on = false;
    MatchProfile() { on = getenv("DVM_MATCH_PROFILE") != nullptr; }
    void arm(MatchScratch& s, cudaStream_t st)
    {
        if (!on) return;
        if (!d) cudaMalloc(&d, 64 * 8);
        cudaMemsetAsync(d, 0, 64 * 8, st);
        cudaMemsetAsync(d, 0xff, 8, st);
        s.prof = d;
    }
    void report(const char* what, const MatchScratch& s, cudaStream_t st)
    {
        if (!on) return;
        unsigned long long p[64];
        int rounds = 0;
        cudaStreamSynchronize(st);
        cudaMemcpy(p, d, sizeof(p), cudaMemcpyDeviceToHost);
        cudaMemcpy(&rounds, s.iters, 4, cudaMemcpyDeviceToHost);
        if (p[2] == 0) { fprintf(stderr, "[%s] skipped\n", what); return; }
        auto us = [&](int a, int b) { return p[a] && p[b] ? (double)(long long)(p[b] - p[a]) * 1e-3 : -1.0; };
        fprintf(stderr, "[%s us] walk %.1f gap %.1f | stage %.1f rounds(%d) %.1f owner %.1f tail %.1f | total %.1f | active %llu slow calls %llu full walks %llu\n",
                what, us(0, 1), us(1, 2), us(2, 4), rounds, us(4, 5), p[6] ? us(5, 6) : 0.0, p[6] ? us(6, 7) : us(5, 7), us(0, 7), p[9], p[10], p[8]);
        fprintf(stderr, "    rounds (work / barrier wait, us):");
        for (int r = 2; r <= 12 && p[16 + r]; r++)
            fprintf(stderr, " %.2f/%.2f", (double)(long long)(p[16 + r] - (r == 2 ? p[4] : p[32 + r - 1])) * 1e-3,
                    p[32 + r] ? (double)(long long)(p[32 + r] - p[16 + r]) * 1e-3 : 0.0);
        fprintf(stderr, "\n");
    }
};
static MatchProfile g_match_profile;

void launch_match_last(const FrameDev& cur, const MatchLastArgs& a, const MatchScratch& s_in, int* cur_mp, int* nmatches,
                       cudaStream_t stream)
{
    const int use_smem = match_smem_bytes(1, cur.cap) <= kMatchSmemLimit ? 1 : 0;
    prepare_match_kernels();
    MatchScratch s = s_in;
    g_match_profile.arm(s, stream);
    DVM_LAUNCH_PDL(match_last_walk_kernel, div_up(max(a.last_n, 1), kWalkThreads / 32), kWalkThreads, 0, stream, cur, a, s);
    if (use_smem) DVM_LAUNCH_PDL(match_last_kernel<true>, 1, kMatchThreads, match_smem_bytes(1, cur.cap), stream, cur, a, s, cur_mp, nmatches);
    else DVM_LAUNCH_PDL(match_last_kernel<false>, 1, kMatchThreads, match_smem_bytes(0, cur.cap), stream, cur, a, s, cur_mp, nmatches);
    g_match_profile.report("match-last", s, stream);
}

// --------------------------------------------------------------- SearchByProjection(F, mapPoints)
// phase 1: one warp per in-view map point walks its window once
__global__ void __launch_bounds__(kWalkThreads) match_map_walk_kernel(FrameDev cur, MatchMapArgs a, MatchScratch s)
{
    pdl_trigger(); pdl_wait();
    prof_walk_begin(s.prof);
    struct Stamp { unsigned long long* p; __device__ ~Stamp() { prof_walk_end(p); } } stamp_end{ s.prof };
    const int nq = a.m_ptr ? *a.m_ptr : a.m;
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * kWalkThreads + threadIdx.x) >> 5;
    if (i >= nq) return;
    const FrameLook fl = look_global(cur);
    int lvl;
    float px, py, vcos;
    uint32_t d[8];   // (loaded before the gates: its latency overlaps theirs)
    load_desc(d, a.mp_desc + (size_t)(a.q_index ? a.q_index[i] : i) * 32);
    if (a.use_frustum) { // SearchLocalPoints: isInFrustum decides whether map point i takes part at all
        FrustumPose fp;
        frustum_pose(a.fr.pose, fp);
        if (!frustum_eval(a.fr, fp, i, px, py, lvl, vcos)) return;
    } else {
        lvl = a.level[i]; px = a.projX[i]; py = a.projY[i]; vcos = a.sim3_mode ? 0.f : a.view_cos[i];
        if (a.sim3_mode && lvl < 0) return; // rejected by the projection gates
    }
    int slot = 0;
    if (lane == 0) slot = atomicAdd(s.qcount, 1);
    float r;
    if (a.sim3_mode) r = __fmul_rn(a.th, cur.scale[lvl]);   // const float radius = th * pKF->mvScaleFactors[nPredictedLevel]
    else {
        r = vcos > 0.998f ? 2.5f : 4.0f; // RadiusByViewingCos
        if (a.th != 1.0f) r = __fmul_rn(r, a.th);
        r = __fmul_rn(r, cur.scale[lvl]);
    }
    if (lane == 0) { s.pu[i] = px; s.pv[i] = py; s.pr[i] = r; }
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
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (lane < kMatchCacheK) s.cache[(size_t)slot * kMatchCacheK + lane] = e;
    if (lane < kMatchCacheK) s.cache8[(size_t)slot * kMatchCacheK + lane] = cand_short(e);
    if (lane == 0) s.qmeta[slot] = make_int4(i, lvl, nc, 0);
}

// the acceptance rule of SearchByProjection(F, vpMapPoints) (ORBmatcher.cc:150-170) / of the Sim3-guided search (:470-480)
__device__ inline int map_decide(const MatchMapArgs& a, int bestDist, int bestLevel, int bestDist2, int bestLevel2, int bestIdx)
{
    if (a.sim3_mode) return ((float)bestDist <= a.accept_limit && bestDist < 256) ? bestIdx : -1;
    if (bestDist <= kThHigh) {
        const float lim = __fmul_rn(a.nnratio, (float)bestDist2);
        const bool reject = (bestLevel == bestLevel2) && ((float)bestDist > lim);
        if (!reject && (bestLevel != bestLevel2 || (float)bestDist <= lim)) return bestIdx;
    }
    return -1;
}
struct MapDecide { int sim3_mode; float accept_limit, nnratio; };
__device__ inline int map_decide(const MapDecide& a, int bestDist, int bestLevel, int bestDist2, int bestLevel2, int bestIdx)
{
    if (a.sim3_mode) return ((float)bestDist <= a.accept_limit && bestDist < 256) ? bestIdx : -1;
    if (bestDist <= kThHigh) {
        const float lim = __fmul_rn(a.nnratio, (float)bestDist2);
        const bool reject = (bestLevel == bestLevel2) && ((float)bestDist > lim);
        if (!reject && (bestLevel != bestLevel2 || (float)bestDist <= lim)) return bestIdx;
    }
    return -1;
}

// slow path of one query of SearchByProjection(F, vpMapPoints): all cached candidates, then the full walk
__device__ __noinline__ int slow_pick_map(const SlowCtx* c, const MapDecide* md, int j, int i, int lvl, int nc, const unsigned* claim)
{
    if (c->prof) atomicAdd(&c->prof[10], 1ull);
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    int found = 0;
    // sorted by (distance, walk order): the first two free entries are the reference's best and
    // second best (its scan keeps the earliest of equal distances)
    const unsigned long long* e = c->cache + (size_t)j * kMatchCacheK;
    for (int p = 0; p < kMatchCacheK && p < nc && found < 2; p++) {
        const unsigned long long ev = e[p];
        const int idx = cand_idx(ev);
        if (held_before(claim, idx, i)) continue; // held by a map point with observations
        if (found == 0) { bestDist = cand_dist(ev); bestLevel = cand_oct(ev); bestIdx = idx; }
        else { bestDist2 = cand_dist(ev); bestLevel2 = cand_oct(ev); }
        found++;
    }
    if (found < 2 && nc > kMatchCacheK) { // cache exhausted: full walk against the current claims
        if (c->prof) atomicAdd(&c->prof[8], 1ull);
        uint32_t d[8];
        load_desc(d, c->q_desc + (size_t)(c->q_desc_index ? c->q_desc_index[i] : i) * 32);
        bestDist = 256; bestLevel = -1; bestDist2 = 256; bestLevel2 = -1; bestIdx = -1;
        walk_area(c->fl, c->pu[i], c->pv[i], c->pr[i], lvl - 1, lvl, [&](int idx, int oct) {
            if (c->cur_map ? c->cur_map[idx] >= 0 : (c->cur_blocked && c->cur_blocked[idx])) return;
            if (held_before(claim, idx, i)) return;
            const int dist = hamming256(d, c->fl.desc + (size_t)idx * 32);
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
    return map_decide(*md, bestDist, bestLevel, bestDist2, bestLevel2, bestIdx);
}

template <bool kClaimsInSmem>
__global__ void __launch_bounds__(kMatchThreads, 1)
match_map_kernel(FrameDev cur, MatchMapArgs a, MatchScratch s, int* __restrict__ cur_mp, int* __restrict__ nmatches)
{
    pdl_trigger(); pdl_wait();
    extern __shared__ __align__(16) unsigned char claim_smem[];
    __shared__ int s_events, s_dirty;
    const int tid = threadIdx.x, lane = tid & 31;
    prof_stamp(s.prof, 2);
    const int ncur = min(*cur.n, cur.cap);
    const int nact = *s.qcount; // queries in the frustum (any order: the greedy priority is the query index itself)
    SlowCtx sc;
    sc.fl = look_global(cur);
    sc.cache = s.cache; sc.pu = s.pu; sc.pv = s.pv; sc.pr = s.pr;
    sc.q_desc = a.mp_desc; sc.q_desc_index = a.q_index; sc.cur_map = a.cur_map; sc.cur_blocked = a.cur_blocked; sc.prof = s.prof;
    MapDecide md;
    md.sim3_mode = a.sim3_mode; md.accept_limit = a.accept_limit; md.nnratio = a.nnratio;
    const bool all_obs = a.obs_pos == nullptr;
    auto obs_of = [&](int i) { return a.obs_pos ? a.obs_pos[i] != 0 : true; };
    unsigned* more = reinterpret_cast<unsigned*>(claim_smem);
    unsigned* claim = kClaimsInSmem ? more + kMoreWords : reinterpret_cast<unsigned*>(s.claim_a);
    int* owner = kClaimsInSmem ? reinterpret_cast<int*>(more + kMoreWords + cur.cap) : s.claim_b;

    // register-resident queries: index | candidate count, the best candidate, the choice and the two WATCHED keypoints -- the
    // first two candidates that no earlier query holds, which are all the acceptance rule looks at.  Held keypoints stay held
    // (monotone process), so a round is two shared-memory loads and compares per query; a query is evaluated again only when
    // one of its watched keypoints was taken (or after the rare rebuild of the claims, which re-evaluates everybody).
    int r_in[kRegQ], r_choice[kRegQ];
    unsigned r_c0[kRegQ], r_watch[kRegQ];   // r_watch: first | second << 16
    const MapDecide mdr = { a.sim3_mode, a.accept_limit, a.nnratio };   // register copy (md's address goes to the slow path)
    // best / second best as packed candidates (~0 = none) -> the reference's acceptance rule
    auto decide = [&](unsigned first, unsigned second) {
        return map_decide(mdr, first == ~0u ? 256 : short_dist(first), first == ~0u ? -1 : short_oct(first),
                          second == ~0u ? 256 : short_dist(second), second == ~0u ? -1 : short_oct(second),
                          first == ~0u ? -1 : short_idx(first));
    };
#pragma unroll
    for (int q = 0; q < kRegQ; q++) {
        const int j = tid + q * kMatchThreads;
        r_in[q] = -1; r_c0[q] = ~0u; r_choice[q] = -1; r_watch[q] = kWatchNone | (kWatchNone << 16);
        if (j < nact) {
            const int4 m = s.qmeta[j];
            const uint4 c = reinterpret_cast<const uint4*>(s.cache8)[2 * j], c2 = reinterpret_cast<const uint4*>(s.cache8)[2 * j + 1];
            r_in[q] = m.x | (min(m.z, 255) << kIdxBits);
            r_c0[q] = c.x;
            unsigned* mo = more + 7 * j;
            mo[0] = c.y; mo[1] = c.z; mo[2] = c.w; mo[3] = c2.x; mo[4] = c2.y; mo[5] = c2.z; mo[6] = c2.w;
            // first evaluation, nothing claimed yet: the two best candidates
            const unsigned first = m.z > 0 ? c.x : ~0u, second = m.z > 1 ? c.y : ~0u;
            r_watch[q] = (first == ~0u ? kWatchNone : (unsigned)short_idx(first)) | ((second == ~0u ? kWatchNone : (unsigned)short_idx(second)) << 16);
            r_choice[q] = decide(first, second);
        }
    }
    for (int k = tid; k < ncur; k += kMatchThreads) claim[k] = kNoClaim;
    if (tid == 0) { s_events = 0; s_dirty = 0; }
    for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) s.choice[j] = -1;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kRegQ; q++)
        if (r_choice[q] >= 0 && obs_of(r_in[q] & kIdxMask)) atomicMin(&claim[r_choice[q]], (unsigned)(r_in[q] & kIdxMask));
    __syncthreads();
    prof_stamp(s.prof, 4);
    if (s.prof && tid == 0) s.prof[9] = (unsigned long long)nact;
    // a changed choice: claim the new keypoint; giving up a keypoint this query HOLDS is the one non-monotone event
    // (the claim array is then rebuilt)
    int dirty = 0;
    auto move_to = [&](int i, int prev, int pick) {
        if (!obs_of(i)) return;
        if (prev >= 0 && claim[prev] == (unsigned)i) dirty = 1;
        if (pick >= 0) atomicMin(&claim[pick], (unsigned)i);
    };

    int rounds = 1;
    bool everybody = false;   // after a rebuild of the claims
    while (true) {
        int changed = 0;
        unsigned redo = 0u;
#pragma unroll
        for (int q = 0; q < kRegQ; q++) {
            const unsigned w0 = r_watch[q] & 0xffffu, w1 = r_watch[q] >> 16;
            const unsigned i = (unsigned)(r_in[q] & kIdxMask);
            const bool stale = w0 == kWatchAlways || (w0 != kWatchNone && claim[w0] < i) || (w1 != kWatchNone && claim[w1] < i) ||
                               (everybody && r_in[q] >= 0);
            if (stale) redo |= 1u << q;
        }
        everybody = false;
#pragma unroll 1
        while (redo) { // a watched keypoint was taken by an earlier query: scan the candidates again
            const int q = __ffs(redo) - 1;
            redo &= redo - 1u;
            const int j = tid + q * kMatchThreads;
            int rin = r_in[0];
            unsigned c0 = r_c0[0];
#pragma unroll
            for (int qq = 1; qq < kRegQ; qq++) { rin = qq == q ? r_in[qq] : rin; c0 = qq == q ? r_c0[qq] : c0; }
            const int i = rin & kIdxMask, nc = rin >> kIdxBits;
            unsigned first = ~0u, second = ~0u;
            for (int p = 0; p < kMatchCacheK && p < nc && second == ~0u; p++) {
                const unsigned c = p == 0 ? c0 : more[7 * j + p - 1];
                if (held_before(claim, short_idx(c), i)) continue;
                if (first == ~0u) first = c; else second = c;
            }
            int pick;
            unsigned watch;
            if (second == ~0u && nc > kMatchCacheK) {
                pick = slow_pick_map(&sc, &md, j, i, s.qmeta[j].y, s.qmeta[j].z, claim);
                watch = kWatchAlways | (kWatchNone << 16);
            } else {
                pick = decide(first, second);
                watch = (first == ~0u ? kWatchNone : (unsigned)short_idx(first)) | ((second == ~0u ? kWatchNone : (unsigned)short_idx(second)) << 16);
            }
            int prev = r_choice[0];
#pragma unroll
            for (int qq = 1; qq < kRegQ; qq++) prev = qq == q ? r_choice[qq] : prev;
#pragma unroll
            for (int qq = 0; qq < kRegQ; qq++) { r_choice[qq] = qq == q ? pick : r_choice[qq]; r_watch[qq] = qq == q ? watch : r_watch[qq]; }
            if (pick != prev) { move_to(i, prev, pick); changed = 1; }
        }
#pragma unroll 1
        for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) {
            const int4 m = s.qmeta[j];
            const int pick = slow_pick_map(&sc, &md, j, m.x, m.y, m.z, claim);
            const int prev = s.choice[j];
            if (pick != prev) { move_to(m.x, prev, pick); s.choice[j] = pick; changed = 1; }
        }
        rounds++;
        if (s.prof && tid == 0 && rounds <= 12) s.prof[16 + rounds] = gtimer();
        if (dirty) s_dirty = 1;
        if (!__syncthreads_or(changed)) break;
        if (s.prof && tid == 0 && rounds <= 12) s.prof[32 + rounds] = gtimer();
        if (rounds > nact + 8) break; // (the fixed point is reached within one round per query; never seen, kept as a fuse)
        if (s_dirty) { // a held keypoint was given up: rebuild the claims from the current choices, then look at everybody again
            dirty = 0;
            everybody = true;
            for (int k = tid; k < ncur; k += kMatchThreads) claim[k] = kNoClaim;
            __syncthreads();
            if (tid == 0) s_dirty = 0;
#pragma unroll
            for (int q = 0; q < kRegQ; q++)
                if (r_in[q] >= 0 && r_choice[q] >= 0 && obs_of(r_in[q] & kIdxMask)) atomicMin(&claim[r_choice[q]], (unsigned)(r_in[q] & kIdxMask));
#pragma unroll 1
            for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) {
                const int k = s.choice[j];
                const int i = s.qmeta[j].x;
                if (k >= 0 && obs_of(i)) atomicMin(&claim[k], (unsigned)i);
            }
            __syncthreads();
        }
    }
    prof_stamp(s.prof, 5);
    // the last writer of a keypoint owns it; with observations everywhere the chooser is unique and the claims are the owners
    if (!all_obs) {
        for (int k = tid; k < ncur; k += kMatchThreads) owner[k] = -1;
        __syncthreads();
    }
    int events = 0;
#pragma unroll
    for (int q = 0; q < kRegQ; q++) {
        if (r_in[q] < 0 || r_choice[q] < 0) continue;
        events++;
        if (!all_obs) atomicMax(&owner[r_choice[q]], r_in[q] & kIdxMask);
    }
#pragma unroll 1
    for (int j = tid + kRegQ * kMatchThreads; j < nact; j += kMatchThreads) {
        const int k = s.choice[j];
        if (k < 0) continue;
        events++;
        if (!all_obs) atomicMax(&owner[k], s.qmeta[j].x);
    }
    events = __reduce_add_sync(0xffffffffu, events);
    if (lane == 0 && events) atomicAdd(&s_events, events);
    __syncthreads();
    const int* own = all_obs ? reinterpret_cast<const int*>(claim) : owner;   // kNoClaim reads as -1
    for (int k = tid; k < ncur; k += kMatchThreads) {
        const int o = own[k];
        cur_mp[k] = o;
        if (a.merge_into && o >= 0) a.merge_into[k] = a.q_index ? a.q_index[o] : o;
    }
    if (tid == 0) { *nmatches = s_events; *s.iters = rounds; *s.qcount = 0; }
    prof_stamp(s.prof, 7);
}

void launch_match_map(const FrameDev& cur, const MatchMapArgs& a, const MatchScratch& s_in, int* cur_mp, int* nmatches,
                      cudaStream_t stream)
{
    const int use_smem = match_smem_bytes(1, cur.cap) <= kMatchSmemLimit ? 1 : 0;
    prepare_match_kernels();
    MatchScratch s = s_in;
    g_match_profile.arm(s, stream);
    DVM_LAUNCH_PDL(match_map_walk_kernel, div_up(max(a.m, 1), kWalkThreads / 32), kWalkThreads, 0, stream, cur, a, s);
    if (use_smem) DVM_LAUNCH_PDL(match_map_kernel<true>, 1, kMatchThreads, match_smem_bytes(1, cur.cap), stream, cur, a, s, cur_mp, nmatches);
    else DVM_LAUNCH_PDL(match_map_kernel<false>, 1, kMatchThreads, match_smem_bytes(0, cur.cap), stream, cur, a, s, cur_mp, nmatches);
    g_match_profile.report("match-map", s, stream);
}


static bool prepare_match_kernels()
{
    static std::atomic<bool> done[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || done[dev].load()) return true;
    cudaFuncSetAttribute(match_last_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmemLimit);
    cudaFuncSetAttribute(match_map_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmemLimit);
    cudaFuncSetAttribute(match_last_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmemLimit);
    cudaFuncSetAttribute(match_map_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmemLimit);
    done[dev].store(true);
    return true;
}

} // namespace dvm
