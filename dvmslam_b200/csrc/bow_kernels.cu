// bow_kernels.cu -- sm_100a kernels of the descriptor matchers that do not project a map point.
//
//   bow_node_kernel / bow_finish_kernel    ORBmatcher::SearchByBoW(KF, Frame)  (O3/src/ORBmatcher.cc:214-393)
//                                          ORBmatcher::SearchByBoW(KF, KF)     (O3/src/ORBmatcher.cc:709-834)
//   init_match_kernel                      ORBmatcher::SearchForInitialization (O3/src/ORBmatcher.cc:605-707)
//   triangulation_kernel                   ORBmatcher::SearchForTriangulation  (O3/src/ORBmatcher.cc:836-1058)
//   fuse_search_kernel                     ORBmatcher::Fuse, search half       (O3/src/ORBmatcher.cc:1060-1228)
//
// SearchByBoW walks the vocabulary nodes the two feature vectors share; inside a node it is a sequential
// greedy loop (a feature of side 2 taken by an earlier feature of side 1 is skipped), but a feature
// belongs to exactly one node, so nodes are independent: one warp per shared node, the side-1 features of
// the node in order, the side-2 features across the lanes, nearest / second-nearest by warp reduction.
//
// SearchForInitialization is sequential over the whole first frame (vMatchedDistance is live state).
// As in the projection matchers, the sequential outcome is reached as the fixed point of
//     choice[i1] = best candidate i2 with  min{ dist[j] : j < i1, choice[j] = i2 } > dist(i1, i2)
// iterated from "nothing matched" on one CTA; at a fixed point the rule holds for i1 = 0, 1, 2, ... in
// turn, which is the reference's loop.
#include "bow_kernels.cuh"
#include "sophus_f32.cuh"
#include "track_device.cuh"

namespace dvm {

// ------------------------------------------------------------------------------------- SearchByBoW
constexpr int kBowWarps = 4;
constexpr int kBowStage = 192;   // side-2 features of a node staged per warp beyond the first 32 (6 KB + 768 B per warp)

__device__ inline int find_node(const uint32_t* ids, int n, uint32_t key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ids[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < n && ids[lo] == key) ? lo : -1;
}

// key of a candidate: distance in the high bits, position in the node's list below -- the minimum is the
// reference's "first of the nearest" (strict '<' keeps the earlier of equal distances)
constexpr unsigned kPosBits = 20;
constexpr unsigned kNoKey = (256u << kPosBits);

__global__ void __launch_bounds__(kBowWarps * 32) bow_node_kernel(BowArgs g)
{
    const int lane = threadIdx.x & 31;
    const int ia = blockIdx.x * kBowWarps + (threadIdx.x >> 5);
    if (ia >= g.a.n_nodes) return;
    const int ib = find_node(g.b.node_id, g.b.n_nodes, g.a.node_id[ia]);
    if (ib < 0) return;
    const int a0 = g.a.node_start[ia], a1 = g.a.node_start[ia + 1];
    const int b0 = g.b.node_start[ib], nb = g.b.node_start[ib + 1] - b0;
    if (nb <= 0) return;
    // side 2 of the node: positions 0..31 in registers, 32..32+kBowStage in this warp's shared memory (descriptor
    // + feature index, -1 once taken or invalid), anything beyond through L1 -- the sequential loop over side 1
    // must not wait on global memory per candidate chunk (a 156-feature node took 178 us that way)
    __shared__ uint4 s_desc[kBowWarps][kBowStage][2];
    __shared__ int s_r2[kBowWarps][kBowStage];
    __shared__ float s_ang[kBowWarps][kBowStage];
    const int wslot = threadIdx.x >> 5;
    const int nstage = min(max(nb - 32, 0), kBowStage);
    for (int q = lane; q < nstage; q += 32) {
        int r2 = (int)g.b.feat_idx[b0 + 32 + q];
        if (g.kf_kf && g.b.valid && !g.b.valid[r2]) r2 = -1;
        s_r2[wslot][q] = r2;
        if (r2 >= 0) {
            s_ang[wslot][q] = g.check_ori ? g.b.angle[r2] : 0.f;
            const uint4* dp = reinterpret_cast<const uint4*>(g.b.desc + (size_t)r2 * 32);
            s_desc[wslot][q][0] = __ldg(dp); s_desc[wslot][q][1] = __ldg(dp + 1);
        }
    }
    __syncwarp();
    int r2_0 = -1;
    uint32_t d2_0[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    float ang2_0 = 0.f;
    if (lane < nb) {
        r2_0 = (int)g.b.feat_idx[b0 + lane];
        if (g.kf_kf && g.b.valid && !g.b.valid[r2_0]) r2_0 = -1;   // !pMP2 || pMP2->isBad()
        if (r2_0 >= 0) { load_desc(d2_0, g.b.desc + (size_t)r2_0 * 32); if (g.check_ori) ang2_0 = g.b.angle[r2_0]; }
    }
    bool taken_0 = false;
    // the side-1 features of the node are fetched 32 at a time (lane = feature: index, validity, descriptor) and
    // handed to the whole warp by shuffles, so the sequential loop below never waits on global memory
    int my_r1 = -1;
    float my_ang1 = 0.f;
    uint32_t my_d1[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int p = a0; p < a1; p++) {
        if (((p - a0) & 31) == 0) {
            my_r1 = -1;
            if (p + lane < a1) {
                const int r = (int)g.a.feat_idx[p + lane];
                if (!(g.a.valid && !g.a.valid[r])) {
                    my_r1 = r; load_desc(my_d1, g.a.desc + (size_t)r * 32);
                    if (g.check_ori) my_ang1 = g.a.angle[r];
                }
            }
        }
        const int src = (p - a0) & 31;
        const int r1 = __shfl_sync(0xffffffffu, my_r1, src);
        if (r1 < 0) continue;
        const float ang1 = __shfl_sync(0xffffffffu, my_ang1, src);
        uint32_t d1[8];
#pragma unroll
        for (int q = 0; q < 8; q++) d1[q] = __shfl_sync(0xffffffffu, my_d1[q], src);
        unsigned k1 = kNoKey, dist2 = 256;   // lane-local nearest key / second-nearest distance
        if (r2_0 >= 0 && !taken_0) {
            const unsigned dist = __popc(d1[0] ^ d2_0[0]) + __popc(d1[1] ^ d2_0[1]) + __popc(d1[2] ^ d2_0[2]) +
                                  __popc(d1[3] ^ d2_0[3]) + __popc(d1[4] ^ d2_0[4]) + __popc(d1[5] ^ d2_0[5]) +
                                  __popc(d1[6] ^ d2_0[6]) + __popc(d1[7] ^ d2_0[7]);
            if (dist < 256) k1 = (dist << kPosBits) | (unsigned)lane;
        }
        for (int q = lane; q < nstage; q += 32) {    // larger nodes: the staged part
            if (s_r2[wslot][q] < 0) continue;
            const uint4 v0 = s_desc[wslot][q][0], v1 = s_desc[wslot][q][1];
            const unsigned dist = __popc(d1[0] ^ v0.x) + __popc(d1[1] ^ v0.y) + __popc(d1[2] ^ v0.z) + __popc(d1[3] ^ v0.w) +
                                  __popc(d1[4] ^ v1.x) + __popc(d1[5] ^ v1.y) + __popc(d1[6] ^ v1.z) + __popc(d1[7] ^ v1.w);
            const unsigned key = (dist << kPosBits) | (unsigned)(32 + q);
            if (dist < (k1 >> kPosBits)) { dist2 = k1 >> kPosBits; k1 = key; }
            else if (dist < dist2) dist2 = dist;
        }
        for (int q = 32 + kBowStage + lane; q < nb; q += 32) {   // and whatever does not fit, through L1
            const int r2 = (int)g.b.feat_idx[b0 + q];
            if (g.match21[r2] >= 0) continue;
            if (g.kf_kf && g.b.valid && !g.b.valid[r2]) continue;
            const unsigned dist = (unsigned)hamming256(d1, g.b.desc + (size_t)r2 * 32);
            const unsigned key = (dist << kPosBits) | (unsigned)q;
            if (dist < (k1 >> kPosBits)) { dist2 = k1 >> kPosBits; k1 = key; }
            else if (dist < dist2) dist2 = dist;
        }
        const unsigned best = __reduce_min_sync(0xffffffffu, k1);
        const unsigned mine = (k1 == best) ? dist2 : min(k1 >> kPosBits, 256u);
        const unsigned bestDist2 = __reduce_min_sync(0xffffffffu, mine);
        const unsigned bestDist1 = best >> kPosBits;
        const bool near = g.kf_kf ? bestDist1 < (unsigned)kThLow : bestDist1 <= (unsigned)kThLow;
        if (near && (float)bestDist1 < __fmul_rn(g.nnratio, (float)bestDist2)) {
            const int q = (int)(best & ((1u << kPosBits) - 1));
            if (q < 32) { if (lane == q) taken_0 = true; }
            const int r2 = __shfl_sync(0xffffffffu, q < 32 ? r2_0 : 0, q & 31);
            const float a2 = __shfl_sync(0xffffffffu, ang2_0, q & 31);
            if (lane == 0) {
                // partner index and angle from registers / shared memory: no global load on the sequential path
                const bool staged = q >= 32 && q < 32 + kBowStage;
                const int rr = q < 32 ? r2 : staged ? s_r2[wslot][q - 32] : (int)g.b.feat_idx[b0 + q];
                const float ang2 = q < 32 ? a2 : staged ? s_ang[wslot][q - 32] : (g.check_ori ? g.b.angle[rr] : 0.f);
                if (staged) s_r2[wslot][q - 32] = -1;   // taken
                g.match21[rr] = r1;
                g.match12[r1] = rr;
                atomicAdd(&g.counters[0], 1);
                if (g.check_ori) atomicAdd(&g.histo[rot_bin(ang1, ang2)], 1);
            }
            __syncwarp(); // the write to match21 is visible to the lanes that read it for the next feature
        }
    }
}

// rotation-consistency check over all accepted pairs + the match count
__global__ void __launch_bounds__(1024) bow_finish_kernel(BowArgs g)
{
    __shared__ int s_ind[3], s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) {
        int i1 = -1, i2 = -1, i3 = -1;
        if (g.check_ori) three_maxima(g.histo, kHistoLength, i1, i2, i3);
        s_ind[0] = i1; s_ind[1] = i2; s_ind[2] = i3;
        s_bad = 0;
    }
    __syncthreads();
    if (g.check_ori) {
        int bad = 0;
        for (int r1 = tid; r1 < g.a.n; r1 += 1024) {
            const int r2 = g.match12[r1];
            if (r2 < 0) continue;
            const int bin = rot_bin(g.a.angle[r1], g.b.angle[r2]);
            if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { g.match12[r1] = -1; g.match21[r2] = -1; bad++; }
        }
        if (bad) atomicAdd(&s_bad, bad);
    }
    __syncthreads();
    if (tid == 0) g.counters[1] = g.counters[0] - s_bad;
}

void launch_bow_match(const BowArgs& g, cudaStream_t stream)
{
    if (g.a.n_nodes > 0 && g.b.n_nodes > 0)
        DVM_LAUNCH(bow_node_kernel, div_up(g.a.n_nodes, kBowWarps), kBowWarps * 32, 0, stream, g);
    DVM_LAUNCH(bow_finish_kernel, 1, 1024, 0, stream, g);
}

// --------------------------------------------------------------------------- SearchForInitialization
constexpr int kInitThreads = 1024;
constexpr int kInitCache = 8;     // nearest candidates kept per query, sorted by (distance, traversal order)
constexpr int kInitWalkThreads = 128;

__device__ inline unsigned long long init_pack(int dist, int ord, int idx)
{
    return ((unsigned long long)(unsigned)dist << 48) | ((unsigned long long)(unsigned)min(ord, 0xffffff) << 24) | (unsigned)idx;
}

// phase 1 (many CTAs): one warp per keypoint of F1 walks its window in F2 ONCE and keeps the kInitCache nearest
// candidates; the fixed-point rounds below then only consult this cache (a query whose cached candidates are all
// skipped while it had more falls back to a full walk)
__global__ void __launch_bounds__(kInitWalkThreads) init_walk_kernel(FrameDev f2, InitMatchArgs a)
{
    const int lane = threadIdx.x & 31;
    const int i1 = (blockIdx.x * kInitWalkThreads + threadIdx.x) >> 5;
    if (i1 >= a.n1) return;
    const int level1 = a.kps1[i1].octave;
    if (level1 > 0) { if (lane == 0) a.ncand[i1] = -1; return; }     // level1 > 0: continue
    const FrameLook fl = look_global(f2);
    uint32_t d1[8];
    load_desc(d1, a.desc1 + (size_t)i1 * 32);
    unsigned long long top[kInitCache];
#pragma unroll
    for (int p = 0; p < kInitCache; p++) top[p] = ~0ull;
    int nc = 0;
    walk_area_warp(fl, a.prev_matched[2 * i1], a.prev_matched[2 * i1 + 1], (float)a.window, level1, level1, lane,
                   [&](int i2, int, int ord) {
                       unsigned long long v = init_pack(hamming256(d1, fl.desc + (size_t)i2 * 32), ord, i2);
#pragma unroll
                       for (int p = 0; p < kInitCache; p++)
                           if (v < top[p]) { const unsigned long long t = top[p]; top[p] = v; v = t; }
                       nc++;
                   });
    nc = __reduce_add_sync(0xffffffffu, nc);
    unsigned long long mine = ~0ull;
#pragma unroll
    for (int p = 0; p < kInitCache; p++) {
        const unsigned long long m = warp_min_u64(top[0]);
        if (lane == p) mine = m;
        if (top[0] == m && m != ~0ull) {
#pragma unroll
            for (int q = 0; q + 1 < kInitCache; q++) top[q] = top[q + 1];
            top[kInitCache - 1] = ~0ull;
        }
    }
    if (lane < kInitCache) a.cache[(size_t)i1 * kInitCache + lane] = mine;
    if (lane == 0) a.ncand[i1] = nc;
}

__global__ void __launch_bounds__(kInitThreads, 1) init_match_kernel(FrameDev f2, InitMatchArgs a)
{
    __shared__ int histo[kHistoLength];
    __shared__ int s_ind[3], s_n;
    const int tid = threadIdx.x;
    const int n2 = min(*f2.n, f2.cap);
    const FrameLook fl = look_global(f2);
    int* choice_prev = a.choice_a;  int* choice_next = a.choice_b;
    int* cdist_prev = a.cdist_a;    int* cdist_next = a.cdist_b;
    for (int i = tid; i < a.n1; i += kInitThreads) { choice_prev[i] = -1; cdist_prev[i] = 0; }
    __syncthreads();
    int rounds = 0;
    while (true) {
        // claim lists of the previous round: head[i2] -> i1 -> next[i1] -> ...
        for (int k = tid; k < n2; k += kInitThreads) a.head[k] = -1;
        __syncthreads();
        for (int i = tid; i < a.n1; i += kInitThreads) {
            const int k = choice_prev[i];
            if (k >= 0) a.next[i] = atomicExch(&a.head[k], i);
        }
        __syncthreads();
        int changed = 0;
        for (int i1 = tid; i1 < a.n1; i1 += kInitThreads) {
            int pick = -1, pickDist = 0;
            const int nc = a.ncand[i1];
            if (nc > 0) {
                // vMatchedDistance[i2] as the sequential loop sees it at i1: a candidate is skipped when an earlier
                // query holds it at a distance <= ours
                auto skipped = [&](int i2, int dist) {
                    for (int j = a.head[i2]; j >= 0; j = a.next[j])
                        if (j < i1 && cdist_prev[j] <= dist) return true;
                    return false;
                };
                int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1, found = 0;
                const unsigned long long* e = a.cache + (size_t)i1 * kInitCache;
                for (int p = 0; p < kInitCache && p < nc && found < 2; p++) {
                    const unsigned long long ev = e[p];
                    const int dist = (int)(ev >> 48), i2 = (int)(ev & 0xffffffull);
                    if (skipped(i2, dist)) continue;
                    if (found == 0) { bestDist = dist; bestIdx2 = i2; } else bestDist2 = dist;
                    found++;
                }
                if (found < 2 && nc > kInitCache) {   // cache exhausted: full walk against the current claims
                    uint32_t d1[8];
                    load_desc(d1, a.desc1 + (size_t)i1 * 32);
                    bestDist = INT_MAX; bestDist2 = INT_MAX; bestIdx2 = -1;
                    const int level1 = a.kps1[i1].octave;
                    walk_area(fl, a.prev_matched[2 * i1], a.prev_matched[2 * i1 + 1], (float)a.window, level1, level1,
                              [&](int i2, int) {
                                  const int dist = hamming256(d1, fl.desc + (size_t)i2 * 32);
                                  if (skipped(i2, dist)) return;
                                  if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
                                  else if (dist < bestDist2) bestDist2 = dist;
                              });
                }
                if (bestDist <= kThLow && (float)bestDist < __fmul_rn((float)bestDist2, a.nnratio)) { pick = bestIdx2; pickDist = bestDist; }
            }
            choice_next[i1] = pick; cdist_next[i1] = pickDist;
            if (pick != choice_prev[i1]) changed = 1;
        }
        rounds++;
        const int any = __syncthreads_or(changed);
        int* t = choice_prev; choice_prev = choice_next; choice_next = t;
        t = cdist_prev; cdist_prev = cdist_next; cdist_next = t;
        if (!any) break;
    }
    // vnMatches21[i2] = the last query that took i2; every accepted query left one histogram entry
    if (tid < kHistoLength) histo[tid] = 0;
    if (tid == 0) s_n = 0;
    for (int k = tid; k < n2; k += kInitThreads) a.head[k] = -1;
    __syncthreads();
    for (int i = tid; i < a.n1; i += kInitThreads) {
        const int k = choice_prev[i];
        if (k < 0) continue;
        atomicMax(&a.head[k], i);
        if (a.check_ori) atomicAdd(&histo[rot_bin(a.kps1[i].angle, f2.kps[k].angle)], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int i1 = -1, i2 = -1, i3 = -1;
        if (a.check_ori) three_maxima(histo, kHistoLength, i1, i2, i3);
        s_ind[0] = i1; s_ind[1] = i2; s_ind[2] = i3;
    }
    __syncthreads();
    int cnt = 0;
    for (int i = tid; i < a.n1; i += kInitThreads) {
        int k = choice_prev[i];
        if (k >= 0 && a.head[k] != i) k = -1;   // displaced by a later query
        if (k >= 0 && a.check_ori) {
            const int bin = rot_bin(a.kps1[i].angle, f2.kps[k].angle);
            if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) k = -1;
        }
        a.matches12[i] = k;
        if (k >= 0) {
            cnt++;
            a.prev_matched[2 * i] = f2.kps[k].x;
            a.prev_matched[2 * i + 1] = f2.kps[k].y;
        }
    }
    if (cnt) atomicAdd(&s_n, cnt);
    __syncthreads();
    if (tid == 0) { a.result[0] = s_n; a.result[1] = rounds; }
}

void launch_init_match(const FrameDev& f2, const InitMatchArgs& a, cudaStream_t stream)
{
    if (a.n1 > 0) DVM_LAUNCH(init_walk_kernel, div_up(a.n1, kInitWalkThreads / 32), kInitWalkThreads, 0, stream, f2, a);
    DVM_LAUNCH(init_match_kernel, 1, kInitThreads, 0, stream, f2, a);
}

// --------------------------------------------------------------------------- SearchForTriangulation
// This fork never marks a feature of keyframe 2 as taken (vbMatched2 stays false), so every feature of
// keyframe 1 is independent: a warp takes kTriSlots consecutive slots of keyframe 1's feature vector (whatever
// node they fall in -- a 156-feature node no longer serialises on one warp), the lanes share the keyframe-2
// features of the slot's node.  The reference keeps a candidate when dist <= min(TH_LOW, bestDist) and the
// geometric gates pass, i.e. the LAST of the nearest valid candidates wins: the key is
// distance << 20 | (0xfffff - position) and the warp takes its minimum.
constexpr int kTriSlots = 16;   // keyframe-1 feature-vector slots per warp

__global__ void __launch_bounds__(kBowWarps * 32) triangulation_kernel(TriArgs g)
{
    const int lane = threadIdx.x & 31;
    const int total = g.a.n_nodes ? g.a.node_start[g.a.n_nodes] : 0;
    const int p0 = (blockIdx.x * kBowWarps + (threadIdx.x >> 5)) * kTriSlots;
    if (p0 >= total) return;
    const int p1 = min(p0 + kTriSlots, total);
    // the slots' features, lane = slot: index, keypoint, descriptor (features with a map point drop out here)
    int my_i1 = -1;
    float my_x = 0.f, my_y = 0.f, my_ang = 0.f;
    uint32_t my_d1[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    if (p0 + lane < p1) {
        const int i = (int)g.a.feat_idx[p0 + lane];
        if (!g.a.valid[i]) {
            my_i1 = i; my_x = g.kps1[i].x; my_y = g.kps1[i].y; my_ang = g.kps1[i].angle;
            load_desc(my_d1, g.a.desc + (size_t)i * 32);
        }
    }
    // node of slot p0: the last node whose start is <= p0 (empty nodes are skipped by the upper bound)
    int ia;
    {
        int lo = 0, hi = g.a.n_nodes;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (g.a.node_start[mid + 1] <= p0) lo = mid + 1; else hi = mid; }
        ia = lo;
    }
    int ib = -2, b0 = 0, nb = 0;   // -2: not looked up yet for the current node
    for (int p = p0; p < p1; p++) {
        while (p >= g.a.node_start[ia + 1]) { ia++; ib = -2; }
        if (ib == -2) {
            ib = find_node(g.b.node_id, g.b.n_nodes, g.a.node_id[ia]);
            if (ib >= 0) { b0 = g.b.node_start[ib]; nb = g.b.node_start[ib + 1] - b0; }
        }
        const int src = p - p0;
        const int i1 = __shfl_sync(0xffffffffu, my_i1, src);
        if (i1 < 0 || ib < 0) continue;
        dvm_keypoint kp1;
        kp1.x = __shfl_sync(0xffffffffu, my_x, src); kp1.y = __shfl_sync(0xffffffffu, my_y, src);
        kp1.angle = __shfl_sync(0xffffffffu, my_ang, src);
        uint32_t d1[8];
#pragma unroll
        for (int q = 0; q < 8; q++) d1[q] = __shfl_sync(0xffffffffu, my_d1[q], src);
        // epipolar line of kp1 in image 2: l = x1' F12
        const float la = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, g.F12[0]), __fmul_rn(kp1.y, g.F12[3])), g.F12[6]);
        const float lb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, g.F12[1]), __fmul_rn(kp1.y, g.F12[4])), g.F12[7]);
        const float lc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, g.F12[2]), __fmul_rn(kp1.y, g.F12[5])), g.F12[8]);
        const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
        unsigned best = ~0u;
        for (int q = lane; q < nb; q += 32) {
            const int i2 = (int)g.b.feat_idx[b0 + q];
            if (g.b.valid[i2]) continue;
            const int dist = hamming256(d1, g.b.desc + (size_t)i2 * 32);
            if (dist > kThLow) continue;
            const float x2 = g.kps2[i2].x, y2 = g.kps2[i2].y;
            const int oct2 = g.kps2[i2].octave;
            const float ex = __fsub_rn(g.ep[0], x2), ey = __fsub_rn(g.ep[1], y2);
            if (__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)) < __fmul_rn(100.f, g.scale2[oct2])) continue;
            if (!g.coarse) {
                if (den == 0.f) continue;
                const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, x2), __fmul_rn(lb, y2)), lc);
                const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                if (!((double)dsqr < 3.84 * (double)g.sigma2_2[oct2])) continue;
            }
            best = min(best, ((unsigned)dist << kPosBits) | (0xfffffu - (unsigned)q));
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (best != ~0u && lane == 0) {
            const int q = (int)(0xfffffu - (best & 0xfffffu));
            const int i2 = (int)g.b.feat_idx[b0 + q];
            g.matches12[i1] = i2;
            atomicAdd(&g.counters[0], 1);
            if (g.check_ori) atomicAdd(&g.histo[rot_bin(kp1.angle, g.kps2[i2].angle)], 1);
        }
    }
}

__global__ void __launch_bounds__(1024) triangulation_finish_kernel(TriArgs g)
{
    __shared__ int s_ind[3], s_bad;
    const int tid = threadIdx.x;
    if (tid == 0) {
        int i1 = -1, i2 = -1, i3 = -1;
        if (g.check_ori) three_maxima(g.histo, kHistoLength, i1, i2, i3);
        s_ind[0] = i1; s_ind[1] = i2; s_ind[2] = i3;
        s_bad = 0;
    }
    __syncthreads();
    if (g.check_ori) {
        int bad = 0;
        for (int i1 = tid; i1 < g.a.n; i1 += 1024) {
            const int i2 = g.matches12[i1];
            if (i2 < 0) continue;
            const int bin = rot_bin(g.kps1[i1].angle, g.kps2[i2].angle);
            if (bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) { g.matches12[i1] = -1; bad++; }
        }
        if (bad) atomicAdd(&s_bad, bad);
    }
    __syncthreads();
    if (tid == 0) g.counters[1] = g.counters[0] - s_bad;
}

void launch_triangulation_match(const TriArgs& g, cudaStream_t stream)
{
    if (g.a.n_nodes > 0 && g.b.n_nodes > 0 && g.a.n > 0)   // every feature sits in at most one node: <= a.n slots
        DVM_LAUNCH(triangulation_kernel, div_up(div_up(g.a.n, kTriSlots), kBowWarps), kBowWarps * 32, 0, stream, g);
    DVM_LAUNCH(triangulation_finish_kernel, 1, 1024, 0, stream, g);
}

// ---------------------------------------------------------------------------------------- Fuse (search)
// One warp per map point: the gates of O3/src/ORBmatcher.cc:1100-1150 (depth, image bounds, distance range,
// viewing angle, predicted level), then the window of KeyFrame::GetFeaturesInArea shared by the lanes.
// Strict '<' on the distance: the FIRST of the nearest keypoints in traversal order wins.
constexpr int kFuseWarps = 4;

__global__ void __launch_bounds__(kFuseWarps * 32) fuse_search_kernel(FrameDev kf, FuseArgs a)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * kFuseWarps + (threadIdx.x >> 5);
    if (i >= a.m) return;
    int bestIdx = -1, bestDist = 256, lvl = -1;
    float u = 0.f, v = 0.f;
    do {
        if (a.skip && a.skip[i]) break;
        const float P[3] = { a.xw[3 * i], a.xw[3 * i + 1], a.xw[3 * i + 2] };
        float pc[3];
        so::se3_apply(a.q, a.t, P, pc);                                         // p3Dc = Tcw * p3Dw
        if (a.chain_sim3) { float p2[3]; so::sim3_apply(a.sq, a.st, pc, p2); pc[0] = p2[0]; pc[1] = p2[1]; pc[2] = p2[2]; }
        if (pc[2] < 0.0f) break;
        if (a.proj_invz) {
            const float invz = (float)(1.0 / (double)pc[2]);
            u = __fadd_rn(__fmul_rn(a.K[0], __fmul_rn(pc[0], invz)), a.K[2]);
            v = __fadd_rn(__fmul_rn(a.K[1], __fmul_rn(pc[1], invz)), a.K[3]);
        } else {                                                                // Pinhole::project
            u = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[0], pc[0]), pc[2]), a.K[2]);
            v = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[1], pc[1]), pc[2]), a.K[3]);
        }
        if (!(u >= kf.minX && u < kf.maxX && v >= kf.minY && v < kf.maxY)) break;   // KeyFrame::IsInImage
        const float maxDistance = __fmul_rn(1.2f, a.max_dist[i]), minDistance = __fmul_rn(0.8f, a.min_dist[i]);
        float dist3D;
        if (a.dist_camera) dist3D = so::norm3(pc);
        else {
            // Ow = pKF->GetCameraCenter() = translation of Tcw.inverse() (KeyFrame.cc:224-257)
            float qi[4], Ow[3];
            so::se3_inverse(a.q, a.t, qi, Ow);
            const float PO[3] = { __fsub_rn(P[0], Ow[0]), __fsub_rn(P[1], Ow[1]), __fsub_rn(P[2], Ow[2]) };
            dist3D = so::norm3(PO);
            if (dist3D < minDistance || dist3D > maxDistance) break;
            if (a.check_normal) {
                const float Pn[3] = { a.normal[3 * i], a.normal[3 * i + 1], a.normal[3 * i + 2] };
                if ((double)so::dot3(PO, Pn) < 0.5 * (double)dist3D) break;
            }
        }
        if (dist3D < minDistance || dist3D > maxDistance) break;
        lvl = so::predict_scale(a.max_dist[i], dist3D, a.logScale, a.nlevels);
        if (a.gate_only) break;
        const float radius = __fmul_rn(a.th, kf.scale[lvl]);
        uint32_t d[8];
        load_desc(d, a.mp_desc + (size_t)i * 32);
        const FrameLook fl = look_global(kf);
        unsigned long long best = ~0ull;   // distance << 48 | traversal position << 24 | keypoint
        const int level = lvl;
        walk_area_warp(fl, u, v, radius, -1, -1, lane, [&](int idx, int oct, int ord) {
            if (oct < level - 1 || oct > level) return;
            if (a.check_chi) {
                const float ex = __fsub_rn(u, fl.x(idx)), ey = __fsub_rn(v, fl.y(idx));
                const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                if ((double)__fmul_rn(e2, a.inv_sigma2[oct]) > 5.99) return;
            }
            const unsigned long long dist = (unsigned long long)hamming256(d, fl.desc + (size_t)idx * 32);
            best = min(best, (dist << 48) | ((unsigned long long)min(ord, 0xffffff) << 24) | (unsigned long long)idx);
        });
        best = warp_min_u64(best);
        if (best == ~0ull) break;
        bestDist = (int)(best >> 48);
        bestIdx = (int)(best & 0xffffffull);
    } while (false);
    if (lane == 0) {
        if (a.gate_only) { a.gate_u[i] = u; a.gate_v[i] = v; a.gate_level[i] = lvl; return; }
        const bool ok = bestIdx >= 0 && bestDist <= a.accept_th;
        a.best_idx[i] = ok ? bestIdx : -1;
        a.best_dist[i] = ok ? bestDist : 256;
    }
}

void launch_fuse_search(const FrameDev& kf, const FuseArgs& a, cudaStream_t stream)
{
    if (a.m > 0) DVM_LAUNCH(fuse_search_kernel, div_up(a.m, kFuseWarps), kFuseWarps * 32, 0, stream, kf, a);
}

} // namespace dvm
