// track_pipeline.cu -- Frame::isInFrustum on the GPU and the device-resident per-frame tracker that
// chains ExtractORB -> Frame -> SearchByProjection(last frame) -> PoseOptimization -> isInFrustum +
// SearchByProjection(local map) -> PoseOptimization without host round trips
// (call order of Tracking::TrackWithMotionModel / TrackLocalMap / SearchLocalPoints,
//  O3/src/Tracking.cc:2584-2666, 2668-2768, 3041-3106).
#include "sophus_f32.cuh"
#include "glibc_logf.h"
#include "track_internal.cuh"
#include <cmath>
#include <vector>

using namespace dvm;

namespace {

// The constant-velocity prior of Tracking::TrackWithMotionModel in the reference's own float32 Sophus arithmetic:
// mVelocity = mCurrentFrame.GetPose() * mLastFrame.GetPose().inverse() after a tracked frame (Tracking.cc:1990-1991),
// then mCurrentFrame.SetPose(mVelocity * mLastFrame.GetPose()) for the next one (:2598).  `last` / `prev` are the two
// most recent poses (qx,qy,qz,qw,tx,ty,tz) as their SE3f would hold them.
// Also the per-frame reset of the "seen in this frame" marks and the counters (one launch instead of three).
__global__ void __launch_bounds__(1024) begin_frame_kernel(const float* last, const float* prev, float* prior, int have_prior,
                                                           uint8_t* seen, int map_n, int* cnt)
{
    pdl_trigger(); pdl_wait();
    for (int i = threadIdx.x; i < (map_n + 3) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(seen)[i] = 0u;
    if (threadIdx.x < 8) cnt[threadIdx.x] = 0;
    if (threadIdx.x != 0 || have_prior) return;
    float ql[4], tl[3], qp[4], tp[3], qpi[4], tpi[3], qv[4], tv[3], qo[4], to[3];
    for (int i = 0; i < 4; i++) { ql[i] = last[i]; qp[i] = prev[i]; }
    for (int i = 0; i < 3; i++) { tl[i] = last[4 + i]; tp[i] = prev[4 + i]; }
    so::se3_inverse(qp, tp, qpi, tpi);            // LastTwc
    so::se3_mul(ql, tl, qpi, tpi, qv, tv);        // mVelocity
    so::se3_mul(qv, tv, ql, tl, qo, to);          // mVelocity * mLastFrame.GetPose()
    for (int i = 0; i < 4; i++) prior[i] = qo[i];
    for (int i = 0; i < 3; i++) prior[4 + i] = to[i];
}

} // namespace

// ------------------------------------------------------------------------------------------------
// Two streams per agent: `es` (the extractor handle's stream) runs the Frame constructor (ExtractORB
// straight into the frame's buffers + AssignFeaturesToGrid), `stream` runs the dependency chain of the
// tracked frame.  Frame i+1 does not depend on frame i's pose until the chain starts, so its
// construction overlaps frame i's chain -- and so does frame i+2's: two extractor handles (the caller's and a
// clone) alternate on their own streams, so the extraction chain (about as long as the tracking chain) never
// bounds the frame rate.  Frames live in a ring of four (last, current, two being built); ev_extracted[slot] /
// ev_done[slot] order the hand-over between the streams.
constexpr int kRing = 4;
constexpr int kExtractors = 2;
constexpr int kMaxPending = 2;   // frames extracted (or being extracted) ahead of the chain
struct dvm_tracker {
    int device = 0;
    cudaStream_t stream = nullptr;  // tracking chain
    cudaStream_t es[kExtractors] = { nullptr, nullptr };   // extraction (owned by the extractor handles)
    cudaEvent_t ev_extracted[kRing] = { nullptr, nullptr, nullptr, nullptr };
    cudaEvent_t ev_done[kRing] = { nullptr, nullptr, nullptr, nullptr };
    long long n_extracted = 0;      // frames whose construction has been enqueued
    long long n_tracked = 0;        // frames whose chain has been enqueued (frame 0 = bootstrap)
    dvm_orb* orbs[kExtractors] = { nullptr, nullptr };   // [0] the caller's, [1] an owned clone
    dvm_frame* frames[kRing] = { nullptr, nullptr, nullptr, nullptr };
    int cap = 0, map_n = 0, nlevels = 0;
    float K[4], bounds[4], logScale = 0;
    std::vector<float> inv_sigma2;
    // map snapshot
    float* d_xw = nullptr; uint8_t* d_desc = nullptr; float* d_normal = nullptr; float* d_mind = nullptr; float* d_maxd = nullptr;
    // per-frame association (same ring as the frames)
    int* d_mp[kRing] = { nullptr, nullptr, nullptr, nullptr };
    uint8_t* d_outl[kRing] = { nullptr, nullptr, nullptr, nullptr };
    // scratch
    uint8_t* d_seen = nullptr;
    int* d_cur_mp = nullptr; int* d_cur_mp2 = nullptr;
    int* d_cnt = nullptr;       // [8]: 1 nmatches last, 3 nmatches map
    float* d_pose = nullptr;    // [7] working pose (prior in, optimised out)
    float* d_pose_last = nullptr; float* d_pose_prev = nullptr;
    int* d_res1 = nullptr; int* d_res2 = nullptr;
    uint8_t* d_result = nullptr; // pose[7] float | counts[4] int
    uint8_t* h_result = nullptr; // pinned: [64,92) prior staging
    uint8_t* h_ring = nullptr;   // pinned + mapped, [kRing][64]: pose[7] float | counts[4] int of frame i in slot i % kRing, written by
    uint8_t* d_ring = nullptr;   // the chain's last kernel straight into host memory (its device alias): no read-back copy
    uint8_t* d_img[kExtractors] = { nullptr, nullptr }; size_t img_cap[kExtractors] = { 0, 0 };   // H2D staging
    // per-segment device time of the chain (dvm_tracker_set_profiling): events between the operators of a tracked frame
    bool profiling = false;
    cudaEvent_t pev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
    double prof_ms[5] = { 0, 0, 0, 0, 0 };
    long long prof_frames = 0;
};

static void tracker_free(dvm_tracker* t)
{
    if (!t) return;
    cudaSetDevice(t->device);
    if (t->stream) cudaStreamSynchronize(t->stream);
    for (auto e : t->es) if (e) cudaStreamSynchronize(e);
    for (auto f : t->frames) if (f) dvm_frame_destroy(f);
    if (t->orbs[1]) dvm_orb_destroy(t->orbs[1]);
    void* ptrs[] = { t->d_xw, t->d_desc, t->d_normal, t->d_mind, t->d_maxd, t->d_mp[0], t->d_mp[1], t->d_mp[2], t->d_mp[3],
                     t->d_outl[0], t->d_outl[1], t->d_outl[2], t->d_outl[3], t->d_seen, t->d_cur_mp, t->d_cur_mp2, t->d_cnt,
                     t->d_pose, t->d_pose_last, t->d_pose_prev, t->d_res1, t->d_res2, t->d_result, t->d_img[0], t->d_img[1] };
    for (void* p : ptrs) cudaFree(p);
    if (t->h_result) cudaFreeHost(t->h_result);
    if (t->h_ring) cudaFreeHost(t->h_ring);
    for (auto e : t->ev_extracted) if (e) cudaEventDestroy(e);
    for (auto e : t->ev_done) if (e) cudaEventDestroy(e);
    for (auto e : t->pev) if (e) cudaEventDestroy(e);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

static void fill_frustum_args(FrustumArgs& a, const float* pose_dev, const float* K, const float* bounds, int nlevels,
                              float logScale, float cosLimit)
{
    memset(&a, 0, sizeof(a));
    a.pose = pose_dev;
    for (int i = 0; i < 4; i++) { a.K[i] = K[i]; a.bounds[i] = bounds[i]; }
    a.nlevels = nlevels; a.logScale = logScale; a.cosLimit = cosLimit;
}

extern "C" {

int dvm_frame_is_in_frustum(dvm_frame* f, const float* pose_q, const float* pose_t, const float* K, int m, const float* xw,
                            const float* normal, const float* min_dist, const float* max_dist, const uint8_t* skip,
                            float viewing_cos_limit, uint8_t* in_view, float* proj_x, float* proj_y, int32_t* level,
                            float* view_cos)
{
    DVM_REQUIRE(f && pose_q && pose_t && K, "null argument");
    DVM_REQUIRE(m >= 0, "negative count");
    if (m == 0) return DVM_OK;
    DVM_REQUIRE(xw && normal && min_dist && max_dist && in_view && proj_x && proj_y && level && view_cos, "null arrays");
    DVM_CUDA(cudaSetDevice(f->device));
    const size_t n = (size_t)m;
    // layout: pose | xw | normal | min | max | skip | outputs(in_view, px, py, level, cos)
    size_t off = 0;
    auto take = [&](size_t b) { off = (off + 255) & ~(size_t)255; size_t o = off; off += b; return o; };
    const size_t o_pose = take(28), o_xw = take(n * 12), o_nrm = take(n * 12), o_min = take(n * 4), o_max = take(n * 4),
                 o_skip = take(n);
    const size_t in_bytes = off;
    const size_t o_vis = take(n), o_px = take(n * 4), o_py = take(n * 4), o_lv = take(n * 4), o_cos = take(n * 4);
    int rc = dvm_frame_ensure_bytes(f, off + 256, off + 256);
    if (rc != DVM_OK) return rc;
    uint8_t* hb = f->h_in;
    float pose[7] = { pose_q[0], pose_q[1], pose_q[2], pose_q[3], pose_t[0], pose_t[1], pose_t[2] };
    memcpy(hb + o_pose, pose, 28);
    memcpy(hb + o_xw, xw, n * 12); memcpy(hb + o_nrm, normal, n * 12);
    memcpy(hb + o_min, min_dist, n * 4); memcpy(hb + o_max, max_dist, n * 4);
    if (skip) memcpy(hb + o_skip, skip, n); else memset(hb + o_skip, 0, n);
    DVM_CUDA(cudaMemcpyAsync(f->d_in, hb, in_bytes, cudaMemcpyHostToDevice, f->stream));
    FrustumArgs a;
    const float bounds[4] = { f->dev.minX, f->dev.minY, f->dev.maxX, f->dev.maxY };
    // mfLogScaleFactor = log(mfScaleFactor), O3/src/Frame.cc:401 (scale[1] is the per-level factor)
    const float logScale = dvm_glibc_logf(f->dev.nlevels > 1 ? f->dev.scale[1] : 1.2f);
    fill_frustum_args(a, (const float*)(f->d_in + o_pose), K, bounds, f->dev.nlevels, logScale, viewing_cos_limit);
    a.m = m;
    a.xw = (const float*)(f->d_in + o_xw); a.normal = (const float*)(f->d_in + o_nrm);
    a.min_dist = (const float*)(f->d_in + o_min); a.max_dist = (const float*)(f->d_in + o_max);
    a.skip = f->d_in + o_skip;
    a.in_view = f->d_in + o_vis; a.px = (float*)(f->d_in + o_px); a.py = (float*)(f->d_in + o_py);
    a.level = (int*)(f->d_in + o_lv); a.view_cos = (float*)(f->d_in + o_cos);
    launch_frustum(a, f->stream);
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaMemcpyAsync(f->h_out, f->d_in + o_vis, off - o_vis, cudaMemcpyDeviceToHost, f->stream));
    DVM_CUDA(cudaStreamSynchronize(f->stream));
    const uint8_t* ho = f->h_out;
    memcpy(in_view, ho, n);
    memcpy(proj_x, ho + (o_px - o_vis), n * 4); memcpy(proj_y, ho + (o_py - o_vis), n * 4);
    memcpy(level, ho + (o_lv - o_vis), n * 4); memcpy(view_cos, ho + (o_cos - o_vis), n * 4);
    return DVM_OK;
}

int dvm_tracker_create(dvm_tracker** out, dvm_orb* orb, const float* K, const float* bounds, int map_n, const float* map_xw,
                       const uint8_t* map_desc, const float* map_normal, const float* map_min_dist, const float* map_max_dist)
{
    DVM_REQUIRE(out != nullptr, "null output handle");
    *out = nullptr;
    DVM_REQUIRE(orb && K && bounds && map_n > 0 && map_xw && map_desc && map_normal && map_min_dist && map_max_dist, "null argument");
    dvm_tracker* t = new dvm_tracker;
    t->orbs[0] = orb;
    {
        const int rc = dvm_orb_clone(orb, &t->orbs[1]);
        if (rc != DVM_OK) { delete t; return rc; }
    }
    for (int e = 0; e < kExtractors; e++) t->es[e] = (cudaStream_t)dvm_orb_stream(t->orbs[e]);
    t->cap = dvm_orb_max_keypoints(orb);
    t->map_n = map_n;
    float sc[16], is2[16];
    dvm_orb_tables(orb, &t->nlevels, sc, nullptr, nullptr, is2, nullptr);
    t->inv_sigma2.assign(is2, is2 + t->nlevels);
    t->logScale = dvm_glibc_logf(t->nlevels > 1 ? sc[1] : 1.2f);
    for (int i = 0; i < 4; i++) { t->K[i] = K[i]; t->bounds[i] = bounds[i]; }
    cudaGetDevice(&t->device);
#define DVM_TCREATE(call)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            set_error("%s failed in dvm_tracker_create: %s", #call, cudaGetErrorString(e__));  \
            tracker_free(t);                                                                   \
            return DVM_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)
    {   // the chain is the critical path: its few small CTAs go ahead of the extractor's wide grids
        int lo = 0, hi = 0;
        DVM_TCREATE(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        DVM_TCREATE(cudaStreamCreateWithPriority(&t->stream, cudaStreamNonBlocking, hi));
    }
    for (int i = 0; i < kRing; i++) {
        DVM_TCREATE(cudaEventCreateWithFlags(&t->ev_extracted[i], cudaEventDisableTiming));
        DVM_TCREATE(cudaEventCreateWithFlags(&t->ev_done[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < kRing; i++) {
        int rc = dvm_frame_create(&t->frames[i], t->device, t->stream, t->cap, t->nlevels, sc, is2);
        if (rc != DVM_OK) { tracker_free(t); return rc; }
        rc = dvm_frame_ensure_query_cap(t->frames[i], std::max(map_n, t->cap));
        if (rc != DVM_OK) { tracker_free(t); return rc; }
    }
    const size_t M = (size_t)map_n, C = (size_t)t->cap;
    DVM_TCREATE(cudaMalloc(&t->d_xw, M * 12)); DVM_TCREATE(cudaMalloc(&t->d_desc, M * 32));
    DVM_TCREATE(cudaMalloc(&t->d_normal, M * 12)); DVM_TCREATE(cudaMalloc(&t->d_mind, M * 4)); DVM_TCREATE(cudaMalloc(&t->d_maxd, M * 4));
    DVM_TCREATE(cudaMemcpy(t->d_xw, map_xw, M * 12, cudaMemcpyHostToDevice));
    DVM_TCREATE(cudaMemcpy(t->d_desc, map_desc, M * 32, cudaMemcpyHostToDevice));
    DVM_TCREATE(cudaMemcpy(t->d_normal, map_normal, M * 12, cudaMemcpyHostToDevice));
    DVM_TCREATE(cudaMemcpy(t->d_mind, map_min_dist, M * 4, cudaMemcpyHostToDevice));
    DVM_TCREATE(cudaMemcpy(t->d_maxd, map_max_dist, M * 4, cudaMemcpyHostToDevice));
    for (int i = 0; i < kRing; i++) {
        DVM_TCREATE(cudaMalloc(&t->d_mp[i], C * 4)); DVM_TCREATE(cudaMemset(t->d_mp[i], 0xff, C * 4));
        DVM_TCREATE(cudaMalloc(&t->d_outl[i], C)); DVM_TCREATE(cudaMemset(t->d_outl[i], 0, C));
    }
    DVM_TCREATE(cudaMalloc(&t->d_seen, M + 4));
    DVM_TCREATE(cudaMalloc(&t->d_cur_mp, C * 4)); DVM_TCREATE(cudaMalloc(&t->d_cur_mp2, C * 4));
    DVM_TCREATE(cudaMalloc(&t->d_cnt, 8 * 4)); DVM_TCREATE(cudaMemset(t->d_cnt, 0, 32));
    DVM_TCREATE(cudaMalloc(&t->d_pose, 28)); DVM_TCREATE(cudaMalloc(&t->d_pose_last, 28)); DVM_TCREATE(cudaMalloc(&t->d_pose_prev, 28));
    DVM_TCREATE(cudaMalloc(&t->d_res1, 16)); DVM_TCREATE(cudaMalloc(&t->d_res2, 16));
    DVM_TCREATE(cudaMalloc(&t->d_result, 64)); DVM_TCREATE(cudaMemset(t->d_result, 0, 64));
    DVM_TCREATE(cudaHostAlloc(&t->h_result, 128, cudaHostAllocDefault));
    DVM_TCREATE(cudaHostAlloc(&t->h_ring, kRing * 64, cudaHostAllocMapped));
    memset(t->h_ring, 0, kRing * 64);
    DVM_TCREATE(cudaHostGetDevicePointer(&t->d_ring, t->h_ring, 0));
#undef DVM_TCREATE
    *out = t;
    return DVM_OK;
}

void dvm_tracker_destroy(dvm_tracker* t) { tracker_free(t); }

int dvm_tracker_set_distortion(dvm_tracker* t, const float* dist5)
{
    DVM_REQUIRE(t != nullptr && dist5 != nullptr, "null argument");
    for (auto f : t->frames) dvm_frame_set_distortion(f, t->K, dist5);
    return DVM_OK;
}
void* dvm_tracker_stream(const dvm_tracker* t) { return t ? (void*)t->stream : nullptr; }

int dvm_tracker_set_profiling(dvm_tracker* t, int on)
{
    DVM_REQUIRE(t != nullptr, "null handle");
    DVM_CUDA(cudaSetDevice(t->device));
    for (auto& e : t->pev)
        if (on && !e) DVM_CUDA(cudaEventCreate(&e));
    t->profiling = on != 0;
    for (double& v : t->prof_ms) v = 0;
    t->prof_frames = 0;
    return DVM_OK;
}

int dvm_tracker_get_profile(const dvm_tracker* t, double* segment_ms, long long* frames)
{
    DVM_REQUIRE(t != nullptr && segment_ms && frames, "null argument");
    for (int i = 0; i < 5; i++) segment_ms[i] = t->prof_ms[i];
    *frames = t->prof_frames;
    return DVM_OK;
}

// SearchLocalPoints: isInFrustum + SearchByProjection(local map) + merge into cur_map, two launches
static int enqueue_local_map_search(dvm_tracker* t, dvm_frame* cur, int* cur_map, float th, float nnratio)
{
    MatchMapArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.m = t->map_n;
    ma.mp_desc = t->d_desc; ma.obs_pos = nullptr;
    ma.th = th; ma.nnratio = nnratio; ma.cur_map = cur_map; ma.merge_into = cur_map;
    ma.use_frustum = 1;
    fill_frustum_args(ma.fr, t->d_pose, t->K, t->bounds, t->nlevels, t->logScale, 0.5f);
    ma.fr.m = t->map_n;
    ma.fr.xw = t->d_xw; ma.fr.normal = t->d_normal; ma.fr.min_dist = t->d_mind; ma.fr.max_dist = t->d_maxd; ma.fr.skip = t->d_seen;
    launch_match_map(cur->dev, ma, cur->ms, t->d_cur_mp2, t->d_cnt + 3, t->stream);
    return DVM_OK;
}

// Frame construction on the extraction stream: H2D staging (if the image is on the host), ExtractORB
// into the ring slot of frame n_extracted, AssignFeaturesToGrid
static int enqueue_extract(dvm_tracker* t, const uint8_t* gray, int gray_is_device, int width, int height, int stride)
{
    const long long i = t->n_extracted;
    const int slot = (int)(i % kRing), e = (int)(i % kExtractors);
    cudaStream_t es = t->es[e];
    // the slot was the "last frame" of chain i - (kRing - 1): it is free again once that chain has finished
    DVM_REQUIRE(i - t->n_tracked < kMaxPending, "too many frames pending extraction");
    if (i >= kRing - 1) DVM_CUDA(cudaStreamWaitEvent(es, t->ev_done[(i - (kRing - 1)) % kRing], 0));
    const uint8_t* img = gray;
    int istride = stride;
    if (!gray_is_device) {
        const size_t need = (size_t)width * height;
        if (need > t->img_cap[e]) {
            DVM_CUDA(cudaStreamSynchronize(es));
            cudaFree(t->d_img[e]); t->d_img[e] = nullptr;
            DVM_CUDA(cudaMalloc(&t->d_img[e], need));
            t->img_cap[e] = need;
        }
        DVM_CUDA(cudaMemcpy2DAsync(t->d_img[e], width, gray, stride, width, height, cudaMemcpyHostToDevice, es));
        img = t->d_img[e]; istride = width;
    }
    int rc = dvm_frame_construct_device(t->frames[slot], t->orbs[e], img, width, height, istride, t->bounds[0], t->bounds[1],
                                        t->bounds[2], t->bounds[3]);
    if (rc != DVM_OK) return rc;
    DVM_CUDA(cudaEventRecord(t->ev_extracted[slot], es));
    t->n_extracted = i + 1;
    return DVM_OK;
}

int dvm_tracker_bootstrap(dvm_tracker* t, const uint8_t* gray, int width, int height, int stride, const float* pose_q,
                          const float* pose_t, int* n_matched)
{
    DVM_REQUIRE(t && gray && pose_q && pose_t, "null argument");
    DVM_CUDA(cudaSetDevice(t->device));
    for (auto e : t->es) DVM_CUDA(cudaStreamSynchronize(e));
    DVM_CUDA(cudaStreamSynchronize(t->stream));
    t->n_extracted = 0;
    t->n_tracked = 0;
    memset(t->h_ring, 0, kRing * 64);   // (every stream is idle here)
    int rc = enqueue_extract(t, gray, 0, width, height, stride);
    if (rc != DVM_OK) return rc;
    dvm_frame* cur = t->frames[0];
    DVM_CUDA(cudaStreamWaitEvent(t->stream, t->ev_extracted[0], 0));
    float* pose = reinterpret_cast<float*>(t->h_result + 64);
    for (int i = 0; i < 4; i++) pose[i] = pose_q[i];
    for (int i = 0; i < 3; i++) pose[4 + i] = pose_t[i];
    DVM_CUDA(cudaMemcpyAsync(t->d_pose, pose, 28, cudaMemcpyHostToDevice, t->stream));
    DVM_CUDA(cudaMemcpyAsync(t->d_pose_last, pose, 28, cudaMemcpyHostToDevice, t->stream));
    DVM_CUDA(cudaMemcpyAsync(t->d_pose_prev, pose, 28, cudaMemcpyHostToDevice, t->stream));
    DVM_CUDA(cudaMemsetAsync(t->d_seen, 0, t->map_n, t->stream));
    DVM_CUDA(cudaMemsetAsync(t->d_cnt, 0, 32, t->stream));
    DVM_CUDA(cudaMemsetAsync(t->d_mp[0], 0xff, (size_t)t->cap * 4, t->stream));
    DVM_CUDA(cudaMemsetAsync(t->d_outl[0], 0, t->cap, t->stream));
    rc = enqueue_local_map_search(t, cur, t->d_mp[0], 3.0f, 0.8f);
    if (rc != DVM_OK) return rc;
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaEventRecord(t->ev_done[0], t->stream));
    t->n_tracked = 1;
    int nm = 0;
    DVM_CUDA(cudaMemcpyAsync(&nm, t->d_cnt + 3, 4, cudaMemcpyDeviceToHost, t->stream));
    DVM_CUDA(cudaStreamSynchronize(t->stream));
    if (n_matched) *n_matched = nm;
    return DVM_OK;
}

int dvm_tracker_prefetch(dvm_tracker* t, const uint8_t* gray, int gray_is_device, int width, int height, int stride)
{
    DVM_REQUIRE(t && gray, "null argument");
    DVM_REQUIRE(t->n_tracked >= 1, "bootstrap the tracker first");
    DVM_REQUIRE(t->n_extracted - t->n_tracked < kMaxPending, "two prefetched frames are already pending");
    DVM_CUDA(cudaSetDevice(t->device));
    return enqueue_extract(t, gray, gray_is_device, width, height, stride);
}

int dvm_tracker_track(dvm_tracker* t, const uint8_t* gray, int gray_is_device, int width, int height, int stride,
                      const float* prior_q, const float* prior_t, int sync, float* pose_out, int32_t* counts)
{
    DVM_REQUIRE(t != nullptr, "null handle");
    DVM_REQUIRE(t->n_tracked >= 1, "bootstrap the tracker first");
    const bool prefetched = t->n_extracted > t->n_tracked;
    DVM_REQUIRE(gray != nullptr || prefetched, "null image and no prefetched frame");
    DVM_REQUIRE((prior_q == nullptr) == (prior_t == nullptr), "prior_q and prior_t go together");
    DVM_CUDA(cudaSetDevice(t->device));
    int rc;
    if (!prefetched) {
        rc = enqueue_extract(t, gray, gray_is_device, width, height, stride);
        if (rc != DVM_OK) return rc;
    }
    const long long fi = t->n_tracked;
    const int ci = (int)(fi % kRing), li = (int)((fi - 1) % kRing);
    dvm_frame* last = t->frames[li];
    dvm_frame* cur = t->frames[ci];
    DVM_CUDA(cudaStreamWaitEvent(t->stream, t->ev_extracted[ci], 0));
    // mCurrentFrame.SetPose(mVelocity * mLastFrame.GetPose()); reset of the per-frame marks and counters
    if (prior_q) {
        float* pose = reinterpret_cast<float*>(t->h_result + 64);
        for (int i = 0; i < 4; i++) pose[i] = prior_q[i];
        for (int i = 0; i < 3; i++) pose[4 + i] = prior_t[i];
        // staged through the pinned result block's upper half (a stack array may not outlive the async copy)
        DVM_CUDA(cudaMemcpyAsync(t->d_pose, pose, 28, cudaMemcpyHostToDevice, t->stream));
    }
    auto mark = [&](int i) { if (t->profiling) cudaEventRecord(t->pev[i], t->stream); };
    mark(0);
    // every later frame finds its prior and the cleared marks already there: the previous chain's last kernel left them
    if (fi == 1)
        DVM_LAUNCH_PDL(begin_frame_kernel, 1, 1024, 0, t->stream, t->d_pose_last, t->d_pose_prev, t->d_pose, prior_q ? 1 : 0, t->d_seen,
                   t->map_n, t->d_cnt);
    mark(1);
    // ---- TrackWithMotionModel: SearchByProjection(cur, last, th = 15), retry with 2*th below 20 matches ----
    MatchLastArgs la;
    memset(&la, 0, sizeof(la));
    la.last_n = last->dev.cap; la.n_ptr = last->d_n;
    la.mp_index = t->d_mp[li]; la.outlier = t->d_outl[li];
    la.Xw = t->d_xw; la.mp_desc = t->d_desc; la.obs_pos = nullptr; la.last_kps = last->d_kps;
    la.pose = t->d_pose; la.th = 15.0f; la.check_ori = 1;
    la.retry_th = 30.0f;   // the 2 * th retry runs inside the resolution kernel when it is needed
    la.map_out = t->d_mp[ci];
    for (int i = 0; i < 4; i++) la.K[i] = t->K[i];
    launch_match_last(cur->dev, la, cur->ms, t->d_cur_mp, t->d_cnt + 1, t->stream);
    mark(2);
    // ---- PoseOptimization + discard outliers (fused tail) ----
    PoseOptArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.n = cur->dev.cap; pa.n_ptr = cur->d_n; pa.map_index = t->d_mp[ci]; pa.Xw = t->d_xw; pa.kps = cur->d_kps;
    for (int i = 0; i < t->nlevels; i++) pa.inv_sigma2_table[i] = t->inv_sigma2[i];
    for (int i = 0; i < 4; i++) pa.K[i] = t->K[i];
    pa.pose = t->d_pose; pa.outlier = t->d_outl[ci]; pa.result = t->d_res1;
    pa.seen = t->d_seen; pa.map_index_rw = t->d_mp[ci];
    rc = launch_pose_opt(pa, t->stream);
    if (rc != DVM_OK) return rc;
    mark(3);
    // ---- TrackLocalMap: SearchLocalPoints (isInFrustum, th = 1, nnratio 0.8) + PoseOptimization ----
    rc = enqueue_local_map_search(t, cur, t->d_mp[ci], 1.0f, 0.8f);
    if (rc != DVM_OK) return rc;
    mark(4);
    pa.result = t->d_res2;
    pa.seen = nullptr; pa.map_index_rw = nullptr;
    pa.pose_last = t->d_pose_last; pa.pose_prev = t->d_pose_prev;
    pa.out_pose = (float*)(t->d_ring + 64 * ci); pa.out_counts = (int*)(t->d_ring + 64 * ci + 32);
    pa.nm_last = t->d_cnt + 1; pa.res_first = t->d_res1;
    pa.next_prior = t->d_pose; pa.seen_reset = t->d_seen; pa.seen_n = t->map_n;
    rc = launch_pose_opt(pa, t->stream);
    if (rc != DVM_OK) return rc;
    mark(5);
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaEventRecord(t->ev_done[ci], t->stream));
    t->n_tracked = fi + 1;
    if (t->profiling) {
        DVM_CUDA(cudaEventSynchronize(t->pev[5]));
        for (int i = 0; i < 5; i++) {
            float ms = 0;
            DVM_CUDA(cudaEventElapsedTime(&ms, t->pev[i], t->pev[i + 1]));
            t->prof_ms[i] += ms;
        }
        t->prof_frames++;
    }
    if (sync) return dvm_tracker_result(t, pose_out, counts);
    return DVM_OK;
}

int dvm_tracker_result_lag(dvm_tracker* t, int lag, float* pose_out, int32_t* counts)
{
    DVM_REQUIRE(t != nullptr, "null handle");
    DVM_REQUIRE(lag >= 0 && lag <= kRing - 2, "lag out of range (the result ring keeps the last three frames)");
    DVM_REQUIRE(t->n_tracked - 1 - lag >= 0, "no such frame yet");
    DVM_CUDA(cudaSetDevice(t->device));
    const int slot = (int)((t->n_tracked - 1 - lag) % kRing);
    DVM_CUDA(cudaEventSynchronize(t->ev_done[slot]));   // that frame's chain only: later frames keep running
    if (pose_out) memcpy(pose_out, t->h_ring + 64 * slot, 28);
    if (counts) memcpy(counts, t->h_ring + 64 * slot + 32, 16);
    return DVM_OK;
}

int dvm_tracker_result(dvm_tracker* t, float* pose_out, int32_t* counts) { return dvm_tracker_result_lag(t, 0, pose_out, counts); }

int dvm_tracker_debug_matches(dvm_tracker* t, int32_t* cur_map, uint8_t* outlier, int cap, int* n_out)
{
    DVM_REQUIRE(t && n_out, "null argument");
    DVM_CUDA(cudaSetDevice(t->device));
    DVM_CUDA(cudaStreamSynchronize(t->stream));
    int n = 0;
    const int li = (int)((t->n_tracked - 1) % kRing);
    DVM_CUDA(cudaMemcpy(&n, t->frames[li]->d_n, 4, cudaMemcpyDeviceToHost));
    n = std::min(n, t->cap);
    *n_out = n;
    const int m = std::min(n, cap);
    if (cur_map && m > 0) DVM_CUDA(cudaMemcpy(cur_map, t->d_mp[li], (size_t)m * 4, cudaMemcpyDeviceToHost));
    if (outlier && m > 0) DVM_CUDA(cudaMemcpy(outlier, t->d_outl[li], (size_t)m, cudaMemcpyDeviceToHost));
    return DVM_OK;
}

} // extern "C"
