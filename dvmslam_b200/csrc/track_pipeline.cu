// track_pipeline.cu -- Frame::isInFrustum on the GPU and the device-resident per-frame tracker that
// chains ExtractORB -> Frame -> SearchByProjection(last frame) -> PoseOptimization -> isInFrustum +
// SearchByProjection(local map) -> PoseOptimization without host round trips
// (call order of Tracking::TrackWithMotionModel / TrackLocalMap / SearchLocalPoints,
//  O3/src/Tracking.cc:2584-2666, 2668-2768, 3041-3106).
#include "track_internal.cuh"
#include <cmath>
#include <vector>

using namespace dvm;

namespace {

struct FrustumArgs {
    const float* pose;  // device: qx,qy,qz,qw,tx,ty,tz
    float K[4], bounds[4];
    int nlevels;
    float logScale, cosLimit;
    int m;
    const float* xw; const float* normal; const float* min_dist; const float* max_dist;
    const uint8_t* skip;
    uint8_t* in_view; float* px; float* py; int* level; float* view_cos;
};

// Frame::isInFrustum (mono branch) + MapPoint::PredictScale, one thread per map point
__global__ void __launch_bounds__(256) frustum_kernel(FrustumArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.m) return;
    float R[9];
    quat_to_R_f32(a.pose, R);
    const float t0 = a.pose[4], t1 = a.pose[5], t2 = a.pose[6];
    uint8_t vis = 0;
    float u = -1.f, v = -1.f, vc = 0.f;
    int lvl = -1;
    if (!(a.skip && a.skip[k])) {
        const float X = a.xw[3 * k], Y = a.xw[3 * k + 1], Z = a.xw[3 * k + 2];
        const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[0], X), __fmul_rn(R[1], Y)), __fmul_rn(R[2], Z)), t0);
        const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[3], X), __fmul_rn(R[4], Y)), __fmul_rn(R[5], Z)), t1);
        const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(R[6], X), __fmul_rn(R[7], Y)), __fmul_rn(R[8], Z)), t2);
        if (!(zc < 0.0f)) {
            const float pu = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[0], xc), zc), a.K[2]);
            const float pv = __fadd_rn(__fdiv_rn(__fmul_rn(a.K[1], yc), zc), a.K[3]);
            if (!(pu < a.bounds[0] || pu > a.bounds[2]) && !(pv < a.bounds[1] || pv > a.bounds[3])) {
                u = pu; v = pv;
                const float maxD = __fmul_rn(1.2f, a.max_dist[k]), minD = __fmul_rn(0.8f, a.min_dist[k]);
                float Ow[3];
#pragma unroll
                for (int i = 0; i < 3; i++)
                    Ow[i] = __fadd_rn(__fadd_rn(__fmul_rn(R[i], -t0), __fmul_rn(R[3 + i], -t1)), __fmul_rn(R[6 + i], -t2));
                const float p0 = __fsub_rn(X, Ow[0]), p1 = __fsub_rn(Y, Ow[1]), p2 = __fsub_rn(Z, Ow[2]);
                const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(p0, p0), __fmul_rn(p1, p1)), __fmul_rn(p2, p2)));
                if (!(dist < minD || dist > maxD)) {
                    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(p0, a.normal[3 * k]), __fmul_rn(p1, a.normal[3 * k + 1])),
                                                __fmul_rn(p2, a.normal[3 * k + 2]));
                    const float c = __fdiv_rn(dot, dist);
                    if (!(c < a.cosLimit)) {
                        const float ratio = __fdiv_rn(a.max_dist[k], dist);
                        int nScale = (int)ceilf(__fdiv_rn((float)log((double)ratio), a.logScale));
                        if (nScale < 0) nScale = 0;
                        else if (nScale >= a.nlevels) nScale = a.nlevels - 1;
                        vis = 1; lvl = nScale; vc = c;
                    }
                }
            }
        }
    }
    a.in_view[k] = vis; a.px[k] = u; a.py[k] = v; a.level[k] = lvl; a.view_cos[k] = vc;
}

// ordered compaction of the in-view map points (vpMapPoints order is the greedy priority)
__global__ void __launch_bounds__(1024) compact_inview_kernel(int m, const uint8_t* in_view, const float* px, const float* py,
                                                               const int* level, const float* view_cos, int* q_index,
                                                               float* qx, float* qy, int* qlevel, float* qcos, int* count)
{
    __shared__ int warp_sums[33];
    const int tid = threadIdx.x;
    const int ipt = (m + 1023) / 1024;
    const int i0 = min(tid * ipt, m), i1 = min(i0 + ipt, m);
    int c = 0;
    for (int i = i0; i < i1; i++) c += in_view[i];
    int total;
    int off = block_exclusive_scan(c, warp_sums, &total);
    for (int i = i0; i < i1; i++)
        if (in_view[i]) {
            q_index[off] = i; qx[off] = px[i]; qy[off] = py[i]; qlevel[off] = level[i]; qcos[off] = view_cos[i];
            off++;
        }
    if (tid == 0) *count = total;
}

// cur_map[i] = map point of the last-frame keypoint matched to current keypoint i
__global__ void after_last_kernel(const int* n_ptr, int cap, const int* cur_mp, const int* last_mp, int* cur_map)
{
    const int n = min(*n_ptr, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x)
        cur_map[i] = (i < n && cur_mp[i] >= 0) ? last_mp[cur_mp[i]] : -1;
}

// "Discard outliers" (Tracking.cc:2634-2654): matched map points are marked as seen in this frame,
// outlier matches are dropped
__global__ void discard_kernel(const int* n_ptr, int cap, int* cur_map, uint8_t* outlier, uint8_t* seen, int* nmatches)
{
    const int n = min(*n_ptr, cap);
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int m = cur_map[i];
        if (m < 0) continue;
        seen[m] = 1;
        if (outlier[i]) { cur_map[i] = -1; outlier[i] = 0; }
        else cnt++;
    }
    if (cnt) atomicAdd(nmatches, cnt);
}

__global__ void merge_kernel(const int* n_ptr, int cap, int* cur_map, const int* cur_mp2, const int* q_index)
{
    const int n = min(*n_ptr, cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (cur_map[i] < 0 && cur_mp2[i] >= 0) cur_map[i] = q_index[cur_mp2[i]];
}

// result block: pose[7] (float) | counts[4]; pose history for the constant-velocity prior
__global__ void finish_kernel(const int* n_ptr, int cap, const int* cur_map, const uint8_t* outlier, const float* pose,
                              float* pose_last, float* pose_prev, const int* nm_last, const int* res1, float* out_pose,
                              int* out_counts)
{
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int n = min(*n_ptr, cap);
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) cnt += (cur_map[i] >= 0 && !outlier[i]);
    if (cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x < 7) {
        pose_prev[threadIdx.x] = pose_last[threadIdx.x];
        pose_last[threadIdx.x] = pose[threadIdx.x];
        out_pose[threadIdx.x] = pose[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        out_counts[0] = n;
        out_counts[1] = *nm_last;
        out_counts[2] = res1[0];
        out_counts[3] = s_cnt; // mnMatchesInliers
    }
}

// prior = mVelocity * last, mVelocity = last * prev^-1 (Tracking.cc:1968-1971, 2598), in double then float
__global__ void prior_kernel(const float* last, const float* prev, float* prior)
{
    if (threadIdx.x != 0) return;
    auto qmul = [](const double* a, const double* b, double* r) {
        r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
        r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
        r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
        r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    };
    auto qrot = [](const double* q, const double* v, double* o) {
        double uv[3] = { q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0] };
        uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
        o[0] = v[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]);
        o[1] = v[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]);
        o[2] = v[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]);
    };
    double ql[4], tl[3], qp[4], tp[3];
    for (int i = 0; i < 4; i++) { ql[i] = last[i]; qp[i] = prev[i]; }
    for (int i = 0; i < 3; i++) { tl[i] = last[4 + i]; tp[i] = prev[4 + i]; }
    // V = L * P^-1:  qv = ql * conj(qp), tv = tl - qv * tp ;  prior = V * L
    double qpc[4] = { -qp[0], -qp[1], -qp[2], qp[3] }, qv[4], tmp[3], tv[3], qo[4], to[3];
    qmul(ql, qpc, qv);
    qrot(qv, tp, tmp);
    for (int i = 0; i < 3; i++) tv[i] = tl[i] - tmp[i];
    qmul(qv, ql, qo);
    qrot(qv, tl, tmp);
    for (int i = 0; i < 3; i++) to[i] = tmp[i] + tv[i];
    double n = sqrt(qo[0] * qo[0] + qo[1] * qo[1] + qo[2] * qo[2] + qo[3] * qo[3]);
    if (qo[3] < 0) n = -n;
    for (int i = 0; i < 4; i++) prior[i] = (float)(qo[i] / n);
    for (int i = 0; i < 3; i++) prior[4 + i] = (float)to[i];
}

} // namespace

// ------------------------------------------------------------------------------------------------
struct dvm_tracker {
    int device = 0;
    cudaStream_t stream = nullptr;
    dvm_orb* orb = nullptr;
    dvm_frame* frames[2] = { nullptr, nullptr };
    int idx = 0; // frames[idx] is the last frame
    int cap = 0, map_n = 0, nlevels = 0;
    float K[4], bounds[4], logScale = 0;
    std::vector<float> inv_sigma2;
    // map snapshot
    float* d_xw = nullptr; uint8_t* d_desc = nullptr; float* d_normal = nullptr; float* d_mind = nullptr; float* d_maxd = nullptr;
    // per-frame association (ping-pong with the frames)
    int* d_mp[2] = { nullptr, nullptr };
    uint8_t* d_outl[2] = { nullptr, nullptr };
    // scratch
    uint8_t* d_seen = nullptr; uint8_t* d_inview = nullptr;
    float* d_px = nullptr; float* d_py = nullptr; int* d_level = nullptr; float* d_cos = nullptr;
    int* d_qidx = nullptr; float* d_qx = nullptr; float* d_qy = nullptr; int* d_qlevel = nullptr; float* d_qcos = nullptr;
    int* d_cur_mp = nullptr; int* d_cur_mp2 = nullptr;
    int* d_cnt = nullptr;       // [8]: 0 inview count, 1 nmatches last, 2 nmatches after discard, 3 nmatches map
    float* d_pose = nullptr;    // [7] working pose (prior in, optimised out)
    float* d_pose_last = nullptr; float* d_pose_prev = nullptr;
    int* d_res1 = nullptr; int* d_res2 = nullptr;
    double* d_err = nullptr;
    uint8_t* d_result = nullptr; // pose[7] float | counts[4] int
    uint8_t* h_result = nullptr; // pinned: [0,48) result read-back, [64,92) prior staging
    uint8_t* d_img = nullptr; size_t img_cap = 0;
};

static void tracker_free(dvm_tracker* t)
{
    if (!t) return;
    cudaSetDevice(t->device);
    for (auto f : t->frames) if (f) dvm_frame_destroy(f);
    void* ptrs[] = { t->d_xw, t->d_desc, t->d_normal, t->d_mind, t->d_maxd, t->d_mp[0], t->d_mp[1], t->d_outl[0], t->d_outl[1],
                     t->d_seen, t->d_inview, t->d_px, t->d_py, t->d_level, t->d_cos, t->d_qidx, t->d_qx, t->d_qy, t->d_qlevel,
                     t->d_qcos, t->d_cur_mp, t->d_cur_mp2, t->d_cnt, t->d_pose, t->d_pose_last, t->d_pose_prev, t->d_res1,
                     t->d_res2, t->d_err, t->d_result, t->d_img };
    for (void* p : ptrs) cudaFree(p);
    if (t->h_result) cudaFreeHost(t->h_result);
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
    const float logScale = (float)std::log((double)(f->dev.nlevels > 1 ? f->dev.scale[1] : 1.2f));
    fill_frustum_args(a, (const float*)(f->d_in + o_pose), K, bounds, f->dev.nlevels, logScale, viewing_cos_limit);
    a.m = m;
    a.xw = (const float*)(f->d_in + o_xw); a.normal = (const float*)(f->d_in + o_nrm);
    a.min_dist = (const float*)(f->d_in + o_min); a.max_dist = (const float*)(f->d_in + o_max);
    a.skip = f->d_in + o_skip;
    a.in_view = f->d_in + o_vis; a.px = (float*)(f->d_in + o_px); a.py = (float*)(f->d_in + o_py);
    a.level = (int*)(f->d_in + o_lv); a.view_cos = (float*)(f->d_in + o_cos);
    DVM_LAUNCH(frustum_kernel, div_up(m, 256), 256, 0, f->stream, a);
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
    t->orb = orb;
    t->stream = (cudaStream_t)dvm_orb_stream(orb);
    t->cap = dvm_orb_max_keypoints(orb);
    t->map_n = map_n;
    float sc[16], is2[16];
    dvm_orb_tables(orb, &t->nlevels, sc, nullptr, nullptr, is2, nullptr);
    t->inv_sigma2.assign(is2, is2 + t->nlevels);
    t->logScale = (float)std::log((double)(t->nlevels > 1 ? sc[1] : 1.2f));
    for (int i = 0; i < 4; i++) { t->K[i] = K[i]; t->bounds[i] = bounds[i]; }
    cudaGetDevice(&t->device);
    for (int i = 0; i < 2; i++) {
        int rc = dvm_frame_create(&t->frames[i], t->device, t->stream, t->cap, t->nlevels, sc, is2);
        if (rc != DVM_OK) { tracker_free(t); return rc; }
        rc = dvm_frame_ensure_query_cap(t->frames[i], std::max(map_n, t->cap));
        if (rc != DVM_OK) { tracker_free(t); return rc; }
    }
#define DVM_TCREATE(call)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            set_error("%s failed in dvm_tracker_create: %s", #call, cudaGetErrorString(e__));  \
            tracker_free(t);                                                                   \
            return DVM_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)
    const size_t M = (size_t)map_n, C = (size_t)t->cap;
    DVM_TCREATE(cudaMalloc(&t->d_xw, M * 12)); DVM_TCREATE(cudaMalloc(&t->d_desc, M * 32));
    DVM_TCREATE(cudaMalloc(&t->d_normal, M * 12)); DVM_TCREATE(cudaMalloc(&t->d_mind, M * 4)); DVM_TCREATE(cudaMalloc(&t->d_maxd, M * 4));
    DVM_TCREATE(cudaMemcpy(t->d_xw, map_xw, M * 12, cudaMemcpyHostToDevice));
    DVM_TCREATE(cudaMemcpy(t->d_desc, map_desc, M * 32, cudaMemcpyHostToDevice));
    DVM_TCREATE(cudaMemcpy(t->d_normal, map_normal, M * 12, cudaMemcpyHostToDevice));
    DVM_TCREATE(cudaMemcpy(t->d_mind, map_min_dist, M * 4, cudaMemcpyHostToDevice));
    DVM_TCREATE(cudaMemcpy(t->d_maxd, map_max_dist, M * 4, cudaMemcpyHostToDevice));
    for (int i = 0; i < 2; i++) {
        DVM_TCREATE(cudaMalloc(&t->d_mp[i], C * 4)); DVM_TCREATE(cudaMemset(t->d_mp[i], 0xff, C * 4));
        DVM_TCREATE(cudaMalloc(&t->d_outl[i], C)); DVM_TCREATE(cudaMemset(t->d_outl[i], 0, C));
    }
    DVM_TCREATE(cudaMalloc(&t->d_seen, M)); DVM_TCREATE(cudaMalloc(&t->d_inview, M));
    DVM_TCREATE(cudaMalloc(&t->d_px, M * 4)); DVM_TCREATE(cudaMalloc(&t->d_py, M * 4));
    DVM_TCREATE(cudaMalloc(&t->d_level, M * 4)); DVM_TCREATE(cudaMalloc(&t->d_cos, M * 4));
    DVM_TCREATE(cudaMalloc(&t->d_qidx, M * 4)); DVM_TCREATE(cudaMalloc(&t->d_qx, M * 4)); DVM_TCREATE(cudaMalloc(&t->d_qy, M * 4));
    DVM_TCREATE(cudaMalloc(&t->d_qlevel, M * 4)); DVM_TCREATE(cudaMalloc(&t->d_qcos, M * 4));
    DVM_TCREATE(cudaMalloc(&t->d_cur_mp, C * 4)); DVM_TCREATE(cudaMalloc(&t->d_cur_mp2, C * 4));
    DVM_TCREATE(cudaMalloc(&t->d_cnt, 8 * 4)); DVM_TCREATE(cudaMemset(t->d_cnt, 0, 32));
    DVM_TCREATE(cudaMalloc(&t->d_pose, 28)); DVM_TCREATE(cudaMalloc(&t->d_pose_last, 28)); DVM_TCREATE(cudaMalloc(&t->d_pose_prev, 28));
    DVM_TCREATE(cudaMalloc(&t->d_res1, 16)); DVM_TCREATE(cudaMalloc(&t->d_res2, 16));
    DVM_TCREATE(cudaMalloc(&t->d_err, C * 16));
    DVM_TCREATE(cudaMalloc(&t->d_result, 64)); DVM_TCREATE(cudaMemset(t->d_result, 0, 64));
    DVM_TCREATE(cudaHostAlloc(&t->h_result, 128, cudaHostAllocDefault));
#undef DVM_TCREATE
    *out = t;
    return DVM_OK;
}

void dvm_tracker_destroy(dvm_tracker* t) { tracker_free(t); }

// frustum -> ordered compaction -> SearchByProjection(local map) -> merge, all on the stream
static int enqueue_local_map_search(dvm_tracker* t, dvm_frame* cur, int* cur_map, float th, float nnratio)
{
    FrustumArgs fa;
    fill_frustum_args(fa, t->d_pose, t->K, t->bounds, t->nlevels, t->logScale, 0.5f);
    fa.m = t->map_n;
    fa.xw = t->d_xw; fa.normal = t->d_normal; fa.min_dist = t->d_mind; fa.max_dist = t->d_maxd; fa.skip = t->d_seen;
    fa.in_view = t->d_inview; fa.px = t->d_px; fa.py = t->d_py; fa.level = t->d_level; fa.view_cos = t->d_cos;
    DVM_LAUNCH(frustum_kernel, div_up(t->map_n, 256), 256, 0, t->stream, fa);
    DVM_LAUNCH(compact_inview_kernel, 1, 1024, 0, t->stream, t->map_n, t->d_inview, t->d_px, t->d_py, t->d_level, t->d_cos,
               t->d_qidx, t->d_qx, t->d_qy, t->d_qlevel, t->d_qcos, t->d_cnt + 0);
    MatchMapArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.m = t->map_n; ma.m_ptr = t->d_cnt + 0;
    ma.projX = t->d_qx; ma.projY = t->d_qy; ma.level = t->d_qlevel; ma.view_cos = t->d_qcos;
    ma.mp_desc = t->d_desc; ma.q_index = t->d_qidx; ma.obs_pos = nullptr;
    ma.th = th; ma.nnratio = nnratio; ma.cur_map = cur_map;
    launch_match_map(cur->dev, ma, cur->ms, t->d_cur_mp2, t->d_cnt + 3, t->stream);
    DVM_LAUNCH(merge_kernel, div_up(t->cap, 256), 256, 0, t->stream, cur->d_n, t->cap, cur_map, t->d_cur_mp2, t->d_qidx);
    return DVM_OK;
}

static int stage_image(dvm_tracker* t, const uint8_t* gray, int gray_is_device, int width, int height, int stride,
                       const uint8_t** dev_img, int* dev_stride)
{
    if (gray_is_device) { *dev_img = gray; *dev_stride = stride; return DVM_OK; }
    const size_t need = (size_t)width * height;
    if (need > t->img_cap) {
        DVM_CUDA(cudaStreamSynchronize(t->stream));
        cudaFree(t->d_img); t->d_img = nullptr;
        DVM_CUDA(cudaMalloc(&t->d_img, need));
        t->img_cap = need;
    }
    DVM_CUDA(cudaMemcpy2DAsync(t->d_img, width, gray, stride, width, height, cudaMemcpyHostToDevice, t->stream));
    *dev_img = t->d_img; *dev_stride = width;
    return DVM_OK;
}

int dvm_tracker_bootstrap(dvm_tracker* t, const uint8_t* gray, int width, int height, int stride, const float* pose_q,
                          const float* pose_t, int* n_matched)
{
    DVM_REQUIRE(t && gray && pose_q && pose_t, "null argument");
    DVM_CUDA(cudaSetDevice(t->device));
    const uint8_t* img; int istride;
    int rc = stage_image(t, gray, 0, width, height, stride, &img, &istride);
    if (rc != DVM_OK) return rc;
    rc = dvm_orb_extract_device(t->orb, img, width, height, istride, 0, 1000);
    if (rc != DVM_OK) return rc;
    dvm_frame* cur = t->frames[t->idx];
    rc = dvm_frame_assign_from_orb(cur, t->orb, t->bounds[0], t->bounds[1], t->bounds[2], t->bounds[3]);
    if (rc != DVM_OK) return rc;
    float pose[7] = { pose_q[0], pose_q[1], pose_q[2], pose_q[3], pose_t[0], pose_t[1], pose_t[2] };
    DVM_CUDA(cudaMemcpyAsync(t->d_pose, pose, 28, cudaMemcpyHostToDevice, t->stream));
    DVM_CUDA(cudaMemcpyAsync(t->d_pose_last, pose, 28, cudaMemcpyHostToDevice, t->stream));
    DVM_CUDA(cudaMemcpyAsync(t->d_pose_prev, pose, 28, cudaMemcpyHostToDevice, t->stream));
    DVM_CUDA(cudaMemsetAsync(t->d_seen, 0, t->map_n, t->stream));
    DVM_CUDA(cudaMemsetAsync(t->d_mp[t->idx], 0xff, (size_t)t->cap * 4, t->stream));
    DVM_CUDA(cudaMemsetAsync(t->d_outl[t->idx], 0, t->cap, t->stream));
    rc = enqueue_local_map_search(t, cur, t->d_mp[t->idx], 3.0f, 0.8f);
    if (rc != DVM_OK) return rc;
    DVM_CUDA(cudaGetLastError());
    int nm = 0;
    DVM_CUDA(cudaMemcpyAsync(&nm, t->d_cnt + 3, 4, cudaMemcpyDeviceToHost, t->stream));
    DVM_CUDA(cudaStreamSynchronize(t->stream));
    if (n_matched) *n_matched = nm;
    return DVM_OK;
}

int dvm_tracker_track(dvm_tracker* t, const uint8_t* gray, int gray_is_device, int width, int height, int stride,
                      const float* prior_q, const float* prior_t, int sync, float* pose_out, int32_t* counts)
{
    DVM_REQUIRE(t && gray, "null argument");
    DVM_REQUIRE((prior_q == nullptr) == (prior_t == nullptr), "prior_q and prior_t go together");
    DVM_CUDA(cudaSetDevice(t->device));
    const int li = t->idx, ci = t->idx ^ 1;
    dvm_frame* last = t->frames[li];
    dvm_frame* cur = t->frames[ci];
    const uint8_t* img; int istride;
    int rc = stage_image(t, gray, gray_is_device, width, height, stride, &img, &istride);
    if (rc != DVM_OK) return rc;
    rc = dvm_orb_extract_device(t->orb, img, width, height, istride, 0, 1000);     // Frame::ExtractORB(0, im, 0, 1000)
    if (rc != DVM_OK) return rc;
    rc = dvm_frame_assign_from_orb(cur, t->orb, t->bounds[0], t->bounds[1], t->bounds[2], t->bounds[3]);
    if (rc != DVM_OK) return rc;
    // mCurrentFrame.SetPose(mVelocity * mLastFrame.GetPose())
    if (prior_q) {
        float pose[7] = { prior_q[0], prior_q[1], prior_q[2], prior_q[3], prior_t[0], prior_t[1], prior_t[2] };
        // staged through the pinned result block's upper half (a stack array may not outlive the async copy)
        memcpy(t->h_result + 64, pose, 28);
        DVM_CUDA(cudaMemcpyAsync(t->d_pose, t->h_result + 64, 28, cudaMemcpyHostToDevice, t->stream));
    } else {
        DVM_LAUNCH(prior_kernel, 1, 32, 0, t->stream, t->d_pose_last, t->d_pose_prev, t->d_pose);
    }
    DVM_CUDA(cudaMemsetAsync(t->d_seen, 0, t->map_n, t->stream));
    DVM_CUDA(cudaMemsetAsync(t->d_cnt, 0, 32, t->stream));
    // ---- TrackWithMotionModel: SearchByProjection(cur, last, th = 15), retry with 2*th below 20 matches ----
    MatchLastArgs la;
    memset(&la, 0, sizeof(la));
    la.last_n = t->cap; la.n_ptr = last->d_n;
    la.mp_index = t->d_mp[li]; la.outlier = t->d_outl[li];
    la.Xw = t->d_xw; la.mp_desc = t->d_desc; la.obs_pos = nullptr; la.last_kps = last->d_kps;
    la.pose = t->d_pose; la.th = 15.0f; la.check_ori = 1;
    for (int i = 0; i < 4; i++) la.K[i] = t->K[i];
    launch_match_last(cur->dev, la, cur->ms, t->d_cur_mp, t->d_cnt + 1, t->stream);
    la.th = 30.0f; la.guard = t->d_cnt + 1;
    launch_match_last(cur->dev, la, cur->ms, t->d_cur_mp, t->d_cnt + 1, t->stream);
    DVM_LAUNCH(after_last_kernel, div_up(t->cap, 256), 256, 0, t->stream, cur->d_n, t->cap, t->d_cur_mp, t->d_mp[li], t->d_mp[ci]);
    // ---- PoseOptimization, discard outliers ----
    PoseOptArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.n = t->cap; pa.n_ptr = cur->d_n; pa.map_index = t->d_mp[ci]; pa.Xw = t->d_xw; pa.kps = cur->d_kps;
    for (int i = 0; i < t->nlevels; i++) pa.inv_sigma2_table[i] = t->inv_sigma2[i];
    for (int i = 0; i < 4; i++) pa.K[i] = t->K[i];
    pa.pose = t->d_pose; pa.outlier = t->d_outl[ci]; pa.result = t->d_res1; pa.err = t->d_err;
    launch_pose_opt(pa, t->stream);
    DVM_LAUNCH(discard_kernel, div_up(t->cap, 256), 256, 0, t->stream, cur->d_n, t->cap, t->d_mp[ci], t->d_outl[ci], t->d_seen,
               t->d_cnt + 2);
    // ---- TrackLocalMap: SearchLocalPoints (isInFrustum, th = 1, nnratio 0.8) + PoseOptimization ----
    rc = enqueue_local_map_search(t, cur, t->d_mp[ci], 1.0f, 0.8f);
    if (rc != DVM_OK) return rc;
    pa.result = t->d_res2;
    launch_pose_opt(pa, t->stream);
    DVM_LAUNCH(finish_kernel, 1, 256, 0, t->stream, cur->d_n, t->cap, t->d_mp[ci], t->d_outl[ci], t->d_pose, t->d_pose_last,
               t->d_pose_prev, t->d_cnt + 1, t->d_res1, (float*)t->d_result, (int*)(t->d_result + 32));
    DVM_CUDA(cudaGetLastError());
    t->idx = ci;
    if (sync) return dvm_tracker_result(t, pose_out, counts);
    return DVM_OK;
}

int dvm_tracker_result(dvm_tracker* t, float* pose_out, int32_t* counts)
{
    DVM_REQUIRE(t != nullptr, "null handle");
    DVM_CUDA(cudaSetDevice(t->device));
    DVM_CUDA(cudaMemcpyAsync(t->h_result, t->d_result, 48, cudaMemcpyDeviceToHost, t->stream));
    DVM_CUDA(cudaStreamSynchronize(t->stream));
    if (pose_out) memcpy(pose_out, t->h_result, 28);
    if (counts) memcpy(counts, t->h_result + 32, 16);
    return DVM_OK;
}

int dvm_tracker_debug_matches(dvm_tracker* t, int32_t* cur_map, uint8_t* outlier, int cap, int* n_out)
{
    DVM_REQUIRE(t && n_out, "null argument");
    DVM_CUDA(cudaSetDevice(t->device));
    DVM_CUDA(cudaStreamSynchronize(t->stream));
    int n = 0;
    DVM_CUDA(cudaMemcpy(&n, t->frames[t->idx]->d_n, 4, cudaMemcpyDeviceToHost));
    n = std::min(n, t->cap);
    *n_out = n;
    const int m = std::min(n, cap);
    if (cur_map && m > 0) DVM_CUDA(cudaMemcpy(cur_map, t->d_mp[t->idx], (size_t)m * 4, cudaMemcpyDeviceToHost));
    if (outlier && m > 0) DVM_CUDA(cudaMemcpy(outlier, t->d_outl[t->idx], (size_t)m, cudaMemcpyDeviceToHost));
    return DVM_OK;
}

} // extern "C"
