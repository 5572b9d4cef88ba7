// track_capi.cu -- C-ABI of the Frame grid, the projection matchers and PoseOptimization.
// Host-array entry points upload their flat inputs, run the kernels on the frame's stream and
// read the results back (synchronous, like the reference's calls).
#include "track_internal.cuh"
#include <vector>

using namespace dvm;

static void frame_free(dvm_frame* f)
{
    if (!f) return;
    cudaSetDevice(f->device);
    cudaFree(f->d_kps); cudaFree(f->d_desc); cudaFree(f->d_n); cudaFree(f->d_cell_start); cudaFree(f->d_cell_items); cudaFree(f->d_kxyo);
    cudaFree(f->d_in); cudaFree(f->ms.pu); cudaFree(f->ms.pv); cudaFree(f->ms.pr); cudaFree(f->ms.plevels);
    cudaFree(f->ms.choice); cudaFree(f->ms.cache); cudaFree(f->ms.ncand); cudaFree(f->ms.qmeta); cudaFree(f->ms.cache8); cudaFree(f->ms.claim_a); cudaFree(f->ms.claim_b); cudaFree(f->ms.iters);
    cudaFree(f->d_cur_mp); cudaFree(f->d_err);
    if (f->h_in) cudaFreeHost(f->h_in);
    if (f->h_out) cudaFreeHost(f->h_out);
    if (f->ev) cudaEventDestroy(f->ev);
    if (f->own_stream && f->stream) cudaStreamDestroy(f->stream);
    delete f;
}

int dvm_frame_ensure_query_cap(dvm_frame* f, int nq)
{
    if (nq <= f->q_cap) return DVM_OK;
    DVM_REQUIRE(nq < (1 << 20), "more than 2^20 queries in one projection search");   // claim words carry 20 index bits
    const int cap = nq + nq / 4 + 256;
    DVM_CUDA(cudaStreamSynchronize(f->stream));
    cudaFree(f->ms.pu); cudaFree(f->ms.pv); cudaFree(f->ms.pr); cudaFree(f->ms.plevels); cudaFree(f->ms.choice);
    cudaFree(f->ms.cache); cudaFree(f->ms.ncand); cudaFree(f->ms.qmeta); cudaFree(f->ms.cache8);
    f->ms.qmeta = nullptr; f->ms.cache8 = nullptr;
    f->ms.pu = f->ms.pv = f->ms.pr = nullptr; f->ms.plevels = f->ms.choice = nullptr; f->ms.cache = nullptr; f->ms.ncand = nullptr;
    DVM_CUDA(cudaMalloc(&f->ms.pu, cap * sizeof(float)));
    DVM_CUDA(cudaMalloc(&f->ms.pv, cap * sizeof(float)));
    DVM_CUDA(cudaMalloc(&f->ms.pr, cap * sizeof(float)));
    DVM_CUDA(cudaMalloc(&f->ms.plevels, cap * sizeof(int)));
    DVM_CUDA(cudaMalloc(&f->ms.choice, cap * sizeof(int)));
    DVM_CUDA(cudaMalloc(&f->ms.cache, (size_t)cap * kMatchCacheK * sizeof(unsigned long long)));
    DVM_CUDA(cudaMalloc(&f->ms.ncand, cap * sizeof(int)));
    DVM_CUDA(cudaMalloc(&f->ms.qmeta, cap * sizeof(int4)));
    DVM_CUDA(cudaMalloc(&f->ms.cache8, (size_t)cap * kMatchCacheK * sizeof(unsigned)));
    f->q_cap = cap;
    return DVM_OK;
}

int dvm_frame_ensure_bytes(dvm_frame* f, size_t in_bytes, size_t out_bytes)
{
    if (in_bytes > f->in_cap) {
        DVM_CUDA(cudaStreamSynchronize(f->stream));
        cudaFree(f->d_in); f->d_in = nullptr;
        if (f->h_in) { cudaFreeHost(f->h_in); f->h_in = nullptr; }
        const size_t cap = in_bytes + in_bytes / 4 + 4096;
        DVM_CUDA(cudaMalloc(&f->d_in, cap));
        DVM_CUDA(cudaHostAlloc(&f->h_in, cap, cudaHostAllocDefault));
        f->in_cap = f->h_in_cap = cap;
    }
    if (out_bytes > f->h_out_cap) {
        DVM_CUDA(cudaStreamSynchronize(f->stream));
        if (f->h_out) { cudaFreeHost(f->h_out); f->h_out = nullptr; }
        const size_t cap = out_bytes + out_bytes / 4 + 4096;
        DVM_CUDA(cudaHostAlloc(&f->h_out, cap, cudaHostAllocDefault));
        f->h_out_cap = cap;
    }
    return DVM_OK;
}

// packs host arrays into the pinned staging buffer at 256-byte aligned offsets and returns the
// matching device pointers after one H2D copy
struct Packer {
    dvm_frame* f;
    size_t off = 0;
    explicit Packer(dvm_frame* f_) : f(f_) { }
    template <typename T>
    const T* add(const T* src, size_t count)
    {
        off = (off + 255) & ~(size_t)255;
        if (src) memcpy(f->h_in + off, src, count * sizeof(T));
        const T* d = reinterpret_cast<const T*>(f->d_in + off);
        off += count * sizeof(T);
        return d;
    }
    static size_t need(std::initializer_list<size_t> sizes)
    {
        size_t t = 0;
        for (size_t s : sizes) t = ((t + 255) & ~(size_t)255) + s;
        return t + 256;
    }
};

extern "C" {

int dvm_frame_create(dvm_frame** out, int device, void* cuda_stream, int max_keypoints, int nlevels,
                     const float* scale_factors, const float* inv_level_sigma2)
{
    DVM_REQUIRE(out != nullptr, "null output handle");
    *out = nullptr;
    // the matchers' candidate keys carry the keypoint index in 16 bits (pack_cand, track_kernels.cu)
    DVM_REQUIRE(max_keypoints > 0 && max_keypoints <= 65532 && nlevels >= 1 && nlevels <= kTrackMaxLevels,
                "max_keypoints (1..65532) / nlevels out of range");
    DVM_REQUIRE(scale_factors != nullptr && inv_level_sigma2 != nullptr, "null scale tables");
    int rc = select_device(device);
    if (rc != DVM_OK) return rc;
    dvm_frame* f = new dvm_frame;
    f->device = device;
    f->cap = (max_keypoints + 3) & ~3;
    memset(&f->ms, 0, sizeof(f->ms));
#define DVM_FCREATE(call)                                                                             \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            set_error("%s failed in dvm_frame_create: %s", #call, cudaGetErrorString(e__));           \
            frame_free(f);                                                                            \
            return DVM_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)
    if (cuda_stream) {
        f->stream = (cudaStream_t)cuda_stream;
    } else {
        DVM_FCREATE(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
        f->own_stream = true;
    }
    DVM_FCREATE(cudaEventCreateWithFlags(&f->ev, cudaEventDisableTiming));
    DVM_FCREATE(cudaMalloc(&f->d_kps, f->cap * sizeof(dvm_keypoint)));
    DVM_FCREATE(cudaMalloc(&f->d_desc, (size_t)f->cap * 32));
    DVM_FCREATE(cudaMalloc(&f->d_n, 4 * sizeof(int))); // {n, monoIndex} as the extractor writes them
    DVM_FCREATE(cudaMemset(f->d_n, 0, 4 * sizeof(int)));
    DVM_FCREATE(cudaMalloc(&f->d_cell_start, (kGridCells + 4) * sizeof(int)));
    DVM_FCREATE(cudaMemset(f->d_cell_start, 0, (kGridCells + 4) * sizeof(int)));
    DVM_FCREATE(cudaMalloc(&f->d_cell_items, f->cap * sizeof(int)));
    DVM_FCREATE(cudaMemset(f->d_cell_items, 0, f->cap * sizeof(int)));
    DVM_FCREATE(cudaMalloc(&f->d_kxyo, (size_t)f->cap * 4 * sizeof(int)));
    DVM_FCREATE(cudaMemset(f->d_kxyo, 0, (size_t)f->cap * 4 * sizeof(int)));
    DVM_FCREATE(cudaMalloc(&f->ms.claim_a, f->cap * sizeof(int)));
    DVM_FCREATE(cudaMalloc(&f->ms.claim_b, f->cap * sizeof(int)));
    DVM_FCREATE(cudaMalloc(&f->ms.iters, 4 * sizeof(int)));
    DVM_FCREATE(cudaMemset(f->ms.iters, 0, 4 * sizeof(int)));
    f->ms.qcount = f->ms.iters + 1;
    DVM_FCREATE(cudaMalloc(&f->d_cur_mp, (f->cap + 8) * sizeof(int)));
#undef DVM_FCREATE
    FrameDev& d = f->dev;
    memset(&d, 0, sizeof(d));
    d.kps = f->d_kps; d.desc = f->d_desc; d.n = f->d_n; d.cap = f->cap;
    d.cell_start = f->d_cell_start; d.cell_items = f->d_cell_items; d.cell_rec = reinterpret_cast<int4*>(f->d_kxyo);
    d.nlevels = nlevels;
    for (int i = 0; i < nlevels; i++) { d.scale[i] = scale_factors[i]; d.inv_sigma2[i] = inv_level_sigma2[i]; }
    *out = f;
    return DVM_OK;
}

void dvm_frame_destroy(dvm_frame* f) { frame_free(f); }

static void set_bounds(dvm_frame* f, float min_x, float min_y, float max_x, float max_y)
{
    FrameDev& d = f->dev;
    d.minX = min_x; d.minY = min_y; d.maxX = max_x; d.maxY = max_y;
    // mfGridElementWidthInv = FRAME_GRID_COLS / (mnMaxX - mnMinX), O3/src/Frame.cc:446-447
    d.gwInv = static_cast<float>(kGridCols) / static_cast<float>(max_x - min_x);
    d.ghInv = static_cast<float>(kGridRows) / static_cast<float>(max_y - min_y);
}

int dvm_frame_assign(dvm_frame* f, const dvm_keypoint* kps_un, const uint8_t* desc, int n, float min_x, float min_y,
                     float max_x, float max_y)
{
    DVM_REQUIRE(f != nullptr && n >= 0 && n <= f->cap, "bad frame / too many keypoints");
    DVM_REQUIRE(n == 0 || (kps_un != nullptr && desc != nullptr), "null keypoints");
    DVM_REQUIRE(max_x > min_x && max_y > min_y, "empty image bounds");
    DVM_CUDA(cudaSetDevice(f->device));
    set_bounds(f, min_x, min_y, max_x, max_y);
    DVM_CUDA(cudaMemcpyAsync(f->d_kps, kps_un, (size_t)n * sizeof(dvm_keypoint), cudaMemcpyHostToDevice, f->stream));
    DVM_CUDA(cudaMemcpyAsync(f->d_desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, f->stream));
    f->host_n = n;
    DVM_CUDA(cudaMemcpyAsync(f->d_n, &f->host_n, sizeof(int), cudaMemcpyHostToDevice, f->stream));
    launch_grid_build(f->dev, f->stream);
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaStreamSynchronize(f->stream));
    return DVM_OK;
}

int dvm_frame_assign_from_orb(dvm_frame* f, const dvm_orb* orb, float min_x, float min_y, float max_x, float max_y)
{
    DVM_REQUIRE(f != nullptr && orb != nullptr, "null argument");
    DVM_REQUIRE(dvm_orb_max_keypoints(orb) <= f->cap, "frame capacity below the extractor's maximum");
    DVM_REQUIRE(max_x > min_x && max_y > min_y, "empty image bounds");
    DVM_CUDA(cudaSetDevice(f->device));
    set_bounds(f, min_x, min_y, max_x, max_y);
    const dvm_keypoint* k = nullptr;
    const uint8_t* d = nullptr;
    const int32_t* c = nullptr;
    dvm_orb_result_device(orb, &k, &d, &c);
    cudaStream_t os = (cudaStream_t)dvm_orb_stream(orb);
    if (os != f->stream) {
        DVM_CUDA(cudaEventRecord(f->ev, os));
        DVM_CUDA(cudaStreamWaitEvent(f->stream, f->ev, 0));
    }
    f->host_n = dvm_orb_max_keypoints(orb);
    launch_frame_assign(f->dev, k, d, c, f->stream); // copy + AssignFeaturesToGrid in one launch
    DVM_CUDA(cudaGetLastError());
    return DVM_OK;
}

int dvm_frame_construct_device(dvm_frame* f, dvm_orb* orb, const uint8_t* gray_dev, int width, int height, int stride,
                               float min_x, float min_y, float max_x, float max_y)
{
    DVM_REQUIRE(f != nullptr && orb != nullptr && gray_dev != nullptr, "null argument");
    DVM_REQUIRE(dvm_orb_max_keypoints(orb) <= f->cap, "frame capacity below the extractor's maximum");
    DVM_REQUIRE(max_x > min_x && max_y > min_y, "empty image bounds");
    DVM_CUDA(cudaSetDevice(f->device));
    set_bounds(f, min_x, min_y, max_x, max_y);
    int rc = dvm_orb_extract_device_to(orb, gray_dev, width, height, stride, 0, 1000, f->d_kps, f->d_desc, f->d_n);
    if (rc != DVM_OK) return rc;
    f->host_n = dvm_orb_max_keypoints_current(orb);
    f->dev.cap = f->host_n; // tight bound for this image size: sizes the matchers' shared-memory staging
    if (f->undistort) launch_undistort(f->d_kps, f->d_n, f->host_n, f->und, (cudaStream_t)dvm_orb_stream(orb));
    launch_grid_build(f->dev, (cudaStream_t)dvm_orb_stream(orb));
    DVM_CUDA(cudaGetLastError());
    return DVM_OK;
}

static void fill_undistort(UndistortArgs& u, const float* K, const float* dist5)
{
    for (int i = 0; i < 5; i++) u.k[i] = (double)dist5[i];
    u.fx = K[0]; u.fy = K[1]; u.cx = K[2]; u.cy = K[3];
}

int dvm_frame_set_distortion(dvm_frame* f, const float* K, const float* dist5)
{
    DVM_REQUIRE(f != nullptr, "null frame");
    // Frame::UndistortKeyPoints is the identity when mDistCoef.at<float>(0) == 0 (O3/src/Frame.cc:792-795)
    f->undistort = K != nullptr && dist5 != nullptr && dist5[0] != 0.0f;
    if (f->undistort) fill_undistort(f->und, K, dist5);
    return DVM_OK;
}

int dvm_undistort_keypoints(dvm_frame* ctx, dvm_keypoint* kps, int n, const float* K, const float* dist5)
{
    DVM_REQUIRE(ctx != nullptr && n >= 0 && K && dist5, "bad argument");
    DVM_REQUIRE(n == 0 || kps, "null keypoints");
    if (n == 0 || dist5[0] == 0.0f) return DVM_OK;   // identity, as the reference
    DVM_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)n * sizeof(dvm_keypoint);
    int rc = dvm_frame_ensure_bytes(ctx, bytes + 512, bytes);
    if (rc != DVM_OK) return rc;
    memcpy(ctx->h_in, kps, bytes);
    int* hn = reinterpret_cast<int*>(ctx->h_in + ((bytes + 255) & ~(size_t)255));
    *hn = n;
    DVM_CUDA(cudaMemcpyAsync(ctx->d_in, ctx->h_in, ((bytes + 255) & ~(size_t)255) + 4, cudaMemcpyHostToDevice, ctx->stream));
    UndistortArgs u;
    fill_undistort(u, K, dist5);
    launch_undistort(reinterpret_cast<dvm_keypoint*>(ctx->d_in), reinterpret_cast<const int*>(ctx->d_in + ((bytes + 255) & ~(size_t)255)),
                     n, u, ctx->stream);
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaMemcpyAsync(ctx->h_out, ctx->d_in, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DVM_CUDA(cudaStreamSynchronize(ctx->stream));
    memcpy(kps, ctx->h_out, bytes);
    return DVM_OK;
}

int dvm_image_bounds(dvm_frame* ctx, const float* K, const float* dist5, int width, int height, float* bounds)
{
    DVM_REQUIRE(ctx != nullptr && K && dist5 && bounds && width > 0 && height > 0, "bad argument");
    if (dist5[0] == 0.0f) { bounds[0] = 0.f; bounds[1] = 0.f; bounds[2] = (float)width; bounds[3] = (float)height; return DVM_OK; }
    dvm_keypoint c[4];
    memset(c, 0, sizeof(c));
    c[0].x = 0.f; c[0].y = 0.f; c[1].x = (float)width; c[1].y = 0.f; c[2].x = 0.f; c[2].y = (float)height;
    c[3].x = (float)width; c[3].y = (float)height;
    const int rc = dvm_undistort_keypoints(ctx, c, 4, K, dist5);
    if (rc != DVM_OK) return rc;
    bounds[0] = fminf(c[0].x, c[2].x); bounds[2] = fmaxf(c[1].x, c[3].x);    // O3/src/Frame.cc:838-841
    bounds[1] = fminf(c[0].y, c[1].y); bounds[3] = fmaxf(c[2].y, c[3].y);
    return DVM_OK;
}

int dvm_frame_features_in_area(dvm_frame* f, float x, float y, float r, int min_level, int max_level, int32_t* out,
                               int cap, int* n_out)
{
    DVM_REQUIRE(f != nullptr && n_out != nullptr && cap >= 0, "bad argument");
    DVM_CUDA(cudaSetDevice(f->device));
    int rc = dvm_frame_ensure_bytes(f, 0, (size_t)(f->cap + 8) * sizeof(int));
    if (rc != DVM_OK) return rc;
    int* d_out = f->d_cur_mp;
    launch_features_in_area(f->dev, x, y, r, min_level, max_level, d_out, f->cap, d_out + f->cap, f->stream);
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaMemcpyAsync(f->h_out, d_out, (size_t)(f->cap + 1) * sizeof(int), cudaMemcpyDeviceToHost, f->stream));
    DVM_CUDA(cudaStreamSynchronize(f->stream));
    const int* h = (const int*)f->h_out;
    const int n = h[f->cap];
    *n_out = n;
    for (int i = 0; i < n && i < cap; i++) out[i] = h[i];
    return DVM_OK;
}

int dvm_frame_grid_cell(dvm_frame* f, int ix, int iy, int32_t* out, int cap, int* n_out)
{
    DVM_REQUIRE(f != nullptr && n_out != nullptr && ix >= 0 && ix < kGridCols && iy >= 0 && iy < kGridRows, "bad argument");
    DVM_CUDA(cudaSetDevice(f->device));
    DVM_CUDA(cudaStreamSynchronize(f->stream));
    int se[2];
    DVM_CUDA(cudaMemcpy(se, f->d_cell_start + ix * kGridRows + iy, 2 * sizeof(int), cudaMemcpyDeviceToHost));
    const int n = se[1] - se[0];
    *n_out = n;
    std::vector<int> tmp(n > 0 ? n : 1);
    if (n > 0) DVM_CUDA(cudaMemcpy(tmp.data(), f->d_cell_items + se[0], n * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap; i++) out[i] = tmp[i];
    return DVM_OK;
}

} // extern "C"

int dvm_frame_finish_match(dvm_frame* f, int32_t* cur_mp, int* nmatches)
{
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaMemcpyAsync(f->d_cur_mp + f->cap + 1, f->ms.iters, sizeof(int), cudaMemcpyDeviceToDevice, f->stream));
    DVM_CUDA(cudaMemcpyAsync(f->d_cur_mp + f->cap + 2, f->d_n, sizeof(int), cudaMemcpyDeviceToDevice, f->stream));
    DVM_CUDA(cudaMemcpyAsync(f->h_out, f->d_cur_mp, (size_t)(f->cap + 3) * sizeof(int), cudaMemcpyDeviceToHost, f->stream));
    DVM_CUDA(cudaStreamSynchronize(f->stream));
    const int* h = (const int*)f->h_out;
    f->host_n = h[f->cap + 2] < f->cap ? h[f->cap + 2] : f->cap;
    memcpy(cur_mp, h, (size_t)f->host_n * sizeof(int));
    *nmatches = h[f->cap];
    f->last_rounds = h[f->cap + 1];
    return DVM_OK;
}

extern "C" {

int dvm_match_by_projection_last(dvm_frame* cur, const float* qcw, const float* tcw, const float* K, int last_n,
                                 const uint8_t* has_mp, const uint8_t* outlier, const float* Xw, const uint8_t* mp_desc,
                                 const uint8_t* mp_obs_pos, const int32_t* last_octave, const float* last_angle,
                                 float th, int check_orientation, int32_t* cur_mp, int* nmatches)
{
    DVM_REQUIRE(cur != nullptr && qcw && tcw && K && cur_mp && nmatches, "null argument");
    DVM_REQUIRE(last_n >= 0, "negative count");
    DVM_REQUIRE(last_n == 0 || (has_mp && outlier && Xw && mp_desc && mp_obs_pos && last_octave && last_angle), "null last-frame arrays");
    DVM_CUDA(cudaSetDevice(cur->device));
    const size_t n = (size_t)last_n;
    int rc = dvm_frame_ensure_bytes(cur, Packer::need({ n, n, n * 12, n * 32, n, n * 4, n * 4 }), (size_t)(cur->cap + 8) * sizeof(int));
    if (rc != DVM_OK) return rc;
    rc = dvm_frame_ensure_query_cap(cur, last_n);
    if (rc != DVM_OK) return rc;
    MatchLastArgs a;
    memset(&a, 0, sizeof(a));
    memcpy(a.q, qcw, sizeof(a.q)); memcpy(a.t, tcw, sizeof(a.t)); memcpy(a.K, K, sizeof(a.K));
    a.last_n = last_n; a.th = th; a.check_ori = check_orientation;
    Packer p(cur);
    a.has_mp = p.add(has_mp, n);
    a.outlier = p.add(outlier, n);
    a.Xw = p.add(Xw, n * 3);
    a.mp_desc = p.add(mp_desc, n * 32);
    a.obs_pos = p.add(mp_obs_pos, n);
    a.octave = p.add(last_octave, n);
    a.angle = p.add(last_angle, n);
    if (last_n > 0) {
        for (int i = 0; i < last_n; i++)
            DVM_REQUIRE(!has_mp[i] || (last_octave[i] >= 0 && last_octave[i] < cur->dev.nlevels), "octave out of range");
    }
    DVM_CUDA(cudaMemcpyAsync(cur->d_in, cur->h_in, p.off, cudaMemcpyHostToDevice, cur->stream));
    launch_match_last(cur->dev, a, cur->ms, cur->d_cur_mp, cur->d_cur_mp + cur->cap, cur->stream);
    return dvm_frame_finish_match(cur, cur_mp, nmatches);
}

int dvm_match_by_projection_map(dvm_frame* cur, int m, const float* proj_x, const float* proj_y, const int32_t* level,
                                const float* view_cos, const uint8_t* mp_desc, const uint8_t* mp_obs_pos, float th,
                                float nnratio, const uint8_t* cur_blocked, int32_t* cur_mp, int* nmatches)
{
    DVM_REQUIRE(cur != nullptr && cur_mp && nmatches, "null argument");
    DVM_REQUIRE(m >= 0, "negative count");
    DVM_REQUIRE(m == 0 || (proj_x && proj_y && level && view_cos && mp_desc && mp_obs_pos), "null map-point arrays");
    DVM_CUDA(cudaSetDevice(cur->device));
    const size_t n = (size_t)m, nc = (size_t)cur->host_n;
    int rc = dvm_frame_ensure_bytes(cur, Packer::need({ n * 4, n * 4, n * 4, n * 4, n * 32, n, nc }), (size_t)(cur->cap + 8) * sizeof(int));
    if (rc != DVM_OK) return rc;
    rc = dvm_frame_ensure_query_cap(cur, m);
    if (rc != DVM_OK) return rc;
    for (int i = 0; i < m; i++) DVM_REQUIRE(level[i] >= 0 && level[i] < cur->dev.nlevels, "predicted level out of range");
    MatchMapArgs a;
    memset(&a, 0, sizeof(a));
    a.m = m; a.th = th; a.nnratio = nnratio;
    Packer p(cur);
    a.projX = p.add(proj_x, n);
    a.projY = p.add(proj_y, n);
    a.level = p.add(level, n);
    a.view_cos = p.add(view_cos, n);
    a.mp_desc = p.add(mp_desc, n * 32);
    a.obs_pos = p.add(mp_obs_pos, n);
    a.cur_blocked = cur_blocked ? p.add(cur_blocked, nc) : nullptr;
    DVM_CUDA(cudaMemcpyAsync(cur->d_in, cur->h_in, p.off, cudaMemcpyHostToDevice, cur->stream));
    launch_match_map(cur->dev, a, cur->ms, cur->d_cur_mp, cur->d_cur_mp + cur->cap, cur->stream);
    return dvm_frame_finish_match(cur, cur_mp, nmatches);
}

int dvm_match_last_rounds(dvm_frame* cur) { return cur ? cur->last_rounds : DVM_ERR_INVALID; }

int dvm_pose_optimization(dvm_frame* ctx, float* pose_q, float* pose_t, const float* K, int n, const float* Xw,
                          const float* kp_xy, const float* inv_sigma2, uint8_t* outlier, int* n_inliers, int* stats)
{
    DVM_REQUIRE(ctx != nullptr && pose_q && pose_t && K && n_inliers, "null argument");
    DVM_REQUIRE(n >= 0, "negative count");
    DVM_REQUIRE(n == 0 || (Xw && kp_xy && inv_sigma2 && outlier), "null correspondence arrays");
    DVM_CUDA(cudaSetDevice(ctx->device));
    const size_t sn = (size_t)n;
    // device layout: inputs | pose[7] | result[4] | outlier[n]
    int rc = dvm_frame_ensure_bytes(ctx, Packer::need({ sn * 12, sn * 8, sn * 4, 7 * 4, 4 * 4, sn }), sn + 2048);
    if (rc != DVM_OK) return rc;
    if (n > ctx->err_cap) {
        DVM_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_err); ctx->d_err = nullptr;
        const int cap = n + n / 4 + 256;
        DVM_CUDA(cudaMalloc(&ctx->d_err, (size_t)cap * 2 * sizeof(double)));
        ctx->err_cap = cap;
    }
    PoseOptArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    memcpy(a.K, K, sizeof(a.K));
    Packer p(ctx);
    a.Xw = p.add(Xw, sn * 3);
    a.kp_xy = p.add(kp_xy, sn * 2);
    a.inv_sigma2 = p.add(inv_sigma2, sn);
    float pose[7] = { pose_q[0], pose_q[1], pose_q[2], pose_q[3], pose_t[0], pose_t[1], pose_t[2] };
    const size_t out_begin = (p.off + 255) & ~(size_t)255;
    a.pose = const_cast<float*>(p.add(pose, 7));
    a.result = const_cast<int*>(p.add((const int*)nullptr, 4));
    a.outlier = const_cast<uint8_t*>(p.add((const uint8_t*)nullptr, sn));
    a.valid = nullptr;
    a.err = ctx->d_err;
    DVM_CUDA(cudaMemcpyAsync(ctx->d_in, ctx->h_in, p.off, cudaMemcpyHostToDevice, ctx->stream));
    { const int prc = launch_pose_opt(a, ctx->stream); if (prc != DVM_OK) return prc; }
    DVM_CUDA(cudaGetLastError());
    const size_t out_bytes = p.off - out_begin;
    rc = dvm_frame_ensure_bytes(ctx, 0, out_bytes);
    if (rc != DVM_OK) return rc;
    DVM_CUDA(cudaMemcpyAsync(ctx->h_out, ctx->d_in + out_begin, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DVM_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint8_t* h = ctx->h_out;
    const float* hp = (const float*)(h + ((const uint8_t*)a.pose - (ctx->d_in + out_begin)));
    const int* hr = (const int*)(h + ((const uint8_t*)a.result - (ctx->d_in + out_begin)));
    const uint8_t* ho = h + (a.outlier - (ctx->d_in + out_begin));
    for (int i = 0; i < 4; i++) pose_q[i] = hp[i];
    for (int i = 0; i < 3; i++) pose_t[i] = hp[4 + i];
    if (n > 0) memcpy(outlier, ho, sn);
    *n_inliers = hr[0];
    if (stats) { stats[0] = hr[2]; stats[1] = hr[3]; }
    return DVM_OK;
}

} // extern "C"
