// bow_capi.cu -- C-ABI of SearchByBoW, SearchForInitialization and the exhaustive Hamming search.
// Host-array entry points stage their flat inputs in the frame's pinned buffer, upload them with one
// copy, run the kernels on the frame's stream and read the results back (synchronous, like the
// reference's calls).
#include "sophus_f32.cuh"
#include "glibc_logf.h"
#include "bow_kernels.cuh"
#include "track_internal.cuh"
#include <cmath>
#include <vector>

using namespace dvm;

namespace {

// packs host arrays into the frame's pinned staging buffer at 256-byte aligned offsets
struct Stage {
    dvm_frame* f;
    size_t off = 0;
    explicit Stage(dvm_frame* f_) : f(f_) { }
    template <typename T>
    T* add(const T* src, size_t count)
    {
        off = (off + 255) & ~(size_t)255;
        if (src && count) memcpy(f->h_in + off, src, count * sizeof(T));
        T* d = reinterpret_cast<T*>(f->d_in + off);
        off += count * sizeof(T);
        return d;
    }
};

size_t padded(std::initializer_list<size_t> sizes)
{
    size_t t = 0;
    for (size_t s : sizes) t = ((t + 255) & ~(size_t)255) + s;
    return t + 256;
}

int check_side(const dvm_bow_features* s, std::vector<uint8_t>& seen)
{
    DVM_REQUIRE(s != nullptr && s->n >= 0 && s->n_nodes >= 0, "bad feature set");
    DVM_REQUIRE(s->n == 0 || (s->desc && s->angle), "null descriptors / angles");
    DVM_REQUIRE(s->n_nodes == 0 || (s->node_id && s->node_start && s->feat_idx), "null feature vector");
    seen.assign((size_t)s->n, 0);
    for (int i = 0; i < s->n_nodes; i++) {
        DVM_REQUIRE(i == 0 || s->node_id[i - 1] < s->node_id[i], "feature-vector nodes must ascend");
        DVM_REQUIRE(s->node_start[i] <= s->node_start[i + 1] && s->node_start[i] >= 0, "node_start must not decrease");
    }
    const int total = s->n_nodes ? s->node_start[s->n_nodes] : 0;
    for (int p = s->n_nodes ? s->node_start[0] : 0; p < total; p++) {
        const uint32_t r = s->feat_idx[p];
        DVM_REQUIRE(r < (uint32_t)s->n, "feature index out of range");
        DVM_REQUIRE(!seen[r], "a feature belongs to one vocabulary node only");
        seen[r] = 1;
    }
    return DVM_OK;
}

} // namespace

struct dvm_hamming {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    KnnScratch scratch;
    int mode = 0;   // dvm_hamming_set_mode
    uint8_t* d_buf = nullptr; size_t d_cap = 0;   // host-call staging: a | b | key1 | key2
    uint8_t* h_buf = nullptr; size_t h_cap = 0;
};

extern "C" {

int dvm_match_by_bow(dvm_frame* ctx, int kf_kf, const dvm_bow_features* a, const dvm_bow_features* b, float nnratio,
                     int check_orientation, int32_t* match12, int32_t* match21, int* nmatches)
{
    DVM_REQUIRE(ctx != nullptr && nmatches != nullptr, "null argument");
    std::vector<uint8_t> seen;
    int rc = check_side(a, seen);
    if (rc != DVM_OK) return rc;
    rc = check_side(b, seen);
    if (rc != DVM_OK) return rc;
    DVM_CUDA(cudaSetDevice(ctx->device));
    const dvm_bow_features* S[2] = { a, b };
    size_t in_bytes = 0;
    for (int k = 0; k < 2; k++) {
        const size_t n = (size_t)S[k]->n, nn = (size_t)S[k]->n_nodes, tot = nn ? (size_t)S[k]->node_start[nn] : 0;
        in_bytes += padded({ n * 32, n * 4, n, nn * 4, (nn + 1) * 4, tot * 4 });
    }
    const size_t out_ints = (size_t)a->n + (size_t)b->n + kHistoLength + 2;
    in_bytes += padded({ out_ints * 4 });
    rc = dvm_frame_ensure_bytes(ctx, in_bytes, out_ints * 4 + 256);
    if (rc != DVM_OK) return rc;
    Stage st(ctx);
    BowArgs g;
    memset(&g, 0, sizeof(g));
    BowSide* D[2] = { &g.a, &g.b };
    for (int k = 0; k < 2; k++) {
        const size_t n = (size_t)S[k]->n, nn = (size_t)S[k]->n_nodes, tot = nn ? (size_t)S[k]->node_start[nn] : 0;
        D[k]->n = S[k]->n; D[k]->n_nodes = S[k]->n_nodes;
        D[k]->desc = st.add(S[k]->desc, n * 32);
        D[k]->angle = st.add(S[k]->angle, n);
        D[k]->valid = S[k]->has_mp ? st.add(S[k]->has_mp, n) : nullptr;
        D[k]->node_id = st.add(S[k]->node_id, nn);
        D[k]->node_start = st.add(S[k]->node_start, nn + 1);
        D[k]->feat_idx = st.add(S[k]->feat_idx, tot);
    }
    if (!kf_kf) g.b.valid = nullptr;
    g.kf_kf = kf_kf ? 1 : 0; g.nnratio = nnratio; g.check_ori = check_orientation;
    const size_t in_end = st.off;
    int* d_out = st.add((const int*)nullptr, out_ints);   // match12 | match21 | histo | counters
    g.match12 = d_out; g.match21 = d_out + a->n; g.histo = g.match21 + b->n; g.counters = g.histo + kHistoLength;
    DVM_CUDA(cudaMemcpyAsync(ctx->d_in, ctx->h_in, in_end, cudaMemcpyHostToDevice, ctx->stream));
    DVM_CUDA(cudaMemsetAsync(d_out, 0xff, ((size_t)a->n + (size_t)b->n) * 4, ctx->stream));
    DVM_CUDA(cudaMemsetAsync(g.histo, 0, (kHistoLength + 2) * 4, ctx->stream));
    launch_bow_match(g, ctx->stream);
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaMemcpyAsync(ctx->h_out, d_out, out_ints * 4, cudaMemcpyDeviceToHost, ctx->stream));
    DVM_CUDA(cudaStreamSynchronize(ctx->stream));
    const int* h = (const int*)ctx->h_out;
    if (match12 && a->n) memcpy(match12, h, (size_t)a->n * 4);
    if (match21 && b->n) memcpy(match21, h + a->n, (size_t)b->n * 4);
    *nmatches = h[(size_t)a->n + b->n + kHistoLength + 1];
    return DVM_OK;
}

int dvm_match_for_initialization(dvm_frame* f2, int n1, const dvm_keypoint* kps1_un, const uint8_t* desc1,
                                 float* prev_matched, int window_size, float nnratio, int check_orientation,
                                 int32_t* matches12, int* nmatches)
{
    DVM_REQUIRE(f2 != nullptr && nmatches != nullptr && n1 >= 0 && window_size >= 0, "bad argument");
    DVM_REQUIRE(n1 == 0 || (kps1_un && desc1 && prev_matched && matches12), "null first-frame arrays");
    DVM_CUDA(cudaSetDevice(f2->device));
    const size_t n = (size_t)n1;
    const size_t out_bytes = n * 8 + n * 4 + 8;   // prev_matched | matches12 | result
    int rc = dvm_frame_ensure_bytes(f2, padded({ n * sizeof(dvm_keypoint), n * 32, out_bytes + 512, n * 4, n * 4, n * 4, n * 4, n * 4,
                                                 (size_t)f2->cap * 4, n * 64, n * 4 }), out_bytes + 512);
    if (rc != DVM_OK) return rc;
    Stage st(f2);
    InitMatchArgs a;
    memset(&a, 0, sizeof(a));
    a.n1 = n1; a.window = window_size; a.nnratio = nnratio; a.check_ori = check_orientation;
    a.kps1 = st.add(kps1_un, n);
    a.desc1 = st.add(desc1, n * 32);
    const size_t out_begin = (st.off + 255) & ~(size_t)255;
    a.prev_matched = st.add(prev_matched, n * 2);
    const size_t in_end = st.off;
    a.matches12 = st.add((const int*)nullptr, n);
    a.result = st.add((const int*)nullptr, 2);
    const size_t out_end = st.off;
    a.choice_a = st.add((const int*)nullptr, n); a.choice_b = st.add((const int*)nullptr, n);
    a.cdist_a = st.add((const int*)nullptr, n); a.cdist_b = st.add((const int*)nullptr, n);
    a.next = st.add((const int*)nullptr, n);
    a.head = st.add((const int*)nullptr, (size_t)f2->cap);
    a.cache = st.add((const unsigned long long*)nullptr, n * 8);
    a.ncand = st.add((const int*)nullptr, n);
    DVM_CUDA(cudaMemcpyAsync(f2->d_in, f2->h_in, in_end, cudaMemcpyHostToDevice, f2->stream));
    launch_init_match(f2->dev, a, f2->stream);
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaMemcpyAsync(f2->h_out, f2->d_in + out_begin, out_end - out_begin, cudaMemcpyDeviceToHost, f2->stream));
    DVM_CUDA(cudaStreamSynchronize(f2->stream));
    const uint8_t* h = f2->h_out;
    auto at = [&](const void* dptr) { return h + ((const uint8_t*)dptr - (f2->d_in + out_begin)); };
    if (n1) {
        memcpy(prev_matched, at(a.prev_matched), n * 8);
        memcpy(matches12, at(a.matches12), n * 4);
    }
    const int* r = (const int*)at(a.result);
    *nmatches = r[0];
    f2->last_rounds = r[1];
    return DVM_OK;
}

int dvm_match_for_triangulation(dvm_frame* ctx, const dvm_bow_features* kf1, const dvm_keypoint* kps1,
                                const dvm_bow_features* kf2, const dvm_keypoint* kps2, const float* F12, const float* ep,
                                const float* scale_factors2, const float* level_sigma2_2, int nlevels, int coarse,
                                int check_orientation, int32_t* matches12, int* nmatches)
{
    DVM_REQUIRE(ctx != nullptr && nmatches != nullptr && F12 && ep && scale_factors2 && level_sigma2_2, "null argument");
    DVM_REQUIRE(nlevels >= 1 && nlevels <= kTrackMaxLevels, "nlevels out of range");
    std::vector<uint8_t> seen;
    int rc = check_side(kf1, seen);
    if (rc != DVM_OK) return rc;
    rc = check_side(kf2, seen);
    if (rc != DVM_OK) return rc;
    DVM_REQUIRE((kf1->n == 0 || (kps1 && kf1->has_mp && matches12)) && (kf2->n == 0 || (kps2 && kf2->has_mp)), "null keypoints / has_mp");
    for (int i = 0; i < kf2->n; i++) DVM_REQUIRE(kps2[i].octave >= 0 && kps2[i].octave < nlevels, "octave out of range");
    DVM_CUDA(cudaSetDevice(ctx->device));
    const dvm_bow_features* S[2] = { kf1, kf2 };
    const dvm_keypoint* KP[2] = { kps1, kps2 };
    size_t in_bytes = 0;
    for (int k = 0; k < 2; k++) {
        const size_t n = (size_t)S[k]->n, nn = (size_t)S[k]->n_nodes, tot = nn ? (size_t)S[k]->node_start[nn] : 0;
        in_bytes += padded({ n * 32, n, nn * 4, (nn + 1) * 4, tot * 4, n * sizeof(dvm_keypoint) });
    }
    const size_t out_ints = (size_t)kf1->n + kHistoLength + 2;
    in_bytes += padded({ out_ints * 4 });
    rc = dvm_frame_ensure_bytes(ctx, in_bytes, out_ints * 4 + 256);
    if (rc != DVM_OK) return rc;
    Stage st(ctx);
    TriArgs g;
    memset(&g, 0, sizeof(g));
    BowSide* D[2] = { &g.a, &g.b };
    const dvm_keypoint* dk[2];
    for (int k = 0; k < 2; k++) {
        const size_t n = (size_t)S[k]->n, nn = (size_t)S[k]->n_nodes, tot = nn ? (size_t)S[k]->node_start[nn] : 0;
        D[k]->n = S[k]->n; D[k]->n_nodes = S[k]->n_nodes;
        D[k]->desc = st.add(S[k]->desc, n * 32);
        D[k]->angle = nullptr;
        D[k]->valid = st.add(S[k]->has_mp, n);
        D[k]->node_id = st.add(S[k]->node_id, nn);
        D[k]->node_start = st.add(S[k]->node_start, nn + 1);
        D[k]->feat_idx = st.add(S[k]->feat_idx, tot);
        dk[k] = st.add(KP[k], n);
    }
    g.kps1 = dk[0]; g.kps2 = dk[1];
    memcpy(g.F12, F12, sizeof(g.F12)); g.ep[0] = ep[0]; g.ep[1] = ep[1];
    for (int l = 0; l < nlevels; l++) { g.scale2[l] = scale_factors2[l]; g.sigma2_2[l] = level_sigma2_2[l]; }
    g.coarse = coarse ? 1 : 0; g.check_ori = check_orientation;
    const size_t in_end = st.off;
    int* d_out = st.add((const int*)nullptr, out_ints);   // matches12 | histo | counters
    g.matches12 = d_out; g.histo = d_out + kf1->n; g.counters = g.histo + kHistoLength;
    DVM_CUDA(cudaMemcpyAsync(ctx->d_in, ctx->h_in, in_end, cudaMemcpyHostToDevice, ctx->stream));
    DVM_CUDA(cudaMemsetAsync(d_out, 0xff, (size_t)kf1->n * 4, ctx->stream));
    DVM_CUDA(cudaMemsetAsync(g.histo, 0, (kHistoLength + 2) * 4, ctx->stream));
    launch_triangulation_match(g, ctx->stream);
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaMemcpyAsync(ctx->h_out, d_out, out_ints * 4, cudaMemcpyDeviceToHost, ctx->stream));
    DVM_CUDA(cudaStreamSynchronize(ctx->stream));
    const int* h = (const int*)ctx->h_out;
    if (kf1->n) memcpy(matches12, h, (size_t)kf1->n * 4);
    *nmatches = h[(size_t)kf1->n + kHistoLength + 1];
    return DVM_OK;
}

// stages the candidate arrays, runs fuse_search_kernel with the preset flags of `a` and reads best_idx / best_dist back
static int run_proj_search(dvm_frame* kf, FuseArgs a, int m, const float* xw, const float* normal, const float* min_dist,
                           const float* max_dist, const uint8_t* mp_desc, const uint8_t* skip, float th, int32_t* best_idx,
                           int32_t* best_dist)
{
    DVM_CUDA(cudaSetDevice(kf->device));
    const size_t n = (size_t)m;
    int rc = dvm_frame_ensure_bytes(kf, padded({ n * 12, n * 12, n * 4, n * 4, n * 32, n, n * 4, n * 4 }), n * 8 + 512);
    if (rc != DVM_OK) return rc;
    Stage st(kf);
    a.nlevels = kf->dev.nlevels;
    // mfLogScaleFactor = log(mfScaleFactor) on floats (O3/src/Frame.cc:401, copied into the KeyFrame)
    a.logScale = dvm_glibc_logf(kf->dev.nlevels > 1 ? kf->dev.scale[1] : 1.2f);
    for (int l = 0; l < kf->dev.nlevels; l++) a.inv_sigma2[l] = kf->dev.inv_sigma2[l];
    a.m = m; a.th = th;
    a.xw = st.add(xw, n * 3);
    a.normal = normal ? st.add(normal, n * 3) : nullptr;
    a.min_dist = st.add(min_dist, n); a.max_dist = st.add(max_dist, n);
    a.mp_desc = st.add(mp_desc, n * 32);
    a.skip = skip ? st.add(skip, n) : nullptr;
    const size_t in_end = st.off;
    const size_t out_begin = (st.off + 255) & ~(size_t)255;
    a.best_idx = st.add((const int*)nullptr, n);
    a.best_dist = st.add((const int*)nullptr, n);
    const size_t out_end = st.off;
    DVM_CUDA(cudaMemcpyAsync(kf->d_in, kf->h_in, in_end, cudaMemcpyHostToDevice, kf->stream));
    launch_fuse_search(kf->dev, a, kf->stream);
    DVM_CUDA(cudaGetLastError());
    rc = dvm_frame_ensure_bytes(kf, 0, out_end - out_begin);
    if (rc != DVM_OK) return rc;
    DVM_CUDA(cudaMemcpyAsync(kf->h_out, kf->d_in + out_begin, out_end - out_begin, cudaMemcpyDeviceToHost, kf->stream));
    DVM_CUDA(cudaStreamSynchronize(kf->stream));
    memcpy(best_idx, kf->h_out + ((const uint8_t*)a.best_idx - (kf->d_in + out_begin)), n * 4);
    if (best_dist) memcpy(best_dist, kf->h_out + ((const uint8_t*)a.best_dist - (kf->d_in + out_begin)), n * 4);
    return DVM_OK;
}

int dvm_fuse_search(dvm_frame* kf, const float* pose_q, const float* pose_t, const float* K, int m, const float* xw,
                    const float* normal, const float* min_dist, const float* max_dist, const uint8_t* mp_desc,
                    const uint8_t* skip, float th, int32_t* best_idx, int32_t* best_dist)
{
    DVM_REQUIRE(kf != nullptr && pose_q && pose_t && K && m >= 0, "bad argument");
    if (m == 0) return DVM_OK;
    DVM_REQUIRE(xw && normal && min_dist && max_dist && mp_desc && best_idx && best_dist, "null map-point arrays");
    FuseArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 4; i++) { a.q[i] = pose_q[i]; a.K[i] = K[i]; }
    for (int i = 0; i < 3; i++) a.t[i] = pose_t[i];
    a.check_normal = 1; a.check_chi = 1; a.accept_th = kThLow;
    return run_proj_search(kf, a, m, xw, normal, min_dist, max_dist, mp_desc, skip, th, best_idx, best_dist);
}

int dvm_fuse_search_sim3(dvm_frame* kf, const float* sim3_q, const float* sim3_t, const float* K, int m, const float* xw,
                         const float* normal, const float* min_dist, const float* max_dist, const uint8_t* mp_desc,
                         const uint8_t* skip, float th, int32_t* best_idx, int32_t* best_dist)
{
    DVM_REQUIRE(kf != nullptr && sim3_q && sim3_t && K && m >= 0, "bad argument");
    if (m == 0) return DVM_OK;
    DVM_REQUIRE(xw && normal && min_dist && max_dist && mp_desc && best_idx && best_dist, "null map-point arrays");
    FuseArgs a;
    memset(&a, 0, sizeof(a));
    so::sim3_to_se3(sim3_q, sim3_t, a.q, a.t);   // Tcw = SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale()), :1245
    for (int i = 0; i < 4; i++) a.K[i] = K[i];
    a.check_normal = 1; a.check_chi = 0; a.accept_th = kThLow;
    return run_proj_search(kf, a, m, xw, normal, min_dist, max_dist, mp_desc, skip, th, best_idx, best_dist);
}

int dvm_match_by_sim3(dvm_frame* kf1, dvm_frame* kf2, const float* q1, const float* t1, const float* q2, const float* t2,
                      const float* s12_q, const float* s12_t, const float* K, const uint8_t* skip1, const float* xw1,
                      const float* min_dist1, const float* max_dist1, const uint8_t* mp_desc1, const uint8_t* skip2,
                      const float* xw2, const float* min_dist2, const float* max_dist2, const uint8_t* mp_desc2, float th,
                      int32_t* match12, int* nfound)
{
    DVM_REQUIRE(kf1 && kf2 && q1 && t1 && q2 && t2 && s12_q && s12_t && K && match12 && nfound, "null argument");
    const int n1 = kf1->host_n, n2 = kf2->host_n;
    *nfound = 0;
    for (int i = 0; i < n1; i++) match12[i] = -1;
    if (n1 == 0 || n2 == 0) return DVM_OK;
    DVM_REQUIRE(skip1 && xw1 && min_dist1 && max_dist1 && mp_desc1 && skip2 && xw2 && min_dist2 && max_dist2 && mp_desc2,
                "null map-point arrays");
    float s21_q[4], s21_t[3];
    so::sim3_inverse(s12_q, s12_t, s21_q, s21_t);   // Sophus::Sim3f S21 = S12.inverse(), :1359
    std::vector<int32_t> m1(n1), m2(n2);
    FuseArgs a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 4; i++) a.K[i] = K[i];   // both directions project with pKF1's intrinsics (:1349-1352)
    a.chain_sim3 = 1; a.proj_invz = 1; a.dist_camera = 1; a.accept_th = kThHigh;
    // keyframe 1's map points into keyframe 2: p3Dc2 = S21 * (T1w * p3Dw)
    for (int i = 0; i < 4; i++) { a.q[i] = q1[i]; a.sq[i] = s21_q[i]; }
    for (int i = 0; i < 3; i++) { a.t[i] = t1[i]; a.st[i] = s21_t[i]; }
    int rc = run_proj_search(kf2, a, n1, xw1, nullptr, min_dist1, max_dist1, mp_desc1, skip1, th, m1.data(), nullptr);
    if (rc != DVM_OK) return rc;
    // keyframe 2's map points into keyframe 1: p3Dc1 = S12 * (T2w * p3Dw)
    for (int i = 0; i < 4; i++) { a.q[i] = q2[i]; a.sq[i] = s12_q[i]; }
    for (int i = 0; i < 3; i++) { a.t[i] = t2[i]; a.st[i] = s12_t[i]; }
    rc = run_proj_search(kf1, a, n2, xw2, nullptr, min_dist2, max_dist2, mp_desc2, skip2, th, m2.data(), nullptr);
    if (rc != DVM_OK) return rc;
    int found = 0;
    for (int i1 = 0; i1 < n1; i1++) {   // the agreement check, :1539-1549
        const int idx2 = m1[i1];
        if (idx2 >= 0 && m2[idx2] == i1) { match12[i1] = idx2; found++; }
    }
    *nfound = found;
    return DVM_OK;
}

int dvm_match_by_projection_sim3(dvm_frame* kf, const float* sim3_q, const float* sim3_t, const float* K, int m, const float* xw,
                                 const float* normal, const float* min_dist, const float* max_dist, const uint8_t* mp_desc,
                                 const uint8_t* skip, const uint8_t* kp_matched, int th, float ratio_hamming, int32_t* kp_point,
                                 int* nmatches)
{
    DVM_REQUIRE(kf != nullptr && sim3_q && sim3_t && K && m >= 0 && kp_point && nmatches, "bad argument");
    *nmatches = 0;
    const int nk = kf->host_n;
    for (int k = 0; k < nk; k++) kp_point[k] = -1;
    if (m == 0 || nk == 0) return DVM_OK;
    DVM_REQUIRE(xw && normal && min_dist && max_dist && mp_desc && kp_matched, "null arrays");
    DVM_CUDA(cudaSetDevice(kf->device));
    const size_t n = (size_t)m;
    int rc = dvm_frame_ensure_bytes(kf, padded({ n * 12, n * 12, n * 4, n * 4, n * 32, n, n * 4, n * 4, n * 4, (size_t)nk }),
                                    (size_t)(kf->cap + 8) * sizeof(int));
    if (rc != DVM_OK) return rc;
    rc = dvm_frame_ensure_query_cap(kf, m);
    if (rc != DVM_OK) return rc;
    Stage st(kf);
    FuseArgs a;
    memset(&a, 0, sizeof(a));
    so::sim3_to_se3(sim3_q, sim3_t, a.q, a.t);   // :403
    for (int i = 0; i < 4; i++) a.K[i] = K[i];
    a.check_normal = 1; a.gate_only = 1;
    a.nlevels = kf->dev.nlevels;
    a.logScale = dvm_glibc_logf(kf->dev.nlevels > 1 ? kf->dev.scale[1] : 1.2f);
    a.m = m; a.th = (float)th;
    a.xw = st.add(xw, n * 3); a.normal = st.add(normal, n * 3);
    a.min_dist = st.add(min_dist, n); a.max_dist = st.add(max_dist, n);
    const uint8_t* d_desc = st.add(mp_desc, n * 32);
    a.mp_desc = d_desc;
    a.skip = skip ? st.add(skip, n) : nullptr;
    const uint8_t* d_blocked = st.add(kp_matched, (size_t)nk);
    a.gate_u = st.add((const float*)nullptr, n); a.gate_v = st.add((const float*)nullptr, n);
    a.gate_level = st.add((const int*)nullptr, n);
    DVM_CUDA(cudaMemcpyAsync(kf->d_in, kf->h_in, st.off, cudaMemcpyHostToDevice, kf->stream));
    launch_fuse_search(kf->dev, a, kf->stream);                  // the projection gates, one warp per candidate
    MatchMapArgs ma;
    memset(&ma, 0, sizeof(ma));
    ma.m = m; ma.th = (float)th; ma.nnratio = 1.f;
    ma.projX = a.gate_u; ma.projY = a.gate_v; ma.level = a.gate_level;
    ma.mp_desc = d_desc; ma.cur_blocked = d_blocked;
    ma.sim3_mode = 1;
    ma.accept_limit = (float)kThLow * ratio_hamming;            // bestDist <= TH_LOW * ratioHamming, :487
    launch_match_map(kf->dev, ma, kf->ms, kf->d_cur_mp, kf->d_cur_mp + kf->cap, kf->stream);   // sequential-greedy outcome
    return dvm_frame_finish_match(kf, kp_point, nmatches);
}

int dvm_fundamental_from_poses(const float* q1, const float* t1, const float* q2, const float* t2, const float* K1,
                               const float* K2, float* F12, float* ep)
{
    DVM_REQUIRE(q1 && t1 && q2 && t2 && K1 && K2 && F12 && ep, "null argument");
    float qw2[4], tw2[3], q12[4], t12[3], qw1[4], Cw[3], C2[3], R12[9];
    so::se3_inverse(q2, t2, qw2, tw2);                 // Tw2 = pKF2->GetPoseInverse()
    so::se3_mul(q1, t1, qw2, tw2, q12, t12);           // T12 = T1w * Tw2
    so::quat_to_matrix(q12, R12);
    so::se3_inverse(q1, t1, qw1, Cw);                  // Cw = pKF1->GetCameraCenter()
    so::se3_apply(q2, t2, Cw, C2);                     // C2 = T2w * Cw
    ep[0] = so::fa(so::fd(so::fm(K2[0], C2[0]), C2[2]), K2[2]);
    ep[1] = so::fa(so::fd(so::fm(K2[1], C2[1]), C2[2]), K2[3]);
    const float t12x[9] = { 0.f, -t12[2], t12[1], t12[2], 0.f, -t12[0], -t12[1], t12[0], 0.f };   // Sophus::SO3f::hat
    const float K1T[9] = { K1[0], 0.f, 0.f, 0.f, K1[1], 0.f, K1[2], K1[3], 1.f };                 // toK_().transpose()
    const float K2m[9] = { K2[0], 0.f, K2[2], 0.f, K2[1], K2[3], 0.f, 0.f, 1.f };
    float K1Ti[9], K2i[9], A[9], B[9];
    so::mat_inverse(K1T, K1Ti);
    so::mat_inverse(K2m, K2i);
    so::mat_mul(K1Ti, t12x, A);
    so::mat_mul(A, R12, B);
    so::mat_mul(B, K2i, F12);
    return DVM_OK;
}

int dvm_hamming_create(dvm_hamming** out, int device, void* cuda_stream)
{
    DVM_REQUIRE(out != nullptr, "null output handle");
    *out = nullptr;
    int rc = select_device(device);
    if (rc != DVM_OK) return rc;
    dvm_hamming* h = new dvm_hamming;
    h->device = device;
    if (cuda_stream) h->stream = (cudaStream_t)cuda_stream;
    else {
        cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e)); delete h; return DVM_ERR_CUDA; }
        h->own_stream = true;
    }
    *out = h;
    return DVM_OK;
}

void dvm_hamming_destroy(dvm_hamming* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->scratch.part[0]); cudaFree(h->scratch.part[1]); cudaFree(h->scratch.expanded); cudaFree(h->d_buf);
    if (h->h_buf) cudaFreeHost(h->h_buf);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

int dvm_hamming_knn_device(dvm_hamming* h, const uint8_t* a_dev, int ba, int na, const uint8_t* b_dev, int bb, int nb,
                           uint32_t* key1_dev, uint32_t* key2_dev, int32_t* counts_dev, int th_low, float nnratio)
{
    DVM_REQUIRE(h != nullptr && ba >= 0 && bb >= 0 && na >= 0 && nb >= 0 && nb < (1 << 20), "bad sizes");
    DVM_REQUIRE((long long)ba * bb < (1LL << 31), "too many block pairs");
    DVM_REQUIRE(ba * bb * na == 0 || (a_dev && key1_dev && key2_dev && (nb == 0 || b_dev)), "null device arrays");
    DVM_CUDA(cudaSetDevice(h->device));
    KnnArgs k;
    k.a = a_dev; k.ba = ba; k.na = na; k.b = b_dev; k.bb = bb; k.nb = nb;
    k.key1 = key1_dev; k.key2 = key2_dev; k.counts = counts_dev; k.th_low = th_low; k.nnratio = nnratio;
    return launch_hamming_knn(k, h->scratch, h->stream, h->mode);
}

int dvm_hamming_set_mode(dvm_hamming* h, int mode)
{
    DVM_REQUIRE(h != nullptr && mode >= 0 && mode <= 4, "mode must be 0..4");
    h->mode = mode;
    return DVM_OK;
}

int dvm_hamming_sync(dvm_hamming* h)
{
    DVM_REQUIRE(h != nullptr, "null handle");
    DVM_CUDA(cudaSetDevice(h->device));
    DVM_CUDA(cudaStreamSynchronize(h->stream));
    return DVM_OK;
}

int dvm_hamming_knn(dvm_hamming* h, const uint8_t* a, int na, const uint8_t* b, int nb, int32_t* best_idx,
                    int32_t* best_dist, int32_t* second_dist)
{
    DVM_REQUIRE(h != nullptr && na >= 0 && nb >= 0 && nb < (1 << 20), "bad sizes");
    DVM_REQUIRE(na == 0 || (a && best_idx && best_dist && second_dist), "null query arrays");
    DVM_REQUIRE(nb == 0 || b, "null database");
    if (na == 0) return DVM_OK;
    DVM_CUDA(cudaSetDevice(h->device));
    const size_t ab = ((size_t)na * 32 + 255) & ~(size_t)255, bbytes = ((size_t)nb * 32 + 255) & ~(size_t)255;
    const size_t kb = ((size_t)na * 4 + 255) & ~(size_t)255;
    const size_t need = ab + bbytes + 2 * kb;
    if (need > h->d_cap) {
        DVM_CUDA(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_buf); h->d_buf = nullptr;
        if (h->h_buf) { cudaFreeHost(h->h_buf); h->h_buf = nullptr; }
        h->d_cap = h->h_cap = 0;
        const size_t cap = need + need / 4;
        DVM_CUDA(cudaMalloc(&h->d_buf, cap));
        DVM_CUDA(cudaHostAlloc(&h->h_buf, cap, cudaHostAllocDefault));
        h->d_cap = h->h_cap = cap;
    }
    memcpy(h->h_buf, a, (size_t)na * 32);
    if (nb) memcpy(h->h_buf + ab, b, (size_t)nb * 32);
    DVM_CUDA(cudaMemcpyAsync(h->d_buf, h->h_buf, ab + bbytes, cudaMemcpyHostToDevice, h->stream));
    uint32_t* k1 = reinterpret_cast<uint32_t*>(h->d_buf + ab + bbytes);
    uint32_t* k2 = reinterpret_cast<uint32_t*>(h->d_buf + ab + bbytes + kb);
    int rc = dvm_hamming_knn_device(h, h->d_buf, 1, na, h->d_buf + ab, 1, nb, k1, k2, nullptr, 0, 0.f);
    if (rc != DVM_OK) return rc;
    DVM_CUDA(cudaMemcpyAsync(h->h_buf + ab + bbytes, k1, 2 * kb, cudaMemcpyDeviceToHost, h->stream));
    DVM_CUDA(cudaStreamSynchronize(h->stream));
    const uint32_t* hk1 = reinterpret_cast<const uint32_t*>(h->h_buf + ab + bbytes);
    const uint32_t* hk2 = reinterpret_cast<const uint32_t*>(h->h_buf + ab + bbytes + kb);
    for (int i = 0; i < na; i++) {
        const uint32_t d1 = hk1[i] >> 20, d2 = hk2[i] >> 20;
        best_dist[i] = (int32_t)(d1 < 256 ? d1 : 256);
        best_idx[i] = d1 < 256 ? (int32_t)(hk1[i] & 0xfffffu) : -1;
        second_dist[i] = (int32_t)(d2 < 256 ? d2 : 256);
    }
    return DVM_OK;
}

} // extern "C"
