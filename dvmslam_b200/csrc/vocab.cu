// vocab.cu -- the DBoW2 vocabulary transform (Frame::ComputeBoW, O3/src/Frame.cc:784-789) on sm_100a.
//
// TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)
// (O3/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1106-1146) walks one descriptor down the k-ary tree: at every
// level the nearest child by Hamming distance (FORB::distance, DBoW2/FORB.cpp:80-97), strict '<' so the first of
// equal children wins.  One warp per feature: the lanes take the children of the current node, the winner is a
// warp minimum of distance << 8 | child position.  The tree (ORBvoc: k = 10, L = 6, 1.1 M nodes x 32 B = 35 MB) is
// read through L2; the upper levels stay cache-resident, so a frame costs about L dependent L2 round trips.
// The BowVector / FeatureVector maps are assembled from the per-feature results by the caller (ordered maps,
// sequential double sums in feature order: dvmslam_b200/vocabulary.py, host/vocabulary_adapter.h).
#include <mutex>
#include "common.cuh"
#include <vector>

namespace dvm {

struct VocabDev {
    int n_nodes, L;
    const int* child_start;
    const int* children;
    const uint8_t* desc;
    const double* weight;
    const int* word_id;
};

constexpr int kVocabWarps = 4;

__global__ void __launch_bounds__(kVocabWarps * 32) vocab_transform_kernel(VocabDev v, const uint8_t* __restrict__ feat, int n,
                                                                           int levelsup, int* __restrict__ word,
                                                                           double* __restrict__ w, int* __restrict__ nid_out)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * kVocabWarps + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint4* fp = reinterpret_cast<const uint4*>(feat + (size_t)i * 32);
    const uint4 f0 = __ldg(fp), f1 = __ldg(fp + 1);
    const int nid_level = v.L - levelsup;
    int nid = nid_level <= 0 ? 0 : -1;
    int node = 0, level = 0;
    while (true) {
        const int c0 = v.child_start[node], c1 = v.child_start[node + 1];
        if (c1 <= c0) break;   // leaf
        ++level;
        unsigned best = ~0u;
        for (int c = c0 + lane; c < c1; c += 32) {
            const int id = v.children[c];
            const uint4* dp = reinterpret_cast<const uint4*>(v.desc + (size_t)id * 32);
            const uint4 d0 = __ldg(dp), d1 = __ldg(dp + 1);
            const unsigned dist = __popc(f0.x ^ d0.x) + __popc(f0.y ^ d0.y) + __popc(f0.z ^ d0.z) + __popc(f0.w ^ d0.w) +
                                  __popc(f1.x ^ d1.x) + __popc(f1.y ^ d1.y) + __popc(f1.z ^ d1.z) + __popc(f1.w ^ d1.w);
            best = min(best, (dist << 20) | (unsigned)(c - c0));
        }
        best = __reduce_min_sync(0xffffffffu, best);
        node = v.children[c0 + (int)(best & 0xfffffu)];
        if (level == nid_level) nid = node;
    }
    if (nid < 0) nid = node;
    if (lane == 0) { word[i] = v.word_id[node]; w[i] = v.weight[node]; nid_out[i] = nid; }
}

} // namespace dvm

using namespace dvm;

struct dvm_vocabulary {
    int device = 0;
    cudaStream_t stream = nullptr;
    VocabDev dev;
    int* d_child_start = nullptr; int* d_children = nullptr; uint8_t* d_desc = nullptr; double* d_weight = nullptr;
    int* d_word_id = nullptr;
    uint8_t* d_buf = nullptr; uint8_t* h_buf = nullptr; size_t cap = 0;   // per-call staging: features | word | nid | weight
    // The reference shares ONE vocabulary between the tracking and the local-mapping thread (Frame::ComputeBoW from
    // Tracking.cc:2463 / 3279 and KeyFrame::ComputeBoW from LocalMapping.cc:320 / 375 run concurrently): transform()
    // calls are serialised here because they share the staging buffers and the stream.
    std::mutex mtx;
};

static void vocab_free(dvm_vocabulary* v)
{
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->stream) cudaStreamSynchronize(v->stream);
    cudaFree(v->d_child_start); cudaFree(v->d_children); cudaFree(v->d_desc); cudaFree(v->d_weight); cudaFree(v->d_word_id);
    cudaFree(v->d_buf);
    if (v->h_buf) cudaFreeHost(v->h_buf);
    if (v->stream) cudaStreamDestroy(v->stream);
    delete v;
}

extern "C" {

int dvm_vocabulary_create(dvm_vocabulary** out, int device, int n_nodes, const int32_t* child_start, const int32_t* children,
                          const uint8_t* desc, const double* weight, const int32_t* word_id, int L)
{
    DVM_REQUIRE(out != nullptr, "null output handle");
    *out = nullptr;
    DVM_REQUIRE(n_nodes >= 1 && child_start && children && desc && weight && word_id && L >= 1, "bad vocabulary");
    DVM_REQUIRE(child_start[0] == 0, "child_start must begin at 0");
    std::vector<uint8_t> is_child((size_t)n_nodes, 0);
    for (int i = 0; i < n_nodes; i++) {
        DVM_REQUIRE(child_start[i] <= child_start[i + 1], "child_start must not decrease");
        for (int c = child_start[i]; c < child_start[i + 1]; c++) {
            DVM_REQUIRE(children[c] > 0 && children[c] < n_nodes && !is_child[children[c]], "vocabulary is not a tree");
            is_child[children[c]] = 1;   // one parent per node and the root is nobody's child: every walk ends in a leaf
        }
    }
    int rc = select_device(device);
    if (rc != DVM_OK) return rc;
    dvm_vocabulary* v = new dvm_vocabulary;
    v->device = device;
    const size_t nn = (size_t)n_nodes, nc = (size_t)child_start[n_nodes];
#define DVM_VCREATE(call)                                                                        \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            set_error("%s failed in dvm_vocabulary_create: %s", #call, cudaGetErrorString(e__)); \
            vocab_free(v);                                                                       \
            return DVM_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)
    DVM_VCREATE(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
    DVM_VCREATE(cudaMalloc(&v->d_child_start, (nn + 1) * 4)); DVM_VCREATE(cudaMalloc(&v->d_children, (nc + 1) * 4));
    DVM_VCREATE(cudaMalloc(&v->d_desc, nn * 32)); DVM_VCREATE(cudaMalloc(&v->d_weight, nn * 8));
    DVM_VCREATE(cudaMalloc(&v->d_word_id, nn * 4));
    DVM_VCREATE(cudaMemcpy(v->d_child_start, child_start, (nn + 1) * 4, cudaMemcpyHostToDevice));
    DVM_VCREATE(cudaMemcpy(v->d_children, children, nc * 4, cudaMemcpyHostToDevice));
    DVM_VCREATE(cudaMemcpy(v->d_desc, desc, nn * 32, cudaMemcpyHostToDevice));
    DVM_VCREATE(cudaMemcpy(v->d_weight, weight, nn * 8, cudaMemcpyHostToDevice));
    DVM_VCREATE(cudaMemcpy(v->d_word_id, word_id, nn * 4, cudaMemcpyHostToDevice));
#undef DVM_VCREATE
    v->dev = VocabDev{ n_nodes, L, v->d_child_start, v->d_children, v->d_desc, v->d_weight, v->d_word_id };
    *out = v;
    return DVM_OK;
}

void dvm_vocabulary_destroy(dvm_vocabulary* v) { vocab_free(v); }

int dvm_vocabulary_transform(dvm_vocabulary* v, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* weight,
                             int32_t* node_id)
{
    DVM_REQUIRE(v != nullptr && n >= 0, "bad argument");
    if (n == 0) return DVM_OK;
    DVM_REQUIRE(desc && word_id && weight && node_id, "null arrays");
    std::lock_guard<std::mutex> lock(v->mtx);
    DVM_CUDA(cudaSetDevice(v->device));
    const size_t sn = (size_t)n;
    const size_t o_w = sn * 32, o_word = o_w + sn * 8, o_nid = o_word + sn * 4, total = o_nid + sn * 4;
    if (total > v->cap) {
        DVM_CUDA(cudaStreamSynchronize(v->stream));
        cudaFree(v->d_buf); v->d_buf = nullptr;
        if (v->h_buf) { cudaFreeHost(v->h_buf); v->h_buf = nullptr; }
        v->cap = 0;
        const size_t cap = total + total / 4 + 4096;
        DVM_CUDA(cudaMalloc(&v->d_buf, cap));
        DVM_CUDA(cudaHostAlloc(&v->h_buf, cap, cudaHostAllocDefault));
        v->cap = cap;
    }
    memcpy(v->h_buf, desc, sn * 32);
    DVM_CUDA(cudaMemcpyAsync(v->d_buf, v->h_buf, sn * 32, cudaMemcpyHostToDevice, v->stream));
    DVM_LAUNCH(vocab_transform_kernel, div_up(n, kVocabWarps), kVocabWarps * 32, 0, v->stream, v->dev, v->d_buf, n, levelsup,
               reinterpret_cast<int*>(v->d_buf + o_word), reinterpret_cast<double*>(v->d_buf + o_w),
               reinterpret_cast<int*>(v->d_buf + o_nid));
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaMemcpyAsync(v->h_buf + o_w, v->d_buf + o_w, total - o_w, cudaMemcpyDeviceToHost, v->stream));
    DVM_CUDA(cudaStreamSynchronize(v->stream));
    memcpy(weight, v->h_buf + o_w, sn * 8);
    memcpy(word_id, v->h_buf + o_word, sn * 4);
    memcpy(node_id, v->h_buf + o_nid, sn * 4);
    return DVM_OK;
}

} // extern "C"
