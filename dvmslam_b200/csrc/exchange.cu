// exchange.cu -- the inter-agent loop-closure exchange step (BASELINE.json config C3) behind the C-ABI.
//
// Reference behaviour this replaces (W/ = src/slam_system/): every agent pushes the BoW vectors of its new keyframes to
// its peers (sendNewKeyFrameBows, W/src/orb_slam3_wrapper.cpp:457-534: at least MIN_BOW_SHARE_SIZE new keyframes, each
// keyframe sent to a peer once) and the agent with the LOWER id of a pair looks for merge candidates among its own
// keyframes (receiveNewKeyFrameBows :536-618, isLeadNodeInGroup :1238-1243) before descriptors are compared.
//
// B200-native form: the keyframes' descriptor blocks u8[K][N][32] live in HBM.  One round =
//   1. plan      per peer, the not-yet-sent keyframes go to the rank that owns the agent pair (host logic);
//   2. counts    one int64 per peer, grouped ncclSend / ncclRecv;
//   3. blocks    grouped ncclSend straight out of the database (the unsent keyframes of a peer are one contiguous range)
//                and ncclRecv into the receive buffer -- NVLink / NVSwitch, no staging;
//   4. match     every received block against the whole local database with the exhaustive Hamming search
//                (tcgen05 int8 kernel, hamming_tc.cu) on the same stream, i.e. as soon as the transfers have landed;
//   5. report    keyframe pairs with at least `min_matches` accepted descriptor matches, best first.
// NCCL is resolved at run time (dlopen of libnccl.so.2: the one the process already holds, e.g. PyTorch's, else the
// system's), so the library has no link-time dependency on it.  The communicator is created from an ncclUniqueId that
// the host side distributes by whatever channel it has (the reference's agents talk over DDS).
// For tests of the host logic without GPUs the transport and the matcher can be injected (dvm_exchange_hooks): the
// database then lives in host memory and nothing CUDA is touched.  There is no built-in CPU matcher.
#include "bow_kernels.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <mutex>
#include <vector>

using namespace dvm;

struct Id128 { char bytes[128]; };   // ncclUniqueId

namespace {

// ---- the few NCCL entry points used, bound at run time ----
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128 /* ncclUniqueId, by value */, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclUint8 = 1;   // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1

NcclApi* nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) return;
        auto sym = [&](const char* n) { return dlsym(api.lib, n); };
        api.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
        api.CommInitRank = (int (*)(void**, int, Id128, int))sym("ncclCommInitRank");
        api.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
        api.GroupStart = (int (*)())sym("ncclGroupStart");
        api.GroupEnd = (int (*)())sym("ncclGroupEnd");
        api.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
        api.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
        api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.GroupStart || !api.GroupEnd || !api.Send || !api.Recv) {
            dlclose(api.lib);
            api.lib = nullptr;
        }
    });
    return api.lib ? &api : nullptr;
}

#define DVM_NCCL(call)                                                                                   \
    do {                                                                                                 \
        const int r__ = (call);                                                                          \
        if (r__ != 0) {                                                                                  \
            set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                               \
                      nccl()->GetErrorString ? nccl()->GetErrorString(r__) : "NCCL error");              \
            return DVM_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

// the rank that matches the keyframes of agents i and j: "lead" = the lower id (the reference's rule: only the lead node
// attempts a merge); "balanced" = the lower id when the ids differ by an odd number, else the higher one, so that every
// rank owns about (world - 1) / 2 pairs
int pair_owner(int i, int j, int balanced)
{
    const int lo = std::min(i, j), hi = std::max(i, j);
    return (!balanced || ((hi - lo) & 1)) ? lo : hi;
}

} // namespace

struct dvm_exchange {
    int device = -1;                 // -1: host-memory database (hooks injected, tests only)
    int rank = 0, world = 1, n_feat = 0, cap = 0, balanced = 0;
    int th_low = 50, min_matches = 20, min_share = 5;
    float nnratio = 0.75f;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    void* comm = nullptr;
    dvm_exchange_hooks hooks = {};
    bool hooked = false;
    uint8_t* db = nullptr;           // [cap][n_feat][32], device (or host with hooks)
    int n_kf = 0;
    std::vector<int> sent_upto, recv_upto;
    uint8_t* recvbuf = nullptr; size_t recv_cap = 0;
    long long* d_counts = nullptr;   // [2 * world] send | recv (device)
    long long* h_counts = nullptr;   // pinned mirror
    int* d_match = nullptr; size_t match_cap = 0;   // accepted-match counts [k][n_kf]
    uint32_t* d_keys = nullptr; size_t keys_cap = 0;
    KnnScratch scratch;
    size_t last_bytes_sent = 0;
    float last_match_ms = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

static void exchange_free(dvm_exchange* x)
{
    if (!x) return;
    if (x->device >= 0) {
        cudaSetDevice(x->device);
        if (x->stream) cudaStreamSynchronize(x->stream);
        if (x->comm && nccl()) nccl()->CommDestroy(x->comm);
        cudaFree(x->db); cudaFree(x->recvbuf); cudaFree(x->d_counts); cudaFree(x->d_match); cudaFree(x->d_keys);
        cudaFree(x->scratch.part[0]); cudaFree(x->scratch.part[1]); cudaFree(x->scratch.expanded);
        if (x->h_counts) cudaFreeHost(x->h_counts);
        if (x->ev0) cudaEventDestroy(x->ev0);
        if (x->ev1) cudaEventDestroy(x->ev1);
        if (x->own_stream && x->stream) cudaStreamDestroy(x->stream);
    } else {
        free(x->db); free(x->recvbuf);
    }
    delete x;
}

extern "C" {

int dvm_exchange_unique_id(uint8_t* id128)
{
    DVM_REQUIRE(id128 != nullptr, "null id buffer");
    NcclApi* n = nccl();
    if (!n) { set_error("libnccl.so.2 could not be loaded"); return DVM_ERR_CUDA; }
    DVM_NCCL(n->GetUniqueId(id128));
    return DVM_OK;
}

int dvm_exchange_create(dvm_exchange** out, int device, int rank, int world, const uint8_t* id128, int n_feat, int max_keyframes,
                        int balanced_ownership, void* cuda_stream, const dvm_exchange_hooks* hooks)
{
    DVM_REQUIRE(out != nullptr, "null output handle");
    *out = nullptr;
    DVM_REQUIRE(world >= 1 && rank >= 0 && rank < world && n_feat > 0 && max_keyframes > 0, "bad sizes");
    dvm_exchange* x = new dvm_exchange;
    x->rank = rank; x->world = world; x->n_feat = n_feat; x->cap = max_keyframes; x->balanced = balanced_ownership ? 1 : 0;
    x->sent_upto.assign(world, 0); x->recv_upto.assign(world, 0);
    const size_t db_bytes = (size_t)max_keyframes * n_feat * 32;
    if (hooks) {
        if (!hooks->alltoall || !hooks->match_counts) { set_error("dvm_exchange_hooks needs both callbacks"); delete x; return DVM_ERR_INVALID; }
        x->hooks = *hooks; x->hooked = true; x->device = -1;
        x->db = (uint8_t*)malloc(db_bytes);
        if (!x->db) { set_error("out of host memory"); delete x; return DVM_ERR_CAPACITY; }
        *out = x;
        return DVM_OK;
    }
    int rc = select_device(device);   // no hooks: a B200 is required (there is no CPU matcher)
    if (rc != DVM_OK) { delete x; return rc; }
    x->device = device;
#define DVM_XCREATE(call)                                                                    \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            set_error("%s failed in dvm_exchange_create: %s", #call, cudaGetErrorString(e__)); \
            exchange_free(x);                                                                \
            return DVM_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)
    if (cuda_stream) x->stream = (cudaStream_t)cuda_stream;
    else { DVM_XCREATE(cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking)); x->own_stream = true; }
    DVM_XCREATE(cudaMalloc(&x->db, db_bytes));
    DVM_XCREATE(cudaMalloc(&x->d_counts, sizeof(long long) * 2 * world));
    DVM_XCREATE(cudaHostAlloc(&x->h_counts, sizeof(long long) * 2 * world, cudaHostAllocDefault));
    DVM_XCREATE(cudaEventCreate(&x->ev0));
    DVM_XCREATE(cudaEventCreate(&x->ev1));
#undef DVM_XCREATE
    if (world > 1) {
        NcclApi* n = nccl();
        if (!n || !id128) { set_error(n ? "null ncclUniqueId" : "libnccl.so.2 could not be loaded"); exchange_free(x); return DVM_ERR_CUDA; }
        Id128 id;
        memcpy(id.bytes, id128, 128);
        const int r = n->CommInitRank(&x->comm, world, id, rank);
        if (r != 0) { set_error("ncclCommInitRank failed: %s", n->GetErrorString ? n->GetErrorString(r) : "?"); exchange_free(x); return DVM_ERR_CUDA; }
    }
    *out = x;
    return DVM_OK;
}

void dvm_exchange_destroy(dvm_exchange* x) { exchange_free(x); }

int dvm_exchange_set_policy(dvm_exchange* x, int th_low, float nnratio, int min_matches, int min_share)
{
    DVM_REQUIRE(x != nullptr && th_low >= 0 && min_matches >= 0 && min_share >= 0, "bad argument");
    x->th_low = th_low; x->nnratio = nnratio; x->min_matches = min_matches; x->min_share = min_share;
    return DVM_OK;
}

int dvm_exchange_add_keyframes(dvm_exchange* x, const uint8_t* desc, int n_kf, int desc_is_device, int* first_id)
{
    DVM_REQUIRE(x != nullptr && n_kf >= 0 && (n_kf == 0 || desc), "bad argument");
    if (x->n_kf + n_kf > x->cap) { set_error("keyframe database is full (%d + %d > %d)", x->n_kf, n_kf, x->cap); return DVM_ERR_CAPACITY; }
    const size_t row = (size_t)x->n_feat * 32;
    if (first_id) *first_id = x->n_kf;
    if (n_kf == 0) return DVM_OK;
    uint8_t* dst = x->db + (size_t)x->n_kf * row;
    if (x->device < 0) memcpy(dst, desc, (size_t)n_kf * row);
    else {
        DVM_CUDA(cudaSetDevice(x->device));
        DVM_CUDA(cudaMemcpyAsync(dst, desc, (size_t)n_kf * row, desc_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, x->stream));
        if (!desc_is_device) DVM_CUDA(cudaStreamSynchronize(x->stream));   // the caller's host buffer may go away
    }
    x->n_kf += n_kf;
    return DVM_OK;
}

int dvm_exchange_reset(dvm_exchange* x)
{
    DVM_REQUIRE(x != nullptr, "null handle");
    if (x->device >= 0) { DVM_CUDA(cudaSetDevice(x->device)); DVM_CUDA(cudaStreamSynchronize(x->stream)); }
    x->n_kf = 0;
    std::fill(x->sent_upto.begin(), x->sent_upto.end(), 0);
    std::fill(x->recv_upto.begin(), x->recv_upto.end(), 0);
    x->last_bytes_sent = 0;
    return DVM_OK;
}

int dvm_exchange_keyframes(const dvm_exchange* x) { return x ? x->n_kf : DVM_ERR_INVALID; }
const uint8_t* dvm_exchange_database(const dvm_exchange* x) { return x ? x->db : nullptr; }
size_t dvm_exchange_last_bytes_sent(const dvm_exchange* x) { return x ? x->last_bytes_sent : 0; }
float dvm_exchange_last_match_ms(const dvm_exchange* x) { return x ? x->last_match_ms : -1.f; }

// accepted-match counts [ka][n_kf] of `a` (ka keyframe blocks) against the local database -> host `counts`
int dvm_exchange_match_counts(dvm_exchange* x, const uint8_t* a, int ka, int32_t* counts)
{
    DVM_REQUIRE(x != nullptr && ka >= 0 && (ka == 0 || (a && counts)), "bad argument");
    const int kb = x->n_kf;
    if (ka == 0 || kb == 0) return DVM_OK;
    if (x->hooked) return x->hooks.match_counts(x->hooks.ctx, a, ka, x->db, kb, x->n_feat, x->th_low, x->nnratio, counts) == 0 ? DVM_OK : DVM_ERR_INVALID;
    DVM_CUDA(cudaSetDevice(x->device));
    const size_t nm = (size_t)ka * kb;
    if (nm > x->match_cap) {
        DVM_CUDA(cudaStreamSynchronize(x->stream));
        cudaFree(x->d_match); x->d_match = nullptr;
        DVM_CUDA(cudaMalloc(&x->d_match, (nm + nm / 4 + 64) * sizeof(int)));
        x->match_cap = nm + nm / 4 + 64;
    }
    // the per-descriptor keys are scratch here (only the counts leave the GPU): the batch is chunked to bound them
    const size_t per_a = (size_t)kb * x->n_feat;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)ka, ((size_t)256 << 20) / (per_a * 8)));
    const size_t need = (size_t)chunk * per_a * 2;
    if (need > x->keys_cap) {
        DVM_CUDA(cudaStreamSynchronize(x->stream));
        cudaFree(x->d_keys); x->d_keys = nullptr;
        DVM_CUDA(cudaMalloc(&x->d_keys, need * sizeof(uint32_t)));
        x->keys_cap = need;
    }
    DVM_CUDA(cudaEventRecord(x->ev0, x->stream));
    for (int s = 0; s < ka; s += chunk) {
        const int n = std::min(chunk, ka - s);
        KnnArgs k;
        k.a = a + (size_t)s * x->n_feat * 32; k.ba = n; k.na = x->n_feat;
        k.b = x->db; k.bb = kb; k.nb = x->n_feat;
        k.key1 = x->d_keys; k.key2 = x->d_keys + (size_t)chunk * per_a;
        k.counts = x->d_match + (size_t)s * kb; k.th_low = x->th_low; k.nnratio = x->nnratio;
        const int rc = launch_hamming_knn(k, x->scratch, x->stream, 0);
        if (rc != DVM_OK) return rc;
    }
    DVM_CUDA(cudaEventRecord(x->ev1, x->stream));
    DVM_CUDA(cudaMemcpyAsync(counts, x->d_match, nm * sizeof(int), cudaMemcpyDeviceToHost, x->stream));
    DVM_CUDA(cudaStreamSynchronize(x->stream));
    DVM_CUDA(cudaEventElapsedTime(&x->last_match_ms, x->ev0, x->ev1));
    return DVM_OK;
}

int dvm_exchange_round(dvm_exchange* x, int32_t* candidates, int cap, int* n_candidates)
{
    DVM_REQUIRE(x != nullptr && n_candidates != nullptr && cap >= 0 && (cap == 0 || candidates), "bad argument");
    *n_candidates = 0;
    const int W = x->world;
    const size_t row = (size_t)x->n_feat * 32;
    // ---- 1. plan: only to the owner of the pair, only unsent keyframes, only when enough are new ----
    std::vector<long long> send(W, 0), recv(W, 0);
    for (int p = 0; p < W; p++) {
        if (p == x->rank || pair_owner(x->rank, p, x->balanced) != p) continue;
        const int fresh = x->n_kf - x->sent_upto[p];
        send[p] = fresh >= x->min_share ? fresh : 0;
    }
    // ---- 2. counts ----
    if (W > 1) {
        if (x->hooked) {
            std::vector<size_t> eight(W, sizeof(long long));
            if (x->hooks.alltoall(x->hooks.ctx, send.data(), eight.data(), recv.data(), eight.data()) != 0) { set_error("alltoall hook failed"); return DVM_ERR_INVALID; }
        } else {
            DVM_CUDA(cudaSetDevice(x->device));
            memcpy(x->h_counts, send.data(), sizeof(long long) * W);
            DVM_CUDA(cudaMemcpyAsync(x->d_counts, x->h_counts, sizeof(long long) * W, cudaMemcpyHostToDevice, x->stream));
            NcclApi* n = nccl();
            DVM_NCCL(n->GroupStart());
            for (int p = 0; p < W; p++) {
                if (p == x->rank) continue;
                DVM_NCCL(n->Send(x->d_counts + p, sizeof(long long), kNcclUint8, p, x->comm, x->stream));
                DVM_NCCL(n->Recv(x->d_counts + W + p, sizeof(long long), kNcclUint8, p, x->comm, x->stream));
            }
            DVM_NCCL(n->GroupEnd());
            DVM_CUDA(cudaMemcpyAsync(x->h_counts + W, x->d_counts + W, sizeof(long long) * W, cudaMemcpyDeviceToHost, x->stream));
            DVM_CUDA(cudaStreamSynchronize(x->stream));
            for (int p = 0; p < W; p++) recv[p] = p == x->rank ? 0 : x->h_counts[W + p];
        }
    }
    // ---- 3. blocks ----
    size_t total_recv = 0, total_send = 0;
    for (int p = 0; p < W; p++) { total_recv += (size_t)recv[p]; total_send += (size_t)send[p]; }
    if (total_recv * row > x->recv_cap) {
        const size_t capb = total_recv * row + total_recv * row / 4 + 1024;
        if (x->device < 0) { free(x->recvbuf); x->recvbuf = (uint8_t*)malloc(capb); }
        else {
            DVM_CUDA(cudaStreamSynchronize(x->stream));
            cudaFree(x->recvbuf); x->recvbuf = nullptr;
            DVM_CUDA(cudaMalloc(&x->recvbuf, capb));
        }
        x->recv_cap = capb;
    }
    if (W > 1 && (total_recv || total_send)) {
        if (x->hooked) {
            // the hook takes ONE contiguous send buffer with per-peer sizes: pack the per-peer ranges
            std::vector<size_t> sb(W), rb(W);
            std::vector<uint8_t> pack(total_send * row);
            size_t o = 0;
            for (int p = 0; p < W; p++) {
                sb[p] = (size_t)send[p] * row; rb[p] = (size_t)recv[p] * row;
                if (send[p]) memcpy(pack.data() + o, x->db + (size_t)x->sent_upto[p] * row, sb[p]);
                o += sb[p];
            }
            if (x->hooks.alltoall(x->hooks.ctx, pack.data(), sb.data(), x->recvbuf, rb.data()) != 0) { set_error("alltoall hook failed"); return DVM_ERR_INVALID; }
        } else {
            NcclApi* n = nccl();
            DVM_NCCL(n->GroupStart());
            size_t o = 0;
            for (int p = 0; p < W; p++) {
                if (send[p]) DVM_NCCL(n->Send(x->db + (size_t)x->sent_upto[p] * row, (size_t)send[p] * row, kNcclUint8, p, x->comm, x->stream));
                if (recv[p]) DVM_NCCL(n->Recv(x->recvbuf + o, (size_t)recv[p] * row, kNcclUint8, p, x->comm, x->stream));
                o += (size_t)recv[p] * row;
            }
            DVM_NCCL(n->GroupEnd());
        }
    }
    x->last_bytes_sent = total_send * row;
    for (int p = 0; p < W; p++) x->sent_upto[p] += (int)send[p];
    // ---- 4. match every received block against the local database, 5. report ----
    struct Cand { int peer, peer_kf, own_kf, count; };
    std::vector<Cand> out;
    std::vector<int32_t> counts;
    size_t o = 0;
    for (int p = 0; p < W; p++) {
        const int k = (int)recv[p];
        if (k == 0) continue;
        const int first = x->recv_upto[p];   // ids in the sender's numbering: what it had sent us before this round
        x->recv_upto[p] += k;
        const uint8_t* blocks = x->recvbuf + o * row;
        o += (size_t)k;
        if (x->n_kf == 0) continue;
        counts.assign((size_t)k * x->n_kf, 0);
        const int rc = dvm_exchange_match_counts(x, blocks, k, counts.data());
        if (rc != DVM_OK) return rc;
        for (int a = 0; a < k; a++)
            for (int b = 0; b < x->n_kf; b++) {
                const int c = counts[(size_t)a * x->n_kf + b];
                if (c >= x->min_matches) out.push_back({ p, first + a, b, c });
            }
    }
    std::sort(out.begin(), out.end(), [](const Cand& a, const Cand& b) {
        if (a.count != b.count) return a.count > b.count;
        if (a.peer != b.peer) return a.peer < b.peer;
        if (a.peer_kf != b.peer_kf) return a.peer_kf < b.peer_kf;
        return a.own_kf < b.own_kf;
    });
    *n_candidates = (int)out.size();
    for (int i = 0; i < (int)out.size() && i < cap; i++) {
        candidates[4 * i] = out[i].peer; candidates[4 * i + 1] = out[i].peer_kf;
        candidates[4 * i + 2] = out[i].own_kf; candidates[4 * i + 3] = out[i].count;
    }
    return DVM_OK;
}

} // extern "C"
