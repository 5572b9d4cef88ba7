// sim3.cu -- Optimizer::OptimizeSim3 (O3/src/Optimizer.cc:1960-2212) as one sm_100a kernel.
//
// One Sim3 vertex (7 DoF, or 6 with bFixScale), fixed point vertices, an EdgeSim3ProjectXYZ / EdgeInverseSim3ProjectXYZ pair
// per correspondence (O3/include/OptimizableTypes.h:146-206).  The edges define no linearizeOplus, so g2o differentiates
// them numerically (base_binary_edge.hpp:130-205: central differences, delta 1e-9, through oplus on the vertex); the
// kernel does the same -- the 14 perturbed estimates are shared by all edges, so they are built once per linearisation --
// because an analytic Jacobian would follow a (slightly) different LM path than the reference.
//
// A few hundred correspondences and a 7x7 system: latency-bound, one CTA.  Every thread carries the estimate and the LM
// state in registers and derives them from the same shared-memory sums, so the control flow is uniform without
// broadcasting; the numeric Jacobians (28 projections per correspondence and linearisation) are spread over the threads one
// (correspondence, dimension) at a time, the normal equations are accumulated one correspondence per thread and reduced
// with warp shuffles + one shared-memory stage in a fixed order (deterministic).
#include "sim3_math.cuh"
#include "common.cuh"
#include <math_constants.h>
#include <cmath>

namespace dvm {
using namespace sim3m;
namespace {

constexpr int kSim3Threads = 512;
constexpr int kSim3Warps = kSim3Threads / 32;


struct Sim3Dev {
    int n, fix_scale;
    const float* p1; const float* p2; const float* obs1; const float* obs2; const float* w1; const float* w2;
    double K1[4], K2[4];
    double delta, th2;
    double* err;        // [n*4] the edges' _error: e12 (2), e21 (2)
    double* jac;        // [n*28] numeric Jacobians of the pair's four error components (row-major 4 x 7)
    uint8_t* state;     // [n] bit 0: pair still in the graph, bit 1: still carries its Huber kernel
    double* io;         // [8] q (x,y,z,w), t, s: in/out
    uint8_t* inlier;    // [n] out
    double* stats;      // [8] out: iters pass 1, iters pass 2, trials, nBad, first chi2, last chi2, nIn
};

// VertexSim3Expmap::oplusImpl
__device__ inline Sim3 oplus(const Sim3Dev& P, const Sim3& S, const double* update)
{
    double u[7];
#pragma unroll
    for (int k = 0; k < 7; k++) u[k] = update[k];
    if (P.fix_scale) u[6] = 0;
    return sim3_mul(sim3_exp(u), S);
}

struct Pair { double p1[3], p2[3], o1[2], o2[2]; };
__device__ inline Pair load_pair(const Sim3Dev& P, int i)
{
    Pair c;
#pragma unroll
    for (int k = 0; k < 3; k++) { c.p1[k] = (double)P.p1[3 * i + k]; c.p2[k] = (double)P.p2[3 * i + k]; }
#pragma unroll
    for (int k = 0; k < 2; k++) { c.o1[k] = (double)P.obs1[2 * i + k]; c.o2[k] = (double)P.obs2[2 * i + k]; }
    return c;
}
// EdgeSim3ProjectXYZ::computeError (e[0..1]) and EdgeInverseSim3ProjectXYZ::computeError (e[2..3])
__device__ inline void pair_error(const Sim3Dev& P, const Sim3& S, const Sim3& Sinv, const Pair& c, double e[4])
{
    double rx[3], x[3];
    quat_rotate(S.r, c.p2, rx);
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = S.s * rx[k] + S.t[k];
    e[0] = c.o1[0] - (P.K1[0] * x[0] / x[2] + P.K1[2]);
    e[1] = c.o1[1] - (P.K1[1] * x[1] / x[2] + P.K1[3]);
    quat_rotate(Sinv.r, c.p1, rx);
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = Sinv.s * rx[k] + Sinv.t[k];
    e[2] = c.o2[0] - (P.K2[0] * x[0] / x[2] + P.K2[2]);
    e[3] = c.o2[1] - (P.K2[1] * x[1] / x[2] + P.K2[3]);
}

// deterministic CTA sum of NV doubles per thread; the totals land in out[0..NV) (shared) for every thread to read
template <int NV>
__device__ inline void cta_sum(double* v, double* warp_buf /*[kSim3Warps*NV]*/, double* out /*[NV]*/)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        v[i] = x;
    }
    __syncthreads(); // the previous totals have been read
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; i++) warp_buf[wid * NV + i] = v[i];
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < kSim3Warps; w++) s += warp_buf[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// symmetric 7x7 solve by LDL^T (LinearSolverDense: Eigen::LDLT, solution only if positive)
__device__ inline bool solve7(const double* H /*shared, full 7x7*/, double lambda, const double* b, double* x)
{
    double L[49], D[7];
#pragma unroll
    for (int j = 0; j < 7; j++) {
        double d = H[j * 7 + j] + lambda;
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[j * 7 + k] * L[j * 7 + k] * D[k];
        if (!(d > 0) || !isfinite(d)) return false;
        D[j] = d;
#pragma unroll
        for (int i = j + 1; i < 7; i++) {
            double v = H[i * 7 + j];
#pragma unroll
            for (int k = 0; k < j; k++) v -= L[i * 7 + k] * L[j * 7 + k] * D[k];
            L[i * 7 + j] = v / d;
        }
    }
    double y[7];
#pragma unroll
    for (int i = 0; i < 7; i++) {
        double v = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) v -= L[i * 7 + k] * y[k];
        y[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 7; i++) y[i] /= D[i];
#pragma unroll
    for (int i = 6; i >= 0; i--) {
        double v = y[i];
#pragma unroll
        for (int k = i + 1; k < 7; k++) v -= L[k * 7 + i] * x[k];
        x[i] = v;
    }
    return true;
}

__global__ void __launch_bounds__(kSim3Threads, 1) optimize_sim3_kernel(Sim3Dev P)
{
    __shared__ double warp_buf[kSim3Warps * 35];
    __shared__ double tot[35];
    __shared__ Sim3 pert[28]; // [d]: +delta, [7+d]: its inverse, [14+d]: -delta, [21+d]: its inverse
    __shared__ double Hs[49], bs[7];
    const int tid = threadIdx.x;
    const double dsqr = P.delta * P.delta;

    Sim3 S;
    S.r = { P.io[0], P.io[1], P.io[2], P.io[3] };
    S.t[0] = P.io[4]; S.t[1] = P.io[5]; S.t[2] = P.io[6];
    S.s = P.io[7];
    for (int i = tid; i < P.n; i += kSim3Threads) { P.state[i] = 3; P.inlier[i] = 1; }
    __syncthreads();

    auto chi_of = [&](const double* e, double om) { return e[0] * (om * e[0]) + e[1] * (om * e[1]); };
    // computeActiveErrors + activeRobustChi2 at estimate T: the pairs' errors go to P.err, the robust chi2 to every thread
    auto errors_and_chi = [&](const Sim3& T) -> double {
        const Sim3 Tinv = sim3_inverse(T);
        double part[1] = { 0 };
        for (int i = tid; i < P.n; i += kSim3Threads) {
            const uint8_t st = P.state[i];
            if (!(st & 1)) continue;
            const Pair c = load_pair(P, i);
            double e[4];
            pair_error(P, T, Tinv, c, e);
#pragma unroll
            for (int k = 0; k < 4; k++) P.err[4 * i + k] = e[k];
            const double c1 = chi_of(e, (double)P.w1[i]), c2 = chi_of(e + 2, (double)P.w2[i]);
            const bool rob = st & 2;
            part[0] += (!rob || c1 <= dsqr) ? c1 : 2 * sqrt(c1) * P.delta - dsqr;
            part[0] += (!rob || c2 <= dsqr) ? c2 : 2 * sqrt(c2) * P.delta - dsqr;
        }
        cta_sum<1>(part, warp_buf, tot);
        return tot[0];
    };

    int trials = 0;
    double first_chi = 0, last_chi = 0;
    // optimizer.optimize(iterations): g2o's Levenberg-Marquardt (optimization_algorithm_levenberg.cpp:59-188)
    auto optimize = [&](int iterations, bool record_first) -> int {
        double lambda = -1, ni = 2;
        int nBad = 0, done = 0;
        for (int it = 0; it < iterations; it++) {
            double currentChi = errors_and_chi(S);
            const double iniChi = currentChi;
            if (it == 0 && record_first) first_chi = currentChi;
            // the 14 perturbed estimates of the numeric Jacobian and their inverses
            if (tid < 14) {
                const int d = tid % 7;
                double add[7] = { 0, 0, 0, 0, 0, 0, 0 };
                add[d] = tid < 7 ? 1e-9 : -1e-9;
                const Sim3 T = oplus(P, S, add);
                pert[(tid < 7 ? 0 : 14) + d] = T;
                pert[(tid < 7 ? 7 : 21) + d] = sim3_inverse(T);
            }
            __syncthreads();
            // buildSystem, phase A: one (pair, dimension) per work item -- the pair's four error components at the two
            // perturbed estimates of that dimension give one column of its Jacobians
            const double scalar = 1.0 / (2 * 1e-9);
            for (int w = tid; w < P.n * 7; w += kSim3Threads) {
                const int i = w / 7, d = w - 7 * i;
                if (!(P.state[i] & 1)) continue;
                const Pair c = load_pair(P, i);
                double ep[4], em[4];
                pair_error(P, pert[d], pert[7 + d], c, ep);
                pair_error(P, pert[14 + d], pert[21 + d], c, em);
#pragma unroll
                for (int r = 0; r < 4; r++) P.jac[(size_t)i * 28 + r * 7 + d] = scalar * (ep[r] - em[r]);
            }
            __syncthreads();
            // phase B: one pair per thread -- H (upper triangle, 28) and b (7)
            double acc[35];
#pragma unroll
            for (int k = 0; k < 35; k++) acc[k] = 0;
            for (int i = tid; i < P.n; i += kSim3Threads) {
                const uint8_t st = P.state[i];
                if (!(st & 1)) continue;
                const bool rob = st & 2;
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    double J0[7], J1[7];
#pragma unroll
                    for (int a = 0; a < 7; a++) {
                        J0[a] = P.jac[(size_t)i * 28 + (2 * half) * 7 + a];
                        J1[a] = P.jac[(size_t)i * 28 + (2 * half + 1) * 7 + a];
                    }
                    const double om = half == 0 ? (double)P.w1[i] : (double)P.w2[i];
                    const double e0 = P.err[4 * i + 2 * half], e1 = P.err[4 * i + 2 * half + 1];
                    const double chi = e0 * (om * e0) + e1 * (om * e1);
                    const double w = (!rob || chi <= dsqr) ? 1.0 : P.delta / sqrt(chi);
                    const double r0 = -om * e0 * w, r1 = -om * e1 * w;
                    const double wo = w * om;
                    int idx = 0;
#pragma unroll
                    for (int a = 0; a < 7; a++) {
                        acc[28 + a] += J0[a] * r0 + J1[a] * r1;
#pragma unroll
                        for (int b2 = a; b2 < 7; b2++) acc[idx++] += J0[a] * wo * J0[b2] + J1[a] * wo * J1[b2];
                    }
                }
            }
            cta_sum<35>(acc, warp_buf, tot);
            if (tid < 49) {
                const int a = tid / 7, b2 = tid % 7;
                const int lo = min(a, b2), hi = max(a, b2);
                Hs[tid] = tot[lo * 7 - lo * (lo - 1) / 2 + (hi - lo)];
            }
            if (tid < 7) bs[tid] = tot[28 + tid];
            __syncthreads();
            if (it == 0) {
                double mx = 0;
#pragma unroll
                for (int j = 0; j < 7; j++) mx = fmax(fabs(Hs[j * 8]), mx);
                lambda = 1e-5 * mx;
                ni = 2;
                nBad = 0;
            }
            double bl[7];
#pragma unroll
            for (int j = 0; j < 7; j++) bl[j] = bs[j];
            double rho = 0;
            int qmax = 0;
            do {
                double x[7];
                const bool ok2 = solve7(Hs, lambda, bl, x);
                if (!ok2) {
#pragma unroll
                    for (int j = 0; j < 7; j++) x[j] = 0;
                }
                const Sim3 T = oplus(P, S, x);
                double tempChi = errors_and_chi(T);
                if (!ok2) tempChi = 1.7976931348623157e308;
                rho = currentChi - tempChi;
                double scale = 0;
#pragma unroll
                for (int j = 0; j < 7; j++) scale += x[j] * (lambda * x[j] + bl[j]);
                scale += 1e-3;
                rho /= scale;
                if (rho > 0 && isfinite(tempChi)) {
                    double alpha = 1. - pow((2 * rho - 1), 3.0);
                    alpha = fmin(alpha, 2. / 3.);
                    lambda *= fmax(1. / 3., alpha);
                    ni = 2;
                    currentChi = tempChi;
                    S = T;
                } else {
                    lambda *= ni;
                    ni *= 2;
                }
                qmax++;
                trials++;
            } while (rho < 0 && qmax < 10);
            done++;
            last_chi = currentChi;
            if (qmax == 10 || rho == 0) break;
            if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
            else nBad = 0;
            if (nBad >= 3) break;
        }
        return done;
    };

    const int it1 = optimize(5, true);                                  // :2149-2150
    // inlier check on the stored errors, :2153-2176: outlier pairs leave the graph, the others lose their kernel
    double cnt[1] = { 0 };
    for (int i = tid; i < P.n; i += kSim3Threads) {
        const double* e = P.err + 4 * i;
        const bool out = chi_of(e, (double)P.w1[i]) > P.th2 || chi_of(e + 2, (double)P.w2[i]) > P.th2;
        P.state[i] = out ? 0 : 1;
        if (out) { P.inlier[i] = 0; cnt[0] += 1; }
    }
    cta_sum<1>(cnt, warp_buf, tot);
    const int nBad1 = (int)tot[0];
    __syncthreads();
    int it2 = 0, nIn = 0;
    const bool enough = P.n - nBad1 >= 10;                              // :2183-2184
    if (enough) {
        it2 = optimize(nBad1 > 0 ? 10 : 5, false);                      // :2178-2188
        const Sim3 Sinv = sim3_inverse(S);
        cnt[0] = 0;
        for (int i = tid; i < P.n; i += kSim3Threads) {
            if (!(P.state[i] & 1)) continue;
            const Pair c = load_pair(P, i);
            double e[4];
            pair_error(P, S, Sinv, c, e);                               // e12->computeError(); e21->computeError();
            if (chi_of(e, (double)P.w1[i]) > P.th2 || chi_of(e + 2, (double)P.w2[i]) > P.th2) P.inlier[i] = 0;
            else cnt[0] += 1;
        }
        cta_sum<1>(cnt, warp_buf, tot);
        nIn = (int)tot[0];
    }
    if (tid == 0) {
        if (enough) {
            P.io[0] = S.r.x; P.io[1] = S.r.y; P.io[2] = S.r.z; P.io[3] = S.r.w;
            P.io[4] = S.t[0]; P.io[5] = S.t[1]; P.io[6] = S.t[2];
            P.io[7] = S.s;
        }
        P.stats[0] = it1; P.stats[1] = it2; P.stats[2] = trials; P.stats[3] = nBad1;
        P.stats[4] = first_chi; P.stats[5] = last_chi; P.stats[6] = nIn;
    }
}

} // namespace
} // namespace dvm

using namespace dvm;

struct dvm_sim3 {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint8_t* d_buf = nullptr;
    uint8_t* h_buf = nullptr; // pinned
    size_t cap = 0;
};

extern "C" {

int dvm_sim3_create(dvm_sim3** out, int device)
{
    DVM_REQUIRE(out != nullptr, "null argument");
    *out = nullptr;
    const int rc = select_device(device);
    if (rc != DVM_OK) return rc;
    dvm_sim3* h = new dvm_sim3;
    h->device = device;
    DVM_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    *out = h;
    return DVM_OK;
}

void dvm_sim3_destroy(dvm_sim3* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    cudaFree(h->d_buf);
    if (h->h_buf) cudaFreeHost(h->h_buf);
    delete h;
}

int dvm_optimize_sim3(dvm_sim3* h, int n, const float* p1c, const float* p2c, const float* obs1, const float* obs2,
                      const float* inv_sigma2_1, const float* inv_sigma2_2, const float* K1, const float* K2, double* s12_q,
                      double* s12_t, double* s12_s, float th2, int fix_scale, uint8_t* inlier, int* n_in, double* stats)
{
    DVM_REQUIRE(h != nullptr && n_in != nullptr, "null argument");
    *n_in = 0;
    if (stats) for (int k = 0; k < 6; k++) stats[k] = 0;
    DVM_REQUIRE(n >= 0, "negative size");
    DVM_REQUIRE(K1 && K2 && s12_q && s12_t && s12_s, "null argument");
    DVM_REQUIRE(n == 0 || (p1c && p2c && obs1 && obs2 && inv_sigma2_1 && inv_sigma2_2 && inlier), "null correspondence arrays");
    DVM_REQUIRE(th2 > 0.0f, "th2 must be positive");
    if (n == 0) return DVM_OK;   // no edge: optimize() does nothing and nCorrespondences - nBad < 10 returns 0
    DVM_CUDA(cudaSetDevice(h->device));
    size_t off = 0;
    auto take = [&](size_t bytes) { off = (off + 255) & ~(size_t)255; size_t o = off; off += bytes; return o; };
    const size_t o_p1 = take((size_t)n * 12), o_p2 = take((size_t)n * 12), o_o1 = take((size_t)n * 8), o_o2 = take((size_t)n * 8);
    const size_t o_w1 = take((size_t)n * 4), o_w2 = take((size_t)n * 4), o_io = take(8 * 8);
    const size_t upload = off;
    const size_t o_err = take((size_t)n * 32), o_state = take((size_t)n), o_jac = take((size_t)n * 28 * 8);
    const size_t out_begin = (off + 255) & ~(size_t)255;
    const size_t o_inl = take((size_t)n), o_stats = take(8 * 8), o_ioo = take(8 * 8);
    const size_t total = off + 256;
    if (total > h->cap) {
        DVM_CUDA(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_buf); h->d_buf = nullptr;
        if (h->h_buf) { cudaFreeHost(h->h_buf); h->h_buf = nullptr; }
        h->cap = 0;
        const size_t cap = total * 2;
        DVM_CUDA(cudaMalloc(&h->d_buf, cap));
        DVM_CUDA(cudaHostAlloc(&h->h_buf, cap, cudaHostAllocDefault));
        h->cap = cap;
    }
    uint8_t* hb = h->h_buf;
    memcpy(hb + o_p1, p1c, (size_t)n * 12); memcpy(hb + o_p2, p2c, (size_t)n * 12);
    memcpy(hb + o_o1, obs1, (size_t)n * 8); memcpy(hb + o_o2, obs2, (size_t)n * 8);
    memcpy(hb + o_w1, inv_sigma2_1, (size_t)n * 4); memcpy(hb + o_w2, inv_sigma2_2, (size_t)n * 4);
    double* io = reinterpret_cast<double*>(hb + o_io);
    for (int k = 0; k < 4; k++) io[k] = s12_q[k];
    for (int k = 0; k < 3; k++) io[4 + k] = s12_t[k];
    io[7] = *s12_s;
    DVM_CUDA(cudaMemcpyAsync(h->d_buf, hb, upload, cudaMemcpyHostToDevice, h->stream));
    uint8_t* db = h->d_buf;
    Sim3Dev P;
    memset(&P, 0, sizeof(P));
    P.n = n; P.fix_scale = fix_scale != 0;
    P.p1 = (const float*)(db + o_p1); P.p2 = (const float*)(db + o_p2);
    P.obs1 = (const float*)(db + o_o1); P.obs2 = (const float*)(db + o_o2);
    P.w1 = (const float*)(db + o_w1); P.w2 = (const float*)(db + o_w2);
    for (int k = 0; k < 4; k++) { P.K1[k] = K1[k]; P.K2[k] = K2[k]; }
    P.delta = (double)std::sqrt(th2);   // const float deltaHuber = sqrt(th2), :1997
    P.th2 = (double)th2;
    P.err = (double*)(db + o_err); P.state = db + o_state; P.jac = (double*)(db + o_jac);
    P.io = (double*)(db + o_ioo); P.inlier = db + o_inl; P.stats = (double*)(db + o_stats);
    DVM_CUDA(cudaMemcpyAsync(db + o_ioo, db + o_io, 64, cudaMemcpyDeviceToDevice, h->stream));
    DVM_CUDA(cudaMemsetAsync(db + o_err, 0, (size_t)n * 32, h->stream));
    DVM_LAUNCH(optimize_sim3_kernel, 1, kSim3Threads, 0, h->stream, P);
    DVM_CUDA(cudaGetLastError());
    DVM_CUDA(cudaMemcpyAsync(hb + out_begin, db + out_begin, off - out_begin, cudaMemcpyDeviceToHost, h->stream));
    DVM_CUDA(cudaStreamSynchronize(h->stream));
    const double* st = reinterpret_cast<const double*>(hb + o_stats);
    const double* oo = reinterpret_cast<const double*>(hb + o_ioo);
    memcpy(inlier, hb + o_inl, (size_t)n);
    for (int k = 0; k < 4; k++) s12_q[k] = oo[k];
    for (int k = 0; k < 3; k++) s12_t[k] = oo[4 + k];
    *s12_s = oo[7];
    *n_in = (int)st[6];
    if (stats) for (int k = 0; k < 6; k++) stats[k] = st[k];
    if (!std::isfinite(st[5])) { set_error("OptimizeSim3 produced a non-finite chi2"); return DVM_ERR_NUMERIC; }
    return DVM_OK;
}

} // extern "C"
