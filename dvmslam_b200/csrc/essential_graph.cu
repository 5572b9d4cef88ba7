// essential_graph.cu -- the solve of Optimizer::OptimizeEssentialGraph (O3/src/Optimizer.cc:1389-1651) on one B200.
//
// A pose graph over Sim3 keyframe vertices (VertexSim3Expmap) with EdgeSim3 constraints (information = identity, no
// robust kernel): error = log(Sji * Siw * Sjw^-1) (g2o/types/types_seven_dof_expmap.h:99-106), g2o's NUMERIC Jacobians
// for both vertices (central differences, delta 1e-9, through oplus; g2o/core/base_binary_edge.hpp:130-205),
// Levenberg-Marquardt with setUserLambdaInit(1e-16) (:1400), optimize(20) (:1594).  The host adapter flattens the graph
// (spanning tree, loop edges, covisibility >= minFeat edges, :1423-1592) and applies the corrected poses / map points.
//
// One persistent cooperative kernel, the whole LM loop on the device (the control flow is evaluated identically by every
// thread from deterministic partial sums, as in lba.cu).  Per LM iteration:
//   E   errors and chi2 at the estimate                                         one thread per edge
//   J   numeric Jacobian columns: (edge, side, dimension) -> 7 values           one thread per column, 2 error evaluations
//   H   H += J^T J blocks (lower triangle), b += J^T (-e), FP64 atomics at L2   one warp per edge
//   per trial: Hs = H + lambda I | dense blocked Cholesky of the 7 nf x 7 nf system over the WHOLE grid
//   (grid_cholesky_solve, chol_device.cuh: DMMA trailing updates) | oplus into the trial buffer | errors at the trial
// The reference solves with a sparse LDL^T (LinearSolverEigen); the system is dense here -- at 400 keyframes it is
// 2800 x 2800 = 63 MB and ~7 GFLOP per factorisation, which is where the FP64 tensor pipe has something to do.
#include "chol_device.cuh"
#include "sim3_math.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace cg = cooperative_groups;
using namespace dvm;
using namespace dvm::sim3m;

namespace {

constexpr int kEgThreads = kCholThreads;
constexpr int kEgWarps = kEgThreads / 32;
// per-CTA partial sums, one slot per writer (see lba.cu: a shared slot was a race)
constexpr int kEgPartStride = 4;
constexpr int kEgChi = 0, kEgTrialChi = 1, kEgScale = 2, kEgMaxDiag = 3;

struct EgDev {
    int nv, nf, ne, dim, dimPad, iterations, fix_scale;
    double lambda_init;
    double* V[2];          // [nv][8] q (x,y,z,w), t, s: estimate / trial
    const int* col;        // [nv] free-vertex index or -1
    const int* vi; const int* vj;
    const double* meas;    // [ne][8]
    double* err;           // [ne][7]
    double* J;             // [ne][2][49] row-major 7 x 7: d(error row) / d(update column)
    double* H; double* b;  // [dimPad][dimPad] (lower triangle used), [dimPad]
    double* Hs; double* bs; double* x; double* Linv;
    double* part; int* flags;
    double* out_stats;     // [4] LM iterations, trials, initial chi2, final chi2
};

__device__ inline Sim3 load_sim3(const double* p)
{
    Sim3 S;
    S.r = { p[0], p[1], p[2], p[3] };
    S.t[0] = p[4]; S.t[1] = p[5]; S.t[2] = p[6];
    S.s = p[7];
    return S;
}
__device__ inline void store_sim3(double* p, const Sim3& S)
{
    p[0] = S.r.x; p[1] = S.r.y; p[2] = S.r.z; p[3] = S.r.w;
    p[4] = S.t[0]; p[5] = S.t[1]; p[6] = S.t[2];
    p[7] = S.s;
}
// VertexSim3Expmap::oplusImpl: exp(update) * estimate, the scale component frozen when _fix_scale
__device__ inline Sim3 eg_oplus(int fix_scale, const Sim3& S, const double* update)
{
    double u[7];
#pragma unroll
    for (int k = 0; k < 7; k++) u[k] = update[k];
    if (fix_scale) u[6] = 0;
    return sim3_mul(sim3_exp(u), S);
}
// EdgeSim3::computeError: log(measurement * Siw * Sjw^-1)
__device__ inline void edge_error(const Sim3& M, const Sim3& Si, const Sim3& Sj, double err[7])
{
    sim3_log(sim3_mul(sim3_mul(M, Si), sim3_inverse(Sj)), err);
}

__device__ inline double cta_sum1(double v, double* warp_buf)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) warp_buf[wid] = v;
    __syncthreads();
    double s = 0;
    for (int w = 0; w < kEgWarps; w++) s += warp_buf[w];
    return s;
}

__global__ void __launch_bounds__(kEgThreads, 1) essential_graph_kernel(EgDev P)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double smem[];
    __shared__ double warp_buf[kEgWarps];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x;
    const int gtid = blockIdx.x * kEgThreads + tid, nthreads = G * kEgThreads;
    const int gwarp = blockIdx.x * kEgWarps + wid, nwarps = G * kEgWarps;
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);

    int cur = 0;
    double lambda = -1, ni = 2, first_chi = 0, last_chi = 0;
    int nBad = 0, done = 0, trials = 0;
    bool stop = false;

    auto errors_at = [&](int buf, int slot) {   // errors of every edge at V[buf], chi2 partial into `slot`
        double chi = 0;
        for (int e = gtid; e < P.ne; e += nthreads) {
            double er[7];
            edge_error(load_sim3(P.meas + 8 * e), load_sim3(P.V[buf] + 8 * P.vi[e]), load_sim3(P.V[buf] + 8 * P.vj[e]), er);
#pragma unroll
            for (int k = 0; k < 7; k++) { P.err[7 * e + k] = er[k]; chi += er[k] * er[k]; }
        }
        chi = cta_sum1(chi, warp_buf);
        if (tid == 0) P.part[blockIdx.x * kEgPartStride + slot] = chi;
    };
    auto total = [&](int slot) {
        double s = 0;
        for (int c = 0; c < G; c++) s += P.part[c * kEgPartStride + slot];
        return s;
    };

    for (int it = 0; it < P.iterations && !stop; it++) {
        // ---------------- E: errors at the estimate; clear H and b ----------------
        errors_at(cur, kEgChi);
        for (size_t i = gtid; i < (size_t)P.dimPad * P.dimPad; i += nthreads) P.H[i] = 0.0;
        for (int i = gtid; i < P.dimPad; i += nthreads) P.b[i] = 0.0;
        // ---------------- J: numeric Jacobian columns ----------------
        for (int w = gtid; w < P.ne * 14; w += nthreads) {
            const int e = w / 14, r = w - e * 14, side = r / 7, d = r - side * 7;
            const int a = P.vi[e], c = P.vj[e], v = side == 0 ? a : c;
            if (P.col[v] < 0) continue;   // fixed vertex: no Jacobian (g2o skips it)
            const Sim3 M = load_sim3(P.meas + 8 * e), Sa = load_sim3(P.V[cur] + 8 * a), Sc = load_sim3(P.V[cur] + 8 * c);
            double add[7] = { 0, 0, 0, 0, 0, 0, 0 }, ep[7], em[7];
            add[d] = delta;
            const Sim3 Vp = eg_oplus(P.fix_scale, side == 0 ? Sa : Sc, add);
            add[d] = -delta;
            const Sim3 Vm = eg_oplus(P.fix_scale, side == 0 ? Sa : Sc, add);
            if (side == 0) { edge_error(M, Vp, Sc, ep); edge_error(M, Vm, Sc, em); }
            else { edge_error(M, Sa, Vp, ep); edge_error(M, Sa, Vm, em); }
            double* Jd = P.J + ((size_t)e * 2 + side) * 49;
#pragma unroll
            for (int k = 0; k < 7; k++) Jd[k * 7 + d] = scalar * (ep[k] - em[k]);
        }
        grid.sync();
        // ---------------- H: normal equations, one warp per edge ----------------
        double maxd = 0;
        for (int e = gwarp; e < P.ne; e += nwarps) {
            const int a = P.vi[e], c = P.vj[e], ca = P.col[a], cc = P.col[c];
            const double* Ji = P.J + (size_t)e * 98;
            const double* Jj = Ji + 49;
            const double* er = P.err + 7 * e;
            auto add_block = [&](int rb, int cb, const double* Jr, const double* Jc, bool lower_only) {
                for (int idx = lane; idx < 49; idx += 32) {
                    const int p = idx / 7, q = idx - p * 7;
                    if (lower_only && q > p) continue;
                    double s = 0;
#pragma unroll
                    for (int r = 0; r < 7; r++) s += Jr[r * 7 + p] * Jc[r * 7 + q];
                    atomicAdd(&P.H[(size_t)(7 * rb + p) * P.dimPad + 7 * cb + q], s);
                }
            };
            auto add_rhs = [&](int cb, const double* Jr) {
                if (lane < 7) {
                    double s = 0;
#pragma unroll
                    for (int r = 0; r < 7; r++) s += Jr[r * 7 + lane] * (-er[r]);
                    atomicAdd(&P.b[7 * cb + lane], s);
                }
            };
            if (ca >= 0) { add_block(ca, ca, Ji, Ji, true); add_rhs(ca, Ji); }
            if (cc >= 0) { add_block(cc, cc, Jj, Jj, true); add_rhs(cc, Jj); }
            if (ca >= 0 && cc >= 0 && a != c) {   // off-diagonal block, stored below the diagonal
                if (ca > cc) add_block(ca, cc, Ji, Jj, false); else add_block(cc, ca, Jj, Ji, false);
            }
        }
        grid.sync();
        if (P.lambda_init <= 0 && it == 0) {   // computeLambdaInit without a user value: 1e-5 * max diag(H)
            for (int i = gtid; i < P.dim; i += nthreads) maxd = fmax(maxd, fabs(P.H[(size_t)i * P.dimPad + i]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) maxd = fmax(maxd, __shfl_xor_sync(0xffffffffu, maxd, o));
            __syncthreads();
            if (lane == 0) warp_buf[wid] = maxd;
            __syncthreads();
            if (tid == 0) {
                double m = 0;
                for (int w = 0; w < kEgWarps; w++) m = fmax(m, warp_buf[w]);
                P.part[blockIdx.x * kEgPartStride + kEgMaxDiag] = m;
            }
            grid.sync();
        }
        double currentChi = total(kEgChi);
        if (it == 0) {
            first_chi = currentChi;
            if (P.lambda_init > 0) lambda = P.lambda_init;
            else {
                double m = 0;
                for (int c = 0; c < G; c++) m = fmax(m, P.part[c * kEgPartStride + kEgMaxDiag]);
                lambda = 1e-5 * m;
            }
            ni = 2; nBad = 0;
        }
        const double iniChi = currentChi;
        double rho = 0;
        int qmax = 0;
        do {
            const int trial = cur ^ 1;
            // ---------------- Hs = H + lambda I (lower triangle; identity on the padding) ----------------
            for (size_t i = gtid; i < (size_t)P.dimPad * P.dimPad; i += nthreads) {
                const int r = (int)(i / P.dimPad), c = (int)(i - (size_t)r * P.dimPad);
                double v = 0.0;
                if (r >= P.dim) v = (r == c) ? 1.0 : 0.0;
                else if (c <= r) v = P.H[i] + (r == c ? lambda : 0.0);
                P.Hs[i] = v;
            }
            for (int i = gtid; i < P.dimPad; i += nthreads) P.bs[i] = i < P.dim ? P.b[i] : 0.0;
            grid.sync();
            // ---------------- dense Cholesky solve over the whole grid ----------------
            grid_cholesky_solve(grid, P.dimPad, P.Hs, P.bs, P.x, P.Linv, &P.flags[0], smem);
            grid.sync();
            const bool ok2 = P.flags[0] != 0;
            // ---------------- update into the trial buffer ----------------
            double sc = 0;
            for (int v = gtid; v < P.nv; v += nthreads) {
                const Sim3 S = load_sim3(P.V[cur] + 8 * v);
                const int cf = P.col[v];
                if (cf >= 0) {
                    const double* xv = P.x + 7 * cf;
                    store_sim3(P.V[trial] + 8 * v, eg_oplus(P.fix_scale, S, xv));
#pragma unroll
                    for (int k = 0; k < 7; k++) sc += xv[k] * (lambda * xv[k] + P.b[7 * cf + k]);
                } else store_sim3(P.V[trial] + 8 * v, S);
            }
            sc = cta_sum1(sc, warp_buf);
            if (tid == 0) P.part[blockIdx.x * kEgPartStride + kEgScale] = sc;
            grid.sync();
            // ---------------- errors at the trial estimate ----------------
            errors_at(trial, kEgTrialChi);
            grid.sync();
            double tempChi = total(kEgTrialChi), scale = total(kEgScale);
            if (!ok2) tempChi = 1.7976931348623157e308;
            rho = currentChi - tempChi;
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow((2 * rho - 1), 3.0);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
                cur = trial;
            } else {
                lambda *= ni;
                ni *= 2;
            }
            qmax++;
            trials++;
        } while (rho < 0 && qmax < 10);
        done++;
        last_chi = currentChi;
        if (qmax == 10 || rho == 0) stop = true;
        else {
            if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
            else nBad = 0;
            if (nBad >= 3) stop = true;
        }
        grid.sync();   // part[kEgChi] is rewritten by the next iteration
    }
    // the result sits in V[cur]; make it V[0]
    if (cur != 0)
        for (int i = gtid; i < P.nv * 8; i += nthreads) P.V[0][i] = P.V[1][i];
    if (gtid == 0) { P.out_stats[0] = done; P.out_stats[1] = trials; P.out_stats[2] = first_chi; P.out_stats[3] = last_chi; }
}

} // namespace

struct dvm_essential_graph {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int grid = 0;
    size_t smem_bytes = 0;
    uint8_t* d_buf = nullptr; size_t d_cap = 0;
    float last_ms = 0;
};

static void eg_free(dvm_essential_graph* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_buf);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" {

int dvm_essential_graph_create(dvm_essential_graph** out, int device)
{
    DVM_REQUIRE(out != nullptr, "null output handle");
    *out = nullptr;
    int rc = select_device(device);
    if (rc != DVM_OK) return rc;
    dvm_essential_graph* h = new dvm_essential_graph;
    h->device = device;
#define DVM_ECREATE(call)                                                                            \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            set_error("%s failed in dvm_essential_graph_create: %s", #call, cudaGetErrorString(e__)); \
            eg_free(h);                                                                              \
            return DVM_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)
    DVM_ECREATE(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    DVM_ECREATE(cudaEventCreate(&h->ev0));
    DVM_ECREATE(cudaEventCreate(&h->ev1));
    h->smem_bytes = (size_t)kGridCholSmemDoubles * sizeof(double);
    DVM_ECREATE(cudaFuncSetAttribute(essential_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    int per_sm = 0, sms = 0;
    DVM_ECREATE(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, essential_graph_kernel, kEgThreads, h->smem_bytes));
    DVM_ECREATE(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
#undef DVM_ECREATE
    if (per_sm < 1) { set_error("essential_graph_kernel does not fit on an SM"); eg_free(h); return DVM_ERR_CUDA; }
    h->grid = sms;   // one persistent CTA per SM
    *out = h;
    return DVM_OK;
}

void dvm_essential_graph_destroy(dvm_essential_graph* h) { eg_free(h); }
float dvm_essential_graph_last_kernel_ms(const dvm_essential_graph* h) { return h ? h->last_ms : -1.f; }

int dvm_optimize_essential_graph(dvm_essential_graph* h, int nv, double* sim3, const uint8_t* fixed, int ne, const int32_t* vi,
                                 const int32_t* vj, const double* meas, int fix_scale, int iterations, double lambda_init,
                                 double* stats)
{
    DVM_REQUIRE(h != nullptr && nv >= 0 && ne >= 0 && iterations >= 0, "bad argument");
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (nv == 0 || ne == 0 || iterations == 0) return DVM_OK;
    DVM_REQUIRE(sim3 && fixed && vi && vj && meas, "null arrays");
    std::vector<int> col(nv, -1);
    int nf = 0;
    for (int v = 0; v < nv; v++) if (!fixed[v]) col[v] = nf++;
    for (int e = 0; e < ne; e++) DVM_REQUIRE((unsigned)vi[e] < (unsigned)nv && (unsigned)vj[e] < (unsigned)nv, "edge vertex out of range");
    if (nf == 0) return DVM_OK;
    DVM_CUDA(cudaSetDevice(h->device));
    const int dim = 7 * nf, dimPad = (dim + kNB - 1) / kNB * kNB;
    size_t off = 0;
    auto take = [&](size_t bytes) { off = (off + 255) & ~(size_t)255; size_t o = off; off += bytes; return o; };
    const size_t o_v0 = take((size_t)nv * 64), o_v1 = take((size_t)nv * 64), o_col = take((size_t)nv * 4);
    const size_t o_vi = take((size_t)ne * 4), o_vj = take((size_t)ne * 4), o_meas = take((size_t)ne * 64);
    const size_t o_err = take((size_t)ne * 56), o_J = take((size_t)ne * 98 * 8);
    const size_t o_H = take((size_t)dimPad * dimPad * 8), o_b = take((size_t)dimPad * 8);
    const size_t o_Hs = take((size_t)(dimPad + 8) * dimPad * 8), o_bs = take((size_t)dimPad * 8), o_x = take((size_t)dimPad * 8);
    const size_t o_linv = take((size_t)dimPad * kNB * 8);
    const size_t o_part = take((size_t)h->grid * kEgPartStride * 8), o_flags = take(16), o_stats = take(32);
    const size_t total = off + 256;
    if (total > h->d_cap) {
        DVM_CUDA(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_buf); h->d_buf = nullptr;
        DVM_CUDA(cudaMalloc(&h->d_buf, total + total / 8));
        h->d_cap = total + total / 8;
    }
    uint8_t* db = h->d_buf;
    DVM_CUDA(cudaMemcpyAsync(db + o_v0, sim3, (size_t)nv * 64, cudaMemcpyHostToDevice, h->stream));
    DVM_CUDA(cudaMemcpyAsync(db + o_col, col.data(), (size_t)nv * 4, cudaMemcpyHostToDevice, h->stream));
    DVM_CUDA(cudaMemcpyAsync(db + o_vi, vi, (size_t)ne * 4, cudaMemcpyHostToDevice, h->stream));
    DVM_CUDA(cudaMemcpyAsync(db + o_vj, vj, (size_t)ne * 4, cudaMemcpyHostToDevice, h->stream));
    DVM_CUDA(cudaMemcpyAsync(db + o_meas, meas, (size_t)ne * 64, cudaMemcpyHostToDevice, h->stream));
    DVM_CUDA(cudaMemsetAsync(db + o_J, 0, (size_t)ne * 98 * 8, h->stream));   // Jacobians of fixed vertices are never written
    DVM_CUDA(cudaMemsetAsync(db + o_flags, 0, 16, h->stream));
    EgDev P;
    memset(&P, 0, sizeof(P));
    P.nv = nv; P.nf = nf; P.ne = ne; P.dim = dim; P.dimPad = dimPad; P.iterations = iterations; P.fix_scale = fix_scale;
    P.lambda_init = lambda_init;
    P.V[0] = (double*)(db + o_v0); P.V[1] = (double*)(db + o_v1); P.col = (const int*)(db + o_col);
    P.vi = (const int*)(db + o_vi); P.vj = (const int*)(db + o_vj); P.meas = (const double*)(db + o_meas);
    P.err = (double*)(db + o_err); P.J = (double*)(db + o_J); P.H = (double*)(db + o_H); P.b = (double*)(db + o_b);
    P.Hs = (double*)(db + o_Hs); P.bs = (double*)(db + o_bs); P.x = (double*)(db + o_x); P.Linv = (double*)(db + o_linv);
    P.part = (double*)(db + o_part); P.flags = (int*)(db + o_flags); P.out_stats = (double*)(db + o_stats);
    void* args[] = { &P };
    DVM_CUDA(cudaEventRecord(h->ev0, h->stream));
    DVM_CUDA(cudaLaunchCooperativeKernel((void*)essential_graph_kernel, dim3(h->grid), dim3(kEgThreads), args, h->smem_bytes, h->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    DVM_CUDA(cudaEventRecord(h->ev1, h->stream));
    double st[4];
    DVM_CUDA(cudaMemcpyAsync(sim3, db + o_v0, (size_t)nv * 64, cudaMemcpyDeviceToHost, h->stream));
    DVM_CUDA(cudaMemcpyAsync(st, db + o_stats, 32, cudaMemcpyDeviceToHost, h->stream));
    DVM_CUDA(cudaStreamSynchronize(h->stream));
    DVM_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    if (stats) for (int i = 0; i < 4; i++) stats[i] = st[i];
    if (!std::isfinite(st[3])) { set_error("essential-graph optimisation produced a non-finite chi2"); return DVM_ERR_NUMERIC; }
    return DVM_OK;
}

} // extern "C"
