// lba.cu -- Optimizer::LocalBundleAdjustment on one B200 (O3/src/Optimizer.cc:1030-1387).
//
// One persistent cooperative kernel (one CTA per SM, grid.sync between phases) runs the whole of
// g2o's optimize(10): the Levenberg-Marquardt control flow (lambda, accept/reject, Raul's stop rule,
// g2o/core/optimization_algorithm_levenberg.cpp:59-165) is evaluated redundantly and identically by
// every thread from deterministic partial sums, so the host is never consulted between trials.
//
//   L1  point-major linearisation: error, Huber weight, Jacobians, Hll, bl, Hpl   (block_solver.hpp:502-560,
//       base_binary_edge.hpp:55-120, OptimizableTypes.cpp:136-155)
//   L2  camera-major accumulation of Hpp, bp
//   S0  per-landmark (Hll + lambda I)^-1 and Dinv*bl                               (block_solver.hpp:381-400)
//   S1  Schur complement, one CTA per free camera row, shared-memory accumulation  (:402-437)
//   C   dense reduced-camera solve: blocked Cholesky on one CTA, trailing updates on the FP64 tensor
//       pipe (DMMA m8n8k4)                                                          (linear_solver_eigen.h:89-121)
//   B1  landmark back-substitution, pose/point update into the trial buffers       (:459-483, se3quat.h:212-240)
//   B2  errors and robust chi2 at the trial estimate
#include "common.cuh"
#include "chol_device.cuh"
#include <chrono>
#include <cooperative_groups.h>
#include <math_constants.h>
#include <algorithm>
#include <cstdlib>
#include <vector>
#include <thread>
#include <chrono>

namespace cg = cooperative_groups;
using namespace dvm;

namespace {

inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield");
#endif
}

constexpr int kLbaThreads = kCholThreads;
constexpr int kLbaWarps = kLbaThreads / 32;
// Per-CTA partial sums exchanged through global memory between grid barriers.  Every WRITER has its own slot: the
// linearisation's chi2 (written right at the top of the next LM iteration, with no grid barrier after the previous
// trial's decision was read) must not share a slot with the trial chi2 that decision reads -- with a shared slot a CTA
// without landmarks (np < warps in the grid) overwrote it while slower warps were still summing it, and those warps
// then took a different accept/stop decision than the rest of the grid.
constexpr int kPartStride = 8;
constexpr int kPartLinChi = 0, kPartScale = 1, kPartMaxHll = 2, kPartMaxHpp = 3, kPartTrialChi = 4, kPartExcluded = 5;

struct LbaDev {
    int nc, nf, np, ne, dimP, dimPad, iterations; // dimPad = dimP rounded up to the Cholesky block size
    int iterations2;   // > 0: the welding BA's second pass (no robust kernel, level-0 edges only)
    int grid_chol;     // reduced system too large for one cluster's shared memory: factor it on the whole grid
    uint8_t* level;    // [ne] 1 = edge moved to level 1 before the second pass; null for one-pass calls
    const double* camK;   // [nc][4] fx, fy, cx, cy of every camera (e->pCamera = pKFi->mpCamera, O3/src/Optimizer.cc:1219)
    double delta, dsqr;
    double* camq[2]; double* camt[2];
    const int* cam_col; const int* free_cam;
    double* pts[2];
    const int* ecam; const int* ept; const float* eobs; const float* einfo;
    const int* pt_start; const int* pt_edges;
    const int* cam_start; const int* cam_edges;
    double* err; double* Hpl; double* Hll; double* bl; double* Dinv; double* db;
    double* Hpp; double* bp; double* Hs; double* bs; double* x; double* Linv;
    double* part;  // [gridDim * kPartStride], slots: see kPart*
    int* flags;    // [0] Cholesky ok
    const volatile int* abort_flag;
    float* out_camq; float* out_camt; float* out_pts; double* out_chi2; uint8_t* out_bad; double* out_stats;
    unsigned long long* prof; // [8] ns per phase (L1, L2, S0, S1, C, B1, B2, other), written by thread 0
    // edges grouped by point (as the reference creates them): the two CSR structures are built by the kernel itself
    int build;             // 1 = build pt_start / pt_edges / cam_start / cam_edges in the prologue
    int* pt_start_w; int* pt_edges_w; int* cam_start_w; int* cam_edges_w;
    int* cnt;              // [nf][warps of the grid] per-chunk camera counts, then their exclusive prefix
    int* cam_tot;          // [nf + 1]
    // observation pairs grouped by camera pair (built by the prologue): bin (c2, c1 <= c2) = c2 (c2 + 1) / 2 + c1 holds every
    // pair of edges (e1 on free camera c1, e2 on free camera c2) that see the same landmark; e1 == e2 in the diagonal bins
    // The lists are cut into work items of at most kPairChunk pairs: items[i] = { first, end, c1, c2 }; flags[2] = their number.
    int nbins;
    int* pair_start;       // [nbins + 1]
    int* pair_fill;        // [nbins]
    int4* pairs;           // { e1, e2, landmark, - }
    int4* items;
};
constexpr int kPairChunk = 64;
// Hpl block of an edge (6 x 3, camera rows x landmark columns): rows 0-2 and rows 3-5 as two 16-byte aligned groups of
// 9 (+1 padding) doubles, so that the Schur complement's half-warps read theirs with 128-bit loads
constexpr int kHplStride = 20;
__device__ __forceinline__ int hpl_idx(int a, int k) { return (a / 3) * 10 + (a % 3) * 3 + k; }

// ---- SE3 helpers (same formulas as the pose-only optimiser; g2o/types/se3quat.h) ----
struct Quat { double x, y, z, w; };
__device__ inline void quat_normalize(Quat& q)
{
    if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}
__device__ inline Quat quat_mul(const Quat& a, const Quat& b)
{
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
__device__ inline void quat_rotate(const Quat& q, const double v[3], double out[3])
{
    double uv[3] = { q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0] };
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
    out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
    out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}
__device__ inline void quat_to_matrix(const Quat& q, double R[9])
{
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ inline Quat quat_from_matrix(const double R[9])
{
    Quat q;
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[7] - R[5]) * t; q.y = (R[2] - R[6]) * t; q.z = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 4]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (R[k * 3 + j] - R[j * 3 + k]) * t;
        v[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        v[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}
// T' = exp(u) * T
__device__ inline void se3_update(const double u[6], const double* q_in, const double* t_in, double* q_out, double* t_out)
{
    const double w0 = u[0], w1 = u[1], w2 = u[2];
    const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    const double O[9] = { 0, -w2, w1, w2, 0, -w0, -w1, w0, 0 };
    double O2[9], R[9], V[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    if (theta < 0.00001) {
#pragma unroll
        for (int i = 0; i < 9; i++) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double sa = sin(theta) / theta, sb = (1 - cos(theta)) / (theta * theta);
        const double sc = (theta - sin(theta)) / pow(theta, 3.0);
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const double I = (i % 4 == 0) ? 1.0 : 0.0;
            R[i] = I + sa * O[i] + sb * O2[i];
            V[i] = I + sb * O[i] + sc * O2[i];
        }
    }
    Quat e = quat_from_matrix(R);
    quat_normalize(e);
    double et[3];
#pragma unroll
    for (int i = 0; i < 3; i++) et[i] = V[i * 3] * u[3] + V[i * 3 + 1] * u[4] + V[i * 3 + 2] * u[5];
    const Quat T = { q_in[0], q_in[1], q_in[2], q_in[3] };
    const double tt[3] = { t_in[0], t_in[1], t_in[2] };
    double rt[3];
    quat_rotate(e, tt, rt);
    Quat r = quat_mul(e, T);
    quat_normalize(r);
    q_out[0] = r.x; q_out[1] = r.y; q_out[2] = r.z; q_out[3] = r.w;
    t_out[0] = et[0] + rt[0]; t_out[1] = et[1] + rt[1]; t_out[2] = et[2] + rt[2];
}

__device__ inline double huber_rho0(double delta, double dsqr, double e) { return e <= dsqr ? e : 2 * sqrt(e) * delta - dsqr; }
__device__ inline double huber_rho1(double delta, double dsqr, double e) { return e <= dsqr ? 1.0 : delta / sqrt(e); }

// camera-frame point and reprojection error of edge e at estimate buffer `cur`
__device__ inline void edge_residual(const LbaDev& P, int cur, int e, double xc[3], double r[2], Quat* qout)
{
    const int c = P.ecam[e], l = P.ept[e];
    const double* q = P.camq[cur] + 4 * c;
    const double* t = P.camt[cur] + 3 * c;
    const Quat Q = { q[0], q[1], q[2], q[3] };
    const double X[3] = { P.pts[cur][3 * l], P.pts[cur][3 * l + 1], P.pts[cur][3 * l + 2] };
    quat_rotate(Q, X, xc);
    xc[0] += t[0]; xc[1] += t[1]; xc[2] += t[2];
    const double* kc = P.camK + 4 * c;
    r[0] = (double)P.eobs[2 * e] - (kc[0] * xc[0] / xc[2] + kc[2]);
    r[1] = (double)P.eobs[2 * e + 1] - (kc[1] * xc[1] / xc[2] + kc[3]);
    if (qout) *qout = Q;
}

// deterministic sum of NV doubles per thread over the CTA; result in out[0..NV) for all threads
template <int NV>
__device__ inline void cta_sum(double (&v)[NV], double* warp_buf, double* out)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) warp_buf[wid * NV + i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < kLbaWarps; w++) s += warp_buf[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// The same for up to 32 values per thread with a TRANSPOSED butterfly: 31 shuffles per warp instead of 5 per value (the
// shuffle unit is shared by the SM's 16 warps: 27 values the plain way took 15.8 us per call, measured).  v is clobbered.
template <int NV>
__device__ inline void cta_sum_wide(double (&v)[NV], double* warp_buf, double* out)
{
    static_assert(NV <= 32, "one value per lane");
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double t[32];
#pragma unroll
    for (int i = 0; i < 32; i++) t[i] = i < NV ? v[i] : 0.0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; i++) {
            const double send = upper ? t[i] : t[i + off];
            const double keep = upper ? t[i + off] : t[i];
            t[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    warp_buf[wid * 32 + lane] = t[0];   // lane l: this warp's total of value l
    __syncthreads();
    if (threadIdx.x < NV) {
        double s4[4] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
        for (int w = 0; w < kLbaWarps; w++) s4[w & 3] += warp_buf[w * 32 + threadIdx.x];
        out[threadIdx.x] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    }
    __syncthreads();
}

__device__ inline double warp_sum(double s)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}
__device__ inline double warp_max(double s)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
    return s;
}

// ---- dense SPD solve Hs x = bs on ONE 8-CTA CLUSTER ------------------------------------------------
// Blocked right-looking Cholesky (lower triangle, kNB = 32) of the n x n reduced camera matrix A
// (n a multiple of 32: the caller pads with an identity block), with the right-hand side carried as an
// extra matrix row (row n): after the factorisation that row holds y = L^-1 b, so only the backward
// substitution L^T x = y remains.  Per block column:
//   (a) the 32x32 diagonal block arrives factored and inverted (cta_factor_invert_32) from CTA 0, which did it one block
//       column AHEAD, beside the previous trailing update;
//   (b) the panel rows are split over the 8 CTAs:  L21 = A21 * L11^-T  as a small GEMM with the inverse;
//   (c) after a cluster barrier CTA 0 updates, factors and inverts the next diagonal block while the other seven stage the
//       whole panel in shared memory and share the 8x8 trailing tiles  A22 -= L21 L21^T  on the FP64 tensor pipe
//       (DMMA m8n8k4).
// A has n + 8 rows of leading dimension n.  Linv_g [n/32][32*32] receives the inverted diagonal blocks.
constexpr int kClusterCtas = 8;
constexpr int kLbaClusterFree = 100;   // free keyframes whose reduced system (600 unknowns) fits one SM's shared-memory panel
constexpr int kLbaMaxFree = 2000;


__device__ bool cluster_cholesky_solve(int n, double* __restrict__ A, const double* __restrict__ b, double* __restrict__ x,
                                       double* __restrict__ Linv_g, double* smem, int* s_flag, unsigned long long* prof)
{
    unsigned long long tp = 0;
    auto tk = [&](int slot) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (slot >= 0) prof[slot] += t - tp;
            tp = t;
        }
    };
    tk(-1);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double* panel = smem;                                 // [(n + 8)][kPanelLd]
    double* Ld = panel + (size_t)(n + 8) * kPanelLd;      // [kNB][kDiagLd]  factored diagonal block
    double* Li = Ld + kNB * kDiagLd;                      // [kNB][kDiagLd]  its inverse
    double* stage = Li + kNB * kDiagLd;                   // [64][kPanelLd]  this CTA's panel rows before the TRSM
    const int M = n + 8;                                  // rows including the right-hand-side row group
    if (tid == 0) *s_flag = 1;
    // right-hand side as row n, zero rows n+1 .. n+7
    if (rank == 0)
        for (int i = tid; i < 8 * n; i += kLbaThreads) A[(size_t)n * n + i] = (i < n) ? b[i] : 0.0;
    cluster.sync();
    // block 0 has no predecessor to hide behind
    if (rank == 0) {
        const bool bad = cta_factor_invert_32(A, n, Ld, Li, true, Linv_g, stage);
        if (bad && tid == 0) *s_flag = 0;
    }
    tk(8);
    cluster.sync();
    for (int k0 = 0; k0 < n; k0 += kNB) {
        const int r0 = k0 + kNB;
        // ---- (a) the inverted diagonal block: CTA 0 still holds it from its factorisation, the others fetch it ----
        if (rank != 0) {
            const double* Lg = Linv_g + (size_t)(k0 / kNB) * kNB * kNB;
            for (int t = tid; t < kNB * kNB; t += kLbaThreads) Li[(t >> 5) * kDiagLd + (t & 31)] = Lg[t];
        }
        __syncthreads();
        tk(9);
        // ---- (b) panel rows of this CTA: L21 = A21 * L11^-T ----
        const int mrows = M - r0;
        const int mcols = n - r0;
        // CTA 0 takes the 32 panel rows of the next diagonal block -- they stay in its shared memory (Pn) -- and fetches that
        // block now, under the latency of the panel loads; the other seven CTAs share the rows below
        int i0, i1;
        if (mcols > 0) {
            const int chunk = ((mrows - kNB + kClusterCtas - 2) / (kClusterCtas - 1) + 7) & ~7;   // whole 8-row tiles
            i0 = rank == 0 ? r0 : r0 + kNB + (rank - 1) * chunk;
            i1 = rank == 0 ? r0 + kNB : min(i0 + chunk, M);
        } else {   // last block column: only the right-hand-side rows are left
            const int chunk = 8;
            i0 = r0 + rank * chunk;
            i1 = min(i0 + chunk, M);
        }
        double* Pn = panel;                       // [kNB][kDiagLd]  panel rows r0 .. r0 + 31
        double* Dn = panel + kNB * kDiagLd;       // [kNB][kDiagLd]  the next diagonal block
        const bool own_next = rank == 0 && mcols > 0;
        // (in the layout of the tensor-core tile warp `wid` computes below: 10 lower 8 x 8 tiles of the 32 x 32 block)
        double dnext[2] = { 0.0, 0.0 };
        const int dti = wid < 1 ? 0 : wid < 3 ? 1 : wid < 6 ? 2 : 3, dtj = wid - dti * (dti + 1) / 2;
        const int drow = dti * 8 + (lane >> 2), dcol = dtj * 8 + (lane & 3) * 2;
        if (rank == 0 && mcols > 0 && wid < 10) {
            if (dcol <= drow) dnext[0] = A[(size_t)(r0 + drow) * n + r0 + dcol];
            if (dcol + 1 <= drow) dnext[1] = A[(size_t)(r0 + drow) * n + r0 + dcol + 1];
        }
        for (int base = i0; base < i1; base += 64) {
            const int cnt = min(64, i1 - base);
            for (int t = tid; t < cnt * kNB; t += kLbaThreads) stage[(t >> 5) * kPanelLd + (t & 31)] = A[(size_t)(base + (t >> 5)) * n + k0 + (t & 31)];
            __syncthreads();
            tk(16);
            panel_times_inverse_t(stage, cnt, Li, [&](int r, int c, double v0, double v1) {
                double* dst = A + (size_t)(base + r) * n + k0 + c;
                dst[0] = v0; dst[1] = v1;
                if (own_next) { Pn[(base + r - r0) * kDiagLd + c] = v0; Pn[(base + r - r0) * kDiagLd + c + 1] = v1; }
            });
            tk(17);
            __syncthreads();
        }
        tk(10);
        cluster.sync();
        tk(11);
        // ---- (c) trailing update, with the NEXT diagonal block taken out of it: CTA 0 brings that block up to date from the
        // first 32 panel rows, factors and inverts it while the other seven CTAs update the rest of the trailing matrix, so the
        // 32 dependent pivot steps of a block (the longest chain of the solve) run beside the tensor-pipe work instead of
        // before it (look-ahead of one block column) ----
        if (mcols > 0 && rank == 0) {
            if (wid < 10) {   // D = A(next diagonal block) - Pn Pn^T, lower tiles, on the tensor pipe
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int kk = 0; kk < kNB; kk += 4)
                    dmma_m8n8k4(c0, c1, Pn[drow * kDiagLd + kk + (lane & 3)], Pn[(dtj * 8 + (lane >> 2)) * kDiagLd + kk + (lane & 3)]);
                Dn[drow * kDiagLd + dcol] = dcol <= drow ? dnext[0] - c0 : 0.0;
                Dn[drow * kDiagLd + dcol + 1] = dcol + 1 <= drow ? dnext[1] - c1 : 0.0;
            } else {          // the six tiles above the diagonal
                const int u = wid - 10;
                const int uti = u < 3 ? 0 : u < 5 ? 1 : 2, utj = u < 3 ? u + 1 : u < 5 ? u - 1 : 3;
                Dn[(uti * 8 + (lane >> 2)) * kDiagLd + utj * 8 + (lane & 3) * 2] = 0.0;
                Dn[(uti * 8 + (lane >> 2)) * kDiagLd + utj * 8 + (lane & 3) * 2 + 1] = 0.0;
            }
            __syncthreads();
            tk(12);
            const bool bad = cta_factor_invert_32(Dn, kDiagLd, Ld, Li, true, Linv_g + (size_t)(r0 / kNB) * kNB * kNB, stage);
            if (bad && tid == 0) *s_flag = 0;
            tk(13);
        } else if (mcols > 0) {
            for (int t = tid; t < mrows * kNB; t += kLbaThreads)
                panel[(size_t)(t >> 5) * kPanelLd + (t & 31)] = A[(size_t)(r0 + (t >> 5)) * n + k0 + (t & 31)];
            __syncthreads();
            const int mtc = mcols / 8;
            const int tri = mtc * (mtc + 1) / 2;
            const int ntiles = tri + mtc; // + the right-hand-side row group against every column tile
            // tiles 0 .. 9 (ti < 4) are the next diagonal block: CTA 0's
            for (int t = 10 + (rank - 1) * kLbaWarps + wid; t < ntiles; t += (kClusterCtas - 1) * kLbaWarps) {
                int ti, tj;
                if (t < tri) {
                    ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
                    while (ti * (ti + 1) / 2 > t) ti--;
                    while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
                    tj = t - ti * (ti + 1) / 2;
                } else { ti = mtc; tj = t - tri; }
                const int ar = ti * 8 + (lane >> 2), bc = tj * 8 + (lane >> 2);
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int kk = 0; kk < kNB; kk += 4) {
                    const double a = panel[(size_t)ar * kPanelLd + kk + (lane & 3)];
                    const double bb = panel[(size_t)bc * kPanelLd + kk + (lane & 3)];
                    dmma_m8n8k4(c0, c1, a, bb);
                }
                const int cc = tj * 8 + (lane & 3) * 2;
                double* dst = A + (size_t)(r0 + ar) * n + r0 + cc;
                if (cc <= ar) dst[0] -= c0;
                if (cc + 1 <= ar) dst[1] -= c1;
            }
        }
        cluster.sync();
        tk(14);
    }
    const bool ok = *s_flag != 0;
    // ---- backward substitution L^T x = y on CTA 0 (y = row n of A) ----
    // Per block column (last to first): the inverted diagonal block is fetched into shared memory while the 16 warps sum
    // L(below, block)^T x(below) -- the loads of a warp are issued four rows at a time, a multiply-add waiting for its
    // operand would otherwise hold back the next load (in-order issue: one L2 round trip per row).
    if (rank == 0) {
        double* y = panel;          // [n]
        double* part = panel + n;   // [16][kNB] partial sums
        double* rhs = part + 16 * kNB;
        double* LgS = rhs + kNB;    // [kNB][kNB + 1]   (runs on into the dead diagonal-block buffers when n = 32)
        for (int i = tid; i < n; i += kLbaThreads) y[i] = A[(size_t)n * n + i];
        __syncthreads();
        for (int k0 = n - kNB; k0 >= 0; k0 -= kNB) {
            {
                const double* Lg = Linv_g + (size_t)(k0 / kNB) * kNB * kNB;
                const double g0 = Lg[tid], g1 = Lg[tid + kLbaThreads];
                const int c = tid & 31, g = tid >> 5;
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                int i = k0 + kNB + g;
                for (; i + 3 * kLbaWarps < n; i += 4 * kLbaWarps) {
                    const double a0 = A[(size_t)i * n + k0 + c], a1 = A[(size_t)(i + kLbaWarps) * n + k0 + c];
                    const double a2 = A[(size_t)(i + 2 * kLbaWarps) * n + k0 + c], a3 = A[(size_t)(i + 3 * kLbaWarps) * n + k0 + c];
                    acc0 += a0 * y[i]; acc1 += a1 * y[i + kLbaWarps]; acc2 += a2 * y[i + 2 * kLbaWarps]; acc3 += a3 * y[i + 3 * kLbaWarps];
                }
                {
                    const bool h0 = i < n, h1 = i + kLbaWarps < n, h2 = i + 2 * kLbaWarps < n;
                    const double a0 = h0 ? A[(size_t)i * n + k0 + c] : 0.0, a1 = h1 ? A[(size_t)(i + kLbaWarps) * n + k0 + c] : 0.0;
                    const double a2 = h2 ? A[(size_t)(i + 2 * kLbaWarps) * n + k0 + c] : 0.0;
                    if (h0) acc0 += a0 * y[i];
                    if (h1) acc1 += a1 * y[i + kLbaWarps];
                    if (h2) acc2 += a2 * y[i + 2 * kLbaWarps];
                }
                part[g * kNB + c] = (acc0 + acc1) + (acc2 + acc3);
                LgS[(tid >> 5) * (kNB + 1) + (tid & 31)] = g0;
                LgS[((tid + kLbaThreads) >> 5) * (kNB + 1) + (tid & 31)] = g1;
            }
            __syncthreads();
            if (tid < kNB) {
                double s4[4] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
                for (int g = 0; g < kLbaWarps; g++) s4[g & 3] += part[g * kNB + tid];
                rhs[tid] = y[k0 + tid] - ((s4[0] + s4[1]) + (s4[2] + s4[3]));
            }
            __syncthreads();
            if (tid < kNB) { // x_k = L_kk^-T rhs
                double s4[4] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
                for (int c = 0; c < kNB; c++)
                    if (c >= tid) s4[c & 3] += LgS[c * (kNB + 1) + tid] * rhs[c];
                y[k0 + tid] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            }
            __syncthreads();
        }
        for (int i = tid; i < n; i += kLbaThreads) x[i] = ok ? y[i] : 0.0;
        __syncthreads();
    }
    tk(15);
    return ok;
}

__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kLbaThreads, 1) lba_kernel(LbaDev P)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double smem[];
    __shared__ double warp_buf[kLbaWarps * 32];
    __shared__ double red[28];
    __shared__ int s_flag;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x;
    const int gwarp = blockIdx.x * kLbaWarps + wid, nwarps = G * kLbaWarps;
    const int gtid = blockIdx.x * kLbaThreads + tid, nthreads = G * kLbaThreads;

    unsigned long long t_prev = 0;
    auto tick = [&](int slot) { // phase timer (thread 0 of CTA 0 only; the grid.sync before it makes it meaningful)
        if (gtid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (slot >= 0) P.prof[slot] += t - t_prev;
            t_prev = t;
        }
    };
    tick(-1);
    int cur = 0; // index of the accepted estimate buffers
    double lambda = -1, ni = 2;
    int nBad = 0, done = 0, trials = 0;
    double first_chi = 0, last_chi = 0;
    bool stop = false;

    // pbStopFlag lives in mapped host memory: thread 0 samples it and the whole grid adopts that value,
    // so that every thread leaves the loops at the same point
    auto agree_abort = [&]() -> bool {
        if (!P.abort_flag) return false;
        if (gtid == 0) P.flags[1] = (*P.abort_flag) ? 1 : 0;
        grid.sync();
        const bool a = P.flags[1] != 0;
        grid.sync();
        return a;
    };

    // ---- structure (the analogue of BlockSolver::buildStructure, block_solver.hpp:143-295) on the device, for edges
    // that come grouped by point: rows by point are runs of the edge array; rows by free camera are a STABLE counting
    // sort (ascending edge index inside a row, so that every sum over a row has a fixed order): every warp owns a
    // contiguous chunk of edges, counts its edges per camera, a per-camera scan over the chunks gives each (camera,
    // chunk) its base, and the warps then place their edges in order.
    if (P.build) {
        const int C = (P.ne + nwarps - 1) / nwarps;
        for (size_t i = gtid; i < (size_t)P.nf * nwarps; i += nthreads) P.cnt[i] = 0;
        for (int e = gtid; e < P.ne; e += nthreads) {
            const int l = P.ept[e], prev = e > 0 ? P.ept[e - 1] : -1;
            for (int k = prev + 1; k <= l; k++) P.pt_start_w[k] = e;
            if (e == P.ne - 1) for (int k = l + 1; k <= P.np; k++) P.pt_start_w[k] = P.ne;
            P.pt_edges_w[e] = e;
        }
        grid.sync();
        {
            const int e0 = gwarp * C, e1 = min(e0 + C, P.ne);
            for (int eb = e0; eb < e1; eb += 32) {
                const int e = eb + lane;
                const int cf = e < e1 ? P.cam_col[P.ecam[e]] : -1;
                const unsigned peers = __match_any_sync(0xffffffffu, cf);
                if (cf >= 0 && lane == __ffs(peers) - 1) atomicAdd(&P.cnt[(size_t)cf * nwarps + gwarp], __popc(peers));
            }
        }
        grid.sync();
        for (int cf = gwarp; cf < P.nf; cf += nwarps) { // exclusive prefix over the chunks, one warp per camera
            int* row = P.cnt + (size_t)cf * nwarps;
            int run = 0;
            for (int b0 = 0; b0 < nwarps; b0 += 32) {
                const int v = b0 + lane < nwarps ? row[b0 + lane] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (b0 + lane < nwarps) row[b0 + lane] = run + incl - v;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) P.cam_tot[cf] = run;
        }
        grid.sync();
        if (blockIdx.x == 0 && wid == 0) { // row starts: exclusive prefix of the totals (nf <= 2000)
            int run = 0;
            for (int b0 = 0; b0 < P.nf; b0 += 32) {
                const int v = b0 + lane < P.nf ? P.cam_tot[b0 + lane] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (b0 + lane < P.nf) P.cam_start_w[b0 + lane] = run + incl - v;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) P.cam_start_w[P.nf] = run;
        }
        grid.sync();
        {
            const int e0 = gwarp * C, e1 = min(e0 + C, P.ne);
            for (int eb = e0; eb < e1; eb += 32) {
                const int e = eb + lane;
                const int cf = e < e1 ? P.cam_col[P.ecam[e]] : -1;
                const unsigned peers = __match_any_sync(0xffffffffu, cf);
                const int leader = __ffs(peers) - 1;
                int base = 0;
                if (cf >= 0 && lane == leader) base = atomicAdd(&P.cnt[(size_t)cf * nwarps + gwarp], __popc(peers)); // running base of this chunk
                base = __shfl_sync(0xffffffffu, base, leader);
                if (cf >= 0) P.cam_edges_w[P.cam_start_w[cf] + base + __popc(peers & ((1u << lane) - 1u))] = e;
            }
        }
        grid.sync();
    }

    // ---- the Schur complement's work lists: every pair of observations of one landmark from free cameras, grouped by
    // camera pair (counting sort: count, scan on CTA 0, place).  With them S1 below sums each 6 x 6 block of the reduced
    // system in registers and writes it once, instead of one FP64 atomic per product term (12 M per trial at C4 size).
    {
        __shared__ int s_wsum[kLbaWarps], s_wsum_i[kLbaWarps];
        const int nb = P.nbins;
        for (int i = gtid; i <= nb; i += nthreads) P.pair_start[i] = 0;
        grid.sync();
        auto for_each_pair = [&](auto&& f) {
            for (int l = gwarp; l < P.np; l += nwarps) {
                const int ps = P.pt_start[l], k = P.pt_start[l + 1] - ps;
                const int npairs = k * (k + 1) / 2;
                for (int p = lane; p < npairs; p += 32) {
                    int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
                    while (i * (i + 1) / 2 > p) i--;
                    while ((i + 1) * (i + 2) / 2 <= p) i++;
                    const int j = p - i * (i + 1) / 2;
                    int e1 = P.pt_edges[ps + i], e2 = P.pt_edges[ps + j];
                    int c1 = P.cam_col[P.ecam[e1]], c2 = P.cam_col[P.ecam[e2]];
                    if (c1 < 0 || c2 < 0) continue;
                    if (c1 > c2) { int t = c1; c1 = c2; c2 = t; t = e1; e1 = e2; e2 = t; }
                    f(c2 * (c2 + 1) / 2 + c1, e1, e2);
                }
            }
        };
        for_each_pair([&](int bin, int, int) { atomicAdd(&P.pair_start[bin + 1], 1); });
        grid.sync();
        if (blockIdx.x == 0) {   // scans of the counts (pair_start[b + 1] = end of bin b) and of the bins' item counts
            int run = 0, run_items = 0;
            for (int b0 = 0; b0 < nb; b0 += kLbaThreads) {
                const int b = b0 + tid;
                const int v = b < nb ? P.pair_start[b + 1] : 0;
                const int vi = (v + kPairChunk - 1) / kPairChunk;
                int incl = v, incl_i = vi;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    const int ti = __shfl_up_sync(0xffffffffu, incl_i, o);
                    if (lane >= o) { incl += t; incl_i += ti; }
                }
                if (lane == 31) { s_wsum[wid] = incl; s_wsum_i[wid] = incl_i; }
                __syncthreads();
                int before = 0, total = 0, before_i = 0, total_i = 0;
#pragma unroll
                for (int w2 = 0; w2 < kLbaWarps; w2++) {
                    const int t = s_wsum[w2], ti = s_wsum_i[w2];
                    if (w2 < wid) { before += t; before_i += ti; }
                    total += t; total_i += ti;
                }
                if (b < nb) {
                    const int first = run + before + incl - v;
                    P.pair_start[b + 1] = first + v;
                    P.pair_fill[b] = first;
                    int c2 = (int)((sqrt(8.0 * b + 1.0) - 1.0) * 0.5);
                    while (c2 * (c2 + 1) / 2 > b) c2--;
                    while ((c2 + 1) * (c2 + 2) / 2 <= b) c2++;
                    const int c1 = b - c2 * (c2 + 1) / 2;
                    int4* it = P.items + run_items + before_i + incl_i - vi;
                    for (int ch = 0; ch < vi; ch++) it[ch] = make_int4(first + ch * kPairChunk, min(first + (ch + 1) * kPairChunk, first + v), c1, c2);
                }
                run += total;
                run_items += total_i;
                __syncthreads();
            }
            if (tid == 0) P.flags[2] = run_items;
        }
        grid.sync();
        for_each_pair([&](int bin, int e1, int e2) { P.pairs[atomicAdd(&P.pair_fill[bin], 1)] = make_int4(e1, e2, P.ept[e1], 0); });
        grid.sync();
    }

    // One pass is the reference's optimizer.optimize(iterations).  The welding BA of a map merge (O3/src/Optimizer.cc:
    // 3257-3675) runs two: optimize(5) with the Huber kernel, then -- unless the stop flag is up -- edges with
    // chi2 > 5.991 or non-positive depth go to level 1, every edge loses its robust kernel and optimize(10) runs on the
    // level-0 edges (:3476-3521).  A level-1 edge contributes nothing and keeps the error of the first pass.
    const int npasses = P.iterations2 > 0 ? 2 : 1;
    double delta = P.delta, dsqr = P.dsqr;
    int done_first = 0, excluded = 0;
    for (int pass = 0; pass < npasses; pass++) {
        if (pass == 1) {
            if (agree_abort()) break;
            done_first = done;
            int nex = 0;
            for (int e = gtid; e < P.ne; e += nthreads) {
                const double om = (double)P.einfo[e];
                const double r0 = P.err[2 * e], r1 = P.err[2 * e + 1];
                double xc[3], r[2];
                edge_residual(P, cur, e, xc, r, nullptr);
                const bool out = r0 * (om * r0) + r1 * (om * r1) > 5.991 || !(xc[2] > 0.0);
                P.level[e] = out ? 1 : 0;
                nex += out ? 1 : 0;
            }
            {
                double v[1] = { (double)nex };
                cta_sum<1>(v, warp_buf, red);
                if (tid == 0) P.part[blockIdx.x * kPartStride + kPartExcluded] = red[0];
            }
            grid.sync();
            for (int b2 = 0; b2 < G; b2++) excluded += (int)P.part[b2 * kPartStride + kPartExcluded];
            delta = CUDART_INF; dsqr = CUDART_INF; // (a local copy: writing the parameter struct would move it to the stack)
            stop = false;
        }
        const int iters = pass == 0 ? P.iterations : P.iterations2;
        for (int it = 0; it < iters && !stop; it++) {
            if (agree_abort()) break;
            // ---------------- L1: point-major linearisation ----------------
            double chi_part = 0, maxd = 0;
            for (int l = gwarp; l < P.np; l += nwarps) {
                double h[6] = { 0, 0, 0, 0, 0, 0 }, g[3] = { 0, 0, 0 };
                const int s = P.pt_start[l], e_end = P.pt_start[l + 1];
                for (int k = s + lane; k < e_end; k += 32) {
                    const int e = P.pt_edges[k];
                    if (P.level && P.level[e]) { // level-1 edge: no contribution to H, b or chi2; its error stays
                        if (P.cam_col[P.ecam[e]] >= 0) {
                            double* hp = P.Hpl + (size_t)e * kHplStride;
#pragma unroll
                            for (int i = 0; i < kHplStride; i++) hp[i] = 0.0;
                        }
                        continue;
                    }
                    double xc[3], r[2];
                    Quat Q;
                    edge_residual(P, cur, e, xc, r, &Q);
                    P.err[2 * e] = r[0]; P.err[2 * e + 1] = r[1];
                    const double om = (double)P.einfo[e];
                    const double chi = r[0] * (om * r[0]) + r[1] * (om * r[1]);
                    chi_part += huber_rho0(delta, dsqr, chi);
                    const double w = huber_rho1(delta, dsqr, chi);
                    const double X = xc[0], Y = xc[1], Z = xc[2];
                    const double fx = P.camK[4 * P.ecam[e]], fy = P.camK[4 * P.ecam[e] + 1];
                    const double pj[6] = { -(fx / Z), 0.0, fx * X / (Z * Z), 0.0, -(fy / Z), fy * Y / (Z * Z) };
                    double R[9];
                    quat_to_matrix(Q, R);
                    double A[6];
    #pragma unroll
                    for (int rr = 0; rr < 2; rr++)
    #pragma unroll
                        for (int kk = 0; kk < 3; kk++)
                            A[rr * 3 + kk] = pj[rr * 3] * R[kk] + pj[rr * 3 + 1] * R[3 + kk] + pj[rr * 3 + 2] * R[6 + kk];
                    const double wo = w * om;
                    const double r0 = -om * r[0] * w, r1 = -om * r[1] * w;
                    g[0] += A[0] * r0 + A[3] * r1; g[1] += A[1] * r0 + A[4] * r1; g[2] += A[2] * r0 + A[5] * r1;
                    h[0] += A[0] * wo * A[0] + A[3] * wo * A[3];
                    h[1] += A[0] * wo * A[1] + A[3] * wo * A[4];
                    h[2] += A[0] * wo * A[2] + A[3] * wo * A[5];
                    h[3] += A[1] * wo * A[1] + A[4] * wo * A[4];
                    h[4] += A[1] * wo * A[2] + A[4] * wo * A[5];
                    h[5] += A[2] * wo * A[2] + A[5] * wo * A[5];
                    if (P.cam_col[P.ecam[e]] >= 0) {
                        const double Dv[18] = { 0, Z, -Y, 1, 0, 0, -Z, 0, X, 0, 1, 0, Y, -X, 0, 0, 0, 1 };
                        double* hp = P.Hpl + (size_t)e * kHplStride;
    #pragma unroll
                        for (int a = 0; a < 6; a++) {
                            const double B0 = pj[0] * Dv[a] + pj[1] * Dv[6 + a] + pj[2] * Dv[12 + a];
                            const double B1 = pj[3] * Dv[a] + pj[4] * Dv[6 + a] + pj[5] * Dv[12 + a];
    #pragma unroll
                            for (int b2 = 0; b2 < 3; b2++) hp[hpl_idx(a, b2)] = B0 * wo * A[b2] + B1 * wo * A[3 + b2];
                        }
                        hp[9] = 0.0; hp[19] = 0.0;   // the padding is loaded (never used) by the Schur phase's 128-bit reads: keep it defined
                    }
                }
    #pragma unroll
                for (int i = 0; i < 6; i++) h[i] = warp_sum(h[i]);
    #pragma unroll
                for (int i = 0; i < 3; i++) g[i] = warp_sum(g[i]);
                if (lane == 0) {
    #pragma unroll
                    for (int i = 0; i < 6; i++) P.Hll[(size_t)l * 6 + i] = h[i];
    #pragma unroll
                    for (int i = 0; i < 3; i++) P.bl[(size_t)l * 3 + i] = g[i];
                    maxd = fmax(maxd, fmax(fabs(h[0]), fmax(fabs(h[3]), fabs(h[5]))));
                }
            }
            {
                double v[1] = { chi_part };
                cta_sum<1>(v, warp_buf, red);
                maxd = warp_max(maxd);
                if (lane == 0) warp_buf[wid] = maxd;
                __syncthreads();
                if (tid == 0) {
                    double m = 0;
                    for (int w2 = 0; w2 < kLbaWarps; w2++) m = fmax(m, warp_buf[w2]);
                    P.part[blockIdx.x * kPartStride + kPartLinChi] = red[0];
                    P.part[blockIdx.x * kPartStride + kPartMaxHll] = m;
                }
                __syncthreads();
            }
            grid.sync();
            tick(0);
            // ---------------- L2: camera-major Hpp, bp ----------------
            double maxdp = 0;
            for (int cf = blockIdx.x; cf < P.nf; cf += G) {
                double acc[27];
    #pragma unroll
                for (int i = 0; i < 27; i++) acc[i] = 0;
                const int s = P.cam_start[cf], e_end = P.cam_start[cf + 1];
                for (int k = s + tid; k < e_end; k += kLbaThreads) {
                    const int e = P.cam_edges[k];
                    if (P.level && P.level[e]) continue;
                    double xc[3], r[2];
                    edge_residual(P, cur, e, xc, r, nullptr);
                    const double om = (double)P.einfo[e];
                    const double chi = r[0] * (om * r[0]) + r[1] * (om * r[1]);
                    const double w = huber_rho1(delta, dsqr, chi);
                    const double X = xc[0], Y = xc[1], Z = xc[2];
                    const double fx = P.camK[4 * P.ecam[e]], fy = P.camK[4 * P.ecam[e] + 1];
                    const double pj[6] = { -(fx / Z), 0.0, fx * X / (Z * Z), 0.0, -(fy / Z), fy * Y / (Z * Z) };
                    const double Dv[18] = { 0, Z, -Y, 1, 0, 0, -Z, 0, X, 0, 1, 0, Y, -X, 0, 0, 0, 1 };
                    double B[12];
    #pragma unroll
                    for (int a = 0; a < 6; a++) {
                        B[a] = pj[0] * Dv[a] + pj[1] * Dv[6 + a] + pj[2] * Dv[12 + a];
                        B[6 + a] = pj[3] * Dv[a] + pj[4] * Dv[6 + a] + pj[5] * Dv[12 + a];
                    }
                    const double wo = w * om;
                    const double r0 = -om * r[0] * w, r1 = -om * r[1] * w;
                    int idx = 0;
    #pragma unroll
                    for (int a = 0; a < 6; a++) {
                        acc[21 + a] += B[a] * r0 + B[6 + a] * r1;
    #pragma unroll
                        for (int b2 = a; b2 < 6; b2++) acc[idx++] += B[a] * wo * B[b2] + B[6 + a] * wo * B[6 + b2];
                    }
                }
                cta_sum_wide<27>(acc, warp_buf, red);
                if (tid < 36) {
                    const int a = tid / 6, b2 = tid % 6;
                    const int lo = min(a, b2), hi = max(a, b2);
                    const int idx = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);
                    P.Hpp[(size_t)cf * 36 + tid] = red[idx];
                }
                if (tid < 6) P.bp[cf * 6 + tid] = red[21 + tid];
                if (tid == 0)
                    for (int a = 0; a < 6; a++) maxdp = fmax(maxdp, fabs(red[a * 6 - a * (a - 1) / 2]));
                __syncthreads();
            }
            if (tid == 0) P.part[blockIdx.x * kPartStride + kPartMaxHpp] = maxdp;
            grid.sync();
            tick(1);
            double currentChi = 0;
            {
                double m = 0;
                for (int b2 = 0; b2 < G; b2++) {
                    currentChi += P.part[b2 * kPartStride + kPartLinChi];
                    m = fmax(m, fmax(P.part[b2 * kPartStride + kPartMaxHll], P.part[b2 * kPartStride + kPartMaxHpp]));
                }
                if (it == 0) { lambda = 1e-5 * m; ni = 2; nBad = 0; if (pass == 0) first_chi = currentChi; }
            }
            const double iniChi = currentChi;
            double rho = 0;
            int qmax = 0;
            bool aborted = false;
            do {
                const int trial = cur ^ 1;
                // ---------------- S0: landmark inverses, clear Hs ----------------
                for (int l = gtid; l < P.np; l += nthreads) {
                    const double* h = P.Hll + (size_t)l * 6;
                    const double d00 = h[0] + lambda, d01 = h[1], d02 = h[2], d11 = h[3] + lambda, d12 = h[4], d22 = h[5] + lambda;
                    const double c00 = d11 * d22 - d12 * d12, c01 = d12 * d02 - d01 * d22, c02 = d01 * d12 - d11 * d02;
                    const double det = d00 * c00 + d01 * c01 + d02 * c02;
                    const double id = 1.0 / det;
                    double* Di = P.Dinv + (size_t)l * 6;
                    const double i00 = c00 * id, i01 = c01 * id, i02 = c02 * id;
                    const double i11 = (d00 * d22 - d02 * d02) * id, i12 = (d02 * d01 - d00 * d12) * id, i22 = (d00 * d11 - d01 * d01) * id;
                    Di[0] = i00; Di[1] = i01; Di[2] = i02; Di[3] = i11; Di[4] = i12; Di[5] = i22;
                    const double* g = P.bl + (size_t)l * 3;
                    double* d = P.db + (size_t)l * 3;
                    d[0] = i00 * g[0] + i01 * g[1] + i02 * g[2];
                    d[1] = i01 * g[0] + i11 * g[1] + i12 * g[2];
                    d[2] = i02 * g[0] + i12 * g[1] + i22 * g[2];
                }
                // Hs (lower triangle, padded to dimPad) starts as Hpp + lambda I, identity on the padding
                for (size_t i = gtid; i < (size_t)P.dimPad * P.dimPad; i += nthreads) {
                    const int r = (int)(i / P.dimPad), c = (int)(i - (size_t)r * P.dimPad);
                    double v = 0.0;
                    if (r >= P.dimP) v = (r == c) ? 1.0 : 0.0;
                    else if (c < P.dimP && r / 6 == c / 6) v = P.Hpp[(size_t)(r / 6) * 36 + (r % 6) * 6 + (c % 6)] + (r == c ? lambda : 0.0);
                    P.Hs[i] = v;
                }
                for (int i = gtid; i < P.dimPad; i += nthreads) P.bs[i] = i < P.dimP ? P.bp[i] : 0.0;
                grid.sync();
                tick(2);
                // ---------------- S1: Schur complement, camera-pair-major ----------------
                // Hs(c2, c1) -= sum over the landmarks seen by both of Hpl_2 Dinv Hpl_1^T, bs(c) -= sum Hpl Dinv bl.  One warp
                // per work item (<= 64 observation pairs of one camera pair); its two half-warps take the left and the right
                // three columns of the 6 x 6 block, 16 pairs per step, sums in registers, one butterfly, and 36 additions
                // at L2 per item instead of 36 per pair.
                {
                    const int half = lane >> 4, sub = lane & 15;
                    const int nitems = P.flags[2];
                    for (int item = gwarp; item < nitems; item += nwarps) {
                        const int4 it = P.items[item];
                        const int c1 = it.z, c2 = it.w;
                        const bool diag = c1 == c2;
                        // the item's pairs first: their loads are in flight together
                        int4 pe[kPairChunk / 16];
#pragma unroll
                        for (int u = 0; u < kPairChunk / 16; u++) {
                            const int t = it.x + sub + 16 * u;
                            pe[u] = t < it.y ? P.pairs[t] : make_int4(-1, -1, -1, 0);
                        }
                        double acc[18], g[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
                        for (int v = 0; v < 18; v++) acc[v] = 0.0;
#pragma unroll
                        for (int u = 0; u < kPairChunk / 16; u++) {
                            if (pe[u].x < 0) continue;
                            const int l = pe[u].z;
                            const double2* Di = reinterpret_cast<const double2*>(P.Dinv + (size_t)l * 6);
                            const double2 Da = Di[0], Db = Di[1], Dc = Di[2];
                            const double D0 = Da.x, D1 = Da.y, D2 = Db.x, D3 = Db.y, D4 = Dc.x, D5 = Dc.y;
                            double B1[10], B2[20];   // rows 3 * half .. 3 * half + 2 of the first block, the whole second block
                            {
                                const double2* q1 = reinterpret_cast<const double2*>(P.Hpl + (size_t)pe[u].x * kHplStride + half * 10);
                                const double2* q2 = reinterpret_cast<const double2*>(P.Hpl + (size_t)pe[u].y * kHplStride);
#pragma unroll
                                for (int i = 0; i < 5; i++) { const double2 v = q1[i]; B1[2 * i] = v.x; B1[2 * i + 1] = v.y; }
#pragma unroll
                                for (int i = 0; i < 10; i++) { const double2 v = q2[i]; B2[2 * i] = v.x; B2[2 * i + 1] = v.y; }
                            }
                            double BD[9];
#pragma unroll
                            for (int a = 0; a < 3; a++) {
                                const double b0 = B1[a * 3], b1 = B1[a * 3 + 1], b2 = B1[a * 3 + 2];
                                BD[a * 3] = b0 * D0 + b1 * D1 + b2 * D2;
                                BD[a * 3 + 1] = b0 * D1 + b1 * D3 + b2 * D4;
                                BD[a * 3 + 2] = b0 * D2 + b1 * D4 + b2 * D5;
                            }
                            if (diag) { // once per observation: the gradient part
                                const double* d = P.db + (size_t)l * 3;
                                const double d0 = d[0], d1 = d[1], d2 = d[2];
#pragma unroll
                                for (int a = 0; a < 3; a++) g[a] += B1[a * 3] * d0 + B1[a * 3 + 1] * d1 + B1[a * 3 + 2] * d2;
                            }
#pragma unroll
                            for (int b2 = 0; b2 < 6; b2++) {
                                const double q0 = B2[hpl_idx(b2, 0)], q1 = B2[hpl_idx(b2, 1)], q2 = B2[hpl_idx(b2, 2)];
#pragma unroll
                                for (int a = 0; a < 3; a++) acc[b2 * 3 + a] += BD[a * 3] * q0 + BD[a * 3 + 1] * q1 + BD[a * 3 + 2] * q2;
                            }
                        }
                        // over the 16 lanes of each half: transposed butterfly for values 0 .. 15 (15 shuffles, lane `sub` ends up
                        // with the total of value `sub`), plain butterflies for the two left over and the gradient
#pragma unroll
                        for (int off = 8; off >= 1; off >>= 1) {
                            const bool upper = (lane & off) != 0;
#pragma unroll
                            for (int i = 0; i < off; i++) {
                                const double send = upper ? acc[i] : acc[i + off];
                                const double keep = upper ? acc[i + off] : acc[i];
                                acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                            }
                        }
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) {
                            acc[16] += __shfl_xor_sync(0xffffffffu, acc[16], o);
                            acc[17] += __shfl_xor_sync(0xffffffffu, acc[17], o);
                        }
                        if (diag) {
#pragma unroll
                            for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
                                for (int a = 0; a < 3; a++) g[a] += __shfl_xor_sync(0xffffffffu, g[a], o);
                            }
                        }
                        double* dst = P.Hs + (size_t)(6 * c2) * P.dimPad + 6 * c1 + 3 * half;
                        {   // value v = 3 * b2 + a; diagonal block: lower triangle only (column <= row)
                            const int b2 = sub / 3, a = sub - 3 * b2;
                            if (!diag || 3 * half + a <= b2) atomicAdd(&dst[(size_t)b2 * P.dimPad + a], -acc[0]);
                            if (sub < 2) atomicAdd(&dst[(size_t)5 * P.dimPad + 1 + sub], -(sub == 0 ? acc[16] : acc[17]));   // row 5: never above the diagonal
                        }
                        if (diag && sub < 3) atomicAdd(&P.bs[6 * c1 + 3 * half + sub], -(sub == 0 ? g[0] : sub == 1 ? g[1] : g[2]));
                    }
                }
                grid.sync();
                tick(3);
                // ---------------- C: reduced camera system on one CTA ----------------
                if (P.grid_chol) {   // hundreds of free keyframes (global BA): blocked Cholesky over every SM
                    grid_cholesky_solve(grid, P.dimPad, P.Hs, P.bs, P.x, P.Linv, &P.flags[0], smem);
                } else if (blockIdx.x < kClusterCtas) { // the first cluster
                    bool ok = true;
                    if (P.dimP > 0) ok = cluster_cholesky_solve(P.dimPad, P.Hs, P.bs, P.x, P.Linv, smem, &s_flag, P.prof);
                    if (blockIdx.x == 0 && tid == 0) P.flags[0] = ok ? 1 : 0;
                }
                grid.sync();
                tick(4);
                const bool ok2 = P.flags[0] != 0;
                // ---------------- B1: landmark back-substitution and update into the trial buffers ----------------
                double scale_part = 0;
                for (int l = gwarp; l < P.np; l += nwarps) {
                    double cl[3] = { 0, 0, 0 };
                    const int s = P.pt_start[l], e_end = P.pt_start[l + 1];
                    for (int k = s + lane; k < e_end; k += 32) {
                        const int e = P.pt_edges[k];
                        const int c1 = P.cam_col[P.ecam[e]];
                        if (c1 < 0) continue;
                        const double* B1 = P.Hpl + (size_t)e * kHplStride;
                        const double* xp = P.x + 6 * c1;
    #pragma unroll
                        for (int a = 0; a < 6; a++) { cl[0] -= B1[hpl_idx(a, 0)] * xp[a]; cl[1] -= B1[hpl_idx(a, 1)] * xp[a]; cl[2] -= B1[hpl_idx(a, 2)] * xp[a]; }
                    }
                    cl[0] = warp_sum(cl[0]); cl[1] = warp_sum(cl[1]); cl[2] = warp_sum(cl[2]);
                    if (lane == 0) {
                        const double* g = P.bl + (size_t)l * 3;
                        const double* Di = P.Dinv + (size_t)l * 6;
                        const double c0 = g[0] + cl[0], c1v = g[1] + cl[1], c2v = g[2] + cl[2];
                        double xl[3];
                        xl[0] = Di[0] * c0 + Di[1] * c1v + Di[2] * c2v;
                        xl[1] = Di[1] * c0 + Di[3] * c1v + Di[4] * c2v;
                        xl[2] = Di[2] * c0 + Di[4] * c1v + Di[5] * c2v;
                        if (!ok2) { xl[0] = xl[1] = xl[2] = 0.0; }
    #pragma unroll
                        for (int a = 0; a < 3; a++) {
                            P.x[P.dimP + 3 * l + a] = xl[a];
                            P.pts[trial][3 * l + a] = P.pts[cur][3 * l + a] + xl[a];
                            scale_part += xl[a] * (lambda * xl[a] + g[a]);
                        }
                    }
                }
                for (int c = gtid; c < P.nc; c += nthreads) {
                    const int cf = P.cam_col[c];
                    if (cf >= 0) {
                        const double* xp = P.x + 6 * cf;
                        se3_update(xp, P.camq[cur] + 4 * c, P.camt[cur] + 3 * c, P.camq[trial] + 4 * c, P.camt[trial] + 3 * c);
    #pragma unroll
                        for (int a = 0; a < 6; a++) scale_part += xp[a] * (lambda * xp[a] + P.bp[6 * cf + a]);
                    } else {
                        for (int a = 0; a < 4; a++) P.camq[trial][4 * c + a] = P.camq[cur][4 * c + a];
                        for (int a = 0; a < 3; a++) P.camt[trial][3 * c + a] = P.camt[cur][3 * c + a];
                    }
                }
                {
                    double v[1] = { scale_part };
                    cta_sum<1>(v, warp_buf, red);
                    if (tid == 0) P.part[blockIdx.x * kPartStride + kPartScale] = red[0];
                }
                grid.sync();
                tick(5);
                // ---------------- B2: errors at the trial estimate ----------------
                double chi_t = 0;
                for (int e = gtid; e < P.ne; e += nthreads) {
                    if (P.level && P.level[e]) continue;
                    double xc[3], r[2];
                    edge_residual(P, trial, e, xc, r, nullptr);
                    P.err[2 * e] = r[0]; P.err[2 * e + 1] = r[1];
                    const double om = (double)P.einfo[e];
                    chi_t += huber_rho0(delta, dsqr, r[0] * (om * r[0]) + r[1] * (om * r[1]));
                }
                {
                    double v[1] = { chi_t };
                    cta_sum<1>(v, warp_buf, red);
                    if (tid == 0) P.part[blockIdx.x * kPartStride + kPartTrialChi] = red[0];
                }
                grid.sync();
                tick(6);
                double tempChi = 0, scale = 0;
                for (int b2 = 0; b2 < G; b2++) { tempChi += P.part[b2 * kPartStride + kPartTrialChi]; scale += P.part[b2 * kPartStride + kPartScale]; }
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
                    cur = trial; // discardTop: the trial buffers become the estimate
                } else {
                    lambda *= ni;
                    ni *= 2; // pop: keep `cur`
                }
                qmax++;
                trials++;
                aborted = agree_abort();
            } while (rho < 0 && qmax < 10 && !aborted);
            done++;
            last_chi = currentChi;
            if (qmax == 10 || rho == 0) stop = true;
            else {
                if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
                else nBad = 0;
                if (nBad >= 3) stop = true;
            }
            if (aborted) stop = true;
        }
    }
    // ---------------- outlier test and write-back (O3/src/Optimizer.cc:1313-1386) ----------------
    for (int e = gtid; e < P.ne; e += nthreads) {
        const double om = (double)P.einfo[e];
        const double r0 = P.err[2 * e], r1 = P.err[2 * e + 1];
        const double chi = r0 * (om * r0) + r1 * (om * r1);
        double xc[3], r[2];
        edge_residual(P, cur, e, xc, r, nullptr);
        if (P.out_chi2) P.out_chi2[e] = chi;
        P.out_bad[e] = (chi > 5.991 || !(xc[2] > 0.0)) ? 1 : 0;
    }
    for (int c = gtid; c < P.nc; c += nthreads) {
        for (int a = 0; a < 4; a++) P.out_camq[4 * c + a] = (float)P.camq[cur][4 * c + a];
        for (int a = 0; a < 3; a++) P.out_camt[3 * c + a] = (float)P.camt[cur][3 * c + a];
    }
    for (int i = gtid; i < P.np * 3; i += nthreads) P.out_pts[i] = (float)P.pts[cur][i];
    if (gtid == 0) {
        P.out_stats[0] = done; P.out_stats[1] = trials; P.out_stats[2] = first_chi; P.out_stats[3] = last_chi;
        P.out_stats[4] = done_first; P.out_stats[5] = excluded;
    }
}

} // namespace

struct dvm_lba {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_done = nullptr;
    int max_free = 0;
    int grid = 0;
    size_t smem_bytes = 0;
    uint8_t* d_buf = nullptr; size_t d_cap = 0;
    uint8_t* h_buf = nullptr; size_t h_cap = 0;
    int* h_abort = nullptr; int* d_abort = nullptr; // mapped pinned
    std::vector<int> pt_start, cam_start, fill;     // host-side structure scratch, kept between calls
    std::vector<float> next_cam_K;                  // dvm_lba_set_camera_intrinsics: [nc][4] for the next call
    float last_ms = 0;
};

static void lba_free(dvm_lba* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_buf);
    if (h->h_buf) cudaFreeHost(h->h_buf);
    if (h->h_abort) cudaFreeHost(h->h_abort);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" {

int dvm_lba_create(dvm_lba** out, int device, int max_free_cameras)
{
    DVM_REQUIRE(out != nullptr, "null output handle");
    *out = nullptr;
    DVM_REQUIRE(max_free_cameras >= 1 && max_free_cameras <= kLbaMaxFree, "max_free_cameras must be in 1..2000");
    int rc = select_device(device);
    if (rc != DVM_OK) return rc;
    dvm_lba* h = new dvm_lba;
    h->device = device;
    h->max_free = max_free_cameras;
#define DVM_LCREATE(call)                                                                  \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            set_error("%s failed in dvm_lba_create: %s", #call, cudaGetErrorString(e__));  \
            lba_free(h);                                                                   \
            return DVM_ERR_CUDA;                                                           \
        }                                                                                  \
    } while (0)
    DVM_LCREATE(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    DVM_LCREATE(cudaEventCreate(&h->ev0));
    DVM_LCREATE(cudaEventCreate(&h->ev1));
    DVM_LCREATE(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
    DVM_LCREATE(cudaHostAlloc(&h->h_abort, sizeof(int), cudaHostAllocMapped));
    *h->h_abort = 0;
    DVM_LCREATE(cudaHostGetDevicePointer(&h->d_abort, h->h_abort, 0));
    // shared memory: max(Schur row [nf*36+6], Cholesky panel [n*kNB] + diag [kNB*(kNB+1)]) doubles
    // up to kLbaClusterFree free keyframes the reduced system is factored by ONE cluster with its panel in shared memory
    // (local BA: lowest latency); above that by the whole grid (grid_cholesky_solve), whose shared memory need is fixed
    const size_t n = ((size_t)6 * std::min(max_free_cameras, kLbaClusterFree) + kNB - 1) / kNB * kNB;
    h->smem_bytes = std::max(((n + 8) * kPanelLd + 2 * kNB * kDiagLd + 64 * kPanelLd) * sizeof(double),
                             (size_t)kGridCholSmemDoubles * sizeof(double));
    DVM_REQUIRE(h->smem_bytes <= 227 * 1024, "max_free_cameras needs more shared memory than one SM has");
    DVM_LCREATE(cudaFuncSetAttribute(lba_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
    {   // persistent CTAs in clusters of 8: as many clusters as can be co-resident (cooperative launch)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kClusterCtas * 64);
        cfg.blockDim = dim3(kLbaThreads);
        cfg.dynamicSmemBytes = h->smem_bytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kClusterCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = 0;
        DVM_LCREATE(cudaOccupancyMaxActiveClusters(&nclusters, lba_kernel, &cfg));
        int sms = 0;
        DVM_LCREATE(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        nclusters = std::min(nclusters, sms / kClusterCtas);
        if (nclusters < 1) { set_error("lba kernel: no 8-CTA cluster fits on this device"); lba_free(h); return DVM_ERR_CUDA; }
        h->grid = nclusters * kClusterCtas;
    }
#undef DVM_LCREATE
    *out = h;
    return DVM_OK;
}

void dvm_lba_destroy(dvm_lba* h) { lba_free(h); }

float dvm_lba_last_kernel_ms(const dvm_lba* h) { return h ? h->last_ms : -1.f; }

int dvm_lba_set_camera_intrinsics(dvm_lba* h, int nc, const float* cam_K)
{
    DVM_REQUIRE(h != nullptr && nc >= 0 && (nc == 0 || cam_K), "bad argument");
    h->next_cam_K.assign(cam_K, cam_K + (size_t)nc * 4);
    return DVM_OK;
}

int dvm_local_ba(dvm_lba* h, int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne,
                 const int32_t* edge_cam, const int32_t* edge_pt, const float* edge_obs, const float* edge_inv_sigma2,
                 const float* K, int iterations, const volatile uint8_t* abort_flag, double* edge_chi2, uint8_t* edge_bad,
                 double* stats, int* iters_done)
{
    // const float thHuberMono = sqrt(5.991), O3/src/Optimizer.cc:1178
    return dvm_bundle_adjustment(h, nc, cam_q, cam_t, cam_fixed, np, pts, ne, edge_cam, edge_pt, edge_obs, edge_inv_sigma2, K,
                                 iterations, (float)std::sqrt(5.991), abort_flag, edge_chi2, edge_bad, stats, iters_done);
}

// iterations2 > 0 adds the welding BA's second pass; nstats = entries of `stats` the caller holds (4 or 6)
static int run_ba(dvm_lba* h, int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne,
                  const int32_t* edge_cam, const int32_t* edge_pt, const float* edge_obs, const float* edge_inv_sigma2,
                  const float* K, int iterations, float huber_delta, int iterations2, const volatile uint8_t* abort_flag,
                  double* edge_chi2, uint8_t* edge_bad, double* stats, int nstats, int* iters_done)
{
    DVM_REQUIRE(h != nullptr && iters_done != nullptr, "null argument");
    *iters_done = -1;
    if (stats) for (int i = 0; i < nstats; i++) stats[i] = 0;
    DVM_REQUIRE(nc >= 0 && np >= 0 && ne >= 0 && iterations >= 0 && iterations2 >= 0, "negative size");
    DVM_REQUIRE(nc == 0 || (cam_q && cam_t && cam_fixed), "null camera arrays");
    DVM_REQUIRE(np == 0 || pts, "null point array");
    DVM_REQUIRE(ne == 0 || (edge_cam && edge_pt && edge_obs && edge_inv_sigma2 && edge_bad), "null edge arrays");
    DVM_REQUIRE(K != nullptr, "null intrinsics");
    DVM_REQUIRE(huber_delta > 0.0f, "huber_delta must be positive (infinity = no robust kernel)");
    const bool hprof = getenv("DVM_LBA_PROFILE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    const auto ht0 = now();
    // host-side structure building (the analogue of BlockSolver::buildStructure, block_solver.hpp:143-295)
    std::vector<int> cam_col(nc, -1), free_cam;
    int nfixed = 0;
    for (int c = 0; c < nc; c++) {
        if (cam_fixed[c]) nfixed++;
        else { cam_col[c] = (int)free_cam.size(); free_cam.push_back(c); }
    }
    const int nf = (int)free_cam.size();
    if (nfixed == 0) return DVM_OK;                 // "LBA aborted": no fixed keyframe (:1088-1091)
    if (abort_flag && *abort_flag) return DVM_OK;   // :1306-1308
    if (ne == 0 || nf + np == 0) return DVM_OK;
    // max_free_cameras of dvm_lba_create is a sizing hint, not a limit: a window with more free keyframes than the
    // shared-memory panel of the one-cluster solve was sized for goes to the grid-wide solve
    if (nf > kLbaMaxFree) { set_error("%d free keyframes exceed the dense reduced-camera solver's limit of %d", nf, kLbaMaxFree); return DVM_ERR_CAPACITY; }
    // one pass over the edges: validation, both histograms, and whether the edges already come grouped by point
    // (the reference creates them point by point, O3/src/Optimizer.cc:1182-1232, and so does the host adapter)
    std::vector<int>& pt_start = h->pt_start;
    std::vector<int>& cam_start = h->cam_start;
    bool by_point = true;
    size_t pairs_cap = 0;   // capacity of the Schur complement's pair lists: sum over the landmarks of k (k + 1) / 2 (k = observations)
    {   // validation, and whether the edges come grouped by point: then the device builds both structures itself
        int prev = -1;
        unsigned bad = 0;
        size_t run = 0;
        for (int e = 0; e < ne; e++) {
            const int c = edge_cam[e], l = edge_pt[e];
            bad |= (unsigned)((unsigned)c >= (unsigned)nc) | (unsigned)((unsigned)l >= (unsigned)np);
            by_point = by_point && l >= prev;
            if (l != prev) { pairs_cap += run * (run + 1) / 2; run = 0; }
            run++;
            prev = l;
        }
        pairs_cap += run * (run + 1) / 2;
        DVM_REQUIRE(bad == 0, "edge index out of range");
    }
    static const bool host_structure = getenv("DVM_LBA_HOST_STRUCTURE") != nullptr;   // (diagnostics: force the host build)
    const bool device_build = by_point && !host_structure;
    if (!device_build) {
        pt_start.assign((size_t)np + 1, 0);
        cam_start.assign((size_t)nf + 1, 0);
        for (int e = 0; e < ne; e++) {
            pt_start[edge_pt[e] + 1]++;
            const int cf = cam_col[edge_cam[e]];
            if (cf >= 0) cam_start[cf + 1]++;
        }
        pairs_cap = 0;
        for (int l = 0; l < np; l++) {
            const size_t k = (size_t)pt_start[l + 1];
            pairs_cap += k * (k + 1) / 2;
            pt_start[l + 1] += pt_start[l];
        }
        for (int c = 0; c < nf; c++) cam_start[c + 1] += cam_start[c];
    }
    if (pairs_cap >= ((size_t)1 << 31)) { set_error("the landmarks' observation pairs (%zu) exceed the Schur complement's work list", pairs_cap); return DVM_ERR_CAPACITY; }
    const size_t nbins = (size_t)nf * (nf + 1) / 2;
    const size_t n_cam_edges = (size_t)std::max(ne, 1);
    const auto ht1 = now();
    DVM_CUDA(cudaSetDevice(h->device));
    const int dimP = 6 * nf;
    const int dimPad = (dimP + kNB - 1) / kNB * kNB;
    // ---- device layout ----
    size_t off = 0;
    auto take = [&](size_t bytes) { off = (off + 255) & ~(size_t)255; size_t o = off; off += bytes; return o; };
    // uploaded block
    const size_t o_camq = take((size_t)nc * 4 * 8), o_camt = take((size_t)nc * 3 * 8), o_pts = take((size_t)np * 3 * 8);
    const size_t o_col = take((size_t)nc * 4), o_free = take((size_t)std::max(nf, 1) * 4), o_camk = take((size_t)std::max(nc, 1) * 4 * 8);
    const size_t o_ecam = take((size_t)ne * 4), o_ept = take((size_t)ne * 4), o_obs = take((size_t)ne * 8), o_info = take((size_t)ne * 4);
    const size_t o_pst = take((size_t)(np + 1) * 4), o_ped = take((size_t)ne * 4);
    const size_t o_cst = take((size_t)(nf + 1) * 4), o_ced = take(n_cam_edges * 4);
    const size_t upload_bytes = device_build ? o_pst : off;
    // output block (contiguous, one D2H); the pinned host mirror spans [0, out_end) only
    const size_t out_begin = (off + 255) & ~(size_t)255;
    const size_t o_oq = take((size_t)nc * 4 * 4), o_ot = take((size_t)nc * 3 * 4), o_op = take((size_t)np * 3 * 4);
    const size_t o_ochi = take((size_t)ne * 8), o_obad = take((size_t)ne), o_ostats = take(6 * 8);
    const size_t out_end = off;
    // work block (device only)
    const size_t o_camq1 = take((size_t)nc * 4 * 8), o_camt1 = take((size_t)nc * 3 * 8), o_pts1 = take((size_t)np * 3 * 8);
    const size_t o_err = take((size_t)ne * 2 * 8), o_hpl = take((size_t)ne * kHplStride * 8);
    const size_t o_hll = take((size_t)np * 6 * 8), o_bl = take((size_t)np * 3 * 8), o_dinv = take((size_t)np * 6 * 8), o_db = take((size_t)np * 3 * 8);
    const size_t o_hpp = take((size_t)std::max(nf, 1) * 36 * 8), o_bp = take((size_t)std::max(dimP, 1) * 8);
    const size_t o_hs = take((size_t)std::max((size_t)(dimPad + 8) * dimPad, (size_t)1) * 8), o_bs = take((size_t)std::max(dimPad, 1) * 8);
    const size_t o_linv = take((size_t)std::max(dimPad * kNB, 1) * 8);
    const size_t o_x = take((size_t)(dimPad + np * 3 + 1) * 8);
    const size_t o_part = take((size_t)h->grid * kPartStride * 8), o_flags = take(4 * 4), o_prof = take(32 * 8);
    const size_t o_level = take((size_t)std::max(ne, 1));
    const size_t o_cnt = take(device_build ? (size_t)std::max(nf, 1) * h->grid * kLbaWarps * 4 : 4), o_ctot = take((size_t)(nf + 1) * 4);
    const size_t o_pstart = take((nbins + 1) * 4), o_pfill = take(std::max(nbins, (size_t)1) * 4), o_pairs = take(std::max(pairs_cap, (size_t)1) * 16),
                 o_items = take((pairs_cap / kPairChunk + nbins + 1) * 16);
    const size_t total = off + 256;
    if (total > h->d_cap) {
        DVM_CUDA(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_buf); h->d_buf = nullptr;
        const size_t cap = total + total / 4;
        DVM_CUDA(cudaMalloc(&h->d_buf, cap));
        DVM_CUDA(cudaMemsetAsync(h->d_buf, 0, std::min(cap, out_end + 4096), h->stream)); // the alignment gaps of the output block travel in its one D2H copy
        h->d_cap = cap;
    }
    if (out_end + 256 > h->h_cap) {
        DVM_CUDA(cudaStreamSynchronize(h->stream));
        if (h->h_buf) { cudaFreeHost(h->h_buf); h->h_buf = nullptr; }
        const size_t cap = out_end + out_end / 4 + 4096;
        DVM_CUDA(cudaHostAlloc(&h->h_buf, cap, cudaHostAllocDefault));
        h->h_cap = cap;
    }
    // ---- stage inputs (float -> double conversions as the reference's .cast<double>()) ----
    uint8_t* hb = h->h_buf;
    {
        double* q = (double*)(hb + o_camq); double* t = (double*)(hb + o_camt); double* p = (double*)(hb + o_pts);
        for (int c = 0; c < nc; c++) {
            double x = cam_q[4 * c], y = cam_q[4 * c + 1], z = cam_q[4 * c + 2], w = cam_q[4 * c + 3];
            if (w < 0) { x = -x; y = -y; z = -z; w = -w; }           // SE3Quat ctor: normalizeRotation()
            const double n = std::sqrt(x * x + y * y + z * z + w * w);
            q[4 * c] = x / n; q[4 * c + 1] = y / n; q[4 * c + 2] = z / n; q[4 * c + 3] = w / n;
            for (int i = 0; i < 3; i++) t[3 * c + i] = cam_t[3 * c + i];
        }
        for (int i = 0; i < np * 3; i++) p[i] = pts[i];
        {   // intrinsics per camera: the caller's K for all, or the per-camera table set for this call
            double* kc = (double*)(hb + o_camk);
            const bool per_cam = (int)h->next_cam_K.size() == nc * 4 && nc > 0;
            for (int c = 0; c < nc; c++)
                for (int i = 0; i < 4; i++) kc[4 * c + i] = per_cam ? (double)h->next_cam_K[4 * c + i] : (double)K[i];
            h->next_cam_K.clear();
        }
        memcpy(hb + o_col, cam_col.data(), (size_t)nc * 4);
        memcpy(hb + o_free, free_cam.data(), (size_t)nf * 4);
        memcpy(hb + o_ecam, edge_cam, (size_t)ne * 4);
        memcpy(hb + o_ept, edge_pt, (size_t)ne * 4);
        memcpy(hb + o_obs, edge_obs, (size_t)ne * 8);
        memcpy(hb + o_info, edge_inv_sigma2, (size_t)ne * 4);
        if (!device_build) {
            memcpy(hb + o_pst, pt_start.data(), (size_t)(np + 1) * 4);
            memcpy(hb + o_cst, cam_start.data(), (size_t)(nf + 1) * 4);
            // the two edge lists are filled straight into the pinned upload buffer (stable: ascending edge index per row)
            int* pt_edges = reinterpret_cast<int*>(hb + o_ped);
            int* cam_edges = reinterpret_cast<int*>(hb + o_ced);
            if (by_point) {
                for (int e = 0; e < ne; e++) pt_edges[e] = e;
            } else {
                std::vector<int>& fill = h->fill;
                fill.assign(pt_start.begin(), pt_start.end() - 1);
                for (int e = 0; e < ne; e++) pt_edges[fill[edge_pt[e]]++] = e;
            }
            {
                std::vector<int>& fill = h->fill;
                fill.assign(cam_start.begin(), cam_start.end() - 1);
                for (int e = 0; e < ne; e++) {
                    const int cf = cam_col[edge_cam[e]];
                    if (cf >= 0) cam_edges[fill[cf]++] = e;
                }
            }
        }
    }
    const auto ht2 = now();
    DVM_CUDA(cudaMemcpyAsync(h->d_buf, hb, upload_bytes, cudaMemcpyHostToDevice, h->stream));
    uint8_t* db = h->d_buf;
    LbaDev P;
    memset(&P, 0, sizeof(P));
    P.nc = nc; P.nf = nf; P.np = np; P.ne = ne; P.dimP = dimP; P.dimPad = dimPad; P.iterations = iterations;
    P.iterations2 = iterations2;
    {
        const size_t n = (size_t)((6 * nf + kNB - 1) / kNB * kNB);
        const size_t cluster_need = ((n + 8) * kPanelLd + 2 * kNB * kDiagLd + 64 * kPanelLd) * sizeof(double);
        P.grid_chol = (nf > kLbaClusterFree || cluster_need > h->smem_bytes) ? 1 : 0;
    }
    P.level = iterations2 > 0 ? db + o_level : nullptr;
    P.camK = (const double*)(db + o_camk);
    P.delta = (double)huber_delta;   // the caller's float delta; +infinity = no robust kernel
    P.dsqr = P.delta * P.delta;
    P.camq[0] = (double*)(db + o_camq); P.camq[1] = (double*)(db + o_camq1);
    P.camt[0] = (double*)(db + o_camt); P.camt[1] = (double*)(db + o_camt1);
    P.pts[0] = (double*)(db + o_pts); P.pts[1] = (double*)(db + o_pts1);
    P.cam_col = (const int*)(db + o_col); P.free_cam = (const int*)(db + o_free);
    P.ecam = (const int*)(db + o_ecam); P.ept = (const int*)(db + o_ept);
    P.eobs = (const float*)(db + o_obs); P.einfo = (const float*)(db + o_info);
    P.pt_start = (const int*)(db + o_pst); P.pt_edges = (const int*)(db + o_ped);
    P.cam_start = (const int*)(db + o_cst); P.cam_edges = (const int*)(db + o_ced);
    P.build = device_build ? 1 : 0;
    P.pt_start_w = (int*)(db + o_pst); P.pt_edges_w = (int*)(db + o_ped); P.cam_start_w = (int*)(db + o_cst); P.cam_edges_w = (int*)(db + o_ced);
    P.cnt = (int*)(db + o_cnt); P.cam_tot = (int*)(db + o_ctot);
    P.nbins = (int)nbins; P.pair_start = (int*)(db + o_pstart); P.pair_fill = (int*)(db + o_pfill); P.pairs = (int4*)(db + o_pairs); P.items = (int4*)(db + o_items);
    P.err = (double*)(db + o_err); P.Hpl = (double*)(db + o_hpl); P.Hll = (double*)(db + o_hll);
    P.bl = (double*)(db + o_bl); P.Dinv = (double*)(db + o_dinv); P.db = (double*)(db + o_db);
    P.Hpp = (double*)(db + o_hpp); P.bp = (double*)(db + o_bp); P.Hs = (double*)(db + o_hs); P.bs = (double*)(db + o_bs);
    P.Linv = (double*)(db + o_linv);
    P.x = (double*)(db + o_x); P.part = (double*)(db + o_part); P.flags = (int*)(db + o_flags);
    P.out_camq = (float*)(db + o_oq); P.out_camt = (float*)(db + o_ot); P.out_pts = (float*)(db + o_op);
    P.out_chi2 = (double*)(db + o_ochi); P.out_bad = db + o_obad; P.out_stats = (double*)(db + o_ostats);
    *h->h_abort = 0;
    P.abort_flag = abort_flag ? h->d_abort : nullptr;
    DVM_CUDA(cudaMemsetAsync(db + o_flags, 0, 16, h->stream));
    DVM_CUDA(cudaMemsetAsync(db + o_prof, 0, 256, h->stream));
    P.prof = (unsigned long long*)(db + o_prof);
    DVM_CUDA(cudaMemsetAsync(db + o_err, 0, (size_t)ne * 2 * 8, h->stream));
    if (iterations2 > 0) DVM_CUDA(cudaMemsetAsync(db + o_level, 0, (size_t)ne, h->stream));
    void* args[] = { &P };
    DVM_CUDA(cudaEventRecord(h->ev0, h->stream));
    DVM_CUDA(cudaLaunchCooperativeKernel((void*)lba_kernel, dim3(h->grid), dim3(kLbaThreads), args, h->smem_bytes, h->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    DVM_CUDA(cudaEventRecord(h->ev1, h->stream));
    const size_t out_bytes = out_end - out_begin;
    DVM_CUDA(cudaMemcpyAsync(hb + out_begin, db + out_begin, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (abort_flag) { // mirror the caller's pbStopFlag into mapped memory while the kernel runs
        while (cudaEventQuery(h->ev1) == cudaErrorNotReady) {
            if (*abort_flag) *h->h_abort = 1;
            std::this_thread::sleep_for(std::chrono::microseconds(20));   // (an LM trial takes ~0.25 ms: the flag is seen in time)
        }
    }
    const auto ht3 = now();
    // The calling thread has nothing to do for the length of the kernel (~1.4 ms at C4).  It polls the completion event with
    // PAUSE bursts in between instead of spinning flat out in the driver: with one agent per GPU and eight agents on a
    // 16-thread host, the hyper-thread siblings of spinning waiters otherwise slow the other agents' staging and unpacking
    // loops down (1.64 -> 1.99 ms per BA at N = 8, measured).  (A blocking wait -- cudaEventBlockingSync -- wakes up 0.4 ms
    // late on this box: measured, not used.)
    DVM_CUDA(cudaEventRecord(h->ev_done, h->stream));
    for (;;) {
        const cudaError_t q = cudaEventQuery(h->ev_done);
        if (q == cudaSuccess) break;
        if (q != cudaErrorNotReady) { DVM_CUDA(q); }
        for (int i = 0; i < 32; i++) cpu_relax();
    }
    const auto ht4 = now();
    DVM_CUDA(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    if (getenv("DVM_LBA_PROFILE")) {
        unsigned long long pr[32];
        cudaMemcpy(pr, db + o_prof, 256, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[chol us, CTA 0, inside trsm] panel rows loaded %.1f multiplied %.1f\n", pr[16] * 1e-3, pr[17] * 1e-3);
        fprintf(stderr, "[chol us, CTA 0] first factor %.1f fetch Linv %.1f trsm %.1f sync %.1f next-diagonal update %.1f look-ahead factor %.1f wait for the trailing tiles %.1f backsub %.1f\n", pr[8] * 1e-3,
                pr[9] * 1e-3, pr[10] * 1e-3, pr[11] * 1e-3, pr[12] * 1e-3, pr[13] * 1e-3, pr[14] * 1e-3, pr[15] * 1e-3);
        fprintf(stderr, "[lba phases us] L1 %.1f L2 %.1f S0 %.1f S1 %.1f C %.1f B1 %.1f B2 %.1f total %.1f\n", pr[0] * 1e-3, pr[1] * 1e-3,
                pr[2] * 1e-3, pr[3] * 1e-3, pr[4] * 1e-3, pr[5] * 1e-3, pr[6] * 1e-3, h->last_ms * 1e3);
    }
    const double* ost = (const double*)(hb + o_ostats);
    const float* oq = (const float*)(hb + o_oq);
    const float* ot = (const float*)(hb + o_ot);
    for (int c = 0; c < nc; c++) {
        if (cam_col[c] < 0) continue; // only optimised keyframes are written back
        for (int i = 0; i < 4; i++) cam_q[4 * c + i] = oq[4 * c + i];
        for (int i = 0; i < 3; i++) cam_t[3 * c + i] = ot[3 * c + i];
    }
    memcpy(pts, hb + o_op, (size_t)np * 3 * 4);
    if (edge_chi2) memcpy(edge_chi2, hb + o_ochi, (size_t)ne * 8);
    memcpy(edge_bad, hb + o_obad, (size_t)ne);
    if (stats) for (int i = 0; i < nstats; i++) stats[i] = ost[i];
    *iters_done = (int)ost[0];
    if (hprof) {
        auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
        fprintf(stderr, "[lba host us] structure %.1f staging %.1f enqueue %.1f wait %.1f (kernel %.1f) unpack %.1f\n", us(ht0, ht1),
                us(ht1, ht2), us(ht2, ht3), us(ht3, ht4), h->last_ms * 1e3, us(ht4, now()));
    }
    if (!std::isfinite(ost[3])) { set_error("local BA produced a non-finite chi2"); return DVM_ERR_NUMERIC; }
    return DVM_OK;
}

int dvm_bundle_adjustment(dvm_lba* h, int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne,
                          const int32_t* edge_cam, const int32_t* edge_pt, const float* edge_obs,
                          const float* edge_inv_sigma2, const float* K, int iterations, float huber_delta,
                          const volatile uint8_t* abort_flag, double* edge_chi2, uint8_t* edge_bad, double* stats,
                          int* iters_done)
{
    return run_ba(h, nc, cam_q, cam_t, cam_fixed, np, pts, ne, edge_cam, edge_pt, edge_obs, edge_inv_sigma2, K, iterations,
                  huber_delta, 0, abort_flag, edge_chi2, edge_bad, stats, 4, iters_done);
}

int dvm_merge_ba(dvm_lba* h, int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne,
                 const int32_t* edge_cam, const int32_t* edge_pt, const float* edge_obs, const float* edge_inv_sigma2,
                 const float* K, const volatile uint8_t* abort_flag, double* edge_chi2, uint8_t* edge_bad, double* stats,
                 int* iters_done)
{
    return run_ba(h, nc, cam_q, cam_t, cam_fixed, np, pts, ne, edge_cam, edge_pt, edge_obs, edge_inv_sigma2, K, 5,
                  (float)std::sqrt(5.99), 10, abort_flag, edge_chi2, edge_bad, stats, 6, iters_done);
}

} // extern "C"
