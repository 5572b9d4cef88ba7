// chol_device.cuh -- device pieces shared by the dense SPD solvers of the optimisation kernels (lba.cu: the reduced camera
// system on one 8-CTA cluster; grid_cholesky_solve below: the same factorisation spread over a whole cooperative grid for
// systems too large for one SM's shared memory -- global bundle adjustment over hundreds of keyframes, essential graph).
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>

namespace dvm {

constexpr int kCholThreads = 512;            // every kernel using these helpers runs 512-thread CTAs (16 warps)
constexpr int kCholWarps = kCholThreads / 32;
constexpr int kNB = 32;                      // Cholesky block size
constexpr int kPanelLd = 36;   // shared-memory row stride of a panel (doubles): conflict-free 8-byte fragment loads
constexpr int kDiagLd = kPanelLd;   // diagonal block and its inverse: also read as tensor-core fragments

// D = C - A * B for one m8n8k4 FP64 tensor-core tile step (a: A[r=lane/4][k=lane%4], b: B[k=lane%4][c=lane/4])
__device__ inline void dmma_m8n8k4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Reciprocal square root on the critical path of every column: hardware seed (MUFU.RSQ64H, ~20 bits) and two
// Newton steps -- within 1-2 ulp for the positive, well-scaled pivots of a damped normal matrix.
__device__ inline double fast_rsqrt(double d)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = 0.5 * d;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    return fma(y, e, y);
}

// Factorisation + inversion of a 32x32 diagonal block by ALL threads of the CTA (kCholThreads = 512 = 16 warps): a lone
// warp cannot hide its own latencies (a one-warp version with the block in registers ran at ~7 cycles per instruction,
// 21 us per block, and was 2/3 of the whole solve; DESIGN.md, negative results).  Every thread OWNS two elements of the block and of its inverse in registers for the
// whole sweep: lane = row, warp w = columns w and w + 16.  Step j:
//     the warp that owns column j takes the pivot by shuffle, scales its column by rsqrt(pivot) and publishes it
//     (col[j & 1][:], rinv[j & 1]) | ONE CTA barrier |
//     a[r][c] -= L[r][j] L[c][j]  (c > j);   X[j][c] = x[j][c] * rinv (by shuffle inside the column's warp),
//     x[r][c] -= L[r][j] X[j][c]  (r > j)
// i.e. the right-looking Cholesky step and the column sweep of the triangular inversion share one barrier per
// step (the published column is double-buffered), and the owner of column j + 1 publishes it before doing its share of
// step j's inverse updates.  Measured (tools/chol_probe.cu, cycles per 32 x 32 block): one warp 40 000; this scheme
// 18 900; with the validity test of the pivot off the chain 16 300; with the loop unrolled and the pivots carried by every
// lane 14 200 (this code).  A dependent FP64 operation costs ~24 cycles here and a 512-thread barrier ~55, so a step cannot
// go far below 300 cycles; variants that did NOT help: four warps owning eight columns each (15 000), a named barrier the
// publisher only arrives at (18 100), the inverse on warps of its own (16 000), a third-order rsqrt (slower).
// `scratch` holds 4 * kNB + 4 doubles.  Same outputs as the warp version.
// (No __restrict__ here: the published column is exchanged BETWEEN threads, and with restrict-qualified pointers
// nvcc keeps values read from it across the barriers -- measured: wrong factors.)
__device__ inline bool cta_factor_invert_32(double* A, int ld, double* Ld, double* Li, bool write_back, double* Linv_out,
                                     double* scratch)
{
    const int r = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = w, c1 = w + 16;
    double* colbuf = scratch;                 // [2][kNB]  L[:, j]
    double* s_rinv = scratch + 2 * kNB;       // [2]
    double* s_bad = s_rinv + 2;
    double a0 = A[(size_t)r * ld + c0], a1 = A[(size_t)r * ld + c1];
    // the pivots of the two owned columns, carried by EVERY lane (one more multiply-add per step, but the pivot -> rsqrt ->
    // publish -> barrier chain that bounds a step starts without a shuffle)
    double d0 = A[(size_t)c0 * ld + c0], d1 = A[(size_t)c1 * ld + c1];
    double x0 = (r == c0) ? 1.0 : 0.0, x1 = (r == c1) ? 1.0 : 0.0;
    if (threadIdx.x == 0) *s_bad = 0.0;   // (thread 0 is also the lane that raises it for column 0)
    // the warp that owns column j scales it by rsqrt(pivot) and publishes it in buffer j & 1.  A pivot that is not positive
    // and finite only raises the flag (off the chain): the factor is garbage from there on and the callers discard it.
    auto publish = [&](int j) {
        const bool hi = j >= 16;
        const double djj = hi ? d1 : d0;
        const double rinv = fast_rsqrt(djj);
        const double l = (hi ? a1 : a0) * rinv;   // a[j][j] * rinv = sqrt(a[j][j])
        if (hi) a1 = l; else a0 = l;
        if (r >= j) colbuf[(j & 1) * kNB + r] = l;
        if (r == 0) { s_rinv[j & 1] = rinv; if (!(djj > 0) || !isfinite(djj)) *s_bad = 1.0; }
    };
    if (w == 0) publish(0);
#pragma unroll
    for (int j = 0; j < kNB; j++) {
        __syncthreads();   // column j is published
        const double* col = colbuf + (j & 1) * kNB;
        const double lr = (r >= j) ? col[r] : 0.0;
        const double l0 = col[c0], l1 = col[c1];
        // Cholesky update of the columns right of j
        if (c0 > j && r >= c0) a0 -= lr * l0;
        if (c1 > j && r >= c1) a1 -= lr * l1;
        if (c0 > j) d0 -= l0 * l0;
        if (c1 > j) d1 -= l1 * l1;
        // column j + 1 is complete now: its owner publishes it BEFORE the inverse updates of this step, which keeps
        // them off the pivot -> rsqrt -> publish -> barrier chain that bounds a step
        if (j + 1 < kNB && w == ((j + 1) & 15)) publish(j + 1);
        // inverse: row j is scaled, rows below it are swept (columns <= j)
        const double rinv = s_rinv[j & 1];
        if (c0 <= j) {
            const double xj = __shfl_sync(0xffffffffu, x0, j) * rinv;
            if (r == j) x0 = xj; else if (r > j) x0 -= lr * xj;
        }
        if (c1 <= j) {
            const double xj = __shfl_sync(0xffffffffu, x1, j) * rinv;
            if (r == j) x1 = xj; else if (r > j) x1 -= lr * xj;
        }
    }
    Ld[r * kDiagLd + c0] = (c0 <= r) ? a0 : 0.0;
    Ld[r * kDiagLd + c1] = (c1 <= r) ? a1 : 0.0;
    Li[r * kDiagLd + c0] = (c0 <= r) ? x0 : 0.0;
    Li[r * kDiagLd + c1] = (c1 <= r) ? x1 : 0.0;
    if (write_back) {
        if (c0 <= r) A[(size_t)r * ld + c0] = a0;
        if (c1 <= r) A[(size_t)r * ld + c1] = a1;
        Linv_out[r * kNB + c0] = (c0 <= r) ? x0 : 0.0;
        Linv_out[r * kNB + c1] = (c1 <= r) ? x1 : 0.0;
    }
    __syncthreads();
    return *s_bad != 0.0;
}



// out[rows x 32] = S[rows x 32] * Linv^T on the FP64 tensor pipe (the panel step L21 = A21 * L11^-T as a product with the
// inverted diagonal block).  S is staged in shared memory with row stride kPanelLd, Linv with kDiagLd; rows is a multiple
// of 8; every warp takes 8 x 8 output tiles and hands store(row, col, v(row, col), v(row, col + 1)) its two results per
// lane.  Linv is lower triangular, so column tile tc only needs k < 8 (tc + 1).
// (One thread per output with two shared-memory loads per multiply-add was bound by the shared-memory pipe: 2.7 us for a
// 32 x 32 panel; as tensor-core fragments every loaded value feeds 8 multiply-adds.)
template <class Store>
__device__ inline void panel_times_inverse_t(const double* S, int rows, const double* Linv, Store&& store)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ntiles = (rows >> 3) * (kNB / 8);
    for (int t = wid; t < ntiles; t += kCholWarps) {
        const int tr = t >> 2, tc = t & 3;
        const int ar = tr * 8 + (lane >> 2), bc = tc * 8 + (lane >> 2);
        double c0 = 0.0, c1 = 0.0;
        for (int kk = 0; kk < (tc + 1) * 8; kk += 4)
            dmma_m8n8k4(c0, c1, S[ar * kPanelLd + kk + (lane & 3)], Linv[bc * kDiagLd + kk + (lane & 3)]);
        store(ar, tc * 8 + (lane & 3) * 2, c0, c1);
    }
}

// ---- dense SPD solve A x = b on a whole cooperative grid ------------------------------------------------------------
// Blocked right-looking Cholesky (lower triangle, kNB = 32) of the n x n matrix A (n a multiple of 32: the caller pads with
// an identity block; leading dimension n; n + 8 rows, the right-hand side rides along as row n so that after the
// factorisation it holds y = L^-1 b).  Per block column, three grid barriers:
//   (a) CTA 0 factors and inverts the 32 x 32 diagonal block (cta_factor_invert_32) and publishes L11, L11^-1;
//   (b) the panel rows below it, in chunks of 64 rows dealt round-robin to the CTAs:  L21 = A21 * L11^-T;
//   (c) the trailing update A22 -= L21 L21^T in 64 x 64 tiles dealt round-robin, both panel slices staged in shared
//       memory, 8 x 8 sub-tiles on the FP64 tensor pipe (DMMA m8n8k4), lower triangle only.
// Then CTA 0 solves L^T x = y block by block.  *g_ok (global) is 1 on success, 0 when a pivot was not positive / finite
// (x is then zero).  The caller must grid.sync() before reading x.  smem: kGridCholSmemDoubles doubles.
constexpr int kGridCholSmemDoubles = 2 * kNB * kDiagLd + 64 * kPanelLd + 2 * 64 * kPanelLd + 4 * kNB + 8;

__device__ inline void grid_cholesky_solve(cooperative_groups::grid_group& grid, int n, double* A, const double* b, double* x,
                                           double* Linv_g, int* g_ok, double* smem)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    double* Ld = smem;                            // [kNB][kDiagLd]
    double* Li = Ld + kNB * kDiagLd;              // [kNB][kDiagLd]
    double* stage = Li + kNB * kDiagLd;           // [64][kPanelLd]
    double* Pi = stage + 64 * kPanelLd;           // [64][kPanelLd]
    double* Pj = Pi + 64 * kPanelLd;              // [64][kPanelLd]
    double* scratch = Pj + 64 * kPanelLd;         // [4 * kNB + 8]
    const int M = n + 8;
    if (cta == 0) {
        for (int i = tid; i < 8 * n; i += kCholThreads) A[(size_t)n * n + i] = (i < n) ? b[i] : 0.0;
        if (tid == 0) *g_ok = 1;
    }
    grid.sync();
    for (int k0 = 0; k0 < n; k0 += kNB) {
        const int r0 = k0 + kNB;
        // ---- (a) diagonal block ----
        if (cta == 0) {
            const bool bad = cta_factor_invert_32(A + (size_t)k0 * n + k0, n, Ld, Li, true, Linv_g + (size_t)(k0 / kNB) * kNB * kNB, scratch);
            if (bad && tid == 0) *g_ok = 0;
        }
        grid.sync();
        // ---- (b) panel: L21 = A21 * L11^-T ----
        const int mrows = M - r0;
        if (cta * 64 < mrows) {
            const double* Lg = Linv_g + (size_t)(k0 / kNB) * kNB * kNB;
            for (int t = tid; t < kNB * kNB; t += kCholThreads) Li[(t >> 5) * kDiagLd + (t & 31)] = Lg[t];
            __syncthreads();
            for (int base = r0 + cta * 64; base < M; base += G * 64) {
                const int cnt = min(64, M - base);
                for (int t = tid; t < cnt * kNB; t += kCholThreads) stage[(t >> 5) * kPanelLd + (t & 31)] = A[(size_t)(base + (t >> 5)) * n + k0 + (t & 31)];
                __syncthreads();
                panel_times_inverse_t(stage, cnt, Li, [&](int r, int c, double v0, double v1) {
                    double* dst = A + (size_t)(base + r) * n + k0 + c;
                    dst[0] = v0; dst[1] = v1;
                });
                __syncthreads();
            }
        }
        grid.sync();
        // ---- (c) trailing update in 64 x 64 tiles ----
        const int mcols = n - r0;
        if (mcols > 0) {
            const int mt = (mcols + 63) / 64, mrt = (mrows + 63) / 64;
            for (int t = cta; t < mrt * mt; t += G) {
                const int I = t / mt, J = t - I * mt;
                if (J > I) continue;   // strictly above the diagonal
                const int ri = r0 + I * 64, rj = r0 + J * 64;
                __syncthreads();
                for (int e = tid; e < 64 * kNB; e += kCholThreads) {
                    const int r = e >> 5, c = e & 31;
                    Pi[r * kPanelLd + c] = (ri + r < M) ? A[(size_t)(ri + r) * n + k0 + c] : 0.0;
                    Pj[r * kPanelLd + c] = (rj + r < n) ? A[(size_t)(rj + r) * n + k0 + c] : 0.0;
                }
                __syncthreads();
                for (int sub = wid; sub < 64; sub += kCholWarps) {
                    const int si = sub >> 3, sj = sub & 7;
                    if (I == J && sj > si) continue;
                    const int ar = si * 8 + (lane >> 2), bc = sj * 8 + (lane >> 2);
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int kk = 0; kk < kNB; kk += 4)
                        dmma_m8n8k4(c0, c1, Pi[ar * kPanelLd + kk + (lane & 3)], Pj[bc * kPanelLd + kk + (lane & 3)]);
                    const int row = ri + ar, col = rj + sj * 8 + (lane & 3) * 2;
                    if (row < M) {
                        double* dst = A + (size_t)row * n + col;
                        if (col < n && col <= row) dst[0] -= c0;
                        if (col + 1 < n && col + 1 <= row) dst[1] -= c1;
                    }
                }
            }
        }
        grid.sync();
    }
    // ---- backward substitution L^T x = y on CTA 0 (y = row n of A; x doubles as the work vector) ----
    if (cta == 0) {
        const bool ok = *g_ok != 0;
        double* part = stage;         // [16][kNB]
        double* rhs = scratch;        // [kNB]
        for (int i = tid; i < n; i += kCholThreads) x[i] = A[(size_t)n * n + i];
        __syncthreads();
        for (int k0 = n - kNB; k0 >= 0; k0 -= kNB) {
            {
                const int c = tid & 31, g = tid >> 5;
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                int i = k0 + kNB + g;
                for (; i + 3 * kCholWarps < n; i += 4 * kCholWarps) {
                    acc0 += A[(size_t)i * n + k0 + c] * x[i];
                    acc1 += A[(size_t)(i + kCholWarps) * n + k0 + c] * x[i + kCholWarps];
                    acc2 += A[(size_t)(i + 2 * kCholWarps) * n + k0 + c] * x[i + 2 * kCholWarps];
                    acc3 += A[(size_t)(i + 3 * kCholWarps) * n + k0 + c] * x[i + 3 * kCholWarps];
                }
                for (; i < n; i += kCholWarps) acc0 += A[(size_t)i * n + k0 + c] * x[i];
                part[g * kNB + c] = (acc0 + acc1) + (acc2 + acc3);
            }
            __syncthreads();
            if (tid < kNB) {
                double s4[4] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
                for (int g = 0; g < kCholWarps; g++) s4[g & 3] += part[g * kNB + tid];
                rhs[tid] = x[k0 + tid] - ((s4[0] + s4[1]) + (s4[2] + s4[3]));
            }
            __syncthreads();
            if (tid < kNB) {   // x_k = L_kk^-T rhs
                const double* Lg = Linv_g + (size_t)(k0 / kNB) * kNB * kNB;
                double s4[4] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll 8
                for (int c = 0; c < kNB; c++)
                    if (c >= tid) s4[c & 3] += Lg[c * kNB + tid] * rhs[c];
                x[k0 + tid] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            }
            __syncthreads();
        }
        if (!ok) for (int i = tid; i < n; i += kCholThreads) x[i] = 0.0;
        __syncthreads();
    }
}

} // namespace dvm
