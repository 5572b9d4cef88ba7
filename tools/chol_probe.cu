// chol_probe.cu -- where does a step of the 32x32 diagonal-block factorisation go?  (diagnostics, not product)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dvmslam_b200/csrc tools/chol_probe.cu -o tools/_build/chol_probe
#include "chol_device.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
using namespace dvm;

template <bool kInverse, int kNewton>
__device__ inline bool factor_variant(double* A, int ld, double* Ld, double* Li, double* scratch)
{
    const int r = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = w, c1 = w + 16;
    double* colbuf = scratch;
    double* s_rinv = scratch + 2 * kNB;
    double a0 = A[(size_t)r * ld + c0], a1 = A[(size_t)r * ld + c1];
    double x0 = (r == c0) ? 1.0 : 0.0, x1 = (r == c1) ? 1.0 : 0.0;
    auto publish = [&](int j) {
        const bool hi = j >= 16;
        const double djj = __shfl_sync(0xffffffffu, hi ? a1 : a0, j);
        double rinv;
        if (kNewton == 2) rinv = fast_rsqrt(djj);
        else {
            double y;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(djj));
            if (kNewton == 1) { const double e = fma(-0.5 * djj * y, y, 0.5); y = fma(y, e, y); }
            rinv = y;
        }
        if (r >= j) {
            const double l = (hi ? a1 : a0) * rinv;
            if (hi) a1 = l; else a0 = l;
            colbuf[(j & 1) * kNB + r] = l;
        }
        if (r == 0) s_rinv[j & 1] = rinv;
    };
    if (w == 0) publish(0);
    for (int j = 0; j < kNB; j++) {
        __syncthreads();
        const double* col = colbuf + (j & 1) * kNB;
        const double rinv = s_rinv[j & 1];
        const double lr = (r >= j) ? col[r] : 0.0;
        if (c0 > j && r >= c0) a0 -= lr * col[c0];
        if (c1 > j && r >= c1) a1 -= lr * col[c1];
        if (j + 1 < kNB && w == ((j + 1) & 15)) publish(j + 1);
        if (kInverse) {
            if (c0 <= j) {
                const double xj = __shfl_sync(0xffffffffu, x0, j) * rinv;
                if (r == j) x0 = xj; else if (r > j) x0 -= lr * xj;
            }
            if (c1 <= j) {
                const double xj = __shfl_sync(0xffffffffu, x1, j) * rinv;
                if (r == j) x1 = xj; else if (r > j) x1 -= lr * xj;
            }
        }
    }
    Ld[r * kDiagLd + c0] = (c0 <= r) ? a0 : 0.0;
    Ld[r * kDiagLd + c1] = (c1 <= r) ? a1 : 0.0;
    Li[r * kDiagLd + c0] = (c0 <= r) ? x0 : 0.0;
    Li[r * kDiagLd + c1] = (c1 <= r) ? x1 : 0.0;
    __syncthreads();
    return false;
}


// unrolled over the column index; kDiag: the pivot of an owned column is carried by every lane (no pivot shuffle)
template <bool kInverse, bool kDiag>
__device__ inline bool factor_unrolled(double* A, int ld, double* Ld, double* Li, double* scratch)
{
    const int r = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = w, c1 = w + 16;
    double* colbuf = scratch;
    double* s_rinv = scratch + 2 * kNB;
    double a0 = A[(size_t)r * ld + c0], a1 = A[(size_t)r * ld + c1];
    double d0 = A[(size_t)c0 * ld + c0], d1 = A[(size_t)c1 * ld + c1];
    double x0 = (r == c0) ? 1.0 : 0.0, x1 = (r == c1) ? 1.0 : 0.0;
#define PUBLISH(j)                                                                              \
    {                                                                                           \
        const bool hi = (j) >= 16;                                                              \
        const double djj = kDiag ? (hi ? d1 : d0) : __shfl_sync(0xffffffffu, hi ? a1 : a0, (j)); \
        const double rinv = fast_rsqrt(djj);                                                    \
        const double l = (hi ? a1 : a0) * rinv;                                                 \
        if (hi) a1 = l; else a0 = l;                                                            \
        if (r >= (j)) colbuf[((j) & 1) * kNB + r] = l;                                          \
        if (r == 0) s_rinv[(j) & 1] = rinv;                                                     \
    }
    if (w == 0) PUBLISH(0);
#pragma unroll
    for (int j = 0; j < kNB; j++) {
        __syncthreads();
        const double* col = colbuf + (j & 1) * kNB;
        const double lr = (r >= j) ? col[r] : 0.0;
        const double l0 = col[c0], l1 = col[c1];
        if (c0 > j && r >= c0) a0 -= lr * l0;
        if (c1 > j && r >= c1) a1 -= lr * l1;
        if (kDiag) { if (c0 > j) d0 -= l0 * l0; if (c1 > j) d1 -= l1 * l1; }
        if (j + 1 < kNB && w == ((j + 1) & 15)) {
            PUBLISH(j + 1)
        }
        if (kInverse) {
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
    }
#undef PUBLISH
    Ld[r * kDiagLd + c0] = (c0 <= r) ? a0 : 0.0;
    Ld[r * kDiagLd + c1] = (c1 <= r) ? a1 : 0.0;
    Li[r * kDiagLd + c0] = (c0 <= r) ? x0 : 0.0;
    Li[r * kDiagLd + c1] = (c1 <= r) ? x1 : 0.0;
    __syncthreads();
    return false;
}


// four warps (one per SM sub-partition), lane = row, warp w owns columns w, w + 4, ..., w + 28; the other warps of the CTA wait
template <bool kInverse>
__device__ inline bool factor_w4(double* A, int ld, double* Ld, double* Li, double* scratch)
{
    const int r = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w < 4) {
        double* colbuf = scratch;
        double* s_rinv = scratch + 2 * kNB;
        double a[8], x[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { a[i] = A[(size_t)r * ld + w + 4 * i]; x[i] = (r == w + 4 * i) ? 1.0 : 0.0; }
        auto publish = [&](int j) {   // j is a compile-time constant after unrolling
            const int slot = j >> 2;
            const double djj = __shfl_sync(0xffffffffu, a[slot], j);
            const double rinv = fast_rsqrt(djj);
            const double l = a[slot] * rinv;
            a[slot] = l;
            if (r >= j) colbuf[(j & 1) * kNB + r] = l;
            if (r == 0) s_rinv[j & 1] = rinv;
        };
        if (w == 0) publish(0);
#pragma unroll
        for (int j = 0; j < kNB; j++) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const double* col = colbuf + (j & 1) * kNB;
            const double lr = (r >= j) ? col[r] : 0.0;
            // the column published next goes first
            if (j + 1 < kNB) {
                const int slot = (j + 1) >> 2;
                const int c = w + 4 * slot;
                if (c > j && r >= c) a[slot] -= lr * col[c];
                if (w == ((j + 1) & 3)) publish(j + 1);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int c = w + 4 * i;
                if (i != ((j + 1) >> 2) && 4 * i + 3 > j) { if (c > j && r >= c) a[i] -= lr * col[c]; }
            }
            if (kInverse) {
                const double rinv = s_rinv[j & 1];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int c = w + 4 * i;
                    if (4 * i <= j) {   // some warp's column of this slot is <= j
                        if (c <= j) {
                            const double xj = __shfl_sync(0xffffffffu, x[i], j) * rinv;
                            if (r == j) x[i] = xj; else if (r > j) x[i] -= lr * xj;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int c = w + 4 * i;
            Ld[r * kDiagLd + c] = (c <= r) ? a[i] : 0.0;
            Li[r * kDiagLd + c] = (c <= r) ? x[i] : 0.0;
        }
    }
    __syncthreads();
    return false;
}


// the publisher's chain with time stamps (no inverse): seg[0] barrier -> loads + update done, [1] -> pivot by shuffle,
// [2] -> rsqrt, [3] -> scaled + stored, [4] -> next barrier passed
__device__ inline long long stamp_after(double v)
{
    long long t;
    if (__double2hiint(v) == 0x7ff12345) asm volatile("trap;");
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}
__device__ inline void factor_stamped(double* A, int ld, double* scratch, long long* seg_out)
{
    const int r = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = w, c1 = w + 16;
    double* colbuf = scratch;
    double* s_rinv = scratch + 2 * kNB;
    double a0 = A[(size_t)r * ld + c0], a1 = A[(size_t)r * ld + c1];
    long long seg[5] = { 0, 0, 0, 0, 0 };
    long long tprev = 0;
    bool was_pub = false;
    if (w == 0) {
        const double djj = __shfl_sync(0xffffffffu, a0, 0);
        const double rinv = fast_rsqrt(djj);
        a0 *= rinv;
        colbuf[r] = a0;
        if (r == 0) s_rinv[0] = rinv;
    }
    for (int j = 0; j < kNB; j++) {
        __syncthreads();
        long long tb;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(tb)::"memory");
        if (was_pub) seg[4] += tb - tprev;
        was_pub = false;
        const double* col = colbuf + (j & 1) * kNB;
        const double lr = (r >= j) ? col[r] : 0.0;
        if (c0 > j && r >= c0) a0 -= lr * col[c0];
        if (c1 > j && r >= c1) a1 -= lr * col[c1];
        if (j + 1 < kNB && w == ((j + 1) & 15)) {
            const int jj = j + 1;
            const bool hi = jj >= 16;
            const long long t1 = stamp_after(hi ? a1 : a0);
            const double djj = __shfl_sync(0xffffffffu, hi ? a1 : a0, jj);
            const long long t2 = stamp_after(djj);
            const double rinv = fast_rsqrt(djj);
            const long long t3 = stamp_after(rinv);
            if (r >= jj) {
                const double l = (hi ? a1 : a0) * rinv;
                if (hi) a1 = l; else a0 = l;
                colbuf[(jj & 1) * kNB + r] = l;
            }
            if (r == 0) s_rinv[jj & 1] = rinv;
            long long t4;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t4)::"memory");
            seg[0] += t1 - tb; seg[1] += t2 - t1; seg[2] += t3 - t2; seg[3] += t4 - t3;
            tprev = t4;
            was_pub = true;
        }
    }
    __syncthreads();
    if (r == 0)
        for (int i = 0; i < 5; i++) atomicAdd((unsigned long long*)&seg_out[i], (unsigned long long)seg[i]);
    if (a0 + a1 == 123.456) scratch[0] = a0;
}


__device__ inline double rsqrt_halley(double d)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double q = y * e;
    return fma(q, p, y);
}
__device__ inline void named_arrive(int id) { asm volatile("bar.arrive %0, 512;" ::"r"(id) : "memory"); }
__device__ inline void named_sync(int id) { asm volatile("bar.sync %0, 512;" ::"r"(id) : "memory"); }

// v2: third-order rsqrt (4 dependent FP64 operations instead of 6), validity check off the chain, the publishing warp only
// ARRIVES at the step barrier (its own inverse updates run after it released the others), next column's update first
template <bool kInverse>
__device__ inline bool factor_v2(double* A, int ld, double* Ld, double* Li, double* scratch)
{
    const int r = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = w, c1 = w + 16;
    double* colbuf = scratch;
    double* s_rinv = scratch + 2 * kNB;
    double* s_bad = s_rinv + 2;
    double a0 = A[(size_t)r * ld + c0], a1 = A[(size_t)r * ld + c1];
    double x0 = (r == c0) ? 1.0 : 0.0, x1 = (r == c1) ? 1.0 : 0.0;
    if (threadIdx.x == 0) *s_bad = 0.0;
    __syncthreads();
    auto publish = [&](int j) {
        const bool hi = j >= 16;
        const double mine = hi ? a1 : a0;
        const double djj = __shfl_sync(0xffffffffu, mine, j);
        const double rinv = rsqrt_halley(djj);
        const double l = mine * rinv;
        if (r >= j) {
            if (hi) a1 = l; else a0 = l;
            colbuf[(j & 1) * kNB + r] = l;
        }
        if (r == 0) { s_rinv[j & 1] = rinv; if (!(djj > 0) || !isfinite(djj)) *s_bad = 1.0; }
        __syncwarp();
        named_arrive(1 + (j & 1));
    };
    if (w == 0) publish(0);
    for (int j = 0; j < kNB; j++) {
        if (w != (j & 15)) named_sync(1 + (j & 1));
        const double* col = colbuf + (j & 1) * kNB;
        const double rinv = s_rinv[j & 1];
        const double lr = (r >= j) ? col[r] : 0.0;
        const double l0 = col[c0], l1 = col[c1];
        const int nx = j + 1;
        if (nx < kNB && w == (nx & 15)) {
            if (nx >= 16) { if (r >= c1) a1 -= lr * l1; }
            else { if (r >= c0) a0 -= lr * l0; }
            publish(nx);
            if (nx < 16 && r >= c1) a1 -= lr * l1;   // (nx >= 16: column c0 = nx - 16 <= j is finished)
        } else {
            if (c0 > j && r >= c0) a0 -= lr * l0;
            if (c1 > j && r >= c1) a1 -= lr * l1;
        }
        if (kInverse) {
            if (c0 <= j) {
                const double xj = __shfl_sync(0xffffffffu, x0, j) * rinv;
                if (r == j) x0 = xj; else if (r > j) x0 -= lr * xj;
            }
            if (c1 <= j) {
                const double xj = __shfl_sync(0xffffffffu, x1, j) * rinv;
                if (r == j) x1 = xj; else if (r > j) x1 -= lr * xj;
            }
        }
    }
    Ld[r * kDiagLd + c0] = (c0 <= r) ? a0 : 0.0;
    Ld[r * kDiagLd + c1] = (c1 <= r) ? a1 : 0.0;
    Li[r * kDiagLd + c0] = (c0 <= r) ? x0 : 0.0;
    Li[r * kDiagLd + c1] = (c1 <= r) ? x1 : 0.0;
    __syncthreads();
    return *s_bad != 0.0;
}


// split roles: warps 0-3 (one per SM sub-partition) factor -- lane = row, warp w owns columns w, w + 4, ... --, warps 4-11
// carry the inverse (warp v = w - 4 owns columns v, v + 8, v + 16, v + 24), warps 12-15 only keep the barrier company
template <bool kHalley>
__device__ inline bool factor_split(double* A, int ld, double* Ld, double* Li, double* scratch)
{
    const int r = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* colbuf = scratch;
    double* s_rinv = scratch + 2 * kNB;
    double* s_bad = s_rinv + 2;
    double a[8], x[4];
    if (w < 4) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = A[(size_t)r * ld + w + 4 * i];
    } else if (w < 12) {
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = (r == w - 4 + 8 * i) ? 1.0 : 0.0;
    }
    if (threadIdx.x == 0) *s_bad = 0.0;
    auto publish = [&](int j) {
        const int slot = j >> 2;
        const double djj = __shfl_sync(0xffffffffu, a[slot], j);
        const double rinv = kHalley ? rsqrt_halley(djj) : fast_rsqrt(djj);
        const double l = a[slot] * rinv;
        if (r >= j) { a[slot] = l; colbuf[(j & 1) * kNB + r] = l; }
        if (r == 0) { s_rinv[j & 1] = rinv; if (!(djj > 0) || !isfinite(djj)) *s_bad = 1.0; }
    };
    __syncthreads();
    if (w == 0) publish(0);
#pragma unroll
    for (int j = 0; j < kNB; j++) {
        __syncthreads();
        const double* col = colbuf + (j & 1) * kNB;
        if (w < 4) {
            const double lr = (r >= j) ? col[r] : 0.0;
            if (j + 1 < kNB) {
                const int slot = (j + 1) >> 2;
                const int c = w + 4 * slot;
                if (c > j && r >= c) a[slot] -= lr * col[c];
                if (w == ((j + 1) & 3)) publish(j + 1);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int c = w + 4 * i;
                if (i != ((j + 1) >> 2) && 4 * i + 3 > j) { if (c > j && r >= c) a[i] -= lr * col[c]; }
            }
        } else if (w < 12) {
            const double lr = (r >= j) ? col[r] : 0.0;
            const double rinv = s_rinv[j & 1];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = w - 4 + 8 * i;
                if (8 * i <= j) {
                    if (c <= j) {
                        const double xj = __shfl_sync(0xffffffffu, x[i], j) * rinv;
                        if (r == j) x[i] = xj; else if (r > j) x[i] -= lr * xj;
                    }
                }
            }
        }
    }
    if (w < 4) {
#pragma unroll
        for (int i = 0; i < 8; i++) { const int c = w + 4 * i; Ld[r * kDiagLd + c] = (c <= r) ? a[i] : 0.0; }
    } else if (w < 12) {
#pragma unroll
        for (int i = 0; i < 4; i++) { const int c = w - 4 + 8 * i; Li[r * kDiagLd + c] = (c <= r) ? x[i] : 0.0; }
    }
    __syncthreads();
    return *s_bad != 0.0;
}

// 32 steps of barrier + shared-memory round trip only
__device__ inline void barrier_floor(double* scratch)
{
    const int r = threadIdx.x & 31, w = threadIdx.x >> 5;
    double v = r;
    for (int j = 0; j < kNB; j++) {
        __syncthreads();
        v += scratch[(j & 1) * kNB + r];
        if (w == ((j + 1) & 15)) scratch[((j + 1) & 1) * kNB + r] = v;
    }
    if (v == 123.456) scratch[0] = v;
    __syncthreads();
}

// one warp, 8x8 block: lane = (row r = lane >> 2, column pair q = lane & 3 -> columns 2q, 2q + 1)
__device__ inline void warp_factor_8(double& e0, double& e1, int lane)
{
    const int r = lane >> 2, q = lane & 3;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int src = j * 4 + (j >> 1);
        const double pj = __shfl_sync(0xffffffffu, (j & 1) ? e1 : e0, src);
        const double rinv = fast_rsqrt(pj);
        // column j scaled (held by lanes with q == j >> 1)
        double mine = ((j & 1) ? e1 : e0) * rinv;
        if (q == (j >> 1) && r >= j) { if (j & 1) e1 = mine; else e0 = mine; }
        const double lr = __shfl_sync(0xffffffffu, mine, r * 4 + (j >> 1));
        const double lc0 = __shfl_sync(0xffffffffu, mine, (2 * q) * 4 + (j >> 1));
        const double lc1 = __shfl_sync(0xffffffffu, mine, (2 * q + 1) * 4 + (j >> 1));
        if (2 * q > j && r >= 2 * q) e0 -= lr * lc0;
        if (2 * q + 1 > j && r >= 2 * q + 1) e1 -= lr * lc1;
    }
}

__global__ void __launch_bounds__(512, 1) probe(const double* A0, double* out, long long* cyc, int reps)
{
    __shared__ double As[kNB * kDiagLd], Ld[kNB * kDiagLd], Li[kNB * kDiagLd], scratch[4 * kNB + 4];
    const int tid = threadIdx.x;
    auto reload = [&]() {
        for (int t = tid; t < kNB * kNB; t += 512) As[(t >> 5) * kDiagLd + (t & 31)] = A0[t];
        __syncthreads();
    };
    long long t0, t1;
    // reload cost
    t0 = clock64();
    for (int i = 0; i < reps; i++) reload();
    t1 = clock64();
    if (tid == 0) cyc[0] = (t1 - t0) / reps;
#define RUN(slot, call)                                   \
    t0 = clock64();                                       \
    for (int i = 0; i < reps; i++) { reload(); call; }    \
    t1 = clock64();                                       \
    if (tid == 0) cyc[slot] = (t1 - t0) / reps;
    RUN(1, cta_factor_invert_32(As, kDiagLd, Ld, Li, false, nullptr, scratch));
    for (int t = tid; t < kNB * kNB; t += 512) { out[t] = Ld[(t >> 5) * kDiagLd + (t & 31)]; out[1024 + t] = Li[(t >> 5) * kDiagLd + (t & 31)]; }
    RUN(2, (factor_variant<true, 2>(As, kDiagLd, Ld, Li, scratch)));
    RUN(3, (factor_variant<false, 2>(As, kDiagLd, Ld, Li, scratch)));
    RUN(4, (factor_variant<true, 1>(As, kDiagLd, Ld, Li, scratch)));
    RUN(5, (factor_variant<true, 0>(As, kDiagLd, Ld, Li, scratch)));
    RUN(6, barrier_floor(scratch));
    RUN(13, (factor_unrolled<true, false>(As, kDiagLd, Ld, Li, scratch)));
    RUN(14, (factor_unrolled<true, true>(As, kDiagLd, Ld, Li, scratch)));
    for (int t = tid; t < kNB * kNB; t += 512) { out[t] = Ld[(t >> 5) * kDiagLd + (t & 31)]; out[1024 + t] = Li[(t >> 5) * kDiagLd + (t & 31)]; }
    RUN(15, (factor_unrolled<false, true>(As, kDiagLd, Ld, Li, scratch)));
    RUN(16, (factor_w4<true>(As, kDiagLd, Ld, Li, scratch)));
    for (int t = tid; t < kNB * kNB; t += 512) { out[t] = Ld[(t >> 5) * kDiagLd + (t & 31)]; out[1024 + t] = Li[(t >> 5) * kDiagLd + (t & 31)]; }
    RUN(17, (factor_w4<false>(As, kDiagLd, Ld, Li, scratch)));
    RUN(18, (factor_v2<true>(As, kDiagLd, Ld, Li, scratch)));
    for (int t = tid; t < kNB * kNB; t += 512) { out[t] = Ld[(t >> 5) * kDiagLd + (t & 31)]; out[1024 + t] = Li[(t >> 5) * kDiagLd + (t & 31)]; }
    RUN(19, (factor_v2<false>(As, kDiagLd, Ld, Li, scratch)));
    RUN(20, (factor_split<false>(As, kDiagLd, Ld, Li, scratch)));
    RUN(21, (factor_split<true>(As, kDiagLd, Ld, Li, scratch)));
    for (int t = tid; t < kNB * kNB; t += 512) { out[t] = Ld[(t >> 5) * kDiagLd + (t & 31)]; out[1024 + t] = Li[(t >> 5) * kDiagLd + (t & 31)]; }
    {   // 8x8 in one warp (other warps idle at the barrier)
        t0 = clock64();
        double acc = 0;
        for (int i = 0; i < reps; i++) {
            if (tid < 32) {
                const int r = tid >> 2, q = tid & 3;
                double e0 = A0[r * 32 + 2 * q] + (r == 2 * q ? 8.0 : 0.0), e1 = A0[r * 32 + 2 * q + 1] + (r == 2 * q + 1 ? 8.0 : 0.0);
                warp_factor_8(e0, e1, tid);
                acc += e0 + e1;
            }
            __syncthreads();
        }
        t1 = clock64();
        if (tid == 0) cyc[7] = (t1 - t0) / reps;
        if (acc == 123.456) out[0] = acc;
    }
    {   // chain of dependent DFMA / shuffles / rsqrt in one warp, to read the latencies
        double v = A0[tid & 31] + 2.0;
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < 256; i++) v = fma(v, 1.0000001, 1e-9);
        t1 = clock64();
        if (tid == 0) cyc[8] = (t1 - t0) / 256;
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < 256; i++) v = __shfl_sync(0xffffffffu, v, (tid + 1) & 31);
        t1 = clock64();
        if (tid == 0) cyc[9] = (t1 - t0) / 256;
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < 256; i++) v = fast_rsqrt(v) + 1.5;
        t1 = clock64();
        if (tid == 0) cyc[10] = (t1 - t0) / 256;
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < 256; i++) { __syncthreads(); }
        t1 = clock64();
        if (tid == 0) cyc[11] = (t1 - t0) / 256;
        t0 = clock64();
#pragma unroll 1
        for (int i = 0; i < 256; i++) { scratch[tid & 63] = v; __syncthreads(); v += scratch[(tid + 1) & 63]; }
        t1 = clock64();
        if (tid == 0) cyc[12] = (t1 - t0) / 256;
        if (v == 123.456) out[1] = v;
        // FP64 issue rate: 8 independent accumulators per thread, 64 rounds = 512 DFMA per warp
        for (int nw = 1; nw <= 16; nw *= 2) {
            double q[8];
#pragma unroll
            for (int k = 0; k < 8; k++) q[k] = v + k;
            __syncthreads();
            t0 = clock64();
            if ((tid >> 5) < nw) {
#pragma unroll 1
                for (int i = 0; i < 64; i++) {
#pragma unroll
                    for (int k = 0; k < 8; k++) q[k] = fma(q[k], 1.0000001, 1e-9);
                }
            }
            __syncthreads();
            t1 = clock64();
            double sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) sum += q[k];
            if (sum == 123.456) out[2] = sum;
            int slot = 25 + (nw == 1 ? 0 : nw == 2 ? 1 : nw == 4 ? 2 : nw == 8 ? 3 : 4);
            if (tid == 0) cyc[slot] = t1 - t0;
        }
    }
}

int main()
{
    std::vector<double> B(1024), A(1024, 0.0);
    srand(1);
    for (auto& v : B) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 32; j++) {
            double s = 0;
            for (int k = 0; k < 32; k++) s += B[i * 32 + k] * B[j * 32 + k];
            A[i * 32 + j] = s + (i == j ? 4.0 : 0.0);
        }
    double *dA, *dout; long long* dc;
    cudaMalloc(&dA, 8192); cudaMalloc(&dout, 16384); cudaMalloc(&dc, 32 * 8);
    cudaMemcpy(dA, A.data(), 8192, cudaMemcpyHostToDevice);
    cudaMemset(dc, 0, 256);
    probe<<<1, 512>>>(dA, dout, dc, 200);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[32];
    cudaMemcpy(c, dc, 256, cudaMemcpyDeviceToHost);
    std::vector<double> out(2048);
    cudaMemcpy(out.data(), dout, 16384, cudaMemcpyDeviceToHost);
    // check L L^T = A
    double err = 0;
    for (int i = 0; i < 32; i++)
        for (int j = 0; j <= i; j++) {
            double s = 0;
            for (int k = 0; k <= j; k++) s += out[i * 32 + k] * out[j * 32 + k];
            err = fmax(err, fabs(s - A[i * 32 + j]));
        }
    double erri = 0;
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 32; j++) {
            double s2 = 0;
            for (int k = 0; k < 32; k++) s2 += out[i * 32 + k] * out[1024 + k * 32 + j];
            erri = fmax(erri, fabs(s2 - (i == j ? 1.0 : 0.0)));
        }
    printf("status %s  |LL^T - A| %.3g  |L Linv - I| %.3g\n", cudaGetErrorString(e), err, erri);
    const char* names[] = { "reload", "current factor+invert", "copy of it", "without inverse", "1 Newton step", "0 Newton steps", "barrier floor (32 steps)", "8x8 in one warp",
                            "DFMA latency", "64-bit shuffle latency", "fast_rsqrt + add latency", "__syncthreads (512 thr)", "STS + barrier + LDS + add",
                            "unrolled", "unrolled + diag everywhere", "unrolled + diag, no inverse", "4 warps", "4 warps, no inverse", "v2", "v2, no inverse", "split roles", "split roles + 3rd-order rsqrt" };
    for (int i = 0; i < 22; i++) printf("%-28s %lld cycles%s\n", names[i], c[i], (i >= 1 && i <= 6) ? " (incl. reload)" : "");
    printf("publisher chain, cycles per step: barrier->update %lld, pivot shuffle %lld, rsqrt %lld, scale+store %lld, to next barrier %lld\n",
           c[20] / 3100, c[21] / 3100, c[22] / 3100, c[23] / 3100, c[24] / 3100);
    printf("512 independent DFMA per warp, cycles with 1/2/4/8/16 warps: %lld %lld %lld %lld %lld\n", c[25], c[26], c[27], c[28], c[29]);
    return 0;
}
