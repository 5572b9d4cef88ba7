// pose_opt.cu -- Optimizer::PoseOptimization on one 8-CTA cluster (O3/src/Optimizer.cc:744-1028).
// Float64 work inside the stated tolerance: this file is compiled with FMA contraction enabled.
#include "track_kernels.cuh"
#include <cooperative_groups.h>
#include <cstdlib>

namespace dvm {

// ----------------------------------------------------------------------------- PoseOptimization

struct Quat { double x, y, z, w; };
struct SE3d { Quat r; double t[3]; };

// Branch-free reciprocal / reciprocal square root: hardware seed (MUFU.RCP64H / RSQ64H, ~20 bits) and two
// Newton steps.  For the normal, well-scaled operands of this file (depths, Hessian pivots, squared
// norms) the result is within 1 ulp of the IEEE quotient, far inside the stated pose tolerance, and
// the dependent chain is about half of the compiler's division with its special-case slow path.
__device__ inline double fast_rcp(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}
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

__device__ inline void quat_normalize(Quat& q)
{
    double s = fast_rsqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    if (q.w < 0) s = -s;
    q.x *= s; q.y *= s; q.z *= s; q.w *= s;
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
// SE3Quat::exp (g2o/types/se3quat.h:212-240): R = I + A*O + B*O^2, V = I + B'*O + C*O^2 with O = [w]x;
// below theta = 1e-5 the reference takes A = B = B' = C = 1.  Its Quaternion(R) constructor goes through
// the trace branch whenever trace(R) > 0 (rotation steps below 120 degrees):
//   trace = 3 - 2*B*theta^2,  R[2][1] - R[1][2] = 2*A*w0, ...   (O^2 is symmetric)
// so q = (A*w / tt, tt / 2) with tt = sqrt(trace + 1), then normalised -- evaluated here without forming
// the matrices.  A step of 120 degrees or more cannot come out of a damped LM solve on a tracked frame;
// it would only make this kernel's result differ from the reference, never fail.
__device__ inline SE3d se3_exp(const double u[6])
{
    const double w0 = u[0], w1 = u[1], w2 = u[2];
    const double th2 = w0 * w0 + w1 * w1 + w2 * w2;
    double A = 1.0, B = 1.0, Bv = 1.0, C = 1.0;
    if (!(th2 < 1e-10) && th2 < 0.25) {
        // 1e-5 <= theta < 0.5 rad, every LM step on a tracked frame: even power series of
        // sin(t)/t, (1-cos t)/t^2, (t-sin t)/t^3 in t^2 (truncation below 1e-19) -- no sqrt, no sincos
        const double x = th2;
        A = 1.0 + x * (-1.0 / 6 + x * (1.0 / 120 + x * (-1.0 / 5040 + x * (1.0 / 362880 + x * (-1.0 / 39916800 + x * (1.0 / 6227020800.0 + x * (-1.0 / 1307674368000.0)))))));
        B = 0.5 + x * (-1.0 / 24 + x * (1.0 / 720 + x * (-1.0 / 40320 + x * (1.0 / 3628800 + x * (-1.0 / 479001600 + x * (1.0 / 87178291200.0 + x * (-1.0 / 20922789888000.0)))))));
        C = 1.0 / 6 + x * (-1.0 / 120 + x * (1.0 / 5040 + x * (-1.0 / 362880 + x * (1.0 / 39916800 + x * (-1.0 / 6227020800.0 + x * (1.0 / 1307674368000.0 + x * (-1.0 / 355687428096000.0)))))));
        Bv = B;
    } else if (!(th2 < 1e-10)) {
        const double inv = fast_rsqrt(th2), theta = th2 * inv;
        double sn, cs;
        sincos(theta, &sn, &cs);
        A = sn * inv;
        B = (1.0 - cs) * inv * inv;
        Bv = B;
        C = (theta - sn) * inv * inv * inv;
    }
    SE3d T;
    const double tt2 = 4.0 - 2.0 * B * th2;     // trace + 1
    const double itt = fast_rsqrt(tt2);
    T.r.w = 0.5 * tt2 * itt;
    const double sx = A * itt;
    T.r.x = sx * w0; T.r.y = sx * w1; T.r.z = sx * w2;
    quat_normalize(T.r);
    // t = V * v = v + B'*(w x v) + C*(w x (w x v))
    const double v0 = u[3], v1 = u[4], v2 = u[5];
    const double c0 = w1 * v2 - w2 * v1, c1 = w2 * v0 - w0 * v2, c2 = w0 * v1 - w1 * v0;
    const double d0 = w1 * c2 - w2 * c1, d1 = w2 * c0 - w0 * c2, d2 = w0 * c1 - w1 * c0;
    T.t[0] = v0 + Bv * c0 + C * d0;
    T.t[1] = v1 + Bv * c1 + C * d1;
    T.t[2] = v2 + Bv * c2 + C * d2;
    return T;
}
__device__ inline SE3d se3_mul(const SE3d& a, const SE3d& b)
{
    SE3d r = a;
    double rt[3];
    quat_rotate(a.r, b.t, rt);
    r.t[0] += rt[0]; r.t[1] += rt[1]; r.t[2] += rt[2];
    r.r = quat_mul(a.r, b.r);
    quat_normalize(r.r);
    return r;
}

// VertexSE3Expmap::oplusImpl: exp(u) * T (g2o/types/types_six_dof_expmap.h:71-74) as ONE dependent chain instead of the
// three the reference's call sequence has (quaternion of exp, its normalisation, product, normalisation).  With
// exp(u) = (R_e, V v), R_e = I + A*O + B*O^2 and its quaternion q_e = (A*w, 2 - B*theta^2) / sqrt(trace + 1):
//   rotation     normalize(q_e (x) T.r): the scale of q_e drops out of the normalisation, so its own square root is skipped;
//   translation  R_e * T.t + V v with R_e p = p + A*(w x p) + B*(w x (w x p)), independent of the quaternion chain.
// Same value as se3_mul(se3_exp(u), T) up to rounding (1e-16), about 20 dependent FP64 operations shorter.
__device__ inline SE3d se3_left_update(const double u[6], const SE3d& T)
{
    const double w0 = u[0], w1 = u[1], w2 = u[2];
    const double th2 = w0 * w0 + w1 * w1 + w2 * w2;
    double A = 1.0, B = 1.0, Bv = 1.0, C = 1.0;
    if (!(th2 < 1e-10) && th2 < 0.25) {
        const double x = th2;
        A = 1.0 + x * (-1.0 / 6 + x * (1.0 / 120 + x * (-1.0 / 5040 + x * (1.0 / 362880 + x * (-1.0 / 39916800 + x * (1.0 / 6227020800.0 + x * (-1.0 / 1307674368000.0)))))));
        B = 0.5 + x * (-1.0 / 24 + x * (1.0 / 720 + x * (-1.0 / 40320 + x * (1.0 / 3628800 + x * (-1.0 / 479001600 + x * (1.0 / 87178291200.0 + x * (-1.0 / 20922789888000.0)))))));
        C = 1.0 / 6 + x * (-1.0 / 120 + x * (1.0 / 5040 + x * (-1.0 / 362880 + x * (1.0 / 39916800 + x * (-1.0 / 6227020800.0 + x * (1.0 / 1307674368000.0 + x * (-1.0 / 355687428096000.0)))))));
        Bv = B;
    } else if (!(th2 < 1e-10)) {
        const double inv = fast_rsqrt(th2), theta = th2 * inv;
        double sn, cs;
        sincos(theta, &sn, &cs);
        A = sn * inv;
        B = (1.0 - cs) * inv * inv;
        Bv = B;
        C = (theta - sn) * inv * inv * inv;
    }
    SE3d R;
    // below theta = 1e-5 the reference takes A = B = 1 in R_e (se3quat.h:222-226): q_e = (w, 2 - theta^2) / sqrt(trace + 1) there too
    Quat qe;
    qe.x = A * w0; qe.y = A * w1; qe.z = A * w2; qe.w = 2.0 - B * th2;
    R.r = quat_mul(qe, T.r);
    quat_normalize(R.r);
    const double p0 = T.t[0], p1 = T.t[1], p2 = T.t[2];
    const double a0 = w1 * p2 - w2 * p1, a1 = w2 * p0 - w0 * p2, a2 = w0 * p1 - w1 * p0;        // w x p
    const double b0 = w1 * a2 - w2 * a1, b1 = w2 * a0 - w0 * a2, b2 = w0 * a1 - w1 * a0;        // w x (w x p)
    const double v0 = u[3], v1 = u[4], v2 = u[5];
    const double c0 = w1 * v2 - w2 * v1, c1 = w2 * v0 - w0 * v2, c2 = w0 * v1 - w1 * v0;        // w x v
    const double d0 = w1 * c2 - w2 * c1, d1 = w2 * c0 - w0 * c2, d2 = w0 * c1 - w1 * c0;        // w x (w x v)
    // the point is moved by the (normalised) QUATERNION of exp(u) (SE3Quat::operator*, se3quat.h:105-111).  For theta >= 1e-5
    // that is R_e itself; below it R_e = I + O + O^2 is not a rotation and its normalised quaternion
    // (w, 2 - theta^2) / n turns by p + (2 s / n^2) (w x p) + (2 / n^2) (w x (w x p)) with s = 2 - theta^2, n^2 = theta^2 + s^2
    double Ar = A, Br = B;
    if (th2 < 1e-10) {
        const double sq = 2.0 - th2, in2 = fast_rcp(th2 + sq * sq);
        Ar = 2.0 * sq * in2; Br = 2.0 * in2;
    }
    R.t[0] = (p0 + Ar * a0 + Br * b0) + (v0 + Bv * c0 + C * d0);
    R.t[1] = (p1 + Ar * a1 + Br * b1) + (v1 + Bv * c1 + C * d1);
    R.t[2] = (p2 + Ar * a2 + Br * b2) + (v2 + Bv * c2 + C * d2);
    return R;
}

// Symmetric positive-definite 6x6 solve by 3x3 blocks (rotation block A11, translation block A22):
//   S = A22 - A12^T A11^-1 A12,  x2 = S^-1 (b2 - A12^T A11^-1 b1),  x1 = A11^-1 b1 - A11^-1 A12 x2
// with closed-form (adjugate) 3x3 inverses.  Same solution as the LDL^T the reference's dense solver
// computes (g2o/solvers/linear_solver_dense.h:57-104) up to rounding, but the dependent chain is a third
// as long, which is what bounds one LM trial here.  Returns false unless the matrix is positive definite
// (all six leading principal minors positive <=> A11 and its Schur complement positive definite).
struct Sym3 { double m00, m01, m02, m11, m12, m22; };
__device__ inline bool sym3_inverse(const Sym3& m, Sym3& inv)
{
    const double c00 = m.m11 * m.m22 - m.m12 * m.m12, c01 = m.m02 * m.m12 - m.m01 * m.m22, c02 = m.m01 * m.m12 - m.m02 * m.m11;
    const double c11 = m.m00 * m.m22 - m.m02 * m.m02, c12 = m.m01 * m.m02 - m.m00 * m.m12, c22 = m.m00 * m.m11 - m.m01 * m.m01;
    const double det = m.m00 * c00 + m.m01 * c01 + m.m02 * c02;
    const double id = fast_rcp(det);
    inv.m00 = c00 * id; inv.m01 = c01 * id; inv.m02 = c02 * id; inv.m11 = c11 * id; inv.m12 = c12 * id; inv.m22 = c22 * id;
    return m.m00 > 0 && c22 > 0 && det > 0;
}
__device__ inline bool spd6_solve(const double* A, const double* b, double* x)
{
    const Sym3 A11 = { A[0], A[1], A[2], A[7], A[8], A[14] };
    Sym3 I1;
    const bool ok1 = sym3_inverse(A11, I1);
    const double i1[9] = { I1.m00, I1.m01, I1.m02, I1.m01, I1.m11, I1.m12, I1.m02, I1.m12, I1.m22 };
    double W[9], u[3]; // W = A11^-1 A12, u = A11^-1 b1
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 3; c++) W[r * 3 + c] = i1[r * 3] * A[0 * 6 + 3 + c] + i1[r * 3 + 1] * A[1 * 6 + 3 + c] + i1[r * 3 + 2] * A[2 * 6 + 3 + c];
        u[r] = i1[r * 3] * b[0] + i1[r * 3 + 1] * b[1] + i1[r * 3 + 2] * b[2];
    }
    auto a12 = [&](int r, int c) { return A[r * 6 + 3 + c]; };
    Sym3 S;
    S.m00 = A[21] - (a12(0, 0) * W[0] + a12(1, 0) * W[3] + a12(2, 0) * W[6]);
    S.m01 = A[22] - (a12(0, 0) * W[1] + a12(1, 0) * W[4] + a12(2, 0) * W[7]);
    S.m02 = A[23] - (a12(0, 0) * W[2] + a12(1, 0) * W[5] + a12(2, 0) * W[8]);
    S.m11 = A[28] - (a12(0, 1) * W[1] + a12(1, 1) * W[4] + a12(2, 1) * W[7]);
    S.m12 = A[29] - (a12(0, 1) * W[2] + a12(1, 1) * W[5] + a12(2, 1) * W[8]);
    S.m22 = A[35] - (a12(0, 2) * W[2] + a12(1, 2) * W[5] + a12(2, 2) * W[8]);
    double r2[3];
#pragma unroll
    for (int c = 0; c < 3; c++) r2[c] = b[3 + c] - (a12(0, c) * u[0] + a12(1, c) * u[1] + a12(2, c) * u[2]);
    Sym3 IS;
    const bool ok2 = sym3_inverse(S, IS);
    x[3] = IS.m00 * r2[0] + IS.m01 * r2[1] + IS.m02 * r2[2];
    x[4] = IS.m01 * r2[0] + IS.m11 * r2[1] + IS.m12 * r2[2];
    x[5] = IS.m02 * r2[0] + IS.m12 * r2[1] + IS.m22 * r2[2];
#pragma unroll
    for (int r = 0; r < 3; r++) x[r] = u[r] - (W[r * 3] * x[3] + W[r * 3 + 1] * x[4] + W[r * 3 + 2] * x[5]);
    return ok1 && ok2;
}

struct PoseCam { double fx, fy, cx, cy, delta, dsqr; };

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

// One thread-block CLUSTER (8 CTAs x 128 threads) runs the whole of Optimizer::PoseOptimization:
// 4 rounds x optimize(10) of g2o's LM on one SE3 vertex with outlier re-classification between
// rounds.  Every thread keeps its edges (world point, observation, information, state, last error) in
// REGISTERS -- EPT edges per thread, edge k of the frame lives in thread k % 1024, slot k / 1024 -- so
// the ~50 dependent passes of one call touch no global memory.  Every thread carries the pose and the
// 6x6 system redundantly (identical arithmetic), so nothing is broadcast; per-edge terms are reduced in
// a fixed order (transposed warp butterfly -> CTA -> the 8 CTA partials exchanged through distributed
// shared memory), so the result is deterministic.  Each LM trial is ONE pass over the edges: errors,
// robust chi2 and the linearisation at the trial estimate are accumulated together; if the trial is
// accepted the next iteration's buildSystem() is already there (same values g2o would recompute).
constexpr int kPoseCtas = 8;
constexpr int kPoseThreads = 128;
constexpr int kPoseWarps = kPoseThreads / 32;
constexpr int kPoseStride = kPoseCtas * kPoseThreads;
// reduction slots: 0..20 H upper triangle (row-major), 21..26 b, 27 robust chi2, 28 active edges
constexpr int kPoseNV = 32;

struct PoseShared {
    double warp_buf[kPoseWarps][kPoseNV];
    // partials of every CTA of the cluster, double-buffered; filled by the peers' st.async stores, whose
    // bytes are counted by bar[parity] (transaction barrier): consumers sleep on the barrier, nobody polls
    double recv[2][kPoseCtas][kPoseNV];
    unsigned long long bar[2];
};

__device__ inline uint32_t cvta_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr unsigned kPoseTxBytes = kPoseCtas * kPoseNV * 8;

// Sums v[s] over all threads of the cluster; on return v[s] holds the total of slot s in every thread
// (same summation tree everywhere).  v is clobbered during the exchange.  `pass` counts the calls (>= 0).
//   warp butterfly -> CTA partial (one __syncthreads) -> warp 0 sends its 32 partials to the 8 CTAs with
//   st.async (remote store + complete_tx on the receiver's mbarrier) -> every warp waits on the local
//   mbarrier (try_wait sleeps in hardware) and adds the 8 partials in rank order.
__device__ inline void cluster_sum(double (&v)[kPoseNV], PoseShared& sh, unsigned& pass)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned rank;
    asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int parity = (int)(pass & 1u);
    const unsigned phase = (pass >> 1) & 1u;
    // transposed butterfly: after the 5 steps lane l holds the warp total of slot l (31 shuffles)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; i++) {
            const double send = upper ? v[i] : v[i + off];
            const double keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    sh.warp_buf[wid][lane] = v[0];
    __syncthreads(); // also: every warp of this CTA has finished reading recv[parity] of pass - 2
    const uint32_t bar = cvta_smem(&sh.bar[parity]);
    if (wid == 0) {
        // arm this pass's phase (peers' bytes may already have been counted: the tx-count is signed)
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kPoseTxBytes) : "memory");
        double s = sh.warp_buf[0][lane];
#pragma unroll
        for (int w = 1; w < kPoseWarps; w++) s += sh.warp_buf[w][lane];
        const uint32_t local = cvta_smem(&sh.recv[parity][rank][lane]);
#pragma unroll
        for (unsigned r = 0; r < kPoseCtas; r++) {
            uint32_t rdst, rbar;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(local), "r"(r));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(bar), "r"(r));
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                         ::"r"(rdst), "l"(__double_as_longlong(s)), "r"(rbar) : "memory");
        }
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    }
    double tot = sh.recv[parity][0][lane];
#pragma unroll
    for (int r = 1; r < kPoseCtas; r++) tot += sh.recv[parity][r][lane];
#pragma unroll
    for (int i = 0; i < 29; i++) v[i] = __shfl_sync(0xffffffffu, tot, i);
    pass++;
}

template <int EPT>
__global__ void __cluster_dims__(kPoseCtas, 1, 1) __launch_bounds__(kPoseThreads, 1) pose_opt_kernel(PoseOptArgs a)
{
    pdl_trigger(); pdl_wait();
    __shared__ PoseShared sh;
    const int tid = blockIdx.x * kPoseThreads + threadIdx.x;
    unsigned parity = 0; // pass counter of cluster_sum
    const int n = a.n_ptr ? min(*a.n_ptr, a.n) : a.n;
    PoseCam cam;
    cam.fx = a.K[0]; cam.fy = a.K[1]; cam.cx = a.K[2]; cam.cy = a.K[3];
    cam.delta = (double)(float)sqrt(5.991);
    cam.dsqr = cam.delta * cam.delta;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cvta_smem(&sh.bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cvta_smem(&sh.bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cooperative_groups::this_cluster().sync(); // nobody sends before every CTA has initialised its barriers

    // ---- gather this CTA's edges into registers, compacted ----
    // Potential edge k of the frame belongs to CTA (k % kPoseStride) / kPoseThreads.  Only a part of the keypoints carries a
    // map point, so the CTA first packs its valid ones (block scan, order kept) and thread p then takes entries p,
    // p + 128, ...: with at most 128 valid edges per CTA -- a tracked frame of 2000 keypoints -- every thread holds ONE edge
    // and the second register slot is skipped by whole warps in each of the ~25 dependent passes.
    // state: bit0 excluded (level 1), bit1 robust kernel removed, bit2 not an edge
    __shared__ int s_list[EPT * kPoseThreads];
    __shared__ int s_scan[33];
    double X[EPT][3], ox[EPT], oy[EPT], om[EPT], e0[EPT], e1[EPT];
    int st[EPT], mis[EPT], kk[EPT];
    double acc[kPoseNV];
#pragma unroll
    for (int i = 0; i < kPoseNV; i++) acc[i] = 0;
    {
        unsigned vmask = 0u;
        int mine = 0;
#pragma unroll
        for (int s = 0; s < EPT; s++) {
            const int k = tid + s * kPoseStride;
            if (k < n) {
                const bool valid = a.map_index ? a.map_index[k] >= 0 : (a.valid ? a.valid[k] != 0 : true);
                if (valid) { vmask |= 1u << s; mine++; }
                else a.outlier[k] = 0; // not an edge
            }
        }
        int total;
        int off = block_exclusive_scan(mine, s_scan, &total);
#pragma unroll
        for (int s = 0; s < EPT; s++)
            if (vmask & (1u << s)) s_list[off++] = tid + s * kPoseStride;
        __syncthreads();
#pragma unroll
        for (int s = 0; s < EPT; s++) {
            const int slot = threadIdx.x + s * kPoseThreads;
            st[s] = 4; mis[s] = -1; kk[s] = -1;
            X[s][0] = X[s][1] = X[s][2] = 0; ox[s] = oy[s] = om[s] = 0; e0[s] = e1[s] = 0;
            if (slot < total) {
                const int k = s_list[slot];
                const int mi = a.map_index ? a.map_index[k] : k;
                st[s] = 0; mis[s] = mi; kk[s] = k;
                const float* xp = a.Xw + 3 * (size_t)mi;
                X[s][0] = (double)xp[0]; X[s][1] = (double)xp[1]; X[s][2] = (double)xp[2];
                if (a.kps) {
                    ox[s] = (double)a.kps[k].x; oy[s] = (double)a.kps[k].y;
                    om[s] = (double)a.inv_sigma2_table[a.kps[k].octave];
                } else {
                    ox[s] = (double)a.kp_xy[2 * k]; oy[s] = (double)a.kp_xy[2 * k + 1];
                    om[s] = (double)a.inv_sigma2[k];
                }
                acc[28] += 1;
            }
        }
    }
    cluster_sum(acc, sh, parity);
    const int nedges = (int)acc[28];

    SE3d T0;
    T0.r.x = a.pose[0]; T0.r.y = a.pose[1]; T0.r.z = a.pose[2]; T0.r.w = a.pose[3];
    T0.t[0] = a.pose[4]; T0.t[1] = a.pose[5]; T0.t[2] = a.pose[6];
    quat_normalize(T0.r);
    SE3d T = T0;
    int nBadEdges = 0, total_iters = 0, total_trials = 0;

    unsigned long long pr_t = 0, pr_edges = 0, pr_reduce = 0, pr_solve = 0, pr_pass = 0;
    auto stamp = [&](unsigned long long& bucket) {
        if (a.prof) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            bucket += now - pr_t;
            pr_t = now;
        }
    };
    // errors + robust chi2 + linearisation of every active edge at estimate Tx -> acc (all threads)
    auto linearize_at = [&](const SE3d& Tx) {
        stamp(pr_solve);
        double R[9];
        quat_to_matrix(Tx.r, R);
#pragma unroll
        for (int i = 0; i < kPoseNV; i++) acc[i] = 0;
#pragma unroll
        for (int s = 0; s < EPT; s++) {
            if (st[s] & 5) continue;
            const double x = R[0] * X[s][0] + R[1] * X[s][1] + R[2] * X[s][2] + Tx.t[0];
            const double y = R[3] * X[s][0] + R[4] * X[s][1] + R[5] * X[s][2] + Tx.t[1];
            const double z = R[6] * X[s][0] + R[7] * X[s][1] + R[8] * X[s][2] + Tx.t[2];
            const double iz = fast_rcp(z);
            const double ax = cam.fx * iz, ay = cam.fy * iz;      // d u / d x, d v / d y
            const double bx = -ax * x * iz, by = -ay * y * iz;    // d u / d z, d v / d z
            const double ex = ox[s] - (ax * x + cam.cx), ey = oy[s] - (ay * y + cam.cy);
            e0[s] = ex; e1[s] = ey;
            const double chi = om[s] * (ex * ex + ey * ey);
            const bool robust = !(st[s] & 2);
            double w = 1.0, rho = chi;
            if (robust && chi > cam.dsqr) {
                const double isq = fast_rsqrt(chi);
                rho = 2 * (chi * isq) * cam.delta - cam.dsqr;
                w = cam.delta * isq;
            }
            acc[27] += rho;
            acc[28] += 1;
            // J = -dproj * [ -[Xc]x | I ]  (O3/src/OptimizableTypes.cpp:51-63); J0[4] = J1[3] = 0
            const double J0[6] = { -bx * y, bx * x - ax * z, ax * y, -ax, 0.0, -bx };
            const double J1[6] = { ay * z - by * y, by * x, -ay * x, 0.0, -ay, -by };
            const double sw = w * om[s];
            double S0[6], S1[6];
#pragma unroll
            for (int c = 0; c < 6; c++) { S0[c] = sw * J0[c]; S1[c] = sw * J1[c]; }
            int idx = 0;
#pragma unroll
            for (int c = 0; c < 6; c++) {
                if (c != 4) acc[21 + c] -= S0[c] * ex;
                if (c != 3) acc[21 + c] -= S1[c] * ey;
#pragma unroll
                for (int d = c; d < 6; d++) {
                    if (c != 4 && d != 4) acc[idx] += S0[c] * J0[d];
                    if (c != 3 && d != 3) acc[idx] += S1[c] * J1[d];
                    idx++;
                }
            }
        }
        stamp(pr_edges);
        cluster_sum(acc, sh, parity);
        stamp(pr_reduce);
        pr_pass++;
    };
    auto unpack = [&](double* H, double* b) {
        int idx = 0;
#pragma unroll
        for (int c = 0; c < 6; c++)
#pragma unroll
            for (int d = c; d < 6; d++) { H[c * 6 + d] = acc[idx]; H[d * 6 + c] = acc[idx]; idx++; }
#pragma unroll
        for (int c = 0; c < 6; c++) b[c] = acc[21 + c];
    };

    if (a.prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(pr_t));
    if (nedges >= 3) {
        for (int round = 0; round < 4; round++) {
            T = T0; // the frame's pose is only written back at the end (O3/src/Optimizer.cc:935-936)
            // ---- optimize(10) ----
            double lambda = -1, ni = 2;
            int nBadIter = 0;
            double H[36], b[6], currentChi = 0;
            bool have_system = false; // H, b, currentChi valid for the current T
            for (int it = 0; it < 10; it++) {
                if (!have_system) {
                    linearize_at(T);
                    if ((int)acc[28] == 0) break; // no active edge: optimize() returns without touching anything
                    unpack(H, b);
                    currentChi = acc[27];
                }
                const double iniChi = currentChi;
                if (it == 0) {
                    double mx = 0;
#pragma unroll
                    for (int j = 0; j < 6; j++) mx = fmax(fabs(H[j * 6 + j]), mx);
                    lambda = 1e-5 * mx; // computeLambdaInit, tau = 1e-5
                    ni = 2;
                    nBadIter = 0;
                }
                double rho = 0;
                int qmax = 0;
                have_system = false;
                do {
                    const SE3d backup = T;
                    double Hl[36], x[6];
#pragma unroll
                    for (int j = 0; j < 36; j++) Hl[j] = H[j];
#pragma unroll
                    for (int j = 0; j < 6; j++) Hl[j * 6 + j] += lambda;
                    const bool ok2 = spd6_solve(Hl, b, x);
                    if (!ok2) {
#pragma unroll
                        for (int j = 0; j < 6; j++) x[j] = 0;
                    }
                    T = se3_left_update(x, T);
                    linearize_at(T); // computeActiveErrors at the trial (+ speculative buildSystem)
                    double tempChi = acc[27];
                    if (!ok2) tempChi = 1.7976931348623157e308;
                    rho = currentChi - tempChi;
                    double scale = 0;
#pragma unroll
                    for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
                    scale += 1e-3;
                    rho *= fast_rcp(scale);
                    if (rho > 0 && isfinite(tempChi)) {
                        const double c = 2 * rho - 1;
                        double alpha = 1. - c * c * c;
                        alpha = fmin(alpha, 2. / 3.);
                        const double sf = fmax(1. / 3., alpha);
                        lambda *= sf;
                        ni = 2;
                        currentChi = tempChi;
                        unpack(H, b); // the accepted trial's linearisation is the next iteration's system
                        have_system = true;
                    } else {
                        lambda *= ni;
                        ni *= 2;
                        T = backup;
                    }
                    qmax++;
                    total_trials++;
                } while (rho < 0 && qmax < 10);
                total_iters++;
                if (qmax == 10 || rho == 0) break;
                if ((iniChi - currentChi) * 1e3 < iniChi) nBadIter++;
                else nBadIter = 0;
                if (nBadIter >= 3) break;
            }
            // ---- re-classify every edge (O3/src/Optimizer.cc:941-965): inliers keep the error of the
            // last computeActiveErrors(), outliers are re-evaluated at the final estimate ----
            double R[9];
            quat_to_matrix(T.r, R);
#pragma unroll
            for (int i = 0; i < kPoseNV; i++) acc[i] = 0;
#pragma unroll
            for (int s = 0; s < EPT; s++) {
                if (st[s] & 4) continue;
                if (st[s] & 1) {
                    const double x = R[0] * X[s][0] + R[1] * X[s][1] + R[2] * X[s][2] + T.t[0];
                    const double y = R[3] * X[s][0] + R[4] * X[s][1] + R[5] * X[s][2] + T.t[1];
                    const double z = R[6] * X[s][0] + R[7] * X[s][1] + R[8] * X[s][2] + T.t[2];
                    const double iz = fast_rcp(z);
                    e0[s] = ox[s] - (cam.fx * iz * x + cam.cx);
                    e1[s] = oy[s] - (cam.fy * iz * y + cam.cy);
                }
                const float chi2 = (float)(om[s] * (e0[s] * e0[s] + e1[s] * e1[s]));
                if (chi2 > 5.991f) { st[s] |= 1; acc[0] += 1; }
                else st[s] &= ~1;
                if (round == 2) st[s] |= 2;
            }
            cluster_sum(acc, sh, parity);
            nBadEdges = (int)acc[0];
            if (nedges < 10) break;
        }
    }
#pragma unroll
    for (int s = 0; s < EPT; s++) {
        const int k = kk[s];
        if (k < 0) continue;
        int out = st[s] & 1;
        if (a.seen) { // "discard outliers": Tracking.cc:2634-2654
            a.seen[mis[s]] = 1;
            if (out) { a.map_index_rw[k] = -1; out = 0; }
        }
        a.outlier[k] = (uint8_t)out;
    }
    if (tid == 0) {
        if (nedges >= 3) {
            // pFrame->SetPose(Sophus::SE3<float>(rotation().cast<float>(), translation().cast<float>())), Optimizer.cc:1021-1024:
            // the SE3f constructor normalises the float quaternion (so3.hpp:480-487; 4-float norm (x^2 + z^2) + (y^2 + w^2))
            {
                const float qx = (float)T.r.x, qy = (float)T.r.y, qz = (float)T.r.z, qw = (float)T.r.w;
                const float qn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qz, qz)), __fadd_rn(__fmul_rn(qy, qy), __fmul_rn(qw, qw))));
                a.pose[0] = __fdiv_rn(qx, qn); a.pose[1] = __fdiv_rn(qy, qn); a.pose[2] = __fdiv_rn(qz, qn); a.pose[3] = __fdiv_rn(qw, qn);
            }
            a.pose[4] = (float)T.t[0]; a.pose[5] = (float)T.t[1]; a.pose[6] = (float)T.t[2];
        }
        a.result[0] = nedges >= 3 ? nedges - nBadEdges : 0;
        a.result[1] = nedges;
        a.result[2] = total_iters;
        a.result[3] = total_trials;
        if (a.prof) { a.prof[0] = pr_edges; a.prof[1] = pr_reduce; a.prof[2] = pr_solve; a.prof[3] = pr_pass; }
        if (a.out_pose) { // frame hand-over: pose history for the constant-velocity prior + result block
            float cur[7], prev[7];
            for (int i = 0; i < 7; i++) {
                cur[i] = a.pose[i];
                prev[i] = a.pose_last[i];
                a.pose_prev[i] = prev[i];
                a.pose_last[i] = cur[i];
                a.out_pose[i] = cur[i];
            }
            if (a.next_prior) { // mVelocity = Tcw * LastTwc, next prior = mVelocity * Tcw, in the reference's float32 Sophus arithmetic
                float qpi[4], tpi[3], qv[4], tv[3], qo[4], to[3];
                so::se3_inverse(prev, prev + 4, qpi, tpi);
                so::se3_mul(cur, cur + 4, qpi, tpi, qv, tv);
                so::se3_mul(qv, tv, cur, cur + 4, qo, to);
                for (int i = 0; i < 4; i++) a.next_prior[i] = qo[i];
                for (int i = 0; i < 3; i++) a.next_prior[4 + i] = to[i];
            }
            a.out_counts[0] = n;
            a.out_counts[1] = *a.nm_last;
            a.out_counts[2] = a.res_first[0];
            a.out_counts[3] = nedges >= 3 ? nedges - nBadEdges : nedges; // mnMatchesInliers
        }
    }
    if (a.seen_reset)
        for (int i = tid; i < (a.seen_n + 3) / 4; i += kPoseStride) reinterpret_cast<uint32_t*>(a.seen_reset)[i] = 0u;
    cooperative_groups::this_cluster().sync(); // no CTA may exit while peers can still write into its shared memory
}

int launch_pose_opt(const PoseOptArgs& a_in, cudaStream_t stream)
{
    PoseOptArgs a = a_in;
    static const bool profile = getenv("DVM_POSE_PROFILE") != nullptr;
    static unsigned long long* d_prof = nullptr;
    if (profile) {
        if (!d_prof) cudaMalloc(&d_prof, 32);
        a.prof = d_prof;
    }
    struct Report { // prints the phase timers of this launch when profiling (synchronises: diagnostics only)
        cudaStream_t st; bool on; unsigned long long* d;
        ~Report()
        {
            if (!on) return;
            unsigned long long pr[4] = { 0, 0, 0, 0 };
            cudaStreamSynchronize(st);
            cudaMemcpy(pr, d, 32, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[pose-opt us] edges %.1f reduce %.1f solve %.1f passes %llu\n", pr[0] * 1e-3, pr[1] * 1e-3,
                    pr[2] * 1e-3, pr[3]);
        }
    } report{ stream, profile, d_prof };
    // a.n bounds the number of potential edges (the device count may be smaller): pick the register tile
    if (a.n <= 2 * kPoseStride) DVM_LAUNCH_PDL(pose_opt_kernel<2>, kPoseCtas, kPoseThreads, 0, stream, a);
    else if (a.n <= 4 * kPoseStride) DVM_LAUNCH_PDL(pose_opt_kernel<4>, kPoseCtas, kPoseThreads, 0, stream, a);
    else if (a.n <= 8 * kPoseStride) DVM_LAUNCH_PDL(pose_opt_kernel<8>, kPoseCtas, kPoseThreads, 0, stream, a);
    else if (a.n <= 16 * kPoseStride) DVM_LAUNCH_PDL(pose_opt_kernel<16>, kPoseCtas, kPoseThreads, 0, stream, a);
    else {
        set_error("PoseOptimization: %d correspondences exceed the supported %d", a.n, 16 * kPoseStride);
        return DVM_ERR_CAPACITY;
    }
    return DVM_OK;
}

} // namespace dvm
