// pose_opt.cu -- Optimizer::PoseOptimization on one 8-CTA cluster (O3/src/Optimizer.cc:744-1028).
// Float64 work inside the stated tolerance: this file is compiled with FMA contraction enabled.
#include "track_kernels.cuh"
#include <cooperative_groups.h>

namespace dvm {

// ----------------------------------------------------------------------------- PoseOptimization

struct Quat { double x, y, z, w; };
struct SE3d { Quat r; double t[3]; };

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
__device__ inline Quat quat_from_matrix(const double R[9])
{
    Quat q;
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[7] - R[5]) * t;
        q.y = (R[2] - R[6]) * t;
        q.z = (R[3] - R[1]) * t;
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
// SE3Quat::exp (g2o/types/se3quat.h:212-240)
__device__ inline SE3d se3_exp(const double u[6])
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
        const double sc = (theta - sin(theta)) / (theta * theta * theta);
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const double I = (i % 4 == 0) ? 1.0 : 0.0;
            R[i] = I + sa * O[i] + sb * O2[i];
            V[i] = I + sb * O[i] + sc * O2[i];
        }
    }
    SE3d T;
    T.r = quat_from_matrix(R);
    quat_normalize(T.r);
#pragma unroll
    for (int i = 0; i < 3; i++) T.t[i] = V[i * 3] * u[3] + V[i * 3 + 1] * u[4] + V[i * 3 + 2] * u[5];
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

// unpivoted LDL^T of a symmetric 6x6 system (full storage); false if a pivot is not positive
__device__ inline bool ldlt6_solve(const double* A, const double* b, double* x)
{
    double L[36], D[6], Dinv[6], y[6];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[j * 6 + j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
        ok = ok && (d > 0);
        D[j] = d;
        Dinv[j] = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double s = A[i * 6 + j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k] * D[k];
            L[i * 6 + j] = s * Dinv[j];
        }
    }
    if (!ok) return false;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= L[i * 6 + k] * y[k];
        y[i] = s;
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double s = y[i] * Dinv[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * x[k];
        x[i] = s;
    }
    return true;
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
    double recv[2][kPoseCtas][kPoseNV]; // partials of every CTA of the cluster, double-buffered
};

// Sums v[s] over all threads of the cluster; on return v[s] holds the total of slot s in every thread
// (same summation tree everywhere).  v is clobbered during the exchange.
__device__ inline void cluster_sum(double (&v)[kPoseNV], PoseShared& sh, int& parity)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned rank = cluster.block_rank();
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
    __syncthreads();
    if (wid == 0) {
        double s = sh.warp_buf[0][lane];
#pragma unroll
        for (int w = 1; w < kPoseWarps; w++) s += sh.warp_buf[w][lane];
#pragma unroll
        for (unsigned r = 0; r < kPoseCtas; r++) *cluster.map_shared_rank(&sh.recv[parity][rank][lane], r) = s;
    }
    cluster.sync();
    double tot = sh.recv[parity][0][lane];
#pragma unroll
    for (int r = 1; r < kPoseCtas; r++) tot += sh.recv[parity][r][lane];
#pragma unroll
    for (int i = 0; i < 29; i++) v[i] = __shfl_sync(0xffffffffu, tot, i);
    parity ^= 1;
}

template <int EPT>
__global__ void __cluster_dims__(kPoseCtas, 1, 1) __launch_bounds__(kPoseThreads, 1) pose_opt_kernel(PoseOptArgs a)
{
    __shared__ PoseShared sh;
    const int tid = blockIdx.x * kPoseThreads + threadIdx.x;
    int parity = 0;
    const int n = a.n_ptr ? min(*a.n_ptr, a.n) : a.n;
    PoseCam cam;
    cam.fx = a.K[0]; cam.fy = a.K[1]; cam.cx = a.K[2]; cam.cy = a.K[3];
    cam.delta = (double)(float)sqrt(5.991);
    cam.dsqr = cam.delta * cam.delta;

    // ---- gather this thread's edges into registers ----
    // state: bit0 excluded (level 1), bit1 robust kernel removed, bit2 not an edge
    double X[EPT][3], ox[EPT], oy[EPT], om[EPT], e0[EPT], e1[EPT];
    int st[EPT], mis[EPT];
    double acc[kPoseNV];
#pragma unroll
    for (int i = 0; i < kPoseNV; i++) acc[i] = 0;
#pragma unroll
    for (int s = 0; s < EPT; s++) {
        const int k = tid + s * kPoseStride;
        st[s] = 4; mis[s] = -1;
        X[s][0] = X[s][1] = X[s][2] = 0; ox[s] = oy[s] = om[s] = 0; e0[s] = e1[s] = 0;
        if (k < n) {
            const int mi = a.map_index ? a.map_index[k] : k;
            const bool valid = a.map_index ? mi >= 0 : (a.valid ? a.valid[k] != 0 : true);
            if (valid) {
                st[s] = 0; mis[s] = mi;
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

    // errors + robust chi2 + linearisation of every active edge at estimate Tx -> acc (all threads)
    auto linearize_at = [&](const SE3d& Tx) {
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
            const double iz = 1.0 / z;
            const double ax = cam.fx * iz, ay = cam.fy * iz;      // d u / d x, d v / d y
            const double bx = -ax * x * iz, by = -ay * y * iz;    // d u / d z, d v / d z
            const double ex = ox[s] - (ax * x + cam.cx), ey = oy[s] - (ay * y + cam.cy);
            e0[s] = ex; e1[s] = ey;
            const double chi = om[s] * (ex * ex + ey * ey);
            const bool robust = !(st[s] & 2);
            double w = 1.0, rho = chi;
            if (robust && chi > cam.dsqr) {
                const double sq = sqrt(chi);
                rho = 2 * sq * cam.delta - cam.dsqr;
                w = cam.delta / sq;
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
        cluster_sum(acc, sh, parity);
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
                    const bool ok2 = ldlt6_solve(Hl, b, x);
                    if (!ok2) {
#pragma unroll
                        for (int j = 0; j < 6; j++) x[j] = 0;
                    }
                    T = se3_mul(se3_exp(x), T);
                    linearize_at(T); // computeActiveErrors at the trial (+ speculative buildSystem)
                    double tempChi = acc[27];
                    if (!ok2) tempChi = 1.7976931348623157e308;
                    rho = currentChi - tempChi;
                    double scale = 0;
#pragma unroll
                    for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
                    scale += 1e-3;
                    rho /= scale;
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
                    const double iz = 1.0 / z;
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
        const int k = tid + s * kPoseStride;
        if (k >= n) continue;
        int out = (st[s] & 4) ? 0 : (st[s] & 1);
        if (a.seen && !(st[s] & 4)) { // "discard outliers": Tracking.cc:2634-2654
            a.seen[mis[s]] = 1;
            if (out) { a.map_index_rw[k] = -1; out = 0; }
        }
        a.outlier[k] = (uint8_t)out;
    }
    if (tid == 0) {
        if (nedges >= 3) {
            a.pose[0] = (float)T.r.x; a.pose[1] = (float)T.r.y; a.pose[2] = (float)T.r.z; a.pose[3] = (float)T.r.w;
            a.pose[4] = (float)T.t[0]; a.pose[5] = (float)T.t[1]; a.pose[6] = (float)T.t[2];
        }
        a.result[0] = nedges >= 3 ? nedges - nBadEdges : 0;
        a.result[1] = nedges;
        a.result[2] = total_iters;
        a.result[3] = total_trials;
        if (a.out_pose) { // frame hand-over: pose history for the constant-velocity prior + result block
            for (int i = 0; i < 7; i++) {
                const float v = a.pose[i];
                a.pose_prev[i] = a.pose_last[i];
                a.pose_last[i] = v;
                a.out_pose[i] = v;
            }
            a.out_counts[0] = n;
            a.out_counts[1] = *a.nm_last;
            a.out_counts[2] = a.res_first[0];
            a.out_counts[3] = nedges >= 3 ? nedges - nBadEdges : nedges; // mnMatchesInliers
        }
    }
    cooperative_groups::this_cluster().sync(); // no CTA may exit while peers can still write into its shared memory
}

int launch_pose_opt(const PoseOptArgs& a, cudaStream_t stream)
{
    // a.n bounds the number of potential edges (the device count may be smaller): pick the register tile
    if (a.n <= 2 * kPoseStride) DVM_LAUNCH(pose_opt_kernel<2>, kPoseCtas, kPoseThreads, 0, stream, a);
    else if (a.n <= 4 * kPoseStride) DVM_LAUNCH(pose_opt_kernel<4>, kPoseCtas, kPoseThreads, 0, stream, a);
    else if (a.n <= 8 * kPoseStride) DVM_LAUNCH(pose_opt_kernel<8>, kPoseCtas, kPoseThreads, 0, stream, a);
    else if (a.n <= 16 * kPoseStride) DVM_LAUNCH(pose_opt_kernel<16>, kPoseCtas, kPoseThreads, 0, stream, a);
    else {
        set_error("PoseOptimization: %d correspondences exceed the supported %d", a.n, 16 * kPoseStride);
        return DVM_ERR_CAPACITY;
    }
    return DVM_OK;
}

} // namespace dvm
