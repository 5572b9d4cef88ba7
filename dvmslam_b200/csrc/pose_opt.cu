// pose_opt.cu -- Optimizer::PoseOptimization on one CTA (O3/src/Optimizer.cc:744-1028).
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
        const double sc = (theta - sin(theta)) / pow(theta, 3.0);
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
    double L[36], D[6], y[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = A[j * 6 + j];
        for (int k = 0; k < j; k++) d -= L[j * 6 + k] * L[j * 6 + k] * D[k];
        if (!(d > 0)) return false;
        D[j] = d;
        for (int i = j + 1; i < 6; i++) {
            double s = A[i * 6 + j];
            for (int k = 0; k < j; k++) s -= L[i * 6 + k] * L[j * 6 + k] * D[k];
            L[i * 6 + j] = s / d;
        }
    }
    for (int i = 0; i < 6; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L[i * 6 + k] * y[k];
        y[i] = s;
    }
    for (int i = 5; i >= 0; i--) {
        double s = y[i] / D[i];
        for (int k = i + 1; k < 6; k++) s -= L[k * 6 + i] * x[k];
        x[i] = s;
    }
    return true;
}

struct PoseCam { double fx, fy, cx, cy, delta, dsqr; };

__device__ inline void pose_edge_error(const PoseCam& c, const SE3d& T, const float* Xw, const float* obs, double e[2], double xc[3])
{
    const double X[3] = { (double)Xw[0], (double)Xw[1], (double)Xw[2] };
    quat_rotate(T.r, X, xc);
    xc[0] += T.t[0]; xc[1] += T.t[1]; xc[2] += T.t[2];
    e[0] = (double)obs[0] - (c.fx * xc[0] / xc[2] + c.cx);
    e[1] = (double)obs[1] - (c.fy * xc[1] / xc[2] + c.cy);
}
__device__ inline double pose_chi2(const double e[2], double info) { return e[0] * (info * e[0]) + e[1] * (info * e[1]); }
__device__ inline double huber_rho0(const PoseCam& c, double e) { return e <= c.dsqr ? e : 2 * sqrt(e) * c.delta - c.dsqr; }
__device__ inline double huber_rho1(const PoseCam& c, double e) { return e <= c.dsqr ? 1.0 : c.delta / sqrt(e); }

// One thread-block CLUSTER (8 CTAs x 128 threads, one edge per thread up to 1024 edges) runs the whole
// of Optimizer::PoseOptimization: 4 rounds x optimize(10) of g2o's LM on one SE3 vertex with outlier
// re-classification between rounds.  Every thread carries the pose and the 6x6 system redundantly
// (identical arithmetic), so nothing is broadcast; per-edge terms are reduced in a fixed order
// (warp shuffle -> CTA -> the 8 CTA partials exchanged through distributed shared memory), so the
// result is deterministic.  Each LM trial is ONE pass over the edges: errors, robust chi2 and the
// linearisation at the trial estimate are accumulated together; if the trial is accepted the next
// iteration's buildSystem() is already there (same values g2o would recompute).
constexpr int kPoseCtas = 8;
constexpr int kPoseThreads = 128;
constexpr int kPoseWarps = kPoseThreads / 32;
constexpr int kPoseNV = 30; // 21 H + 6 b + chi + active count + spare

struct PoseShared {
    double warp_buf[kPoseWarps][kPoseNV];
    double recv[2][kPoseCtas][kPoseNV]; // partials of every CTA of the cluster, double-buffered
    double tot[kPoseNV];
};

// sums v over all threads of the cluster; result in sh.tot (valid until the next call)
__device__ inline void cluster_sum(double (&v)[kPoseNV], PoseShared& sh, int& parity)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned rank = cluster.block_rank();
#pragma unroll
    for (int i = 0; i < kPoseNV; i++) {
        double s = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) sh.warp_buf[wid][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < kPoseNV) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < kPoseWarps; w++) s += sh.warp_buf[w][threadIdx.x];
        for (unsigned r = 0; r < kPoseCtas; r++) {
            double* dst = cluster.map_shared_rank(&sh.recv[parity][rank][threadIdx.x], r);
            *dst = s;
        }
    }
    cluster.sync();
    if (threadIdx.x < kPoseNV) {
        double s = 0;
#pragma unroll
        for (int r = 0; r < kPoseCtas; r++) s += sh.recv[parity][r][threadIdx.x];
        sh.tot[threadIdx.x] = s;
    }
    __syncthreads();
    parity ^= 1;
}

__global__ void __cluster_dims__(kPoseCtas, 1, 1) __launch_bounds__(kPoseThreads, 1) pose_opt_kernel(PoseOptArgs a)
{
    __shared__ PoseShared sh;
    const int tid = blockIdx.x * kPoseThreads + threadIdx.x;
    constexpr int kStride = kPoseCtas * kPoseThreads;
    int parity = 0;
    const int n = a.n_ptr ? min(*a.n_ptr, a.n) : a.n;
    auto edge_valid = [&](int k) { return a.map_index ? a.map_index[k] >= 0 : (a.valid ? a.valid[k] != 0 : true); };
    auto edge_X = [&](int k) { return a.Xw + 3 * (size_t)(a.map_index ? a.map_index[k] : k); };
    auto edge_info = [&](int k) { return (double)(a.kps ? a.inv_sigma2_table[a.kps[k].octave] : a.inv_sigma2[k]); };
    auto edge_err = [&](const PoseCam& c, const SE3d& T, int k, double e[2], double xc[3]) {
        float o[2];
        if (a.kps) { o[0] = a.kps[k].x; o[1] = a.kps[k].y; }
        else { o[0] = a.kp_xy[2 * k]; o[1] = a.kp_xy[2 * k + 1]; }
        pose_edge_error(c, T, edge_X(k), o, e, xc);
    };
    PoseCam cam;
    cam.fx = a.K[0]; cam.fy = a.K[1]; cam.cx = a.K[2]; cam.cy = a.K[3];
    cam.delta = (double)(float)sqrt(5.991);
    cam.dsqr = cam.delta * cam.delta;

    // edge state byte (kept in a.outlier until the end): bit0 excluded (level 1), bit1 robust kernel
    // removed, bit2 not an edge
    double acc[kPoseNV];
#pragma unroll
    for (int i = 0; i < kPoseNV; i++) acc[i] = 0;
    for (int k = tid; k < n; k += kStride) {
        const bool valid = edge_valid(k);
        a.outlier[k] = valid ? 0 : 4;
        acc[28] += valid;
    }
    cluster_sum(acc, sh, parity);
    const int nedges = (int)sh.tot[28];

    SE3d T0;
    T0.r.x = a.pose[0]; T0.r.y = a.pose[1]; T0.r.z = a.pose[2]; T0.r.w = a.pose[3];
    T0.t[0] = a.pose[4]; T0.t[1] = a.pose[5]; T0.t[2] = a.pose[6];
    quat_normalize(T0.r);
    SE3d T = T0;
    int nBadEdges = 0, total_iters = 0, total_trials = 0;

    // errors + robust chi2 + linearisation of every active edge at estimate Tx -> sh.tot
    // (tot[0..20] H upper, [21..26] b, [27] robust chi2, [28] active edges)
    auto linearize_at = [&](const SE3d& Tx) {
#pragma unroll
        for (int i = 0; i < kPoseNV; i++) acc[i] = 0;
        for (int k = tid; k < n; k += kStride) {
            const int st = a.outlier[k];
            if (st & 5) continue;
            acc[28] += 1;
            double e[2], xc[3];
            edge_err(cam, Tx, k, e, xc);
            a.err[2 * k] = e[0]; a.err[2 * k + 1] = e[1];
            const double om = edge_info(k);
            const double chi = pose_chi2(e, om);
            const bool robust = !(st & 2);
            acc[27] += robust ? huber_rho0(cam, chi) : chi;
            const double w = robust ? huber_rho1(cam, chi) : 1.0;
            const double x = xc[0], y = xc[1], z = xc[2];
            const double pj[6] = { cam.fx / z, 0, -cam.fx * x / (z * z), 0, cam.fy / z, -cam.fy * y / (z * z) };
            const double D[18] = { 0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1 };
            double J[12];
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 6; c++)
                    J[r * 6 + c] = -(pj[r * 3] * D[c] + pj[r * 3 + 1] * D[6 + c] + pj[r * 3 + 2] * D[12 + c]);
            int idx = 0;
#pragma unroll
            for (int c = 0; c < 6; c++) {
                acc[21 + c] -= w * (J[c] * (om * e[0]) + J[6 + c] * (om * e[1]));
#pragma unroll
                for (int d = c; d < 6; d++) acc[idx++] += J[c] * (w * om) * J[d] + J[6 + c] * (w * om) * J[6 + d];
            }
        }
        cluster_sum(acc, sh, parity);
    };
    auto unpack = [&](double* H, double* b) {
        int idx = 0;
        for (int c = 0; c < 6; c++)
            for (int d = c; d < 6; d++) { H[c * 6 + d] = sh.tot[idx]; H[d * 6 + c] = sh.tot[idx]; idx++; }
        for (int c = 0; c < 6; c++) b[c] = sh.tot[21 + c];
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
                    if ((int)sh.tot[28] == 0) break; // no active edge: optimize() returns without touching anything
                    unpack(H, b);
                    currentChi = sh.tot[27];
                }
                const double iniChi = currentChi;
                if (it == 0) {
                    double mx = 0;
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
                    for (int j = 0; j < 36; j++) Hl[j] = H[j];
                    for (int j = 0; j < 6; j++) Hl[j * 6 + j] += lambda;
                    const bool ok2 = ldlt6_solve(Hl, b, x);
                    if (!ok2) for (int j = 0; j < 6; j++) x[j] = 0;
                    T = se3_mul(se3_exp(x), T);
                    linearize_at(T); // computeActiveErrors at the trial (+ speculative buildSystem)
                    double tempChi = sh.tot[27];
                    if (!ok2) tempChi = 1.7976931348623157e308;
                    rho = currentChi - tempChi;
                    double scale = 0;
                    for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
                    scale += 1e-3;
                    rho /= scale;
                    if (rho > 0 && isfinite(tempChi)) {
                        double alpha = 1. - pow((2 * rho - 1), 3.0);
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
            // ---- re-classify every edge (O3/src/Optimizer.cc:941-965) ----
#pragma unroll
            for (int i = 0; i < kPoseNV; i++) acc[i] = 0;
            for (int k = tid; k < n; k += kStride) {
                int st = a.outlier[k];
                if (st & 4) continue;
                double e[2] = { a.err[2 * k], a.err[2 * k + 1] };
                if (st & 1) { double xc[3]; edge_err(cam, T, k, e, xc); a.err[2 * k] = e[0]; a.err[2 * k + 1] = e[1]; }
                const float chi2 = (float)pose_chi2(e, edge_info(k));
                if (chi2 > 5.991f) { st |= 1; acc[0] += 1; }
                else st &= ~1;
                if (round == 2) st |= 2;
                a.outlier[k] = (uint8_t)st;
            }
            cluster_sum(acc, sh, parity);
            nBadEdges = (int)sh.tot[0];
            if (nedges < 10) break;
        }
    }
    for (int k = tid; k < n; k += kStride) a.outlier[k] = (a.outlier[k] & 4) ? 0 : (a.outlier[k] & 1);
    if (tid == 0) {
        if (nedges >= 3) {
            a.pose[0] = (float)T.r.x; a.pose[1] = (float)T.r.y; a.pose[2] = (float)T.r.z; a.pose[3] = (float)T.r.w;
            a.pose[4] = (float)T.t[0]; a.pose[5] = (float)T.t[1]; a.pose[6] = (float)T.t[2];
        }
        a.result[0] = nedges >= 3 ? nedges - nBadEdges : 0;
        a.result[1] = nedges;
        a.result[2] = total_iters;
        a.result[3] = total_trials;
    }
    cooperative_groups::this_cluster().sync(); // no CTA may exit while peers can still write into its shared memory
}

void launch_pose_opt(const PoseOptArgs& a, cudaStream_t stream) { DVM_LAUNCH(pose_opt_kernel, kPoseCtas, kPoseThreads, 0, stream, a); }

} // namespace dvm
