/*
 * oracle/track_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement, on flat arrays, of the per-frame tracking operators of the reference
 * (O3/ = /root/reference/src/slam_system/orb_slam3/):
 *   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   O3/src/Frame.cc:481-506,712-782
 *   ORBmatcher::DescriptorDistance                                  O3/src/ORBmatcher.cc:1900-1914
 *   ORBmatcher::SearchByProjection(Frame&, const Frame&, th, mono)  O3/src/ORBmatcher.cc:1553-1748
 *   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)  O3/src/ORBmatcher.cc:44-212
 *   ORBmatcher::SearchForInitialization                             O3/src/ORBmatcher.cc:605-707
 *   ORBmatcher::Fuse (search half, mono)                            O3/src/ORBmatcher.cc:1060-1228
 *   ORBmatcher::ComputeThreeMaxima                                  O3/src/ORBmatcher.cc:1862-1896
 *   Optimizer::PoseOptimization                                     O3/src/Optimizer.cc:744-1028
 *     with g2o's Levenberg-Marquardt (g2o/core/optimization_algorithm_levenberg.cpp:59-188),
 *     unary-edge quadratic form (g2o/core/base_unary_edge.hpp:43-72), Huber kernel
 *     (g2o/core/robust_kernel_impl.cpp:68-81), SE3Quat::exp (g2o/types/se3quat.h:212-240),
 *     EdgeSE3ProjectXYZOnlyPose (O3/include/OptimizableTypes.h:32-57, O3/src/OptimizableTypes.cpp:51-63)
 *     and Pinhole::project / projectJac (O3/src/CameraModels/Pinhole.cpp:38-79).
 *
 * Parity status: PINNED to the reference source.  The reference has no tests or fixtures for this path, so its own
 * ORBmatcher.cc (+ the Frame / MapPoint / Pinhole bodies it calls) and Optimizer::PoseOptimization (+ vendored g2o) are
 * compiled unmodified into oracle/_ref/libref_matcher.so / libref_opt.so (oracle/Makefile `ref`); tests/test_ref_matchers.py
 * requires identical match index arrays, tests/test_ref_optimizer.py float32-identical poses and identical outlier flags.
 * Conventions fixed here: float32 transforms in Sophus' evaluation order (oracle/sophus_order.h), no FMA; the
 * mono path only (Nleft == -1, mvuRight < 0); the 6x6 solve is an unpivoted LDL^T in double
 * (the reference uses Eigen::LDLT, which pivots: results agree to rounding, tolerance in tests).
 */
#include <climits>
#include "sophus_order.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct KeyPt { float x, y, size, angle, response; int32_t octave, class_id; };

const int GRID_COLS = 64, GRID_ROWS = 48; // O3/include/Frame.h:44-45
const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;

struct Frame {
    int n = 0;
    std::vector<KeyPt> kps;       // mvKeysUn
    std::vector<uint8_t> desc;    // mDescriptors
    float minX, minY, maxX, maxY; // mnMinX ...
    float gwInv, ghInv;           // mfGridElementWidthInv / HeightInv
    std::vector<float> scaleFactors;
    std::vector<int> grid[GRID_COLS][GRID_ROWS];
};

/* DescriptorDistance */
int descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    const int32_t* pa = (const int32_t*)a;
    const int32_t* pb = (const int32_t*)b;
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        unsigned int v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

/* Frame::PosInGrid + AssignFeaturesToGrid */
void assign_grid(Frame& F)
{
    for (int i = 0; i < F.n; i++) {
        const KeyPt& kp = F.kps[i];
        int px = (int)std::round((kp.x - F.minX) * F.gwInv);
        int py = (int)std::round((kp.y - F.minY) * F.ghInv);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
        F.grid[px][py].push_back(i);
    }
}

/* Frame::GetFeaturesInArea (mono: bRight = false) */
void features_in_area(const Frame& F, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out)
{
    out.clear();
    const float factorX = r, factorY = r;
    const int nMinCellX = std::max(0, (int)std::floor((x - F.minX - factorX) * F.gwInv));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - F.minX + factorX) * F.gwInv));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - F.minY - factorY) * F.ghInv));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - F.minY + factorY) * F.ghInv));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const std::vector<int>& cell = F.grid[ix][iy];
            for (int idx : cell) {
                const KeyPt& kp = F.kps[idx];
                if (bCheckLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float dx = kp.x - x, dy = kp.y - y;
                if (std::fabs(dx) < factorX && std::fabs(dy) < factorY) out.push_back(idx);
            }
        }
}

/* ComputeThreeMaxima */
void three_maxima(const int* sizes, int L, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = sizes[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

/* ---------------------------------------------------------------- pose-only optimisation */
struct Quat { double x, y, z, w; };
struct SE3 { Quat r; double t[3]; };

void quat_normalize(Quat& q) /* SE3Quat::normalizeRotation */
{
    if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}
Quat quat_mul(const Quat& a, const Quat& b)
{
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
void quat_rotate(const Quat& q, const double v[3], double out[3])
{
    /* Eigen's QuaternionBase::_transformVector: v + 2w (u x v) + 2 u x (u x v) */
    double uv[3] = { q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0] };
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
    out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
    out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}
Quat quat_from_matrix(const double R[9]) /* Eigen quaternion from a rotation matrix (row-major R) */
{
    Quat q;
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[7] - R[5]) * t;
        q.y = (R[2] - R[6]) * t;
        q.z = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 4]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
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
void mat3_mul(const double A[9], const double B[9], double C[9])
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
/* SE3Quat::exp */
SE3 se3_exp(const double u[6])
{
    const double omega[3] = { u[0], u[1], u[2] }, ups[3] = { u[3], u[4], u[5] };
    const double theta = std::sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
    const double O[9] = { 0, -omega[2], omega[1], omega[2], 0, -omega[0], -omega[1], omega[0], 0 };
    double O2[9], R[9], V[9];
    mat3_mul(O, O, O2);
    const double I[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    if (theta < 0.00001) {
        for (int i = 0; i < 9; i++) { R[i] = I[i] + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
        const double c = (theta - std::sin(theta)) / std::pow(theta, 3);
        for (int i = 0; i < 9; i++) { R[i] = I[i] + a * O[i] + b * O2[i]; V[i] = I[i] + b * O[i] + c * O2[i]; }
    }
    SE3 T;
    T.r = quat_from_matrix(R);
    quat_normalize(T.r);
    for (int i = 0; i < 3; i++) T.t[i] = V[i * 3] * ups[0] + V[i * 3 + 1] * ups[1] + V[i * 3 + 2] * ups[2];
    return T;
}
/* SE3Quat::operator* */
SE3 se3_mul(const SE3& a, const SE3& b)
{
    SE3 r = a;
    double rt[3];
    quat_rotate(a.r, b.t, rt);
    for (int i = 0; i < 3; i++) r.t[i] += rt[i];
    r.r = quat_mul(a.r, b.r);
    quat_normalize(r.r);
    return r;
}

struct PoseProblem {
    int n;
    const double* Xw;   // [n*3]
    const double* obs;  // [n*2]
    const double* info; // [n] invSigma2
    double fx, fy, cx, cy; // float intrinsics promoted to double (GeometricCamera.h:105)
    double delta, dsqr;
    std::vector<uint8_t> level;  // 0 = active, 1 = excluded
    std::vector<uint8_t> robust; // robust kernel still attached
    std::vector<double> err;     // [n*2], the edge's _error (last computeError)
};

void edge_error(const PoseProblem& P, const SE3& T, int k, double e[2], double xc[3])
{
    quat_rotate(T.r, P.Xw + 3 * k, xc);
    xc[0] += T.t[0]; xc[1] += T.t[1]; xc[2] += T.t[2];
    e[0] = P.obs[2 * k] - (P.fx * xc[0] / xc[2] + P.cx);
    e[1] = P.obs[2 * k + 1] - (P.fy * xc[1] / xc[2] + P.cy);
}
double edge_chi2(const PoseProblem& P, int k) /* _error.dot(information * _error) */
{
    const double* e = &P.err[2 * k];
    return e[0] * (P.info[k] * e[0]) + e[1] * (P.info[k] * e[1]);
}
void huber(const PoseProblem& P, double e, double rho[3])
{
    if (e <= P.dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
    else {
        const double sq = std::sqrt(e);
        rho[0] = 2 * sq * P.delta - P.dsqr;
        rho[1] = P.delta / sq;
        rho[2] = -0.5 * rho[1] / e;
    }
}
void compute_active_errors(PoseProblem& P, const SE3& T)
{
    double xc[3];
    for (int k = 0; k < P.n; k++)
        if (P.level[k] == 0) edge_error(P, T, k, &P.err[2 * k], xc);
}
double active_robust_chi2(const PoseProblem& P)
{
    double chi = 0, rho[3];
    for (int k = 0; k < P.n; k++) {
        if (P.level[k] != 0) continue;
        const double c = edge_chi2(P, k);
        if (P.robust[k]) { huber(P, c, rho); chi += rho[0]; }
        else chi += c;
    }
    return chi;
}
/* unpivoted LDL^T of a symmetric n x n system; false if a pivot is not positive */
bool ldlt_solve(int n, const double* A, const double* b, double* x)
{
    std::vector<double> L(n * n, 0.0), D(n);
    for (int j = 0; j < n; j++) {
        double d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k] * D[k];
        if (!(d > 0)) return false;
        D[j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k] * D[k];
            L[i * n + j] = s / d;
        }
    }
    std::vector<double> y(n);
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k];
        y[i] = s;
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = y[i] / D[i];
        for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
        x[i] = s;
    }
    return true;
}

/* SparseOptimizer::optimize(10) with OptimizationAlgorithmLevenberg on one SE3 vertex */
int optimize_pose(PoseProblem& P, SE3& T, int iterations, int* lm_trials_total)
{
    int nact = 0;
    for (int k = 0; k < P.n; k++) nact += (P.level[k] == 0);
    if (nact == 0) return -1;
    double lambda = -1, ni = 2;
    int nBad = 0, done = 0;
    for (int it = 0; it < iterations; it++) {
        compute_active_errors(P, T);
        double currentChi = active_robust_chi2(P);
        const double iniChi = currentChi;
        double tempChi;
        /* buildSystem */
        double H[36] = { 0 }, b[6] = { 0 };
        for (int k = 0; k < P.n; k++) {
            if (P.level[k] != 0) continue;
            double xc[3];
            quat_rotate(T.r, P.Xw + 3 * k, xc);
            xc[0] += T.t[0]; xc[1] += T.t[1]; xc[2] += T.t[2];
            const double x = xc[0], y = xc[1], z = xc[2];
            /* -projectJac(xyz) * SE3deriv */
            const double pj[6] = { P.fx / z, 0, -P.fx * x / (z * z), 0, P.fy / z, -P.fy * y / (z * z) };
            const double D[18] = { 0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1 };
            double J[12];
            for (int r = 0; r < 2; r++)
                for (int c = 0; c < 6; c++)
                    J[r * 6 + c] = -(pj[r * 3] * D[c] + pj[r * 3 + 1] * D[6 + c] + pj[r * 3 + 2] * D[12 + c]);
            const double* e = &P.err[2 * k];
            double w = 1.0;
            if (P.robust[k]) { double rho[3]; huber(P, edge_chi2(P, k), rho); w = rho[1]; }
            const double om = P.info[k];
            for (int c = 0; c < 6; c++) {
                b[c] -= w * (J[c] * (om * e[0]) + J[6 + c] * (om * e[1]));
                for (int d = 0; d < 6; d++) H[c * 6 + d] += J[c] * (w * om) * J[d] + J[6 + c] * (w * om) * J[6 + d];
            }
        }
        if (it == 0) {
            double mx = 0;
            for (int j = 0; j < 6; j++) mx = std::max(std::fabs(H[j * 6 + j]), mx);
            lambda = 1e-5 * mx;
            ni = 2;
            nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const SE3 backup = T; /* push */
            double Hl[36], x[6];
            memcpy(Hl, H, sizeof(H));
            for (int j = 0; j < 6; j++) Hl[j * 6 + j] += lambda;
            const bool ok2 = ldlt_solve(6, Hl, b, x);
            if (!ok2) for (int j = 0; j < 6; j++) x[j] = 0; /* LinearSolverDense leaves x untouched (zero-initialised) */
            T = se3_mul(se3_exp(x), T);
            compute_active_errors(P, T);
            tempChi = active_robust_chi2(P);
            if (!ok2) tempChi = std::numeric_limits<double>::max();
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                const double sf = std::max(1. / 3., alpha);
                lambda *= sf;
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                T = backup; /* pop */
            }
            qmax++;
            if (lm_trials_total) ++*lm_trials_total;
        } while (rho < 0 && qmax < 10);
        done++;
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
        else nBad = 0;
        if (nBad >= 3) break;
    }
    return done;
}

} // namespace

extern "C" {

void* trko_frame_create(const void* kps, const uint8_t* desc, int n, float minX, float minY, float maxX, float maxY,
                        const float* scaleFactors, int nlevels)
{
    Frame* F = new Frame;
    F->n = n;
    F->kps.assign((const KeyPt*)kps, (const KeyPt*)kps + n);
    F->desc.assign(desc, desc + (size_t)n * 32);
    F->minX = minX; F->minY = minY; F->maxX = maxX; F->maxY = maxY;
    F->gwInv = static_cast<float>(GRID_COLS) / static_cast<float>(maxX - minX);
    F->ghInv = static_cast<float>(GRID_ROWS) / static_cast<float>(maxY - minY);
    F->scaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    assign_grid(*F);
    return F;
}
void trko_frame_destroy(void* f) { delete (Frame*)f; }

int trko_grid_cell(void* f, int ix, int iy, int* out, int cap)
{
    const std::vector<int>& c = ((Frame*)f)->grid[ix][iy];
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) out[i] = c[i];
    return (int)c.size();
}

int trko_features_in_area(void* f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap)
{
    std::vector<int> v;
    features_in_area(*(Frame*)f, x, y, r, minLevel, maxLevel, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

int trko_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

/* SearchByProjection(CurrentFrame, LastFrame, th, bMono = true).
 * qcw (x, y, z, w) / tcw: the current frame's pose as its SE3f holds it (used as is, not renormalised).
 * Last frame, one entry per keypoint i: has_mp, outlier, world position, the map point's
 * descriptor, whether the map point has Observations() > 0, and the last frame's keypoint
 * octave / angle.  cur_mp[cur.n] (in: all -1) receives the last-frame index matched to each current
 * keypoint.  Returns nmatches. */
int trko_search_by_projection_last(void* fcur, const float* qcw, const float* tcw, const float* K, int lastN,
                                   const uint8_t* has_mp, const uint8_t* outlier, const float* Xw,
                                   const uint8_t* mp_desc, const uint8_t* mp_obs_pos, const int* last_octave,
                                   const float* last_angle, float th, int checkOri, int* cur_mp)
{
    Frame& C = *(Frame*)fcur;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> idx;
    for (int i = 0; i < C.n; i++) cur_mp[i] = -1;
    for (int i = 0; i < lastN; i++) {
        if (!has_mp[i] || outlier[i]) continue;
        float x3Dc[3];
        so::se3_apply(qcw, tcw, Xw + 3 * i, x3Dc);   /* x3Dc = Tcw * x3Dw (:1577): Sophus' quaternion action */
        const float xc = x3Dc[0], yc = x3Dc[1], zc = x3Dc[2];
        const float invzc = (float)(1.0 / zc);
        if (invzc < 0) continue;
        const float u = K[0] * xc / zc + K[2];
        const float v = K[1] * yc / zc + K[3];
        if (u < C.minX || u > C.maxX) continue;
        if (v < C.minY || v > C.maxY) continue;
        const int nLastOctave = last_octave[i];
        const float radius = th * C.scaleFactors[nLastOctave];
        features_in_area(C, u, v, radius, nLastOctave - 1, nLastOctave + 1, idx);
        if (idx.empty()) continue;
        const uint8_t* dMP = mp_desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : idx) {
            if (cur_mp[i2] >= 0 && mp_obs_pos[cur_mp[i2]]) continue;
            const int dist = descriptor_distance(dMP, &C.desc[(size_t)i2 * 32]);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            cur_mp[bestIdx2] = i;
            nmatches++;
            if (checkOri) {
                float rot = last_angle[i] - C.kps[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1, sizes[HISTO_LENGTH];
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j : rotHist[i]) { cur_mp[j] = -1; nmatches--; }
    }
    return nmatches;
}

/* SearchByProjection(F, vpMapPoints, th, bFarPoints = false): the caller lists the map points
 * with mbTrackInView set (and not bad), in vpMapPoints order, with the fields isInFrustum stored.
 * cur_blocked[i] != 0: keypoint i already holds a map point with Observations() > 0.
 * cur_mp[cur.n] (in/out): -1 or the index (into this call's arrays) of the map point assigned here. */
int trko_search_by_projection_map(void* fcur, int M, const float* projX, const float* projY, const int* level,
                                  const float* viewCos, const uint8_t* mp_desc, const uint8_t* mp_obs_pos, float th,
                                  float nnratio, const uint8_t* cur_blocked, int* cur_mp)
{
    Frame& F = *(Frame*)fcur;
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> idx;
    for (int i = 0; i < F.n; i++) cur_mp[i] = -1;
    for (int m = 0; m < M; m++) {
        const int nPredictedLevel = level[m];
        float r = viewCos[m] > 0.998 ? 2.5f : 4.0f;
        if (bFactor) r *= th;
        features_in_area(F, projX[m], projY[m], r * F.scaleFactors[nPredictedLevel], nPredictedLevel - 1,
                         nPredictedLevel, idx);
        if (idx.empty()) continue;
        const uint8_t* d0 = mp_desc + (size_t)m * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int i2 : idx) {
            if (cur_blocked[i2]) continue;
            if (cur_mp[i2] >= 0 && mp_obs_pos[cur_mp[i2]]) continue;
            const int dist = descriptor_distance(d0, &F.desc[(size_t)i2 * 32]);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = F.kps[i2].octave;
                bestIdx = i2;
            } else if (dist < bestDist2) {
                bestLevel2 = F.kps[i2].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            if (bestLevel != bestLevel2 || bestDist <= nnratio * bestDist2) {
                cur_mp[bestIdx] = m;
                nmatches++;
            }
        }
    }
    return nmatches;
}

/* Frame::isInFrustum (mono branch, O3/src/Frame.cc:576-636) + MapPoint::PredictScale (O3/src/MapPoint.cc:573-587)
 * over a batch.  The pose (q, t) is the frame's SE3f as stored; Frame::UpdatePoseMatrices (:553-559) derives
 * mRcw = toRotationMatrix(q), mtcw = t and mOw = translation of Tcw.inverse() (Sophus: normalised conjugate quaternion
 * applied to -t).  Pc = mRcw * P + mtcw is Eigen's coefficient product (x0 + (x1 + x2)), PO.norm() / PO.dot(Pn) use the
 * same 3-term order; min_dist / max_dist are mfMinDistance / mfMaxDistance. */
void trko_is_in_frustum(const float* q, const float* t, const float* K, const float* bounds, int nlevels,
                        float scaleFactor, int m, const float* xw, const float* normal, const float* min_dist,
                        const float* max_dist, const uint8_t* skip, float cosLimit, uint8_t* in_view, float* px,
                        float* py, int* level, float* viewCos)
{
    float R[9], qi[4], Ow[3];
    so::quat_to_matrix(q, R);
    so::se3_inverse(q, t, qi, Ow);
    const float logScale = std::log(scaleFactor);   /* Frame.cc:399 mfLogScaleFactor = log(mfScaleFactor) on floats */
    for (int k = 0; k < m; k++) {
        in_view[k] = 0; px[k] = -1; py[k] = -1; level[k] = -1; viewCos[k] = 0;
        if (skip && skip[k]) continue;
        const float* P = xw + 3 * k;
        float Pc[3];
        so::mat_vec(R, P, Pc);
        for (int i = 0; i < 3; i++) Pc[i] = Pc[i] + t[i];
        if (Pc[2] < 0.0f) continue;
        const float u = K[0] * Pc[0] / Pc[2] + K[2];
        const float v = K[1] * Pc[1] / Pc[2] + K[3];
        if (u < bounds[0] || u > bounds[2]) continue;
        if (v < bounds[1] || v > bounds[3]) continue;
        px[k] = u; py[k] = v;
        const float maxD = 1.2f * max_dist[k], minD = 0.8f * min_dist[k];
        const float PO[3] = { P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2] };
        const float dist = so::norm3(PO);
        if (dist < minD || dist > maxD) continue;
        const float vc = so::dot3(PO, normal + 3 * k) / dist;
        if (vc < cosLimit) continue;
        in_view[k] = 1; level[k] = so::predict_scale(max_dist[k], dist, logScale, nlevels); viewCos[k] = vc;
    }
}

/* Optimizer::PoseOptimization, mono observations only.
 * pose_q = (x,y,z,w) of Tcw.unit_quaternion(), pose_t = Tcw.translation() (float, in/out);
 * K = fx,fy,cx,cy (float); per correspondence: world point (float), undistorted keypoint (float),
 * mvInvLevelSigma2[octave] (float).  outlier[n] receives mvbOutlier.  Returns
 * nInitialCorrespondences - nBad.  stats (optional): [0] = total LM iterations, [1] = total trials. */
int trko_pose_optimization(float* pose_q, float* pose_t, const float* K, int n, const float* Xw, const float* kp_xy,
                           const float* inv_sigma2, uint8_t* outlier, int* stats)
{
    if (stats) { stats[0] = 0; stats[1] = 0; }
    for (int i = 0; i < n; i++) outlier[i] = 0;
    if (n < 3) return 0;
    std::vector<double> X(3 * n), O(2 * n), W(n);
    for (int i = 0; i < 3 * n; i++) X[i] = Xw[i];
    for (int i = 0; i < 2 * n; i++) O[i] = kp_xy[i];
    for (int i = 0; i < n; i++) W[i] = inv_sigma2[i];
    PoseProblem P;
    P.n = n; P.Xw = X.data(); P.obs = O.data(); P.info = W.data();
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    const float deltaMono = (float)std::sqrt(5.991);
    P.delta = deltaMono; P.dsqr = P.delta * P.delta;
    P.level.assign(n, 0); P.robust.assign(n, 1); P.err.assign(2 * n, 0.0);
    SE3 T0;
    T0.r = { pose_q[0], pose_q[1], pose_q[2], pose_q[3] };
    T0.t[0] = pose_t[0]; T0.t[1] = pose_t[1]; T0.t[2] = pose_t[2];
    quat_normalize(T0.r);
    SE3 T = T0;
    const float chi2Mono[4] = { 5.991f, 5.991f, 5.991f, 5.991f };
    int nBad = 0;
    for (int it = 0; it < 4; it++) {
        T = T0; /* the frame's pose is only written back at the end: every round restarts from it */
        int trials = 0;
        int iters = optimize_pose(P, T, 10, &trials);
        if (stats && iters > 0) { stats[0] += iters; stats[1] += trials; }
        nBad = 0;
        for (int k = 0; k < n; k++) {
            if (outlier[k]) { double xc[3]; edge_error(P, T, k, &P.err[2 * k], xc); }
            const float chi2 = (float)edge_chi2(P, k);
            if (chi2 > chi2Mono[it]) { outlier[k] = 1; P.level[k] = 1; nBad++; }
            else { outlier[k] = 0; P.level[k] = 0; }
            if (it == 2) P.robust[k] = 0;
        }
        if (n < 10) break;
    }
    /* pFrame->SetPose(Sophus::SE3<float>(rotation().cast<float>(), translation().cast<float>())), Optimizer.cc:1021-1024:
     * the SE3f constructor normalises the float quaternion */
    pose_q[0] = (float)T.r.x; pose_q[1] = (float)T.r.y; pose_q[2] = (float)T.r.z; pose_q[3] = (float)T.r.w;
    so::quat_normalize(pose_q);
    pose_t[0] = (float)T.t[0]; pose_t[1] = (float)T.t[1]; pose_t[2] = (float)T.t[2];
    return n - nBad;
}

/* SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  (O3/src/ORBmatcher.cc:605-707).
 * f2 = F2 (grid queries); F1 is given by its undistorted keypoints and descriptors.  prev_matched[n1*2]
 * (in/out) = vbPrevMatched; matches12[n1] receives vnMatches12.  Returns nmatches. */
int trko_search_for_initialization(int n1, const void* kps1_, const uint8_t* desc1, void* f2, float* prev_matched,
                                   int windowSize, float nnratio, int checkOri, int* matches12)
{
    const KeyPt* kps1 = (const KeyPt*)kps1_;
    Frame& F2 = *(Frame*)f2;
    int nmatches = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vMatchedDistance(F2.n, std::numeric_limits<int>::max());
    std::vector<int> vnMatches21(F2.n, -1);
    std::vector<int> idx;
    for (int i1 = 0; i1 < n1; i1++) {
        const int level1 = kps1[i1].octave;
        if (level1 > 0) continue;
        features_in_area(F2, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)windowSize, level1, level1, idx);
        if (idx.empty()) continue;
        const uint8_t* d1 = desc1 + (size_t)i1 * 32;
        int bestDist = std::numeric_limits<int>::max(), bestDist2 = std::numeric_limits<int>::max(), bestIdx2 = -1;
        for (int i2 : idx) {
            const int dist = descriptor_distance(d1, &F2.desc[(size_t)i2 * 32]);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) { bestDist2 = dist; }
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { matches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOri) {
                    float rot = kps1[i1].angle - F2.kps[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1, sizes[HISTO_LENGTH];
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j : rotHist[i])
                    if (matches12[j] >= 0) { matches12[j] = -1; nmatches--; }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (matches12[i1] >= 0) {
            prev_matched[2 * i1] = F2.kps[matches12[i1]].x;
            prev_matched[2 * i1 + 1] = F2.kps[matches12[i1]].y;
        }
    return nmatches;
}

/* The search half of ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight = false)  (O3/src/ORBmatcher.cc:1060-1228): per map
 * point the keyframe keypoint it would be fused with (best_idx, best_dist; -1 when a gate rejects it or bestDist >
 * TH_LOW).  The search does not read the keyframe's map points, so it is independent of the side effects
 * (Replace / AddObservation / AddMapPoint, :1209-1222), which the caller applies afterwards in vpMapPoints order.
 * fkf = the keyframe's keypoints/descriptors/grid (KeyFrame::GetFeaturesInArea has no level filter).
 * skip[i]: !pMP || isBad() || IsInKeyFrame(pKF).  min_dist / max_dist = mfMinDistance / mfMaxDistance.
 * The pose (q, t) is the keyframe's SE3f as stored: p3Dc = Tcw * p3Dw is Sophus' quaternion action, Ow =
 * pKF->GetCameraCenter() = translation of Tcw.inverse() (KeyFrame.cc:224-257). */
void trko_fuse_search(void* fkf, const float* q, const float* t, const float* K, int nlevels, float logScaleFactor,
                      const float* invLevelSigma2, int m, const float* xw, const float* normal, const float* min_dist,
                      const float* max_dist, const uint8_t* mp_desc, const uint8_t* skip, float th, int* best_idx,
                      int* best_dist)
{
    Frame& F = *(Frame*)fkf;
    float qi[4], Ow[3];
    so::se3_inverse(q, t, qi, Ow);
    std::vector<int> idx;
    for (int i = 0; i < m; i++) {
        best_idx[i] = -1; best_dist[i] = 256;
        if (skip && skip[i]) continue;
        const float* P = xw + 3 * i;
        float pc[3];
        so::se3_apply(q, t, P, pc);
        if (pc[2] < 0.0f) continue;
        const float u = K[0] * pc[0] / pc[2] + K[2], v = K[1] * pc[1] / pc[2] + K[3];     /* Pinhole::project */
        if (!(u >= F.minX && u < F.maxX && v >= F.minY && v < F.maxY)) continue;          /* KeyFrame::IsInImage */
        const float maxDistance = 1.2f * max_dist[i], minDistance = 0.8f * min_dist[i];
        const float PO[3] = { P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2] };
        const float dist3D = so::norm3(PO);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        if (so::dot3(PO, normal + 3 * i) < 0.5 * dist3D) continue;
        const int nPredictedLevel = so::predict_scale(max_dist[i], dist3D, logScaleFactor, nlevels);   /* PredictScale */
        const float radius = th * F.scaleFactors[nPredictedLevel];
        features_in_area(F, u, v, radius, -1, -1, idx);
        if (idx.empty()) continue;
        const uint8_t* dMP = mp_desc + (size_t)i * 32;
        int bestDist = 256, bestIdx = -1;
        for (int k : idx) {
            const KeyPt& kp = F.kps[k];
            if (kp.octave < nPredictedLevel - 1 || kp.octave > nPredictedLevel) continue;
            const float ex = u - kp.x, ey = v - kp.y;
            const float e2 = ex * ex + ey * ey;
            if (e2 * invLevelSigma2[kp.octave] > 5.99) continue;
            const int dist = descriptor_distance(dMP, &F.desc[(size_t)k * 32]);
            if (dist < bestDist) { bestDist = dist; bestIdx = k; }
        }
        if (bestDist <= TH_LOW) { best_idx[i] = bestIdx; best_dist[i] = bestDist; }
    }
}


/* ---- Sim3-guided matchers of loop closing / map merging (LoopClosing::DetectCommonRegionsFromBoW, O3/src/LoopClosing.cc:823-847) ----
 * Common to all: the candidate map point is projected into a keyframe (its keypoints / grid = fkf), gated by depth,
 * image bounds (KeyFrame::IsInImage), the scale-invariance range and -- except in SearchBySim3 -- the viewing angle; the
 * keypoints in a window of radius th * scale[predicted level] with octave in [level - 1, level] are compared and the first
 * of the nearest wins.  min_dist / max_dist are mfMinDistance / mfMaxDistance. */
namespace {
struct ProjGate { float u, v; int level; };
/* the gates of O3/src/ORBmatcher.cc:421-454 (= :524-563, :1264-1297): returns false when the point is rejected */
bool sim3_project_gate(const Frame& F, const float q[4], const float t[3], const float Ow[3], const float* K, int nlevels,
                       float logScale, const float* P, const float* Pn, float minD, float maxD, ProjGate& g)
{
    float pc[3];
    so::se3_apply(q, t, P, pc);
    if (pc[2] < 0.0f) return false;
    g.u = K[0] * pc[0] / pc[2] + K[2];
    g.v = K[1] * pc[1] / pc[2] + K[3];
    if (!(g.u >= F.minX && g.u < F.maxX && g.v >= F.minY && g.v < F.maxY)) return false;
    const float maxDistance = 1.2f * maxD, minDistance = 0.8f * minD;
    const float PO[3] = { P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2] };
    const float dist = so::norm3(PO);
    if (dist < minDistance || dist > maxDistance) return false;
    if (so::dot3(PO, Pn) < 0.5 * dist) return false;
    g.level = so::predict_scale(maxD, dist, logScale, nlevels);
    return true;
}
} // namespace

/* int ORBmatcher::SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints,
 * vector<MapPoint*>& vpMatched, int th, float ratioHamming)  (:395-494; the overload with vpPointsKFs, :496-603, matches
 * identically).  sq / st: Scw.quaternion() (x, y, z, w; |q|^2 = scale) and Scw.translation().  skip[i]: isBad() or already
 * in vpMatched.  kp_matched[kf n]: in = vpMatched[k] != NULL, out = unchanged for those, else the index of the candidate
 * now held (-1 none) in kp_point.  Sequential: a keypoint taken by an earlier candidate is not available to later ones. */
int trko_search_by_projection_sim3(void* fkf, const float* sq, const float* st, const float* K, int nlevels, float logScale,
                                   int m, const float* xw, const float* normal, const float* min_dist, const float* max_dist,
                                   const uint8_t* mp_desc, const uint8_t* skip, const uint8_t* kp_matched, int th,
                                   float ratioHamming, int* kp_point)
{
    Frame& F = *(Frame*)fkf;
    float q[4], t[3], qi[4], Ow[3];
    so::sim3_to_se3(sq, st, q, t);
    so::se3_inverse(q, t, qi, Ow);
    std::vector<uint8_t> taken(kp_matched, kp_matched + F.n);
    for (int k = 0; k < F.n; k++) kp_point[k] = -1;
    int nmatches = 0;
    std::vector<int> idx;
    for (int i = 0; i < m; i++) {
        if (skip && skip[i]) continue;
        ProjGate g;
        if (!sim3_project_gate(F, q, t, Ow, K, nlevels, logScale, xw + 3 * i, normal + 3 * i, min_dist[i], max_dist[i], g)) continue;
        const float radius = th * F.scaleFactors[g.level];
        features_in_area(F, g.u, g.v, radius, -1, -1, idx);
        if (idx.empty()) continue;
        const uint8_t* dMP = mp_desc + (size_t)i * 32;
        int bestDist = 256, bestIdx = -1;
        for (int k : idx) {
            if (taken[k]) continue;
            const int lv = F.kps[k].octave;
            if (lv < g.level - 1 || lv > g.level) continue;
            const int dist = descriptor_distance(dMP, &F.desc[(size_t)k * 32]);
            if (dist < bestDist) { bestDist = dist; bestIdx = k; }
        }
        if (bestDist <= TH_LOW * ratioHamming) { taken[bestIdx] = 1; kp_point[bestIdx] = i; nmatches++; }
    }
    return nmatches;
}

/* The search half of int ORBmatcher::Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints, float th,
 * vector<MapPoint*>& vpReplacePoint)  (:1236-1345): per candidate the keypoint it is fused with (-1 none).  skip[i]:
 * isBad() or already among pKF->GetMapPoints().  No chi-square gate here (unlike Fuse(pKF, vpMapPoints, th)). */
void trko_fuse_search_sim3(void* fkf, const float* sq, const float* st, const float* K, int nlevels, float logScale, int m,
                           const float* xw, const float* normal, const float* min_dist, const float* max_dist,
                           const uint8_t* mp_desc, const uint8_t* skip, float th, int* best_idx, int* best_dist)
{
    Frame& F = *(Frame*)fkf;
    float q[4], t[3], qi[4], Ow[3];
    so::sim3_to_se3(sq, st, q, t);
    so::se3_inverse(q, t, qi, Ow);
    std::vector<int> idx;
    for (int i = 0; i < m; i++) {
        best_idx[i] = -1; best_dist[i] = 256;
        if (skip && skip[i]) continue;
        ProjGate g;
        if (!sim3_project_gate(F, q, t, Ow, K, nlevels, logScale, xw + 3 * i, normal + 3 * i, min_dist[i], max_dist[i], g)) continue;
        const float radius = th * F.scaleFactors[g.level];
        features_in_area(F, g.u, g.v, radius, -1, -1, idx);
        if (idx.empty()) continue;
        const uint8_t* dMP = mp_desc + (size_t)i * 32;
        int bestDist = INT_MAX, bestIdx = -1;
        for (int k : idx) {
            const int lv = F.kps[k].octave;
            if (lv < g.level - 1 || lv > g.level) continue;
            const int dist = descriptor_distance(dMP, &F.desc[(size_t)k * 32]);
            if (dist < bestDist) { bestDist = dist; bestIdx = k; }
        }
        if (bestDist <= TH_LOW) { best_idx[i] = bestIdx; best_dist[i] = bestDist; }
    }
}

/* int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const Sophus::Sim3f& S12,
 * float th)  (:1347-1551).  Both directions project with pKF1's intrinsics (:1349-1352, the reference's choice) in the
 * u = fx * (x * invz) + cx form with invz = 1.0 / z taken in double; the distance is the norm of the point in the target
 * camera frame; no viewing-angle gate; TH_HIGH; a pair is kept when the two directions agree.
 * Side s (1, 2): n_s keypoints of keyframe s with the map point behind each (skip_s[i]: no map point, bad, or already
 * matched -- vbAlreadyMatched), its world position, distance range and descriptor.  match12[n1]: keypoint of keyframe 2
 * whose map point feature i1 is matched with (-1 none).  Returns nFound. */
int trko_search_by_sim3(void* f1, void* f2, const float* q1, const float* t1, const float* q2, const float* t2,
                        const float* s12q, const float* s12t, const float* K, int nlevels, float logScale,
                        const uint8_t* skip1, const float* xw1, const float* min1, const float* max1, const uint8_t* desc1,
                        const uint8_t* skip2, const float* xw2, const float* min2, const float* max2, const uint8_t* desc2,
                        float th, int* match12)
{
    Frame &F1 = *(Frame*)f1, &F2 = *(Frame*)f2;
    float s21q[4], s21t[3];
    so::sim3_inverse(s12q, s12t, s21q, s21t);
    auto direction = [&](const Frame& Fsrc, const Frame& Fdst, const float* qs, const float* ts, const float* sq, const float* st,
                         const uint8_t* skip, const float* xw, const float* mind, const float* maxd, const uint8_t* desc,
                         std::vector<int>& out) {
        out.assign(Fsrc.n, -1);
        std::vector<int> idx;
        for (int i = 0; i < Fsrc.n; i++) {
            if (skip[i]) continue;
            float pa[3], pb[3];
            so::se3_apply(qs, ts, xw + 3 * i, pa);
            so::sim3_apply(sq, st, pa, pb);
            if (pb[2] < 0.0) continue;
            const float invz = (float)(1.0 / pb[2]);
            const float x = pb[0] * invz, y = pb[1] * invz;
            const float u = K[0] * x + K[2], v = K[1] * y + K[3];
            if (!(u >= Fdst.minX && u < Fdst.maxX && v >= Fdst.minY && v < Fdst.maxY)) continue;
            const float maxDistance = 1.2f * maxd[i], minDistance = 0.8f * mind[i];
            const float dist3D = so::norm3(pb);
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            const int level = so::predict_scale(maxd[i], dist3D, logScale, nlevels);
            const float radius = th * Fdst.scaleFactors[level];
            features_in_area(Fdst, u, v, radius, -1, -1, idx);
            if (idx.empty()) continue;
            int bestDist = INT_MAX, bestIdx = -1;
            for (int k : idx) {
                const int lv = Fdst.kps[k].octave;
                if (lv < level - 1 || lv > level) continue;
                const int dist = descriptor_distance(desc + (size_t)i * 32, &Fdst.desc[(size_t)k * 32]);
                if (dist < bestDist) { bestDist = dist; bestIdx = k; }
            }
            if (bestDist <= TH_HIGH) out[i] = bestIdx;
        }
    };
    std::vector<int> m1, m2;
    direction(F1, F2, q1, t1, s21q, s21t, skip1, xw1, min1, max1, desc1, m1);
    direction(F2, F1, q2, t2, s12q, s12t, skip2, xw2, min2, max2, desc2, m2);
    int nFound = 0;
    for (int i1 = 0; i1 < F1.n; i1++) {
        match12[i1] = -1;
        const int idx2 = m1[i1];
        if (idx2 >= 0 && m2[idx2] == i1) { match12[i1] = idx2; nFound++; }
    }
    return nFound;
}

/* The constant-velocity prior of Tracking::TrackWithMotionModel in the reference's float32 Sophus arithmetic:
 * mVelocity = mCurrentFrame.GetPose() * mLastFrame.GetPose().inverse() once a frame is tracked (O3/src/Tracking.cc:1990-1991)
 * and mCurrentFrame.SetPose(mVelocity * mLastFrame.GetPose()) for the next one (:2598).  last / prev = the two most recent
 * poses (qx,qy,qz,qw,tx,ty,tz); prior receives the seven floats of the predicted pose. */
void trko_velocity_prior(const float* last, const float* prev, float* prior)
{
    float qpi[4], tpi[3], qv[4], tv[3];
    so::se3_inverse(prev, prev + 4, qpi, tpi);
    so::se3_mul(last, last + 4, qpi, tpi, qv, tv);
    so::se3_mul(qv, tv, last, last + 4, prior, prior + 4);
}

} // extern "C"
