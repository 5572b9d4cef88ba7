/*
 * oracle/lba_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (float64, flat arrays) of the reference's local bundle adjustment
 *   Optimizer::LocalBundleAdjustment          O3/src/Optimizer.cc:1030-1387 (graph, optimize(10), outlier test, write-back)
 * on g2o's machinery (g2o/ = O3/Thirdparty/g2o/g2o/):
 *   EdgeSE3ProjectXYZ::computeError/linearizeOplus  O3/include/OptimizableTypes.h:98-109, O3/src/OptimizableTypes.cpp:136-155
 *   BaseBinaryEdge::constructQuadraticForm          g2o/core/base_binary_edge.hpp:55-120
 *   RobustKernelHuber::robustify                    g2o/core/robust_kernel_impl.cpp:68-81
 *   BlockSolver<6,3>::buildSystem/setLambda/solve   g2o/core/block_solver.hpp:502-604, 354-486 (Schur complement)
 *   OptimizationAlgorithmLevenberg::solve           g2o/core/optimization_algorithm_levenberg.cpp:59-188
 *   SparseOptimizer::optimize                       g2o/core/sparse_optimizer.cpp:349-413
 *   VertexSE3Expmap / VertexSBAPointXYZ oplus       g2o/types/types_six_dof_expmap.h:71-74, g2o/types/types_sba.h:49-52
 *
 * Parity status: PINNED to the reference source.  The reference has no tests or fixtures for this path, so its own
 * Optimizer::LocalBundleAdjustment (local mapping and welding BA) and BundleAdjustment, its edge types and its vendored g2o
 * are compiled unmodified over a mini Eigen into oracle/_ref/libref_opt.so (oracle/Makefile `ref`, oracle/g2oshim) and
 * tests/test_ref_optimizer.py requires float32-identical poses, points within 2.4e-7 m and identical erased-observation
 * sets up to C4 size.  Differences by construction, all at rounding level: the reduced camera system
 * is solved with an unpivoted dense LDL^T (reference: Eigen::SimplicialLDLT with AMD ordering);
 * summation order over edges is the input order (reference: allocation/id order).  Self-checks in
 * tests/test_lba_oracle.py: finite-difference Jacobians, zero-noise convergence, chi2 monotonicity.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct Quat { double x, y, z, w; };
struct SE3 { Quat r; double t[3]; };

void quat_normalize(Quat& q)
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
    double uv[3] = { q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0] };
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
    out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
    out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}
void quat_to_matrix(const Quat& q, double R[9]) /* Eigen toRotationMatrix */
{
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
Quat quat_from_matrix(const double R[9])
{
    Quat q;
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[7] - R[5]) * t; q.y = (R[2] - R[6]) * t; q.z = (R[3] - R[1]) * t;
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
SE3 se3_exp(const double u[6])
{
    const double w0 = u[0], w1 = u[1], w2 = u[2];
    const double theta = std::sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    const double O[9] = { 0, -w2, w1, w2, 0, -w0, -w1, w0, 0 };
    double O2[9], R[9], V[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    if (theta < 0.00001) {
        for (int i = 0; i < 9; i++) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
    } else {
        const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
        const double c = (theta - std::sin(theta)) / std::pow(theta, 3);
        for (int i = 0; i < 9; i++) {
            const double I = (i % 4 == 0) ? 1.0 : 0.0;
            R[i] = I + a * O[i] + b * O2[i];
            V[i] = I + b * O[i] + c * O2[i];
        }
    }
    SE3 T;
    T.r = quat_from_matrix(R);
    quat_normalize(T.r);
    for (int i = 0; i < 3; i++) T.t[i] = V[i * 3] * u[3] + V[i * 3 + 1] * u[4] + V[i * 3 + 2] * u[5];
    return T;
}
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

bool ldlt_solve(int n, std::vector<double>& A /* destroyed */, const double* b, double* x)
{
    /* in-place unpivoted LDL^T on the lower triangle (row-major full storage) */
    std::vector<double> D(n);
    for (int j = 0; j < n; j++) {
        double d = A[(size_t)j * n + j];
        for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k] * D[k];
        if (!(d > 0) || !std::isfinite(d)) return false;
        D[j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[(size_t)i * n + j];
            for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k] * D[k];
            A[(size_t)i * n + j] = s / d;
        }
    }
    std::vector<double> y(n);
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= A[(size_t)i * n + k] * y[k];
        y[i] = s;
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = y[i] / D[i];
        for (int k = i + 1; k < n; k++) s -= A[(size_t)k * n + i] * x[k];
        x[i] = s;
    }
    return true;
}

void inv3(const double* D, double* Di) /* Eigen 3x3 inverse: cofactors / determinant */
{
    const double c00 = D[4] * D[8] - D[5] * D[7], c01 = D[5] * D[6] - D[3] * D[8], c02 = D[3] * D[7] - D[4] * D[6];
    const double det = D[0] * c00 + D[1] * c01 + D[2] * c02;
    const double id = 1.0 / det;
    Di[0] = c00 * id; Di[1] = (D[2] * D[7] - D[1] * D[8]) * id; Di[2] = (D[1] * D[5] - D[2] * D[4]) * id;
    Di[3] = c01 * id; Di[4] = (D[0] * D[8] - D[2] * D[6]) * id; Di[5] = (D[2] * D[3] - D[0] * D[5]) * id;
    Di[6] = c02 * id; Di[7] = (D[1] * D[6] - D[0] * D[7]) * id; Di[8] = (D[0] * D[4] - D[1] * D[3]) * id;
}

struct Problem {
    int nc, np, ne;
    std::vector<SE3> cam;
    std::vector<uint8_t> fixed;
    std::vector<int> cam_col; /* free-camera index or -1 */
    int nfree;
    std::vector<double> pt;   /* np*3 */
    const int* ecam; const int* ept;
    std::vector<double> obs, info;
    double fx, fy, cx, cy, delta, dsqr;
    std::vector<double> camK; /* [nc][4]: every camera's own intrinsics (e->pCamera = pKFi->mpCamera, Optimizer.cc:1219) */
    std::vector<double> err; /* ne*2: the edges' _error */
    std::vector<uint8_t> active; /* level-0 edges (all of them unless the welding BA moved some to level 1) */
};

std::vector<float> g_next_cam_K;   /* lbao_set_camera_intrinsics: per-camera K for the next call (mirrors dvm_lba_set_camera_intrinsics) */

void edge_project(const Problem& P, const SE3& T, const double* X, double xc[3], double uv[2], int cam = -1)
{
    quat_rotate(T.r, X, xc);
    xc[0] += T.t[0]; xc[1] += T.t[1]; xc[2] += T.t[2];
    const bool per = cam >= 0 && !P.camK.empty();
    const double fx = per ? P.camK[4 * cam] : P.fx, fy = per ? P.camK[4 * cam + 1] : P.fy;
    const double cx = per ? P.camK[4 * cam + 2] : P.cx, cy = per ? P.camK[4 * cam + 3] : P.cy;
    uv[0] = fx * xc[0] / xc[2] + cx;
    uv[1] = fy * xc[1] / xc[2] + cy;
}
void compute_errors(Problem& P)
{
    for (int e = 0; e < P.ne; e++) {
        if (!P.active[e]) continue; /* computeActiveErrors: a level-1 edge keeps its last _error */
        double xc[3], uv[2];
        edge_project(P, P.cam[P.ecam[e]], &P.pt[3 * P.ept[e]], xc, uv, P.ecam[e]);
        P.err[2 * e] = P.obs[2 * e] - uv[0];
        P.err[2 * e + 1] = P.obs[2 * e + 1] - uv[1];
    }
}
double edge_chi2(const Problem& P, int e)
{
    const double* r = &P.err[2 * e];
    return r[0] * (P.info[e] * r[0]) + r[1] * (P.info[e] * r[1]);
}
double robust_chi2(const Problem& P)
{
    double chi = 0;
    for (int e = 0; e < P.ne; e++) {
        if (!P.active[e]) continue;
        const double c = edge_chi2(P, e);
        chi += c <= P.dsqr ? c : 2 * std::sqrt(c) * P.delta - P.dsqr;
    }
    return chi;
}

} // namespace

extern "C" {

/* Optimizer::LocalBundleAdjustment on a flattened local window.
 *   cam_q[nc*4] (x,y,z,w), cam_t[nc*3], cam_fixed[nc]: keyframe poses Tcw (float, in/out) -- local
 *     keyframes first or in any order; fixed = lFixedCameras or the map's initial keyframe;
 *   pts[np*3]: map point positions (float, in/out);
 *   edge_cam/edge_pt/edge_obs/edge_inv_sigma2: one mono observation per edge;
 *   K = fx, fy, cx, cy (float);  abort_flag may be NULL (pbStopFlag).
 * Outputs: edge_chi2[ne] (double), edge_bad[ne] (chi2 > 5.991 || depth <= 0 -> observation to
 * erase), stats[4] = {LM iterations run, LM trials, initial robust chi2, final robust chi2}.
 * Returns the number of iterations run, or -1 if nothing was optimised. */
/* huber_delta: the robust kernel's delta as the caller's float (LocalBundleAdjustment: sqrt(5.991), Optimizer.cc:1178;
 * BundleAdjustment / GlobalBundleAdjustemnt: sqrt(5.99), :122; bRobust == false: +infinity = no kernel). */
int lbao_bundle_adjustment(int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne,
                           const int* edge_cam, const int* edge_pt, const float* edge_obs, const float* edge_inv_sigma2,
                           const float* K, int iterations, float huber_delta, const volatile int* abort_flag,
                           double* edge_chi2_out, uint8_t* edge_bad, double* stats);

int lbao_local_ba(int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne,
                  const int* edge_cam, const int* edge_pt, const float* edge_obs, const float* edge_inv_sigma2,
                  const float* K, int iterations, const volatile int* abort_flag, double* edge_chi2_out,
                  uint8_t* edge_bad, double* stats)
{
    return lbao_bundle_adjustment(nc, cam_q, cam_t, cam_fixed, np, pts, ne, edge_cam, edge_pt, edge_obs, edge_inv_sigma2, K,
                                  iterations, (float)std::sqrt(5.991), abort_flag, edge_chi2_out, edge_bad, stats);
}

/* iterations2 > 0: the welding BA's second pass (see lbao_merge_ba); stats then has 6 entries */
static int run_ba(int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne,
                  const int* edge_cam, const int* edge_pt, const float* edge_obs, const float* edge_inv_sigma2,
                  const float* K, int iterations, float huber_delta, int iterations2, const volatile int* abort_flag,
                  double* edge_chi2_out, uint8_t* edge_bad, double* stats)
{
    Problem P;
    P.nc = nc; P.np = np; P.ne = ne;
    P.cam.resize(nc); P.fixed.assign(cam_fixed, cam_fixed + nc); P.cam_col.assign(nc, -1);
    P.nfree = 0;
    int nfixed = 0;
    for (int c = 0; c < nc; c++) {
        P.cam[c].r = { cam_q[4 * c], cam_q[4 * c + 1], cam_q[4 * c + 2], cam_q[4 * c + 3] };
        quat_normalize(P.cam[c].r);
        for (int i = 0; i < 3; i++) P.cam[c].t[i] = cam_t[3 * c + i];
        if (!cam_fixed[c]) P.cam_col[c] = P.nfree++;
        else nfixed++;
    }
    if (stats) { stats[0] = stats[1] = stats[2] = stats[3] = 0; if (iterations2 > 0) stats[4] = stats[5] = 0; }
    if (nfixed == 0) return -1;                 /* "LBA aborted": O3/src/Optimizer.cc:1088-1091 */
    if (abort_flag && *abort_flag) return -1;   /* :1306-1308 */
    if (ne == 0 || P.nfree + np == 0) return -1;
    P.pt.resize(3 * np);
    for (int i = 0; i < 3 * np; i++) P.pt[i] = pts[i];
    P.ecam = edge_cam; P.ept = edge_pt;
    P.obs.resize(2 * ne); P.info.resize(ne); P.err.assign(2 * ne, 0.0); P.active.assign(ne, 1);
    for (int i = 0; i < 2 * ne; i++) P.obs[i] = edge_obs[i];
    for (int i = 0; i < ne; i++) P.info[i] = edge_inv_sigma2[i];
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    if ((int)g_next_cam_K.size() == 4 * nc) P.camK.assign(g_next_cam_K.begin(), g_next_cam_K.end());
    g_next_cam_K.clear();
    P.delta = huber_delta; P.dsqr = P.delta * P.delta;

    const int nf = P.nfree, dimP = 6 * nf, dimL = 3 * np;
    std::vector<double> Hpp((size_t)nf * 36), bp(dimP), Hll((size_t)np * 9), bl(dimL), Hpl((size_t)ne * 18);
    std::vector<double> x(dimP + dimL), Hs((size_t)dimP * dimP), bs(dimP), Dinv((size_t)np * 9);
    double lambda = -1, ni = 2;
    int nBad = 0, done = 0, trials = 0;
    double first_chi = 0, last_chi = 0;
    auto terminate = [&]() { return abort_flag && *abort_flag; };

    const int nstages = iterations2 > 0 ? 2 : 1;
    for (int stage = 0; stage < nstages; stage++) {
    if (stage == 1) {
        /* O3/src/Optimizer.cc:3476-3521: unless the stop flag is up, edges with chi2 > 5.991 or non-positive depth go
         * to level 1, EVERY edge loses its robust kernel, initializeOptimization(0), optimize(10) */
        if (terminate()) break;
        if (stats) stats[4] = done;
        int excluded = 0;
        for (int e = 0; e < ne; e++) {
            double xc[3], uv[2];
            edge_project(P, P.cam[P.ecam[e]], &P.pt[3 * P.ept[e]], xc, uv);
            if (edge_chi2(P, e) > 5.991 || !(xc[2] > 0.0)) { P.active[e] = 0; excluded++; }
        }
        if (stats) stats[5] = excluded;
        P.delta = std::numeric_limits<double>::infinity(); P.dsqr = P.delta;
    }
    const int iters = stage == 0 ? iterations : iterations2;
    for (int it = 0; it < iters && !terminate(); it++) {
        compute_errors(P);
        double currentChi = robust_chi2(P);
        const double iniChi = currentChi;
        if (it == 0 && stage == 0) first_chi = currentChi;
        double tempChi;
        /* buildSystem */
        std::fill(Hpp.begin(), Hpp.end(), 0.0); std::fill(bp.begin(), bp.end(), 0.0);
        std::fill(Hll.begin(), Hll.end(), 0.0); std::fill(bl.begin(), bl.end(), 0.0);
        for (int e = 0; e < ne; e++) {
            if (!P.active[e]) continue;
            const int c = P.ecam[e], l = P.ept[e], col = P.cam_col[c];
            const SE3& T = P.cam[c];
            double xc[3], uv[2];
            edge_project(P, T, &P.pt[3 * l], xc, uv, c);
            const double X = xc[0], Y = xc[1], Z = xc[2];
            const double fxc = P.camK.empty() ? P.fx : P.camK[4 * c], fyc = P.camK.empty() ? P.fy : P.camK[4 * c + 1];
            /* projectJac = -pCamera->projectJac(xyz_trans) */
            const double pj[6] = { -(fxc / Z), -0.0, -(-fxc * X / (Z * Z)), -0.0, -(fyc / Z), -(-fyc * Y / (Z * Z)) };
            double R[9];
            quat_to_matrix(T.r, R);
            double A[6], B[12]; /* A: 2x3 wrt the point, B: 2x6 wrt the pose */
            for (int r = 0; r < 2; r++)
                for (int k = 0; k < 3; k++) A[r * 3 + k] = pj[r * 3] * R[k] + pj[r * 3 + 1] * R[3 + k] + pj[r * 3 + 2] * R[6 + k];
            const double Dv[18] = { 0, Z, -Y, 1, 0, 0, -Z, 0, X, 0, 1, 0, Y, -X, 0, 0, 0, 1 };
            for (int r = 0; r < 2; r++)
                for (int k = 0; k < 6; k++) B[r * 6 + k] = pj[r * 3] * Dv[k] + pj[r * 3 + 1] * Dv[6 + k] + pj[r * 3 + 2] * Dv[12 + k];
            const double om = P.info[e];
            const double chi = edge_chi2(P, e);
            const double w = chi <= P.dsqr ? 1.0 : P.delta / std::sqrt(chi);
            const double r0 = -om * P.err[2 * e] * w, r1 = -om * P.err[2 * e + 1] * w; /* omega_r *= rho[1] */
            const double wo = w * om;
            for (int a = 0; a < 3; a++) {
                bl[3 * l + a] += A[a] * r0 + A[3 + a] * r1;
                for (int b2 = 0; b2 < 3; b2++) Hll[(size_t)l * 9 + a * 3 + b2] += A[a] * wo * A[b2] + A[3 + a] * wo * A[3 + b2];
            }
            if (col >= 0) {
                for (int a = 0; a < 6; a++) {
                    bp[6 * col + a] += B[a] * r0 + B[6 + a] * r1;
                    for (int b2 = 0; b2 < 6; b2++) Hpp[(size_t)col * 36 + a * 6 + b2] += B[a] * wo * B[b2] + B[6 + a] * wo * B[6 + b2];
                    for (int b2 = 0; b2 < 3; b2++) Hpl[(size_t)e * 18 + a * 3 + b2] = B[a] * wo * A[b2] + B[6 + a] * wo * A[3 + b2];
                }
            }
        }
        if (it == 0) { /* computeLambdaInit over every non-fixed vertex */
            double mx = 0;
            for (int c = 0; c < nf; c++)
                for (int j = 0; j < 6; j++) mx = std::max(std::fabs(Hpp[(size_t)c * 36 + j * 7]), mx);
            for (int l = 0; l < np; l++)
                for (int j = 0; j < 3; j++) mx = std::max(std::fabs(Hll[(size_t)l * 9 + j * 4]), mx);
            lambda = 1e-5 * mx;
            ni = 2;
            nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const std::vector<SE3> cam_backup = P.cam;   /* push */
            const std::vector<double> pt_backup = P.pt;
            /* solve(): Schur complement with lambda on every diagonal */
            std::fill(Hs.begin(), Hs.end(), 0.0);
            for (int c = 0; c < nf; c++)
                for (int a = 0; a < 6; a++)
                    for (int b2 = 0; b2 < 6; b2++)
                        Hs[(size_t)(6 * c + a) * dimP + 6 * c + b2] = Hpp[(size_t)c * 36 + a * 6 + b2] + (a == b2 ? lambda : 0.0);
            for (int i = 0; i < dimP; i++) bs[i] = bp[i];
            /* per-landmark inverse */
            for (int l = 0; l < np; l++) {
                double D[9];
                for (int k = 0; k < 9; k++) D[k] = Hll[(size_t)l * 9 + k];
                D[0] += lambda; D[4] += lambda; D[8] += lambda;
                inv3(D, &Dinv[(size_t)l * 9]);
            }
            /* edges grouped by landmark: build the groups once per trial (cheap) */
            {
                std::vector<int> start(np + 1, 0), order(ne);
                for (int e = 0; e < ne; e++) start[P.ept[e] + 1]++;
                for (int l = 0; l < np; l++) start[l + 1] += start[l];
                std::vector<int> cur(start.begin(), start.end() - 1);
                for (int e = 0; e < ne; e++) order[cur[P.ept[e]]++] = e;
                for (int l = 0; l < np; l++) {
                    const double* Di = &Dinv[(size_t)l * 9];
                    double db[3];
                    for (int a = 0; a < 3; a++) db[a] = Di[a * 3] * bl[3 * l] + Di[a * 3 + 1] * bl[3 * l + 1] + Di[a * 3 + 2] * bl[3 * l + 2];
                    for (int s1 = start[l]; s1 < start[l + 1]; s1++) {
                        const int e1 = order[s1], c1 = P.cam_col[P.ecam[e1]];
                        if (c1 < 0 || !P.active[e1]) continue;
                        const double* B1 = &Hpl[(size_t)e1 * 18];
                        double BD[18];
                        for (int a = 0; a < 6; a++)
                            for (int b2 = 0; b2 < 3; b2++)
                                BD[a * 3 + b2] = B1[a * 3] * Di[b2] + B1[a * 3 + 1] * Di[3 + b2] + B1[a * 3 + 2] * Di[6 + b2];
                        for (int a = 0; a < 6; a++) bs[6 * c1 + a] -= B1[a * 3] * db[0] + B1[a * 3 + 1] * db[1] + B1[a * 3 + 2] * db[2];
                        for (int s2 = start[l]; s2 < start[l + 1]; s2++) {
                            const int e2 = order[s2], c2 = P.cam_col[P.ecam[e2]];
                            if (c2 < 0 || !P.active[e2]) continue;
                            const double* B2 = &Hpl[(size_t)e2 * 18];
                            for (int a = 0; a < 6; a++)
                                for (int b2 = 0; b2 < 6; b2++)
                                    Hs[(size_t)(6 * c1 + a) * dimP + 6 * c2 + b2] -=
                                        BD[a * 3] * B2[b2 * 3] + BD[a * 3 + 1] * B2[b2 * 3 + 1] + BD[a * 3 + 2] * B2[b2 * 3 + 2];
                        }
                    }
                }
            }
            bool ok2 = dimP == 0 ? true : ldlt_solve(dimP, Hs, bs.data(), x.data());
            if (ok2) {
                /* xl = Dinv (bl - Hpl^T xp) */
                std::vector<double> cl(bl);
                for (int e = 0; e < ne; e++) {
                    const int c1 = P.cam_col[P.ecam[e]], l = P.ept[e];
                    if (c1 < 0 || !P.active[e]) continue;
                    const double* B1 = &Hpl[(size_t)e * 18];
                    for (int b2 = 0; b2 < 3; b2++)
                        for (int a = 0; a < 6; a++) cl[3 * l + b2] -= B1[a * 3 + b2] * x[6 * c1 + a];
                }
                for (int l = 0; l < np; l++) {
                    const double* Di = &Dinv[(size_t)l * 9];
                    for (int a = 0; a < 3; a++)
                        x[dimP + 3 * l + a] = Di[a * 3] * cl[3 * l] + Di[a * 3 + 1] * cl[3 * l + 1] + Di[a * 3 + 2] * cl[3 * l + 2];
                }
            } else {
                std::fill(x.begin(), x.end(), 0.0);
            }
            /* update */
            for (int c = 0; c < nc; c++)
                if (P.cam_col[c] >= 0) P.cam[c] = se3_mul(se3_exp(&x[6 * P.cam_col[c]]), P.cam[c]);
            for (int i = 0; i < dimL; i++) P.pt[i] += x[dimP + i];
            compute_errors(P);
            tempChi = robust_chi2(P);
            if (!ok2) tempChi = std::numeric_limits<double>::max();
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < dimP; j++) scale += x[j] * (lambda * x[j] + bp[j]);
            for (int j = 0; j < dimL; j++) scale += x[dimP + j] * (lambda * x[dimP + j] + bl[j]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                lambda *= std::max(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                P.cam = cam_backup; /* pop */
                P.pt = pt_backup;
            }
            qmax++;
            trials++;
        } while (rho < 0 && qmax < 10 && !terminate());
        done++;
        last_chi = currentChi;
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
        else nBad = 0;
        if (nBad >= 3) break;
    }
    } /* stage */
    /* outlier test on the stored errors + depth at the final estimate (:1313-1329) */
    for (int e = 0; e < ne; e++) {
        const double chi = edge_chi2(P, e);
        double xc[3], uv[2];
        edge_project(P, P.cam[P.ecam[e]], &P.pt[3 * P.ept[e]], xc, uv);
        if (edge_chi2_out) edge_chi2_out[e] = chi;
        if (edge_bad) edge_bad[e] = (chi > 5.991 || !(xc[2] > 0.0)) ? 1 : 0;
    }
    for (int c = 0; c < nc; c++) {
        if (P.cam_col[c] < 0) continue; /* fixed poses are local keyframes only if it is the initial KF: written back unchanged */
        cam_q[4 * c] = (float)P.cam[c].r.x; cam_q[4 * c + 1] = (float)P.cam[c].r.y;
        cam_q[4 * c + 2] = (float)P.cam[c].r.z; cam_q[4 * c + 3] = (float)P.cam[c].r.w;
        for (int i = 0; i < 3; i++) cam_t[3 * c + i] = (float)P.cam[c].t[i];
    }
    for (int i = 0; i < dimL; i++) pts[i] = (float)P.pt[i];
    if (stats) { stats[0] = done; stats[1] = trials; stats[2] = first_chi; stats[3] = last_chi; }
    return done;
}

int lbao_bundle_adjustment(int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne,
                           const int* edge_cam, const int* edge_pt, const float* edge_obs, const float* edge_inv_sigma2,
                           const float* K, int iterations, float huber_delta, const volatile int* abort_flag,
                           double* edge_chi2_out, uint8_t* edge_bad, double* stats)
{
    return run_ba(nc, cam_q, cam_t, cam_fixed, np, pts, ne, edge_cam, edge_pt, edge_obs, edge_inv_sigma2, K, iterations,
                  huber_delta, 0, abort_flag, edge_chi2_out, edge_bad, stats);
}

/* The welding BA of a map merge: Optimizer::LocalBundleAdjustment(pMainKF, vpAdjustKF, vpFixedKF, pbStopFlag),
 * O3/src/Optimizer.cc:3257-3675 (mono observations).  optimize(5) with Huber delta (float)sqrt(5.99) (:3362); then,
 * unless the stop flag is up, edges with chi2 > 5.991 or non-positive depth are moved to level 1, every edge loses its
 * robust kernel and optimize(10) runs on the level-0 edges (:3476-3521).  The final test (:3530-3546) reads e->chi2()
 * of EVERY edge -- a level-1 edge still holds the error of the first pass -- and the depth at the final estimate.
 * stats[6] = {LM iterations of both passes, LM trials, initial robust chi2, final chi2 of the last pass, iterations of
 * the first pass, edges moved to level 1}. */
int lbao_merge_ba(int nc, float* cam_q, float* cam_t, const uint8_t* cam_fixed, int np, float* pts, int ne,
                  const int* edge_cam, const int* edge_pt, const float* edge_obs, const float* edge_inv_sigma2,
                  const float* K, const volatile int* abort_flag, double* edge_chi2_out, uint8_t* edge_bad, double* stats)
{
    return run_ba(nc, cam_q, cam_t, cam_fixed, np, pts, ne, edge_cam, edge_pt, edge_obs, edge_inv_sigma2, K, 5,
                  (float)std::sqrt(5.99), 10, abort_flag, edge_chi2_out, edge_bad, stats);
}

/* per-camera intrinsics camK[nc][4] for the NEXT lbao_bundle_adjustment / lbao_merge_ba call */
void lbao_set_camera_intrinsics(int nc, const float* camK) { g_next_cam_K.assign(camK, camK + (size_t)4 * nc); }

/* analytic Jacobians of one edge, for the finite-difference self-check */
void lbao_edge_jacobians(const float* q, const float* t, const float* X, const float* K, double* A /*2x3*/, double* B /*2x6*/)
{
    Problem P;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    SE3 T;
    T.r = { q[0], q[1], q[2], q[3] };
    quat_normalize(T.r);
    for (int i = 0; i < 3; i++) T.t[i] = t[i];
    const double Xd[3] = { X[0], X[1], X[2] };
    double xc[3], uv[2];
    edge_project(P, T, Xd, xc, uv);
    const double x = xc[0], y = xc[1], z = xc[2];
    const double pj[6] = { -(P.fx / z), 0, P.fx * x / (z * z), 0, -(P.fy / z), P.fy * y / (z * z) };
    double R[9];
    quat_to_matrix(T.r, R);
    for (int r = 0; r < 2; r++)
        for (int k = 0; k < 3; k++) A[r * 3 + k] = pj[r * 3] * R[k] + pj[r * 3 + 1] * R[3 + k] + pj[r * 3 + 2] * R[6 + k];
    const double Dv[18] = { 0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1 };
    for (int r = 0; r < 2; r++)
        for (int k = 0; k < 6; k++) B[r * 6 + k] = pj[r * 3] * Dv[k] + pj[r * 3 + 1] * Dv[6 + k] + pj[r * 3 + 2] * Dv[12 + k];
}

/* error of one edge after a manifold perturbation (pose: exp(d6) * T, point: X + d3) */
void lbao_edge_error_perturbed(const float* q, const float* t, const float* X, const float* K, const float* obs,
                               const double* d6, const double* d3, double* err)
{
    Problem P;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    SE3 T;
    T.r = { q[0], q[1], q[2], q[3] };
    quat_normalize(T.r);
    for (int i = 0; i < 3; i++) T.t[i] = t[i];
    T = se3_mul(se3_exp(d6), T);
    const double Xd[3] = { X[0] + d3[0], X[1] + d3[1], X[2] + d3[2] };
    double xc[3], uv[2];
    edge_project(P, T, Xd, xc, uv);
    err[0] = obs[0] - uv[0];
    err[1] = obs[1] - uv[1];
}

} // extern "C"
