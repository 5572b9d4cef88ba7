/* sim3_oracle.cpp -- CPU restatement of Optimizer::OptimizeSim3 -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Follows (reference = /root/reference/src/slam_system/orb_slam3, "O3"):
 *   O3/src/Optimizer.cc:1960-2212                 OptimizeSim3 (graph assembly is the caller's: this file takes the
 *                                                 correspondences that got an edge pair, i.e. the loop body after :2094)
 *   O3/include/OptimizableTypes.h:146-206         VertexSim3Expmap::oplusImpl, EdgeSim3ProjectXYZ / EdgeInverseSim3ProjectXYZ
 *   O3/Thirdparty/g2o/g2o/types/sim3.h            Sim3(update) exponential, map, inverse, operator*
 *   O3/Thirdparty/g2o/g2o/core/base_binary_edge.hpp:130-205   numeric Jacobian (the edges define no linearizeOplus):
 *                                                 central differences, delta 1e-9, through oplus on the vertex
 *   O3/Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:59-188, sparse_optimizer.cpp:349-413   LM
 *   O3/Thirdparty/g2o/g2o/solvers/linear_solver_dense.h   dense LDL^T of the 7x7 system
 *
 *   O3/src/Optimizer.cc:1389-1651, O3/Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h:93-117   the solve of
 *                                                 OptimizeEssentialGraph (EdgeSim3 pose graph; ORACLE ONLY so far)
 *
 * A property of the reference worth knowing (sim3.h, both in Sim3(update) and in log()): in the small-rotation branch
 * with |sigma| >= 1e-5 the coefficient B = ((0.5 sigma^2 - sigma + 1) s) / sigma^3 has no finite limit (it is ~1/sigma^3),
 * so W = A Omega + B Omega^2 + C I is dominated by B Omega^2 whenever the rotation is small but not zero (log(): below
 * 4.5e-3 rad).  log() then returns an upsilon that is blind to the two translation directions orthogonal to omega: after
 * the first LM iteration of an essential-graph optimisation (errors spread thinly over the edges, scales no longer 1) H is
 * singular to working precision, and with lambda = 1e-16 the next iteration's ten trials all fail and g2o terminates.
 * The restatement keeps the formula as it is: that behaviour IS the reference's.
 *
 * Parity status: PINNED to the reference source.  The reference holds no fixture for these functions, so its own
 * Optimizer::OptimizeSim3 and OptimizeEssentialGraph (bodies cut out of Optimizer.cc at build time) run on its vendored g2o
 * in oracle/_ref/libref_opt.so (oracle/Makefile `ref`, oracle/g2oshim); tests/test_ref_optimizer.py requires the same Sim3
 * (1e-6), inlier sets, iteration counts and corrected keyframe / map-point poses.  Self-checks in tests/test_sim3_oracle.py
 * (recovery of a planted Sim3, fixed-scale column, outlier removal, exp/log round trip, loop-error distribution). */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace {

struct Quat { double x, y, z, w; };
struct Sim3 { Quat r; double t[3]; double s; };

Quat quat_mul(const Quat& a, const Quat& b)
{
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
void quat_rotate(const Quat& q, const double v[3], double out[3]) /* Eigen: v + w*uv + qv x uv, uv = 2 qv x v */
{
    double uv[3] = { q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0] };
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    out[0] = v[0] + q.w * uv[0] + (q.y * uv[2] - q.z * uv[1]);
    out[1] = v[1] + q.w * uv[1] + (q.z * uv[0] - q.x * uv[2]);
    out[2] = v[2] + q.w * uv[2] + (q.x * uv[1] - q.y * uv[0]);
}
Quat quat_from_matrix(const double R[9]) /* Eigen::Quaterniond(Matrix3d) */
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
        const int j = (i + 1) % 3, k = (j + 1) % 3;
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

/* g2o::Sim3(const Vector7d& update): sim3.h */
Sim3 sim3_exp(const double u[7])
{
    const double w0 = u[0], w1 = u[1], w2 = u[2], sigma = u[6];
    const double theta = std::sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    const double O[9] = { 0, -w2, w1, w2, 0, -w0, -w1, w0, 0 };
    double O2[9], R[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    Sim3 S;
    S.s = std::exp(sigma);
    const double eps = 0.00001;
    double A, B, C;
    auto rot_small = [&]() { for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i]; };
    auto rot_full = [&]() {
        const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
        for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * O[i] + b * O2[i];
    };
    if (std::fabs(sigma) < eps) {
        C = 1;
        if (theta < eps) { A = 1. / 2.; B = 1. / 6.; rot_small(); }
        else {
            const double theta2 = theta * theta;
            A = (1 - std::cos(theta)) / theta2;
            B = (theta - std::sin(theta)) / (theta2 * theta);
            rot_full();
        }
    } else {
        C = (S.s - 1) / sigma;
        if (theta < eps) {
            const double sigma2 = sigma * sigma;
            A = ((sigma - 1) * S.s + 1) / sigma2;
            B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
            rot_small();
        } else {
            rot_full();
            const double a = S.s * std::sin(theta), b = S.s * std::cos(theta);
            const double theta2 = theta * theta, sigma2 = sigma * sigma, c = theta2 + sigma2;
            A = (a * sigma + (1 - b) * theta) / (theta * c);
            B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
        }
    }
    S.r = quat_from_matrix(R);
    for (int i = 0; i < 3; i++) {
        double acc = 0;
        for (int j = 0; j < 3; j++) acc += (A * O[i * 3 + j] + B * O2[i * 3 + j] + C * (i == j ? 1.0 : 0.0)) * u[3 + j];
        S.t[i] = acc;
    }
    return S;
}
Sim3 sim3_mul(const Sim3& a, const Sim3& b)
{
    Sim3 r;
    r.r = quat_mul(a.r, b.r);
    double rt[3];
    quat_rotate(a.r, b.t, rt);
    for (int i = 0; i < 3; i++) r.t[i] = a.s * rt[i] + a.t[i];
    r.s = a.s * b.s;
    return r;
}
Sim3 sim3_inverse(const Sim3& a)
{
    Sim3 r;
    r.r = { -a.r.x, -a.r.y, -a.r.z, a.r.w };
    const double v[3] = { (-1. / a.s) * a.t[0], (-1. / a.s) * a.t[1], (-1. / a.s) * a.t[2] };
    quat_rotate(r.r, v, r.t);
    r.s = 1. / a.s;
    return r;
}
void sim3_map(const Sim3& S, const double X[3], double out[3])
{
    double rx[3];
    quat_rotate(S.r, X, rx);
    for (int i = 0; i < 3; i++) out[i] = S.s * rx[i] + S.t[i];
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
/* W.lu().solve(t): 3x3 LU with partial pivoting (Eigen::PartialPivLU) */
void lu_solve3(const double W[9], const double t[3], double x[3])
{
    double A[9], b[3] = { t[0], t[1], t[2] };
    for (int i = 0; i < 9; i++) A[i] = W[i];
    int piv[3] = { 0, 1, 2 };
    for (int k = 0; k < 3; k++) {
        int best = k;
        for (int r = k + 1; r < 3; r++)
            if (std::fabs(A[piv[r] * 3 + k]) > std::fabs(A[piv[best] * 3 + k])) best = r;
        std::swap(piv[k], piv[best]);
        const double* pk = &A[piv[k] * 3];
        for (int r = k + 1; r < 3; r++) {
            double* pr = &A[piv[r] * 3];
            const double f = pr[k] / pk[k];
            pr[k] = f;
            for (int c = k + 1; c < 3; c++) pr[c] -= f * pk[c];
        }
    }
    double y[3];
    for (int k = 0; k < 3; k++) {
        double v = b[piv[k]];
        for (int c = 0; c < k; c++) v -= A[piv[k] * 3 + c] * y[c];
        y[k] = v;
    }
    for (int k = 2; k >= 0; k--) {
        double v = y[k];
        for (int c = k + 1; c < 3; c++) v -= A[piv[k] * 3 + c] * x[c];
        x[k] = v / A[piv[k] * 3 + k];
    }
}
/* g2o::Sim3::log(), sim3.h */
void sim3_log(const Sim3& S, double res[7])
{
    const double sigma = std::log(S.s);
    double R[9];
    quat_to_matrix(S.r, R);
    const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
    const double eps = 0.00001;
    const double dR[3] = { R[7] - R[5], R[2] - R[6], R[3] - R[1] };   /* deltaR(R) */
    double omega[3], A, B, C;
    auto set_omega = [&](double f) { for (int i = 0; i < 3; i++) omega[i] = f * dR[i]; };
    if (std::fabs(sigma) < eps) {
        C = 1;
        if (d > 1 - eps) { set_omega(0.5); A = 1. / 2.; B = 1. / 6.; }
        else {
            const double theta = std::acos(d), theta2 = theta * theta;
            set_omega(theta / (2 * std::sqrt(1 - d * d)));
            A = (1 - std::cos(theta)) / theta2;
            B = (theta - std::sin(theta)) / (theta2 * theta);
        }
    } else {
        C = (S.s - 1) / sigma;
        if (d > 1 - eps) {
            const double sigma2 = sigma * sigma;
            set_omega(0.5);
            A = ((sigma - 1) * S.s + 1) / sigma2;
            B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
        } else {
            const double theta = std::acos(d);
            set_omega(theta / (2 * std::sqrt(1 - d * d)));
            const double theta2 = theta * theta;
            const double a = S.s * std::sin(theta), b = S.s * std::cos(theta), c = theta2 + sigma * sigma;
            A = (a * sigma + (1 - b) * theta) / (theta * c);
            B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
        }
    }
    const double O[9] = { 0, -omega[2], omega[1], omega[2], 0, -omega[0], -omega[1], omega[0], 0 };
    double W[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const double O2 = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
            W[i * 3 + j] = A * O[i * 3 + j] + B * O2 + C * (i == j ? 1.0 : 0.0);
        }
    double ups[3];
    lu_solve3(W, S.t, ups);
    for (int i = 0; i < 3; i++) { res[i] = omega[i]; res[3 + i] = ups[i]; }
    res[6] = sigma;
}

struct Problem {
    int n;
    std::vector<double> p1, p2, obs1, obs2, w1, w2;   /* fixed point vertices (camera frames), measurements, information */
    double K1[4], K2[4];
    bool fix_scale;
    double delta, dsqr;
    std::vector<uint8_t> alive, robust;              /* edge pair still in the graph / still carries its Huber kernel */
    std::vector<double> err;                         /* [n][4]: e12 (2), e21 (2) */
};

void pair_error(const Problem& P, const Sim3& S, const Sim3& Sinv, int i, double e[4])
{
    double x[3];
    sim3_map(S, &P.p2[3 * i], x);                                       /* x1 = S12 * X2 through pCamera1 */
    e[0] = P.obs1[2 * i] - (P.K1[0] * x[0] / x[2] + P.K1[2]);
    e[1] = P.obs1[2 * i + 1] - (P.K1[1] * x[1] / x[2] + P.K1[3]);
    sim3_map(Sinv, &P.p1[3 * i], x);                                    /* x2 = S12^-1 * X1 through pCamera2 */
    e[2] = P.obs2[2 * i] - (P.K2[0] * x[0] / x[2] + P.K2[2]);
    e[3] = P.obs2[2 * i + 1] - (P.K2[1] * x[1] / x[2] + P.K2[3]);
}
void compute_errors(Problem& P, const Sim3& S)
{
    const Sim3 Sinv = sim3_inverse(S);
    for (int i = 0; i < P.n; i++)
        if (P.alive[i]) pair_error(P, S, Sinv, i, &P.err[4 * i]);
}
inline double chi12(const Problem& P, int i) { const double* e = &P.err[4 * i]; return e[0] * (P.w1[i] * e[0]) + e[1] * (P.w1[i] * e[1]); }
inline double chi21(const Problem& P, int i) { const double* e = &P.err[4 * i]; return e[2] * (P.w2[i] * e[2]) + e[3] * (P.w2[i] * e[3]); }
double rho0(const Problem& P, int i, double c) { return (!P.robust[i] || c <= P.dsqr) ? c : 2 * std::sqrt(c) * P.delta - P.dsqr; }
double rho1(const Problem& P, int i, double c) { return (!P.robust[i] || c <= P.dsqr) ? 1.0 : P.delta / std::sqrt(c); }
double active_chi2(const Problem& P)
{
    double chi = 0;
    for (int i = 0; i < P.n; i++)
        if (P.alive[i]) chi += rho0(P, i, chi12(P, i)) + rho0(P, i, chi21(P, i));
    return chi;
}
Sim3 oplus(const Problem& P, const Sim3& S, const double* update)
{
    double u[7];
    for (int k = 0; k < 7; k++) u[k] = update[k];
    if (P.fix_scale) u[6] = 0;
    return sim3_mul(sim3_exp(u), S);
}

/* Eigen::LDLT-equivalent solve of the symmetric 7x7 system (unpivoted LDL^T; the system is positive definite once
 * lambda is on the diagonal) */
bool solve7(const double* H, const double* b, double* x)
{
    double L[49] = { 0 }, D[7];
    for (int j = 0; j < 7; j++) {
        double d = H[j * 7 + j];
        for (int k = 0; k < j; k++) d -= L[j * 7 + k] * L[j * 7 + k] * D[k];
        if (!(d > 0) || !std::isfinite(d)) return false;   /* _cholesky.isPositive() */
        D[j] = d;
        L[j * 7 + j] = 1;
        for (int i = j + 1; i < 7; i++) {
            double v = H[i * 7 + j];
            for (int k = 0; k < j; k++) v -= L[i * 7 + k] * L[j * 7 + k] * D[k];
            L[i * 7 + j] = v / d;
        }
    }
    double y[7];
    for (int i = 0; i < 7; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= L[i * 7 + k] * y[k]; y[i] = v; }
    for (int i = 0; i < 7; i++) y[i] /= D[i];
    for (int i = 6; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < 7; k++) v -= L[k * 7 + i] * x[k]; x[i] = v; }
    return true;
}

/* optimizer.optimize(iterations) on the alive edge pairs; returns iterations run */
int optimize(Problem& P, Sim3& S, int iterations, int* trials_out, double* first_chi, double* last_chi)
{
    double lambda = -1, ni = 2;
    int nBad = 0, done = 0;
    for (int it = 0; it < iterations; it++) {
        compute_errors(P, S);
        double currentChi = active_chi2(P);
        const double iniChi = currentChi;
        if (it == 0 && first_chi) *first_chi = currentChi;
        /* buildSystem: numeric Jacobians w.r.t. the Sim3 vertex (the point vertices are fixed) */
        Sim3 Sp[7], Spi[7], Sm[7], Smi[7];
        const double delta = 1e-9, scalar = 1.0 / (2 * delta);
        for (int d = 0; d < 7; d++) {
            double add[7] = { 0, 0, 0, 0, 0, 0, 0 };
            add[d] = delta;
            Sp[d] = oplus(P, S, add); Spi[d] = sim3_inverse(Sp[d]);
            add[d] = -delta;
            Sm[d] = oplus(P, S, add); Smi[d] = sim3_inverse(Sm[d]);
        }
        double H[49] = { 0 }, b[7] = { 0 };
        for (int i = 0; i < P.n; i++) {
            if (!P.alive[i]) continue;
            double J[4][7];
            for (int d = 0; d < 7; d++) {
                double ep[4], em[4];
                pair_error(P, Sp[d], Spi[d], i, ep);
                pair_error(P, Sm[d], Smi[d], i, em);
                for (int r = 0; r < 4; r++) J[r][d] = scalar * (ep[r] - em[r]);
            }
            const double* e = &P.err[4 * i];
            for (int half = 0; half < 2; half++) {
                const double om = half == 0 ? P.w1[i] : P.w2[i];
                const double chi = half == 0 ? chi12(P, i) : chi21(P, i);
                const double w = rho1(P, i, chi);
                const double r0 = -om * e[2 * half] * w, r1 = -om * e[2 * half + 1] * w;
                const double wo = w * om;
                const double* J0 = J[2 * half];
                const double* J1 = J[2 * half + 1];
                for (int a = 0; a < 7; a++) {
                    b[a] += J0[a] * r0 + J1[a] * r1;
                    for (int c = 0; c < 7; c++) H[a * 7 + c] += J0[a] * wo * J0[c] + J1[a] * wo * J1[c];
                }
            }
        }
        if (it == 0) {
            double mx = 0;
            for (int j = 0; j < 7; j++) mx = std::max(std::fabs(H[j * 8]), mx);
            lambda = 1e-5 * mx;
            ni = 2;
            nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const Sim3 backup = S;
            double Hl[49], x[7];
            for (int k = 0; k < 49; k++) Hl[k] = H[k];
            for (int j = 0; j < 7; j++) Hl[j * 8] += lambda;
            const bool ok2 = solve7(Hl, b, x);
            if (!ok2) for (int j = 0; j < 7; j++) x[j] = 0;
            S = oplus(P, S, x);
            compute_errors(P, S);
            double tempChi = active_chi2(P);
            if (!ok2) tempChi = std::numeric_limits<double>::max();
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < 7; j++) scale += x[j] * (lambda * x[j] + b[j]);
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
                S = backup;
            }
            qmax++;
            if (trials_out) (*trials_out)++;
        } while (rho < 0 && qmax < 10);
        done++;
        if (last_chi) *last_chi = currentChi;
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
        else nBad = 0;
        if (nBad >= 3) break;
    }
    return done;
}

} // namespace

extern "C" {

/* Optimizer::OptimizeSim3 on the flattened correspondences that received an edge pair (:2094-2146):
 *   p1c[n*3], p2c[n*3]: map point positions in the camera frame of KF1 / KF2 as float (P3D1c, P3D2c);
 *   obs1[n*2]: kpUn1.pt; obs2[n*2]: kpUn2.pt, or the normalised projection of P3D2c when the point is not in KF2;
 *   inv_sigma2_1/2[n]: mvInvLevelSigma2 of the keypoints' octaves;  K1, K2: fx, fy, cx, cy of pCamera1 / pCamera2;
 *   s12_q (x,y,z,w), s12_t, s12_s: g2oS12, in/out (double);  th2, fix_scale as the reference.
 * inlier[n] (out): 0 where vpMatches1[idx] is reset (either pass).  Returns nIn; 0 when fewer than 10 correspondences
 * survive the first pass (g2oS12 is then left untouched, :2183-2184).
 * stats[6] = {LM iterations pass 1, pass 2, LM trials, nBad after pass 1, initial robust chi2, final chi2}. */
int sim3o_optimize_sim3(int n, const float* p1c, const float* p2c, const float* obs1, const float* obs2,
                        const float* inv_sigma2_1, const float* inv_sigma2_2, const float* K1, const float* K2, double* s12_q,
                        double* s12_t, double* s12_s, float th2, int fix_scale, uint8_t* inlier, double* stats)
{
    Problem P;
    P.n = n;
    P.p1.assign(p1c, p1c + 3 * n); P.p2.assign(p2c, p2c + 3 * n);
    P.obs1.assign(obs1, obs1 + 2 * n); P.obs2.assign(obs2, obs2 + 2 * n);
    P.w1.assign(inv_sigma2_1, inv_sigma2_1 + n); P.w2.assign(inv_sigma2_2, inv_sigma2_2 + n);
    for (int k = 0; k < 4; k++) { P.K1[k] = K1[k]; P.K2[k] = K2[k]; }
    P.fix_scale = fix_scale != 0;
    const float deltaHuber = std::sqrt(th2);                          /* const float deltaHuber = sqrt(th2), :1997 */
    P.delta = deltaHuber; P.dsqr = P.delta * P.delta;
    P.alive.assign(n, 1); P.robust.assign(n, 1); P.err.assign(4 * (size_t)n, 0.0);
    Sim3 S;
    S.r = { s12_q[0], s12_q[1], s12_q[2], s12_q[3] };
    for (int k = 0; k < 3; k++) S.t[k] = s12_t[k];
    S.s = *s12_s;
    if (stats) for (int k = 0; k < 6; k++) stats[k] = 0;
    for (int i = 0; i < n; i++) inlier[i] = 1;

    int trials = 0;
    double first = 0, last = 0;
    const int it1 = optimize(P, S, 5, &trials, &first, &last);        /* :2149-2150 */
    int nBad = 0;
    for (int i = 0; i < n; i++) {                                     /* :2153-2176 */
        if (chi12(P, i) > th2 || chi21(P, i) > th2) { P.alive[i] = 0; inlier[i] = 0; nBad++; }
        else P.robust[i] = 0;
    }
    if (stats) { stats[0] = it1; stats[2] = trials; stats[3] = nBad; stats[4] = first; stats[5] = last; }
    const int more = nBad > 0 ? 10 : 5;
    if (n - nBad < 10) return 0;
    const int it2 = optimize(P, S, more, &trials, nullptr, &last);    /* :2187-2188 */
    int nIn = 0;
    compute_errors(P, S);                                             /* e12->computeError(); e21->computeError(); */
    for (int i = 0; i < n; i++) {
        if (!P.alive[i]) continue;
        if (chi12(P, i) > th2 || chi21(P, i) > th2) inlier[i] = 0;
        else nIn++;
    }
    s12_q[0] = S.r.x; s12_q[1] = S.r.y; s12_q[2] = S.r.z; s12_q[3] = S.r.w;
    for (int k = 0; k < 3; k++) s12_t[k] = S.t[k];
    *s12_s = S.s;
    if (stats) { stats[1] = it2; stats[2] = trials; stats[5] = last; }
    return nIn;
}

/* Optimizer::OptimizeEssentialGraph's solve (O3/src/Optimizer.cc:1389-1651) on the flattened pose graph the caller
 * assembles (:1423-1592): nv Sim3 vertices (g2o::VertexSim3Expmap, estimate Scw as q (x,y,z,w), t, s), `fixed` = the
 * map's initial keyframe (:1447), ne EdgeSim3 edges with vertex 0 = vi, vertex 1 = vj and measurement Sji, information
 * identity, no robust kernel; error = log(Sji * Siw * Sjw^-1) (types_seven_dof_expmap.h:99-106) with g2o's numeric
 * Jacobians for both vertices; Levenberg-Marquardt with setUserLambdaInit(lambda_init) (1e-16, :1400) for `iterations`
 * (20, :1594) iterations; the reduced system is dense here (the reference uses a sparse LDL^T: same solution).
 * sim3[nv*8] in/out.  stats[4] = {LM iterations, LM trials, initial chi2, final chi2}.  Returns iterations run.
 * NO PRODUCT KERNEL EXISTS FOR THIS YET: the restatement is the head start for it. */
int sim3o_optimize_essential_graph(int nv, double* sim3, const uint8_t* fixed, int ne, const int* vi, const int* vj,
                                   const double* meas, int fix_scale, int iterations, double lambda_init, double* stats)
{
    std::vector<Sim3> V(nv), M(ne);
    auto load = [](const double* p) { Sim3 S; S.r = { p[0], p[1], p[2], p[3] }; S.t[0] = p[4]; S.t[1] = p[5]; S.t[2] = p[6]; S.s = p[7]; return S; };
    for (int v = 0; v < nv; v++) V[v] = load(sim3 + 8 * v);
    for (int e = 0; e < ne; e++) M[e] = load(meas + 8 * e);
    std::vector<int> col(nv, -1);
    int nf = 0;
    for (int v = 0; v < nv; v++) if (!fixed[v]) col[v] = nf++;
    const int dim = 7 * nf;
    Problem Pfs; Pfs.fix_scale = fix_scale != 0;   /* oplus only reads fix_scale */
    auto edge_error = [&](int e, const Sim3& Si, const Sim3& Sj, double err[7]) {
        sim3_log(sim3_mul(sim3_mul(M[e], Si), sim3_inverse(Sj)), err);
    };
    std::vector<double> err(7 * (size_t)ne);
    auto compute_errors = [&]() {
        double chi = 0;
        for (int e = 0; e < ne; e++) {
            edge_error(e, V[vi[e]], V[vj[e]], &err[7 * e]);
            for (int k = 0; k < 7; k++) chi += err[7 * e + k] * err[7 * e + k];
        }
        return chi;
    };
    if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
    std::vector<double> H((size_t)dim * dim), b(dim), Hl, x(dim);
    double lambda = -1, ni = 2;
    int nBad = 0, done = 0, trials = 0;
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    for (int it = 0; it < iterations; it++) {
        double currentChi = compute_errors();
        const double iniChi = currentChi;
        if (it == 0 && stats) stats[2] = currentChi;
        std::fill(H.begin(), H.end(), 0.0); std::fill(b.begin(), b.end(), 0.0);
        for (int e = 0; e < ne; e++) {
            const int a = vi[e], c = vj[e];
            double Ji[49], Jj[49];   /* 7 x 7, row-major: d(error row) / d(update column) */
            for (int side = 0; side < 2; side++) {
                const int v = side == 0 ? a : c;
                if (col[v] < 0) continue;
                double* J = side == 0 ? Ji : Jj;
                for (int d = 0; d < 7; d++) {
                    double add[7] = { 0, 0, 0, 0, 0, 0, 0 }, ep[7], em[7];
                    add[d] = delta;
                    const Sim3 Vp = oplus(Pfs, V[v], add);
                    add[d] = -delta;
                    const Sim3 Vm = oplus(Pfs, V[v], add);
                    if (side == 0) { edge_error(e, Vp, V[c], ep); edge_error(e, Vm, V[c], em); }
                    else { edge_error(e, V[a], Vp, ep); edge_error(e, V[a], Vm, em); }
                    for (int r = 0; r < 7; r++) J[r * 7 + d] = scalar * (ep[r] - em[r]);
                }
            }
            const double* er = &err[7 * e];
            auto add_block = [&](int cr, int cc, const double* Jr, const double* Jc) {
                for (int p = 0; p < 7; p++)
                    for (int q = 0; q < 7; q++) {
                        double sacc = 0;
                        for (int r = 0; r < 7; r++) sacc += Jr[r * 7 + p] * Jc[r * 7 + q];
                        H[(size_t)(7 * cr + p) * dim + 7 * cc + q] += sacc;
                    }
            };
            if (col[a] >= 0) {
                add_block(col[a], col[a], Ji, Ji);
                for (int p = 0; p < 7; p++) { double sacc = 0; for (int r = 0; r < 7; r++) sacc += Ji[r * 7 + p] * (-er[r]); b[7 * col[a] + p] += sacc; }
            }
            if (col[c] >= 0) {
                add_block(col[c], col[c], Jj, Jj);
                for (int p = 0; p < 7; p++) { double sacc = 0; for (int r = 0; r < 7; r++) sacc += Jj[r * 7 + p] * (-er[r]); b[7 * col[c] + p] += sacc; }
            }
            if (col[a] >= 0 && col[c] >= 0 && a != c) { add_block(col[a], col[c], Ji, Jj); add_block(col[c], col[a], Jj, Ji); }
        }
        if (it == 0) {
            if (lambda_init > 0) lambda = lambda_init;   /* computeLambdaInit with setUserLambdaInit */
            else {
                double mx = 0;
                for (int j = 0; j < dim; j++) mx = std::max(std::fabs(H[(size_t)j * dim + j]), mx);
                lambda = 1e-5 * mx;
            }
            ni = 2;
            nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const std::vector<Sim3> backup = V;
            Hl = H;
            for (int j = 0; j < dim; j++) Hl[(size_t)j * dim + j] += lambda;
            /* dense unpivoted LDL^T */
            bool ok2 = true;
            {
                std::vector<double> D(dim);
                for (int j = 0; j < dim && ok2; j++) {
                    double dj = Hl[(size_t)j * dim + j];
                    for (int k = 0; k < j; k++) dj -= Hl[(size_t)j * dim + k] * Hl[(size_t)j * dim + k] * D[k];
                    if (!(dj > 0) || !std::isfinite(dj)) { ok2 = false; break; }
                    D[j] = dj;
                    for (int i = j + 1; i < dim; i++) {
                        double v = Hl[(size_t)i * dim + j];
                        for (int k = 0; k < j; k++) v -= Hl[(size_t)i * dim + k] * Hl[(size_t)j * dim + k] * D[k];
                        Hl[(size_t)i * dim + j] = v / dj;
                    }
                }
                if (ok2) {
                    std::vector<double> y(dim);
                    for (int i = 0; i < dim; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= Hl[(size_t)i * dim + k] * y[k]; y[i] = v; }
                    for (int i = 0; i < dim; i++) y[i] /= D[i];
                    for (int i = dim - 1; i >= 0; i--) { double v = y[i]; for (int k = i + 1; k < dim; k++) v -= Hl[(size_t)k * dim + i] * x[k]; x[i] = v; }
                } else std::fill(x.begin(), x.end(), 0.0);
            }
            for (int v = 0; v < nv; v++)
                if (col[v] >= 0) V[v] = oplus(Pfs, V[v], &x[7 * col[v]]);
            double tempChi = compute_errors();
            if (!ok2) tempChi = std::numeric_limits<double>::max();
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < dim; j++) scale += x[j] * (lambda * x[j] + b[j]);
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
                V = backup;
            }
            qmax++;
            trials++;
        } while (rho < 0 && qmax < 10);
        done++;
        if (stats) stats[3] = currentChi;
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
        else nBad = 0;
        if (nBad >= 3) break;
    }
    for (int v = 0; v < nv; v++) {
        double* p = sim3 + 8 * v;
        p[0] = V[v].r.x; p[1] = V[v].r.y; p[2] = V[v].r.z; p[3] = V[v].r.w;
        p[4] = V[v].t[0]; p[5] = V[v].t[1]; p[6] = V[v].t[2]; p[7] = V[v].s;
    }
    if (stats) { stats[0] = done; stats[1] = trials; }
    return done;
}

/* g2o::Sim3 exp / log, exposed for the round-trip self-check */
void sim3o_exp(const double* u, double* out8)
{
    const Sim3 S = sim3_exp(u);
    out8[0] = S.r.x; out8[1] = S.r.y; out8[2] = S.r.z; out8[3] = S.r.w;
    out8[4] = S.t[0]; out8[5] = S.t[1]; out8[6] = S.t[2]; out8[7] = S.s;
}
void sim3o_log(const double* in8, double* u)
{
    Sim3 S; S.r = { in8[0], in8[1], in8[2], in8[3] }; S.t[0] = in8[4]; S.t[1] = in8[5]; S.t[2] = in8[6]; S.s = in8[7];
    sim3_log(S, u);
}


} // extern "C"
