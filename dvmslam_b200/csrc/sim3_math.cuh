// sim3_math.cuh -- g2o's Sim3 (O3/Thirdparty/g2o/g2o/types/sim3.h) in double on the device: exponential, logarithm,
// product, inverse.  Shared by optimize_sim3_kernel (sim3.cu) and essential_graph_kernel (essential_graph.cu).
#pragma once
#include "common.cuh"

namespace dvm {
namespace sim3m {

struct Quat { double x, y, z, w; };
struct Sim3 { Quat r; double t[3]; double s; };

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
__device__ inline Quat quat_from_matrix(const double R[9]) // Eigen::Quaterniond(Matrix3d)
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

// g2o::Sim3(const Vector7d& update), O3/Thirdparty/g2o/g2o/types/sim3.h
__device__ inline Sim3 sim3_exp(const double u[7])
{
    const double w0 = u[0], w1 = u[1], w2 = u[2], sigma = u[6];
    const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    const double O[9] = { 0, -w2, w1, w2, 0, -w0, -w1, w0, 0 };
    double O2[9], R[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    Sim3 S;
    S.s = exp(sigma);
    const double eps = 0.00001;
    double A, B, C;
    const bool small_rot = theta < eps;
    if (small_rot) {
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i];
    } else {
        const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta);
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * O[i] + b * O2[i];
    }
    if (fabs(sigma) < eps) {
        C = 1;
        if (small_rot) { A = 1. / 2.; B = 1. / 6.; }
        else {
            const double theta2 = theta * theta;
            A = (1 - cos(theta)) / theta2;
            B = (theta - sin(theta)) / (theta2 * theta);
        }
    } else {
        C = (S.s - 1) / sigma;
        if (small_rot) {
            const double sigma2 = sigma * sigma;
            A = ((sigma - 1) * S.s + 1) / sigma2;
            B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
        } else {
            const double a = S.s * sin(theta), b = S.s * cos(theta);
            const double theta2 = theta * theta, sigma2 = sigma * sigma, c = theta2 + sigma2;
            A = (a * sigma + (1 - b) * theta) / (theta * c);
            B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
        }
    }
    S.r = quat_from_matrix(R);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double acc = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) acc += (A * O[i * 3 + j] + B * O2[i * 3 + j] + C * (i == j ? 1.0 : 0.0)) * u[3 + j];
        S.t[i] = acc;
    }
    return S;
}
__device__ inline Sim3 sim3_mul(const Sim3& a, const Sim3& b)
{
    Sim3 r;
    r.r = quat_mul(a.r, b.r);
    double rt[3];
    quat_rotate(a.r, b.t, rt);
#pragma unroll
    for (int i = 0; i < 3; i++) r.t[i] = a.s * rt[i] + a.t[i];
    r.s = a.s * b.s;
    return r;
}
__device__ inline Sim3 sim3_inverse(const Sim3& a)
{
    Sim3 r;
    r.r = { -a.r.x, -a.r.y, -a.r.z, a.r.w };
    const double v[3] = { (-1. / a.s) * a.t[0], (-1. / a.s) * a.t[1], (-1. / a.s) * a.t[2] };
    quat_rotate(r.r, v, r.t);
    r.s = 1. / a.s;
    return r;
}

__device__ inline void quat_to_matrix(const Quat& q, double R[9]) // Eigen toRotationMatrix
{
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// W.lu().solve(t): 3 x 3 LU with partial pivoting (Eigen::PartialPivLU)
__device__ inline void lu_solve3(const double W[9], const double t[3], double x[3])
{
    double A[9], y[3];
#pragma unroll
    for (int i = 0; i < 9; i++) A[i] = W[i];
    int piv[3] = { 0, 1, 2 };
    for (int k = 0; k < 3; k++) {
        int best = k;
        for (int r = k + 1; r < 3; r++)
            if (fabs(A[piv[r] * 3 + k]) > fabs(A[piv[best] * 3 + k])) best = r;
        const int tmp = piv[k]; piv[k] = piv[best]; piv[best] = tmp;
        const double* pk = &A[piv[k] * 3];
        for (int r = k + 1; r < 3; r++) {
            double* pr = &A[piv[r] * 3];
            const double f = pr[k] / pk[k];
            pr[k] = f;
            for (int c = k + 1; c < 3; c++) pr[c] -= f * pk[c];
        }
    }
    for (int k = 0; k < 3; k++) {
        double v = t[piv[k]];
        for (int c = 0; c < k; c++) v -= A[piv[k] * 3 + c] * y[c];
        y[k] = v;
    }
    for (int k = 2; k >= 0; k--) {
        double v = y[k];
        for (int c = k + 1; c < 3; c++) v -= A[piv[k] * 3 + c] * x[c];
        x[k] = v / A[piv[k] * 3 + k];
    }
}
// g2o::Sim3::log(), sim3.h (its small-rotation coefficients kept as they are: see oracle/sim3_oracle.cpp's header)
__device__ inline void sim3_log(const Sim3& S, double res[7])
{
    const double sigma = log(S.s);
    double R[9];
    quat_to_matrix(S.r, R);
    const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
    const double eps = 0.00001;
    const double dR[3] = { R[7] - R[5], R[2] - R[6], R[3] - R[1] };   // deltaR(R)
    double omega[3], A, B, C, f;
    if (fabs(sigma) < eps) {
        C = 1;
        if (d > 1 - eps) { f = 0.5; A = 1. / 2.; B = 1. / 6.; }
        else {
            const double theta = acos(d), theta2 = theta * theta;
            f = theta / (2 * sqrt(1 - d * d));
            A = (1 - cos(theta)) / theta2;
            B = (theta - sin(theta)) / (theta2 * theta);
        }
    } else {
        C = (S.s - 1) / sigma;
        if (d > 1 - eps) {
            const double sigma2 = sigma * sigma;
            f = 0.5;
            A = ((sigma - 1) * S.s + 1) / sigma2;
            B = ((0.5 * sigma2 - sigma + 1) * S.s) / (sigma2 * sigma);
        } else {
            const double theta = acos(d);
            f = theta / (2 * sqrt(1 - d * d));
            const double theta2 = theta * theta;
            const double a = S.s * sin(theta), b = S.s * cos(theta), c = theta2 + sigma * sigma;
            A = (a * sigma + (1 - b) * theta) / (theta * c);
            B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) omega[i] = f * dR[i];
    const double O[9] = { 0, -omega[2], omega[1], omega[2], 0, -omega[0], -omega[1], omega[0], 0 };
    double W[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double O2 = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
            W[i * 3 + j] = A * O[i * 3 + j] + B * O2 + C * (i == j ? 1.0 : 0.0);
        }
    double ups[3];
    lu_solve3(W, S.t, ups);
#pragma unroll
    for (int i = 0; i < 3; i++) { res[i] = omega[i]; res[3 + i] = ups[i]; }
    res[6] = sigma;
}

} // namespace sim3m
} // namespace dvm
