/* oracle/sophus_order.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * float32 restatements of the Sophus / Eigen operations the reference's matchers apply to poses and points, in the
 * operation order of the vendored Sophus (O3/Thirdparty/Sophus/sophus/{so3,se3}.hpp) over Eigen 3.4's fixed-size code
 * paths (3-term reductions are x0 + (x1 + x2), Eigen/src/Core/Redux.h; the 4-float quaternion norm is vectorised:
 * (x^2 + z^2) + (y^2 + w^2)).  Pinned against the reference's own sources compiled with the mini Eigen / Sophus of
 * oracle/slamshim (tests/test_ref_matchers.py).  Quaternions are (x, y, z, w); nothing here renormalises a pose it is
 * given -- a pose is used as the SE3f holds it.  Compile with -ffp-contract=off. */
#ifndef DVM_ORACLE_SOPHUS_ORDER_H
#define DVM_ORACLE_SOPHUS_ORDER_H
#include <cmath>

namespace so {

inline float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
inline float norm3(const float a[3]) { return std::sqrt(dot3(a, a)); }
inline void cross3(const float a[3], const float b[3], float o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
/* SO3::operator*(point), so3.hpp:358-367 */
inline void so3_rotate(const float q[4], const float p[3], float o[3])
{
    float uv[3], c[3];
    cross3(q, p, uv);
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    cross3(q, uv, c);
    for (int i = 0; i < 3; i++) o[i] = (p[i] + q[3] * uv[i]) + c[i];
}
/* SE3::operator*(point), se3.hpp:321-325 */
inline void se3_apply(const float q[4], const float t[3], const float p[3], float o[3])
{
    float r[3];
    so3_rotate(q, p, r);
    for (int i = 0; i < 3; i++) o[i] = r[i] + t[i];
}
inline float quat_norm(const float q[4]) { return std::sqrt((q[0] * q[0] + q[2] * q[2]) + (q[1] * q[1] + q[3] * q[3])); }
/* SO3(quaternion): normalize(), so3.hpp:294-303,480-487 */
inline void quat_normalize(float q[4])
{
    const float n = quat_norm(q);
    for (int i = 0; i < 4; i++) q[i] = q[i] / n;
}
/* SE3::inverse(), se3.hpp:208-211 with SO3::inverse() = SO3(conjugate) (so3.hpp:229-231) */
inline void se3_inverse(const float q[4], const float t[3], float qi[4], float ti[3])
{
    qi[0] = -q[0]; qi[1] = -q[1]; qi[2] = -q[2]; qi[3] = q[3];
    quat_normalize(qi);
    const float mt[3] = { t[0] * -1.f, t[1] * -1.f, t[2] * -1.f };
    so3_rotate(qi, mt, ti);
}
/* SE3 * SE3, se3.hpp:304-309 with the normalising quaternion product of so3.hpp:325-340 */
inline void se3_mul(const float qa[4], const float ta[3], const float qb[4], const float tb[3], float q[4], float t[3])
{
    const float ax = qa[0], ay = qa[1], az = qa[2], aw = qa[3], bx = qb[0], by = qb[1], bz = qb[2], bw = qb[3];
    q[3] = aw * bw - ax * bx - ay * by - az * bz;
    q[0] = aw * bx + ax * bw + ay * bz - az * by;
    q[1] = aw * by + ay * bw + az * bx - ax * bz;
    q[2] = aw * bz + az * bw + ax * by - ay * bx;
    quat_normalize(q);
    float r[3];
    so3_rotate(qa, tb, r);
    for (int i = 0; i < 3; i++) t[i] = ta[i] + r[i];
}
/* QuaternionBase::toRotationMatrix (row-major R) */
inline void quat_to_matrix(const float q[4], float R[9])
{
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w;
    const float txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.f - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.f - (txx + tyy);
}
/* Matrix3f * Vector3f: coefficient-based product, 3-term reduction */
inline void mat_vec(const float R[9], const float p[3], float o[3])
{
    for (int i = 0; i < 3; i++) o[i] = R[3 * i] * p[0] + (R[3 * i + 1] * p[1] + R[3 * i + 2] * p[2]);
}
inline void mat_mul(const float A[9], const float B[9], float C[9])
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + (A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j]);
}
/* Eigen/src/LU/InverseImpl.h, 3x3 */
inline void mat_inverse(const float M[9], float I[9])
{
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return M[3 * i1 + j1] * M[3 * i2 + j2] - M[3 * i1 + j2] * M[3 * i2 + j1];
    };
    const float c0[3] = { cof(0, 0), cof(1, 0), cof(2, 0) };
    const float det = c0[0] * M[0] + (c0[1] * M[3] + c0[2] * M[6]);
    const float invdet = 1.f / det;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) I[3 * r + c] = (r == 0 ? c0[c] : cof(c, r)) * invdet;
}
/* Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other,3,3>: quaternion <- rotation matrix (row-major R) */
inline void quat_from_matrix(const float R[9], float q[4])
{
    float t = R[0] + (R[4] + R[8]);
    if (t > 0.f) {
        t = std::sqrt(t + 1.0f);
        q[3] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0f);
        q[i] = 0.5f * t;
        t = 0.5f / t;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
}
/* Sophus::Sim3f as (RxSO3 quaternion with |q|^2 = scale, translation).  scale(): rxso3.hpp:349-350 */
inline float sim3_scale(const float sq[4]) { return (sq[0] * sq[0] + sq[2] * sq[2]) + (sq[1] * sq[1] + sq[3] * sq[3]); }
/* RxSO3::operator*(point), rxso3.hpp:262-273, then + t (sim3.hpp:226-230) */
inline void sim3_apply(const float sq[4], const float st[3], const float p[3], float o[3])
{
    const float scale = sim3_scale(sq);
    float c2[3], c[3];
    cross3(sq, p, c2);
    c2[0] += c2[0]; c2[1] += c2[1]; c2[2] += c2[2];
    cross3(sq, c2, c);
    for (int i = 0; i < 3; i++) o[i] = (scale * p[i] + (sq[3] * c2[i] + c[i])) + st[i];
}
/* Sim3::inverse(), sim3.hpp:129-132: RxSO3(quaternion().inverse()) (Eigen: conjugate / squaredNorm) applied to -t */
inline void sim3_inverse(const float sq[4], const float st[3], float iq[4], float it[3])
{
    const float n2 = sim3_scale(sq);
    iq[0] = -sq[0] / n2; iq[1] = -sq[1] / n2; iq[2] = -sq[2] / n2; iq[3] = sq[3] / n2;
    const float mt[3] = { st[0] * -1.f, st[1] * -1.f, st[2] * -1.f };
    float z[3] = { 0.f, 0.f, 0.f };
    sim3_apply(iq, z, mt, it);
    /* sim3_apply adds a zero translation: x + 0.f == x for every finite x (and -0.f + 0.f = 0.f, compared equal) */
}
/* Tcw = Sophus::SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale()) (O3/src/ORBmatcher.cc:403,505,1245):
 * rotationMatrix() normalises a copy of the quaternion (rxso3.hpp:341-345), SE3f(R, t) converts the matrix back to a
 * quaternion without normalising (so3.hpp:469-474) */
inline void sim3_to_se3(const float sq[4], const float st[3], float q[4], float t[3])
{
    float nq[4] = { sq[0], sq[1], sq[2], sq[3] }, R[9];
    quat_normalize(nq);
    quat_to_matrix(nq, R);
    quat_from_matrix(R, q);
    const float s = sim3_scale(sq);
    for (int i = 0; i < 3; i++) t[i] = st[i] / s;
}
/* MapPoint::PredictScale, O3/src/MapPoint.cc:557-587: ceil(log(ratio) / mfLogScaleFactor) in float (std::log(float)) */
inline int predict_scale(float maxDistance, float dist, float logScaleFactor, int nlevels)
{
    const float ratio = maxDistance / dist;
    int nScale = (int)std::ceil(std::log(ratio) / logScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= nlevels) nScale = nlevels - 1;
    return nScale;
}

} // namespace so
#endif
