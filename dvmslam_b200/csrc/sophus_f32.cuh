// sophus_f32.cuh -- the float32 pose arithmetic of the reference's matchers, operation by operation.
//
// The reference holds poses as Sophus::SE3f (unit quaternion + translation) and moves points with Sophus' group
// action, not with a rotation matrix: p' = p + w * 2(qv x p) + qv x 2(qv x p) (O3/Thirdparty/Sophus/sophus/so3.hpp:358-367),
// then + t (se3.hpp:321-325).  Frame::isInFrustum is the exception: it uses mRcw * P + mtcw with
// mRcw = toRotationMatrix(q) (O3/src/Frame.cc:553-559,585), which Eigen evaluates coefficient by coefficient with its
// 3-term reduction order x0 + (x1 + x2) (Eigen/src/Core/Redux.h).  Float results differ in the last bit between the two
// forms, and window / bounds tests downstream can flip on that bit, so every step is spelled out with round-to-nearest
// intrinsics (no FMA contraction).  A pose is used as the SE3f holds it: nothing here renormalises its input.
// Pinned to the reference's sources through the oracle (tests/test_ref_matchers.py).
#pragma once
#include "glibc_logf.h"

namespace dvm {
namespace so {

#if defined(__CUDA_ARCH__)
#define DVM_SO_HD __device__ __forceinline__
DVM_SO_HD float fm(float a, float b) { return __fmul_rn(a, b); }
DVM_SO_HD float fa(float a, float b) { return __fadd_rn(a, b); }
DVM_SO_HD float fs(float a, float b) { return __fsub_rn(a, b); }
DVM_SO_HD float fd(float a, float b) { return __fdiv_rn(a, b); }
DVM_SO_HD float fsqrt(float a) { return __fsqrt_rn(a); }
#else   // host code of the library is compiled with -ffp-contract=off
#define DVM_SO_HD inline
DVM_SO_HD float fm(float a, float b) { return a * b; }
DVM_SO_HD float fa(float a, float b) { return a + b; }
DVM_SO_HD float fs(float a, float b) { return a - b; }
DVM_SO_HD float fd(float a, float b) { return a / b; }
DVM_SO_HD float fsqrt(float a) { return sqrtf(a); }
#endif

DVM_SO_HD float dot3(const float a[3], const float b[3]) { return fa(fm(a[0], b[0]), fa(fm(a[1], b[1]), fm(a[2], b[2]))); }
DVM_SO_HD float norm3(const float a[3]) { return fsqrt(dot3(a, a)); }
DVM_SO_HD void cross3(const float a[3], const float b[3], float o[3])
{
    o[0] = fs(fm(a[1], b[2]), fm(a[2], b[1]));
    o[1] = fs(fm(a[2], b[0]), fm(a[0], b[2]));
    o[2] = fs(fm(a[0], b[1]), fm(a[1], b[0]));
}
// SO3::operator*(point), so3.hpp:358-367; q = (x, y, z, w)
DVM_SO_HD void so3_rotate(const float q[4], const float p[3], float o[3])
{
    float uv[3], c[3];
    cross3(q, p, uv);
    uv[0] = fa(uv[0], uv[0]); uv[1] = fa(uv[1], uv[1]); uv[2] = fa(uv[2], uv[2]);
    cross3(q, uv, c);
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = fa(fa(p[i], fm(q[3], uv[i])), c[i]);
}
// SE3::operator*(point), se3.hpp:321-325
DVM_SO_HD void se3_apply(const float q[4], const float t[3], const float p[3], float o[3])
{
    float r[3];
    so3_rotate(q, p, r);
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = fa(r[i], t[i]);
}
// SO3(quaternion) normalises (so3.hpp:294-303,480-487); the 4-float squared norm is a vectorised reduction in Eigen:
// (x^2 + z^2) + (y^2 + w^2)
DVM_SO_HD void quat_normalize(float q[4])
{
    const float n = fsqrt(fa(fa(fm(q[0], q[0]), fm(q[2], q[2])), fa(fm(q[1], q[1]), fm(q[3], q[3]))));
#pragma unroll
    for (int i = 0; i < 4; i++) q[i] = fd(q[i], n);
}
// SE3::inverse(), se3.hpp:208-211: SO3(conjugate) and its action on -t
DVM_SO_HD void se3_inverse(const float q[4], const float t[3], float qi[4], float ti[3])
{
    qi[0] = -q[0]; qi[1] = -q[1]; qi[2] = -q[2]; qi[3] = q[3];
    quat_normalize(qi);
    const float mt[3] = { fm(t[0], -1.f), fm(t[1], -1.f), fm(t[2], -1.f) };
    so3_rotate(qi, mt, ti);
}
// SE3 * SE3, se3.hpp:304-309 (normalising quaternion product, so3.hpp:325-340)
DVM_SO_HD void se3_mul(const float qa[4], const float ta[3], const float qb[4], const float tb[3], float q[4], float t[3])
{
    const float ax = qa[0], ay = qa[1], az = qa[2], aw = qa[3], bx = qb[0], by = qb[1], bz = qb[2], bw = qb[3];
    q[3] = fs(fs(fs(fm(aw, bw), fm(ax, bx)), fm(ay, by)), fm(az, bz));
    q[0] = fs(fa(fa(fm(aw, bx), fm(ax, bw)), fm(ay, bz)), fm(az, by));
    q[1] = fs(fa(fa(fm(aw, by), fm(ay, bw)), fm(az, bx)), fm(ax, bz));
    q[2] = fs(fa(fa(fm(aw, bz), fm(az, bw)), fm(ax, by)), fm(ay, bx));
    quat_normalize(q);
    float r[3];
    so3_rotate(qa, tb, r);
#pragma unroll
    for (int i = 0; i < 3; i++) t[i] = fa(ta[i], r[i]);
}
// QuaternionBase::toRotationMatrix, row-major
DVM_SO_HD void quat_to_matrix(const float q[4], float R[9])
{
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float tx = fm(2.f, x), ty = fm(2.f, y), tz = fm(2.f, z);
    const float twx = fm(tx, w), twy = fm(ty, w), twz = fm(tz, w);
    const float txx = fm(tx, x), txy = fm(ty, x), txz = fm(tz, x);
    const float tyy = fm(ty, y), tyz = fm(tz, y), tzz = fm(tz, z);
    R[0] = fs(1.f, fa(tyy, tzz)); R[1] = fs(txy, twz); R[2] = fa(txz, twy);
    R[3] = fa(txy, twz); R[4] = fs(1.f, fa(txx, tzz)); R[5] = fs(tyz, twx);
    R[6] = fs(txz, twy); R[7] = fa(tyz, twx); R[8] = fs(1.f, fa(txx, tyy));
}
// Matrix3f * Vector3f / Matrix3f * Matrix3f: coefficient-based products, 3-term reduction x0 + (x1 + x2)
DVM_SO_HD void mat_vec(const float R[9], const float p[3], float o[3])
{
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = fa(fm(R[3 * i], p[0]), fa(fm(R[3 * i + 1], p[1]), fm(R[3 * i + 2], p[2])));
}
DVM_SO_HD void mat_mul(const float A[9], const float B[9], float C[9])
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            C[3 * i + j] = fa(fm(A[3 * i], B[j]), fa(fm(A[3 * i + 1], B[3 + j]), fm(A[3 * i + 2], B[6 + j])));
}
// Eigen/src/LU/InverseImpl.h, 3x3: cofactors, det = sum(cofactors_col0 .* col(0)), scaled by 1 / det
DVM_SO_HD void mat_inverse(const float M[9], float I[9])
{
    float cof[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            cof[3 * i + j] = fs(fm(M[3 * i1 + j1], M[3 * i2 + j2]), fm(M[3 * i1 + j2], M[3 * i2 + j1]));
        }
    const float det = fa(fm(cof[0], M[0]), fa(fm(cof[3], M[3]), fm(cof[6], M[6])));
    const float invdet = fd(1.f, det);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) I[3 * r + c] = fm(cof[3 * c + r], invdet);
}
// Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other,3,3>: quaternion <- rotation matrix (row-major R)
DVM_SO_HD void quat_from_matrix(const float R[9], float q[4])
{
    float t = fa(R[0], fa(R[4], R[8]));
    if (t > 0.f) {
        t = fsqrt(fa(t, 1.0f));
        q[3] = fm(0.5f, t);
        t = fd(0.5f, t);
        q[0] = fm(fs(R[7], R[5]), t); q[1] = fm(fs(R[2], R[6]), t); q[2] = fm(fs(R[3], R[1]), t);
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = fsqrt(fa(fs(fs(R[4 * i], R[4 * j]), R[4 * k]), 1.0f));
        q[i] = fm(0.5f, t);
        t = fd(0.5f, t);
        q[3] = fm(fs(R[3 * k + j], R[3 * j + k]), t);
        q[j] = fm(fa(R[3 * j + i], R[3 * i + j]), t);
        q[k] = fm(fa(R[3 * k + i], R[3 * i + k]), t);
    }
}
// Sophus::Sim3f as (RxSO3 quaternion with |q|^2 = scale, translation); scale(): rxso3.hpp:349-350
DVM_SO_HD float sim3_scale(const float sq[4]) { return fa(fa(fm(sq[0], sq[0]), fm(sq[2], sq[2])), fa(fm(sq[1], sq[1]), fm(sq[3], sq[3]))); }
// RxSO3::operator*(point), rxso3.hpp:262-273, then + t (sim3.hpp:226-230)
DVM_SO_HD void sim3_apply(const float sq[4], const float st[3], const float p[3], float o[3])
{
    const float scale = sim3_scale(sq);
    float c2[3], c[3];
    cross3(sq, p, c2);
    c2[0] = fa(c2[0], c2[0]); c2[1] = fa(c2[1], c2[1]); c2[2] = fa(c2[2], c2[2]);
    cross3(sq, c2, c);
#pragma unroll
    for (int i = 0; i < 3; i++) o[i] = fa(fa(fm(scale, p[i]), fa(fm(sq[3], c2[i]), c[i])), st[i]);
}
// Sim3::inverse(), sim3.hpp:129-132: RxSO3(quaternion().inverse()) = conjugate / squaredNorm, applied to -t
DVM_SO_HD void sim3_inverse(const float sq[4], const float st[3], float iq[4], float it[3])
{
    const float n2 = sim3_scale(sq);
    iq[0] = fd(-sq[0], n2); iq[1] = fd(-sq[1], n2); iq[2] = fd(-sq[2], n2); iq[3] = fd(sq[3], n2);
    const float mt[3] = { fm(st[0], -1.f), fm(st[1], -1.f), fm(st[2], -1.f) };
    const float z[3] = { 0.f, 0.f, 0.f };
    sim3_apply(iq, z, mt, it);
}
// Tcw = Sophus::SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale()) (O3/src/ORBmatcher.cc:403,505,1245)
DVM_SO_HD void sim3_to_se3(const float sq[4], const float st[3], float q[4], float t[3])
{
    float nq[4] = { sq[0], sq[1], sq[2], sq[3] }, R[9];
    quat_normalize(nq);
    quat_to_matrix(nq, R);
    quat_from_matrix(R, q);
    const float s = sim3_scale(sq);
#pragma unroll
    for (int i = 0; i < 3; i++) t[i] = fd(st[i], s);
}
// MapPoint::PredictScale, O3/src/MapPoint.cc:557-587: ceil(logf(mfMaxDistance / dist) / mfLogScaleFactor), clamped
DVM_SO_HD int predict_scale(float maxDistance, float dist, float logScaleFactor, int nlevels)
{
    const float ratio = fd(maxDistance, dist);
    int nScale = (int)ceilf(fd(dvm_glibc_logf(ratio), logScaleFactor));
    if (nScale < 0) nScale = 0;
    else if (nScale >= nlevels) nScale = nlevels - 1;
    return nScale;
}

} // namespace so
} // namespace dvm
