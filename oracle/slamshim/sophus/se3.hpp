/* oracle/slamshim/sophus/se3.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The reference vendors Sophus (O3/Thirdparty/Sophus/sophus) but Sophus needs the real Eigen, which is not installed
 * here.  This header restates, operation by operation, the members of Sophus::SO3 / SE3 / RxSO3 / Sim3 that the
 * reference's matcher sources call, over oracle/slamshim/Eigen/Core.  Line numbers are those of the vendored headers. */
#ifndef DVM_SLAMSHIM_SOPHUS_SE3
#define DVM_SLAMSHIM_SOPHUS_SE3
#include <Eigen/Core>

namespace Sophus {

template <typename S> class SO3 {
public:
    typedef Eigen::Quaternion<S> Q;
    typedef Eigen::VecN<S, 3> V3;
    typedef Eigen::Mat33<S> M3;
    SO3() { }
    explicit SO3(const Q& q) : q_(q) { normalize(); }                     /* so3.hpp:480-487: Base::normalize() */
    SO3(const M3& R) : q_(R) { }                                           /* so3.hpp:469-474 (no normalisation) */
    void normalize() { const S len = q_.norm(); for (int i = 0; i < 4; i++) q_.c[i] = q_.c[i] / len; } /* :294-303 */
    const Q& unit_quaternion() const { return q_; }
    /* SHIM ONLY (not Sophus API): an SO3 holding exactly the stored coefficients, the way a Frame's mTcw holds the result of
     * an earlier normalising operation.  Used by the test glue to hand a pose over unchanged. */
    static SO3 fromStored(const Q& q) { SO3 r; r.q_ = q; return r; }
    SO3 inverse() const { return SO3(q_.conjugate()); }                    /* so3.hpp:229-231 */
    M3 matrix() const { return q_.toRotationMatrix(); }                    /* so3.hpp:310-312 */
    SO3 operator*(const SO3& o) const
    {   /* so3.hpp:325-340, then the quaternion constructor normalises */
        const Q &a = q_, &b = o.q_;
        return SO3(Q(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                     a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                     a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                     a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x()));
    }
    V3 operator*(const V3& p) const
    {   /* so3.hpp:358-367 */
        V3 uv = q_.vec().cross(p);
        uv += uv;
        return p + q_.w() * uv + q_.vec().cross(uv);
    }
    static M3 hat(const V3& omega)
    {   /* so3.hpp: hat() */
        M3 O;
        O(0, 0) = S(0); O(0, 1) = -omega(2); O(0, 2) = omega(1);
        O(1, 0) = omega(2); O(1, 1) = S(0); O(1, 2) = -omega(0);
        O(2, 0) = -omega(1); O(2, 1) = omega(0); O(2, 2) = S(0);
        return O;
    }
private:
    Q q_;
};
typedef SO3<float> SO3f;
typedef SO3<double> SO3d;

template <typename S> class SE3 {
public:
    typedef Eigen::VecN<S, 3> V3;
    typedef Eigen::Mat33<S> M3;
    SE3() { }
    SE3(const SO3<S>& so3, const V3& t) : so3_(so3), t_(t) { }            /* se3.hpp: SE3(SO3Base, translation) */
    SE3(const M3& R, const V3& t) : so3_(R), t_(t) { }                    /* se3.hpp: SE3(Matrix3, Point) */
    SE3(const Eigen::Quaternion<S>& q, const V3& t) : so3_(q), t_(t) { }  /* se3.hpp: SE3(Quaternion, Point): normalises */
    const SO3<S>& so3() const { return so3_; }
    const Eigen::Quaternion<S>& unit_quaternion() const { return so3_.unit_quaternion(); }
    const V3& translation() const { return t_; }
    V3& translation() { return t_; }
    M3 rotationMatrix() const { return so3_.matrix(); }
    SE3 inverse() const
    {   /* se3.hpp:208-211 */
        const SO3<S> invR = so3_.inverse();
        return SE3(invR, invR * (t_ * S(-1)));
    }
    SE3 operator*(const SE3& o) const { return SE3(so3_ * o.so3_, t_ + so3_ * o.t_); }   /* se3.hpp:304-309 */
    V3 operator*(const V3& p) const { return so3_ * p + t_; }                             /* se3.hpp:321-325 */
    template <typename T> SE3<T> cast() const
    {
        const Eigen::Quaternion<S>& q = unit_quaternion();
        return SE3<T>(Eigen::Quaternion<T>((T)q.w(), (T)q.x(), (T)q.y(), (T)q.z()), t_.template cast<T>());
    }
private:
    SO3<S> so3_;
    V3 t_;
};
typedef SE3<float> SE3f;
typedef SE3<double> SE3d;

} // namespace Sophus
#endif
