/* oracle/g2oshim/sophus/se3.hpp -- TEST INFRASTRUCTURE.  The few members of Sophus::SE3<T> that the optimisation
 * functions of O3/src/Optimizer.cc touch (construction from quaternion / rotation matrix + translation, accessors, cast,
 * inverse, products), over the mini Eigen of this directory.  The float arithmetic of the pose hand-over is not what this
 * build pins (the matchers' build does, oracle/slamshim): here poses only enter and leave g2o. */
#ifndef DVM_G2OSHIM_SOPHUS_SE3_HPP
#define DVM_G2OSHIM_SOPHUS_SE3_HPP
#include <Eigen/Geometry>

namespace Sophus {
template <class T> class SE3 {
    Eigen::Quaternion<T> q_;
    Eigen::Matrix<T, 3, 1> t_;
public:
    SE3() : q_(1, 0, 0, 0) { t_.setZero(); }
    SE3(const Eigen::Quaternion<T>& q, const Eigen::Matrix<T, 3, 1>& t) : q_(q), t_(t) { q_.normalize(); }   /* so3.hpp:294-303 */
    SE3(const Eigen::Matrix<T, 3, 3>& R, const Eigen::Matrix<T, 3, 1>& t) : q_(R), t_(t) { q_.normalize(); }
    const Eigen::Quaternion<T>& unit_quaternion() const { return q_; }
    const Eigen::Matrix<T, 3, 1>& translation() const { return t_; }
    Eigen::Matrix<T, 3, 1>& translation() { return t_; }
    Eigen::Matrix<T, 3, 3> rotationMatrix() const { return q_.toRotationMatrix(); }
    Eigen::Matrix<T, 4, 4> matrix() const
    {
        Eigen::Matrix<T, 4, 4> m;
        m.setIdentity();
        m.template block<3, 3>(0, 0) = rotationMatrix();
        m.template block<3, 1>(0, 3) = t_;
        return m;
    }
    template <class U> SE3<U> cast() const { return SE3<U>(q_.template cast<U>(), t_.template cast<U>()); }
    SE3 inverse() const { const Eigen::Quaternion<T> qi = q_.conjugate(); return SE3(qi, qi * (t_ * T(-1))); }
    SE3 operator*(const SE3& o) const { return SE3(q_ * o.q_, t_ + q_ * o.t_); }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const { return q_ * p + t_; }
};
typedef SE3<float> SE3f;
typedef SE3<double> SE3d;
} // namespace Sophus
#endif
