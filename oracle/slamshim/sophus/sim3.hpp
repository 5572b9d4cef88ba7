/* oracle/slamshim/sophus/sim3.hpp -- TEST INFRASTRUCTURE (see se3.hpp).  RxSO3 / Sim3 members used by ORBmatcher.cc. */
#ifndef DVM_SLAMSHIM_SOPHUS_SIM3
#define DVM_SLAMSHIM_SOPHUS_SIM3
#include "se3.hpp"

namespace Sophus {

template <typename S> class RxSO3 {
public:
    typedef Eigen::Quaternion<S> Q;
    typedef Eigen::VecN<S, 3> V3;
    typedef Eigen::Mat33<S> M3;
    RxSO3() { }
    explicit RxSO3(const Q& q) : q_(q) { }                                 /* rxso3.hpp:485-491 */
    RxSO3(const S& scale, const M3& R) : q_(R)                             /* rxso3.hpp:460-466 */
    {
        const S s = std::sqrt(scale);
        for (int i = 0; i < 4; i++) q_.c[i] = q_.c[i] * s;
    }
    RxSO3(const S& scale, const SO3<S>& so3) : q_(so3.unit_quaternion())   /* rxso3.hpp:472-478 */
    {
        const S s = std::sqrt(scale);
        for (int i = 0; i < 4; i++) q_.c[i] = q_.c[i] * s;
    }
    const Q& quaternion() const { return q_; }
    RxSO3 inverse() const { return RxSO3(q_.inverse()); }                  /* rxso3.hpp:156-158 */
    S scale() const { return q_.squaredNorm(); }                           /* rxso3.hpp:349-350 */
    M3 rotationMatrix() const { Q n = q_; n.normalize(); return n.toRotationMatrix(); }   /* rxso3.hpp:341-345 */
    V3 operator*(const V3& p) const
    {   /* rxso3.hpp:262-273 */
        const S scale = q_.squaredNorm();
        V3 two_vec_cross_p = q_.vec().cross(p);
        two_vec_cross_p += two_vec_cross_p;
        return scale * p + (q_.w() * two_vec_cross_p + q_.vec().cross(two_vec_cross_p));
    }
private:
    Q q_;
};

template <typename S> class Sim3 {
public:
    typedef Eigen::VecN<S, 3> V3;
    typedef Eigen::Mat33<S> M3;
    Sim3() { }
    Sim3(const RxSO3<S>& r, const V3& t) : r_(r), t_(t) { }               /* sim3.hpp:387-395 */
    Sim3(const Eigen::Quaternion<S>& q, const V3& t) : r_(q), t_(t) { }   /* sim3.hpp:401-409 */
    const RxSO3<S>& rxso3() const { return r_; }
    const Eigen::Quaternion<S>& quaternion() const { return r_.quaternion(); }
    const V3& translation() const { return t_; }
    M3 rotationMatrix() const { return r_.rotationMatrix(); }
    S scale() const { return r_.scale(); }
    Sim3 inverse() const
    {   /* sim3.hpp:129-132 */
        const RxSO3<S> invR = r_.inverse();
        return Sim3(invR, invR * (t_ * S(-1)));
    }
    V3 operator*(const V3& p) const { return r_ * p + t_; }               /* sim3.hpp:226-230 */
private:
    RxSO3<S> r_;
    V3 t_;
};
typedef Sim3<float> Sim3f;
typedef Sim3<double> Sim3d;

} // namespace Sophus
#endif
