// ORACLE / reference pin (test infrastructure only).
// Sophus::SO3 stand-in following the vendored mm-loam/include/sophus/so3.hpp (which needs the real
// Eigen and cannot be compiled here): expAndTheta :585-623, logAndTheta :247-292, normalize
// :302-308, group product :326-340 (no re-normalisation in this vendored version), point action
// :358-371, hat :673-682, constructor from quaternion (normalises) :487-494, epsilon 1e-10
// (common.hpp:117).
#ifndef MML_REF_SOPHUS_H
#define MML_REF_SOPHUS_H
#include "ref_eigen.h"
namespace Sophus {
template <class Scalar> struct Constants { static Scalar epsilon() { return Scalar(1e-10); } static Scalar pi() { return Scalar(3.141592653589793238462643383279502884); } };
template <class Scalar_> class SO3 {
  Eigen::Quaternion<Scalar_> q_;
 public:
  using Scalar = Scalar_;
  using Tangent = Eigen::Matrix<Scalar, 3, 1>;
  using Transformation = Eigen::Matrix<Scalar, 3, 3>;
  SO3() : q_(Scalar(1), Scalar(0), Scalar(0), Scalar(0)) {}
  SO3(const Transformation& R) : q_(R) {}
  explicit SO3(const Eigen::Quaternion<Scalar>& quat) : q_(quat) { normalize(); }
  void normalize() { Scalar length = q_.norm(); q_.coeffs() /= length; }
  const Eigen::Quaternion<Scalar>& unit_quaternion() const { return q_; }
  Transformation matrix() const { return q_.toRotationMatrix(); }
  SO3 inverse() const { return SO3(q_.conjugate()); }
  template <class U> SO3<U> cast() const { return SO3<U>(q_.template cast<U>()); }
  SO3 operator*(const SO3& other) const {
    const Eigen::Quaternion<Scalar>& a = q_; const Eigen::Quaternion<Scalar>& b = other.q_;
    return SO3(Eigen::Quaternion<Scalar>(
        a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
        a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
        a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
        a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x()));
  }
  template <class O> Eigen::Matrix<Scalar, 3, 1> operator*(const Eigen::MatrixBase<O>& p_) const {
    Eigen::Matrix<Scalar, 3, 1> p = p_.eval();
    Eigen::Matrix<Scalar, 3, 1> uv = q_.vec().cross(p);
    uv += uv;
    return p + q_.w() * uv + q_.vec().cross(uv);
  }
  Tangent log() const {
    using std::abs; using std::atan; using std::sqrt;
    Scalar squared_n = q_.vec().squaredNorm();
    Scalar w = q_.w();
    Scalar two_atan_nbyw_by_n;
    if (squared_n < Constants<Scalar>::epsilon() * Constants<Scalar>::epsilon()) {
      Scalar squared_w = w * w;
      two_atan_nbyw_by_n = Scalar(2) / w - Scalar(2.0 / 3.0) * (squared_n) / (w * squared_w);
    } else {
      Scalar n = sqrt(squared_n);
      if (abs(w) < Constants<Scalar>::epsilon()) {
        if (w > Scalar(0)) two_atan_nbyw_by_n = Constants<Scalar>::pi() / n;
        else two_atan_nbyw_by_n = -Constants<Scalar>::pi() / n;
      } else {
        two_atan_nbyw_by_n = Scalar(2) * atan(n / w) / n;
      }
    }
    return two_atan_nbyw_by_n * q_.vec();
  }
  template <class O> static SO3 exp(const Eigen::MatrixBase<O>& omega_) {
    using std::cos; using std::sin; using std::sqrt;
    Tangent omega = omega_.eval();
    Scalar theta_sq = omega.squaredNorm();
    Scalar imag_factor, real_factor;
    if (theta_sq < Constants<Scalar>::epsilon() * Constants<Scalar>::epsilon()) {
      Scalar theta_po4 = theta_sq * theta_sq;
      imag_factor = Scalar(0.5) - Scalar(1.0 / 48.0) * theta_sq + Scalar(1.0 / 3840.0) * theta_po4;
      real_factor = Scalar(1) - Scalar(1.0 / 8.0) * theta_sq + Scalar(1.0 / 384.0) * theta_po4;
    } else {
      Scalar theta = sqrt(theta_sq);
      Scalar half_theta = Scalar(0.5) * theta;
      Scalar sin_half_theta = sin(half_theta);
      imag_factor = sin_half_theta / theta;
      real_factor = cos(half_theta);
    }
    SO3 q;
    q.q_ = Eigen::Quaternion<Scalar>(real_factor, imag_factor * omega.x(), imag_factor * omega.y(), imag_factor * omega.z());
    return q;
  }
  template <class O> static Transformation hat(const Eigen::MatrixBase<O>& omega) {
    Transformation Omega;
    Omega << Scalar(0), -omega(2), omega(1), omega(2), Scalar(0), -omega(0), -omega(1), omega(0), Scalar(0);
    return Omega;
  }
};
using SO3d = SO3<double>;
}  // namespace Sophus
#endif
