// gtsam_lite.h -- header-only subset of the gtsam:: API the reference's wrapper sources and drivers use (SURVEY.md
// section 8b item 2), as host-side VALUE types over Eigen matrices that forward all numeric work to the C ABI
// (include/fg_abi.h).  Nothing here solves anything on the CPU: LevenbergMarquardtOptimizer::optimize, ISAM2::update,
// NonlinearFactorGraph::error, Marginals and PreintegratedCombinedMeasurements run on the GPU through libfg_b200.so.
//
// Its purpose is that gtsam/gtsam_graph.cpp, imu_base.cpp, imu_vn100.cpp and the drivers test_vro_imu_graph.cpp /
// test_ba_imu_graph.cpp of the reference compile UNCHANGED: the headers under compat/gtsam/... all include this file.
// <Eigen/...> resolves to the real Eigen when one is installed, else to compat/mini_eigen.h.
//
// Mirrors (reference file:line of the usage):
//   Pose3/Rot3/Point3/NavState/imuBias::ConstantBias   gtsam/gtsam_graph.cpp:320-368,613-695
//   Values insert/update/at/exists                      gtsam/gtsam_graph.cpp:333,619-623,632-666
//   NonlinearFactorGraph add/push_back/resize/error     gtsam/gtsam_graph.cpp:173-176,341,691-692,1773
//   PriorFactor / BetweenFactor / CombinedImuFactor / OrientedPlane3Factor / GenericProjectionFactor
//   LevenbergMarquardtOptimizer(graph, values).optimize()   gtsam/gtsam_graph.cpp:1786-1787
//   ISAM2::update + calculateEstimate                   gtsam/gtsam_graph.cpp:1770-1772
//   Marginals(...).marginalCovariance                   gtsam/gtsam_graph.cpp:598-601,1357
//   PreintegratedCombinedMeasurements (+Params::MakeSharedD)  gtsam/imu_base.cpp:83-98,258-263; imu_vn100.cpp:16,57-62
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <boost/shared_ptr.hpp>
#include "../../include/fg_abi.h"

// Which of GTSAM's compile-time charts this facade plays (include/fg_abi.h FG_CHART_*): 0 full EXPMAP (default),
// 2 Pose3 FIRST_ORDER over Rot3 EXPMAP, 3 Pose3 FIRST_ORDER over Rot3 CAYLEY (GTSAM 4.0's default build).  One switch for the
// host side (Pose3::ChartAtOrigin -- the 6-vector of the VRO log) and the device side (every context is created with it).
#ifndef FG_POSE3_CHART
#define FG_POSE3_CHART 0
#endif

namespace gtsam {

typedef uint64_t Key;
typedef Eigen::Quaterniond Quaternion;
typedef Eigen::MatrixXd Matrix;
typedef Eigen::VectorXd Vector;
typedef Eigen::Matrix<double, 2, 1> Vector2;
typedef Eigen::Matrix<double, 3, 1> Vector3;
typedef Eigen::Matrix<double, 4, 1> Vector4;
typedef Eigen::Matrix<double, 6, 1> Vector6;
typedef Eigen::Matrix<double, 9, 1> Vector9;
typedef Eigen::Matrix<double, 2, 2> Matrix2;
typedef Eigen::Matrix<double, 2, 2> Matrix22;
typedef Eigen::Matrix<double, 3, 3> Matrix3;
typedef Eigen::Matrix<double, 3, 3> Matrix33;
typedef Eigen::Matrix<double, 3, 2> Matrix32;
typedef Eigen::Matrix<double, 2, 3> Matrix23;
typedef Eigen::Matrix<double, 3, 6> Matrix36;
typedef Eigen::Matrix<double, 4, 4> Matrix4;
typedef Eigen::Matrix<double, 4, 4> Matrix44;
typedef Eigen::Matrix<double, 6, 6> Matrix6;
typedef Eigen::Matrix<double, 6, 6> Matrix66;
typedef Eigen::Matrix<double, 15, 15> Matrix15;
static const Matrix3 I_3x3 = Matrix3::Identity();
static const Matrix3 Z_3x3 = Matrix3::Zero();
static const Matrix6 I_6x6 = Matrix6::Identity();

// Point3 / Point2 are classes in GTSAM 4.0's default build; here thin wrappers so that Values can tell a Point3 landmark
// from a Vector3 velocity
class Point3 : public Vector3 {
 public:
  Point3() { setZero(); }
  Point3(double x, double y, double z) : Vector3(x, y, z) {}
  template <int R, int C, int O, int MR, int MC> Point3(const Eigen::Matrix<double, R, C, O, MR, MC>& v) : Vector3(v) {}
  const Vector3& vector() const { return *this; }
  static Point3 Zero() { return Point3(); }
  void print(const std::string& s = "") const { std::cout << s << " [" << (*this)(0) << ", " << (*this)(1) << ", " << (*this)(2) << "]'" << std::endl; }
};
class Point2 : public Vector2 {
 public:
  Point2() { setZero(); }
  Point2(double x, double y) : Vector2(x, y) {}
  template <int R, int C, int O, int MR, int MC> Point2(const Eigen::Matrix<double, R, C, O, MR, MC>& v) : Vector2(v) {}
  const Vector2& vector() const { return *this; }
};
inline Vector3 vec3(double x, double y, double z) { return Vector3(x, y, z); }

// ------------------------------------------------------------------ Symbol (gtsam_graph.cpp:50-54)
class Symbol {
  unsigned char c_; uint64_t j_;
 public:
  Symbol() : c_(0), j_(0) {}
  Symbol(unsigned char c, uint64_t j) : c_(c), j_(j) {}
  Symbol(Key k) : c_((unsigned char)(k >> 56)), j_(k & ((uint64_t(1) << 56) - 1)) {}
  Key key() const { return (uint64_t(c_) << 56) | j_; }
  operator Key() const { return key(); }
  unsigned char chr() const { return c_; }
  uint64_t index() const { return j_; }
  void print(const std::string& s = "") const { std::cout << s << c_ << j_ << std::endl; }
};
namespace symbol_shorthand {
inline Key X(uint64_t j) { return Symbol('x', j); }
inline Key V(uint64_t j) { return Symbol('v', j); }
inline Key B(uint64_t j) { return Symbol('b', j); }
inline Key L(uint64_t j) { return Symbol('l', j); }
inline Key Q(uint64_t j) { return Symbol('q', j); }
}  // namespace symbol_shorthand

namespace detail {
inline void so3_coeff(double th2, double& a, double& b, double& c) {
  if (th2 < 1e-10) { a = 1 - th2 / 6; b = 0.5 - th2 / 24; c = 1.0 / 6 - th2 / 120; }
  else { double t = std::sqrt(th2), s = std::sin(t), sh = std::sin(0.5 * t); a = s / t; b = 2 * sh * sh / th2; c = (t - s) / (th2 * t); }
}
inline Matrix3 skew(const Vector3& w) { Matrix3 S = Matrix3::Zero(); S(0, 1) = -w[2]; S(0, 2) = w[1]; S(1, 0) = w[2]; S(1, 2) = -w[0]; S(2, 0) = -w[1]; S(2, 1) = w[0]; return S; }
// Rot3::Logmap with GTSAM's three near-pi branches and the small-angle series (the same formulas as the device so3_log)
inline Vector3 so3_log(const Matrix3& R) {
  const double tr = R(0, 0) + R(1, 1) + R(2, 2);
  if (tr + 1.0 < 1e-10) {
    if (std::fabs(R(2, 2) + 1.0) > 1e-5) { const double k = M_PI / std::sqrt(2.0 + 2.0 * R(2, 2)); return Vector3(k * R(0, 2), k * R(1, 2), k * (1.0 + R(2, 2))); }
    if (std::fabs(R(1, 1) + 1.0) > 1e-5) { const double k = M_PI / std::sqrt(2.0 + 2.0 * R(1, 1)); return Vector3(k * R(0, 1), k * (1.0 + R(1, 1)), k * R(2, 1)); }
    const double k = M_PI / std::sqrt(2.0 + 2.0 * R(0, 0));
    return Vector3(k * (1.0 + R(0, 0)), k * R(1, 0), k * R(2, 0));
  }
  double mag;
  const double tr3 = tr - 3.0;
  if (tr3 < -1e-7) { const double th = std::acos(std::min(1.0, std::max(-1.0, 0.5 * (tr - 1.0)))); mag = th / (2.0 * std::sin(th)); }
  else mag = 0.5 - tr3 * tr3 / 12.0;
  return Vector3(mag * (R(2, 1) - R(1, 2)), mag * (R(0, 2) - R(2, 0)), mag * (R(1, 0) - R(0, 1)));
}
inline Matrix3 so3_exp(const Vector3& w) {
  double a, b, c;
  so3_coeff(w.squaredNorm(), a, b, c);
  const Matrix3 W = skew(w);
  return Matrix3(Matrix3::Identity() + W * a + W * W * b);
}
}  // namespace detail

// ------------------------------------------------------------------ Rot3 / Pose3 (A.1)
class Rot3 {
  Matrix3 R_;
 public:
  Rot3() : R_(Matrix3::Identity()) {}
  template <int O, int MR, int MC> Rot3(const Eigen::Matrix<double, 3, 3, O, MR, MC>& m) : R_(m) {}
  Rot3(double r11, double r12, double r13, double r21, double r22, double r23, double r31, double r32, double r33) { R_ << r11, r12, r13, r21, r22, r23, r31, r32, r33; }
  static Rot3 identity() { return Rot3(); }
  static Rot3 RzRyRx(double x, double y, double z) {
    const double cx = std::cos(x), sx = std::sin(x), cy = std::cos(y), sy = std::sin(y), cz = std::cos(z), sz = std::sin(z);
    Matrix3 m;
    m << cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx,
         sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx,
         -sy, cy * sx, cy * cx;
    return Rot3(m);
  }
  static Rot3 RzRyRx(const Vector3& xyz) { return RzRyRx(xyz(0), xyz(1), xyz(2)); }
  static Rot3 Ypr(double y, double p, double r) { return RzRyRx(r, p, y); }
  static Rot3 Rx(double t) { return RzRyRx(t, 0, 0); }
  static Rot3 Ry(double t) { return RzRyRx(0, t, 0); }
  static Rot3 Rz(double t) { return RzRyRx(0, 0, t); }
  static Rot3 Quaternion(double w, double x, double y, double z) { return Rot3(Eigen::Quaterniond(w, x, y, z).toRotationMatrix()); }
  static Rot3 Expmap(const Vector3& w) { return Rot3(detail::so3_exp(w)); }
  static Vector3 Logmap(const Rot3& R) { return detail::so3_log(R.R_); }
  // Rot3::CayleyChart (the Rot3 chart of GTSAM 4.0's default build): the Cayley transform of [w/2]x and its inverse
  static Rot3 CayleyRetract(const Vector3& w) {
    const double x = w(0), y = w(1), z = w(2), x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, xz = x * z, yz = y * z;
    const double f = 1.0 / (4.0 + x2 + y2 + z2), f2 = 2.0 * f;
    return Rot3((4 + x2 - y2 - z2) * f, (xy - 2 * z) * f2, (xz + 2 * y) * f2, (xy + 2 * z) * f2, (4 - x2 + y2 - z2) * f, (yz - 2 * x) * f2,
                (xz - 2 * y) * f2, (yz + 2 * x) * f2, (4 - x2 - y2 + z2) * f);
  }
  static Vector3 CayleyLocal(const Rot3& R) {
    const Matrix3& M = R.R_;
    const double k = 2.0 / (1.0 + M(0, 0) + M(1, 1) + M(2, 2));
    return Vector3(k * (M(2, 1) - M(1, 2)), k * (M(0, 2) - M(2, 0)), k * (M(1, 0) - M(0, 1)));
  }
  const Matrix3& matrix() const { return R_; }
  Matrix3 transpose() const { return R_.transpose(); }
  Rot3 operator*(const Rot3& o) const { return Rot3(Matrix3(R_ * o.R_)); }
  Point3 operator*(const Point3& p) const { return Point3(Vector3(R_ * p)); }
  Vector3 operator*(const Vector3& p) const { return Vector3(R_ * p); }
  Point3 rotate(const Point3& p) const { return Point3(Vector3(R_ * p)); }
  Point3 unrotate(const Point3& p) const { return Point3(Vector3(R_.transpose() * p)); }
  Rot3 inverse() const { return Rot3(Matrix3(R_.transpose())); }
  Rot3 between(const Rot3& o) const { return inverse() * o; }
  Vector3 rpy() const { return Vector3(std::atan2(R_(2, 1), R_(2, 2)), -std::asin(R_(2, 0)), std::atan2(R_(1, 0), R_(0, 0))); }
  Vector3 ypr() const { const Vector3 q = rpy(); return Vector3(q(2), q(1), q(0)); }
  double roll() const { return rpy()(0); }
  double pitch() const { return rpy()(1); }
  double yaw() const { return rpy()(2); }
  Eigen::Quaterniond toQuaternion() const { return Eigen::Quaterniond(R_); }
  Point3 r1() const { return Point3(Vector3(R_.col(0))); }
  Point3 r2() const { return Point3(Vector3(R_.col(1))); }
  Point3 r3() const { return Point3(Vector3(R_.col(2))); }
  void print(const std::string& s = "") const { std::cout << s << "\n" << R_ << std::endl; }
};

class Pose3 {
  Rot3 r_; Point3 t_;
 public:
  Pose3() {}
  Pose3(const Rot3& R, const Point3& T) : r_(R), t_(T) {}
  template <int O, int MR, int MC> explicit Pose3(const Eigen::Matrix<double, 4, 4, O, MR, MC>& m) : r_(Matrix3(m.template block<3, 3>(0, 0))), t_(m(0, 3), m(1, 3), m(2, 3)) {}
  static Pose3 Create(const Rot3& R, const Point3& T) { return Pose3(R, T); }
  static Pose3 identity() { return Pose3(); }
  static Pose3 FromArray12(const double* a) { Matrix3 R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = a[3 * i + j]; return Pose3(Rot3(R), Point3(a[9], a[10], a[11])); }
  void toArray12(double* a) const { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[3 * i + j] = r_.matrix()(i, j); for (int i = 0; i < 3; ++i) a[9 + i] = t_(i); }
  const Rot3& rotation() const { return r_; }
  const Point3& translation() const { return t_; }
  double x() const { return t_(0); } double y() const { return t_(1); } double z() const { return t_(2); }
  Matrix4 matrix() const { Matrix4 m = Matrix4::Identity(); m.block<3, 3>(0, 0) = r_.matrix(); for (int i = 0; i < 3; ++i) m(i, 3) = t_(i); return m; }
  Pose3 operator*(const Pose3& o) const { return Pose3(r_ * o.r_, Point3(Vector3(r_.matrix() * o.t_ + t_))); }
  Point3 operator*(const Point3& p) const { return transform_from(p); }
  Pose3 compose(const Pose3& o) const { return *this * o; }
  Pose3 inverse() const { const Rot3 ri = r_.inverse(); return Pose3(ri, Point3(Vector3(-(ri.matrix() * t_)))); }
  Point3 transform_from(const Point3& p) const { return Point3(Vector3(r_.matrix() * p + t_)); }
  Point3 transform_to(const Point3& p) const { return Point3(Vector3(r_.matrix().transpose() * (p - t_))); }
  Pose3 transform_pose_to(const Pose3& b) const { return inverse() * b; }
  Pose3 between(const Pose3& b) const { return inverse() * b; }
  Matrix6 AdjointMap() const {
    Matrix6 A = Matrix6::Zero();
    const Matrix3 SR = detail::skew(t_) * r_.matrix();
    A.block<3, 3>(0, 0) = r_.matrix(); A.block<3, 3>(3, 3) = r_.matrix(); A.block<3, 3>(3, 0) = SR;
    return A;
  }
  bool equals(const Pose3& o, double tol = 1e-9) const { return (matrix() - o.matrix()).norm() < tol; }
  void print(const std::string& s = "") const { std::cout << s << "\n" << r_.matrix() << "\n[" << t_(0) << ", " << t_(1) << ", " << t_(2) << "]'" << std::endl; }
  // Pose3::Expmap / Logmap, tangent [rot, trans]
  static Pose3 Expmap(const Vector6& xi) {
    const Vector3 w(xi(0), xi(1), xi(2)), v(xi(3), xi(4), xi(5));
    double a, b, c;
    detail::so3_coeff(w.squaredNorm(), a, b, c);
    const Matrix3 W = detail::skew(w), W2 = W * W;
    const Matrix3 Vm = Matrix3::Identity() + W * b + W2 * c;
    return Pose3(Rot3(Matrix3(Matrix3::Identity() + W * a + W2 * b)), Point3(Vector3(Vm * v)));
  }
  static Vector6 Logmap(const Pose3& p) {
    const Vector3 w = detail::so3_log(p.r_.matrix());
    const double th = w.norm();
    Vector6 xi;
    for (int i = 0; i < 3; ++i) xi(i) = w(i);
    if (th < 1e-10) { for (int i = 0; i < 3; ++i) xi(3 + i) = p.t_(i); return xi; }
    const Matrix3 W = detail::skew(Vector3(w / th));
    const Vector3 WT = W * p.t_, WWT = W * WT;
    const double coef = 1.0 - th / (2.0 * std::tan(0.5 * th));
    for (int i = 0; i < 3; ++i) xi(3 + i) = p.t_(i) - 0.5 * th * WT(i) + coef * WWT(i);
    return xi;
  }
  // Pose3::ChartAtOrigin::{Retract, Local}: the full EXPMAP chart (GTSAM_POSE3_EXPMAP build) -- the chart the
  // device uses for BetweenFactor residuals and Values::retract, and the encoding of the VRO log (gtsam_graph.cpp:60,1532)
  struct ChartAtOrigin {
    static Pose3 Retract(const Vector6& xi) {
      if (FG_POSE3_CHART == 0) return Pose3::Expmap(xi);
      const Vector3 w(xi(0), xi(1), xi(2));
      return Pose3(FG_POSE3_CHART == 2 ? Rot3::Expmap(w) : Rot3::CayleyRetract(w), Point3(xi(3), xi(4), xi(5)));
    }
    static Vector6 Local(const Pose3& p) {
      if (FG_POSE3_CHART == 0) return Pose3::Logmap(p);
      const Vector3 w = FG_POSE3_CHART == 2 ? Rot3::Logmap(p.r_) : Rot3::CayleyLocal(p.r_);
      Vector6 xi; for (int i = 0; i < 3; ++i) { xi(i) = w(i); xi(3 + i) = p.t_(i); }
      return xi;
    }
  };
  Pose3 retract(const Vector6& xi) const { return *this * ChartAtOrigin::Retract(xi); }
  Vector6 localCoordinates(const Pose3& o) const { return ChartAtOrigin::Local(between(o)); }
};

// ------------------------------------------------------------------ NavState / ConstantBias / Unit3 / OrientedPlane3
namespace imuBias {
class ConstantBias {
  Vector3 acc_, gyro_;
 public:
  ConstantBias() : acc_(0, 0, 0), gyro_(0, 0, 0) {}
  ConstantBias(const Vector3& a, const Vector3& g) : acc_(a), gyro_(g) {}
  explicit ConstantBias(const Vector6& v) : acc_(v(0), v(1), v(2)), gyro_(v(3), v(4), v(5)) {}
  const Vector3& accelerometer() const { return acc_; }
  const Vector3& gyroscope() const { return gyro_; }
  Vector6 vector() const { Vector6 v; for (int i = 0; i < 3; ++i) { v(i) = acc_(i); v(3 + i) = gyro_(i); } return v; }
  void print(const std::string& s = "") const { std::cout << s << " acc = [" << acc_(0) << " " << acc_(1) << " " << acc_(2) << "] gyro = [" << gyro_(0) << " " << gyro_(1) << " " << gyro_(2) << "]" << std::endl; }
};
}  // namespace imuBias

class NavState {
  Pose3 p_; Vector3 vel_;
 public:
  NavState() : vel_(0, 0, 0) {}
  NavState(const Pose3& pose, const Vector3& v) : p_(pose), vel_(v) {}
  NavState(const Rot3& R, const Point3& t, const Vector3& v) : p_(R, t), vel_(v) {}
  const Pose3& pose() const { return p_; }
  const Rot3& attitude() const { return p_.rotation(); }
  const Point3& position() const { return p_.translation(); }
  const Vector3& v() const { return vel_; }
  const Vector3& velocity() const { return vel_; }
  Matrix3 R() const { return p_.rotation().matrix(); }
  Vector3 t() const { return p_.translation(); }
  void print(const std::string& s = "") const { p_.print(s); std::cout << "v: " << vel_(0) << " " << vel_(1) << " " << vel_(2) << std::endl; }
};

class Unit3 {
  Vector3 p_;
 public:
  Unit3() : p_(1, 0, 0) {}
  Unit3(double x, double y, double z) : p_(x, y, z) { p_.normalize(); }
  explicit Unit3(const Vector3& v) : p_(v) { p_.normalize(); }
  Point3 point3() const { return Point3(p_); }
  Vector3 unitVector() const { return p_; }
  // Unit3::basis(): b1 = normalize(n x axis of the smallest |component| (ties: x, then y)), b2 = n x b1 (A.4)
  Matrix32 basis() const {
    const double mx = std::fabs(p_(0)), my = std::fabs(p_(1)), mz = std::fabs(p_(2));
    Vector3 axis(0, 0, 1);
    if (mx <= my && mx <= mz) axis = Vector3(1, 0, 0);
    else if (my <= mx && my <= mz) axis = Vector3(0, 1, 0);
    Vector3 b1 = p_.cross(axis); b1.normalize();
    const Vector3 b2 = p_.cross(b1);
    Matrix32 B;
    for (int i = 0; i < 3; ++i) { B(i, 0) = b1(i); B(i, 1) = b2(i); }
    return B;
  }
  Vector2 errorVector(const Unit3& q) const { return Vector2(basis().transpose() * q.p_); }
  Vector2 error(const Unit3& q) const { return errorVector(q); }
  Vector2 localCoordinates(const Unit3& q) const {
    const double x = p_.dot(q.p_), z = 1.0 - x * x;
    double y;
    if (z < 2.220446049250313e-16) { if (x > 0) y = 1.0 - (x - 1.0) / 3.0; else return Vector2(M_PI, 0.0); }
    else y = std::acos(std::min(1.0, std::max(-1.0, x))) / std::sqrt(z);
    return Vector2(basis().transpose() * Vector3((q.p_ - p_ * x) * y));
  }
  Unit3 retract(const Vector2& v) const {
    const Vector3 xi = basis() * v;
    const double th = xi.norm();
    if (th < 1e-300) return *this;
    return Unit3(Vector3(p_ * std::cos(th) + xi * (std::sin(th) / th)));
  }
  void print(const std::string& s = "") const { std::cout << s << ":" << p_(0) << " " << p_(1) << " " << p_(2) << std::endl; }
};

class OrientedPlane3 {
  Unit3 n_; double d_;
 public:
  OrientedPlane3() : n_(0, 0, 1), d_(0) {}
  OrientedPlane3(const Unit3& n, double d) : n_(n), d_(d) {}
  OrientedPlane3(double a, double b, double c, double d) : n_(a, b, c), d_(d) {}
  explicit OrientedPlane3(const Vector4& v) : n_(v(0), v(1), v(2)), d_(v(3)) {}
  Vector4 planeCoefficients() const { const Vector3 n = n_.unitVector(); return Vector4(n(0), n(1), n(2), d_); }
  const Unit3& normal() const { return n_; }
  double distance() const { return d_; }
  // n' = R^T n, d' = n.t + d   (gtsam/test/testOrientedPlane3.cpp:61-70)
  OrientedPlane3 transform(const Pose3& xr) const {
    const Vector3 n = n_.unitVector();
    return OrientedPlane3(Unit3(Vector3(xr.rotation().matrix().transpose() * n)), n.dot(xr.translation()) + d_);
  }
  // ... with GTSAM's Jacobians (OrientedPlane3::transform(xr, Hp, Hr); A.4): Hp 3x3 w.r.t. this plane's (2 + 1) tangent,
  // Hr 3x6 w.r.t. the pose [rot, trans]; q = R^T n:  D_q_n = B_q^T R^T B_n,  D_q_R = B_q^T [q]x
  template <int O1, int MR1, int MC1>
  OrientedPlane3 transform(const Pose3& xr, Eigen::Matrix<double, 3, 3, O1, MR1, MC1>& Hp) const {
    const OrientedPlane3 out = transform(xr);
    const Matrix32 Bn = n_.basis(), Bq = out.n_.basis();
    const Matrix22 Dn = Bq.transpose() * xr.rotation().matrix().transpose() * Bn;
    const Vector2 hpp = Bn.transpose() * xr.translation();
    Hp.setZero();
    for (int i = 0; i < 2; ++i) { for (int j = 0; j < 2; ++j) Hp(i, j) = Dn(i, j); Hp(2, i) = hpp(i); }
    Hp(2, 2) = 1.0;
    return out;
  }
  template <int O1, int MR1, int MC1, int O2, int MR2, int MC2>
  OrientedPlane3 transform(const Pose3& xr, Eigen::Matrix<double, 3, 3, O1, MR1, MC1>& Hp, Eigen::Matrix<double, 3, 6, O2, MR2, MC2>& Hr) const {
    const OrientedPlane3 out = transform(xr, Hp);
    const Vector3 q = out.n_.unitVector();
    const Matrix23 Dr = out.n_.basis().transpose() * detail::skew(q);
    Hr.setZero();
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) Hr(i, j) = Dr(i, j);
    for (int j = 0; j < 3; ++j) Hr(2, 3 + j) = q(j);
    return out;
  }
  static OrientedPlane3 Transform(const OrientedPlane3& plane, const Pose3& xr) { return plane.transform(xr); }
  Vector3 errorVector(const OrientedPlane3& o) const { const Vector2 e = n_.errorVector(o.n_); return Vector3(e(0), e(1), d_ - o.d_); }
  Vector3 error(const OrientedPlane3& o) const { const Vector2 e = n_.localCoordinates(o.n_); return Vector3(-e(0), -e(1), d_ - o.d_); }
  OrientedPlane3 retract(const Vector3& v) const { return OrientedPlane3(n_.retract(Vector2(v(0), v(1))), d_ + v(2)); }
  Vector3 localCoordinates(const OrientedPlane3& o) const { const Vector2 e = n_.localCoordinates(o.n_); return Vector3(e(0), e(1), o.d_ - d_); }
  void print(const std::string& s = "") const { const Vector4 c = planeCoefficients(); std::cout << s << " : " << c(0) << " " << c(1) << " " << c(2) << " " << c(3) << std::endl; }
};

// ------------------------------------------------------------------ calibrations
class Cal3DS2 {
 public:
  double K[9];
  typedef boost::shared_ptr<Cal3DS2> shared_ptr;
  Cal3DS2(double fx = 1, double fy = 1, double s = 0, double u0 = 0, double v0 = 0, double k1 = 0, double k2 = 0, double p1 = 0, double p2 = 0) {
    const double k[9] = {fx, fy, s, u0, v0, k1, k2, p1, p2}; for (int i = 0; i < 9; ++i) K[i] = k[i];
  }
  double fx() const { return K[0]; } double fy() const { return K[1]; } double skew() const { return K[2]; }
  double px() const { return K[3]; } double py() const { return K[4]; } double k1() const { return K[5]; } double k2() const { return K[6]; }
};
class Cal3_S2 {
 public:
  double K[9];
  typedef boost::shared_ptr<Cal3_S2> shared_ptr;
  Cal3_S2(double fx = 1, double fy = 1, double s = 0, double u0 = 0, double v0 = 0) { const double k[9] = {fx, fy, s, u0, v0, 0, 0, 0, 0}; for (int i = 0; i < 9; ++i) K[i] = k[i]; }
  double fx() const { return K[0]; } double fy() const { return K[1]; } double px() const { return K[3]; } double py() const { return K[4]; }
};

// ------------------------------------------------------------------ noise models (only Sigma^-1 matters, A.1)
namespace noiseModel {
struct Base {
  std::vector<double> info; int dim_ = 0;   // information matrix, row-major dim x dim
  int dim() const { return dim_; }
  void print(const std::string& s = "") const { std::cout << s << " noise model, dim " << dim_ << std::endl; }
};
typedef boost::shared_ptr<Base> shared_ptr;
inline shared_ptr make(int dim) { shared_ptr b(new Base()); b->dim_ = dim; b->info.assign((size_t)dim * dim, 0.0); return b; }
struct Diagonal : public Base {
  typedef noiseModel::shared_ptr shared_ptr;
  template <int R, int C, int O, int MR, int MC> static shared_ptr Sigmas(const Eigen::Matrix<double, R, C, O, MR, MC>& s) {
    const int n = s.size(); auto b = make(n); for (int i = 0; i < n; ++i) b->info[(size_t)i * n + i] = 1.0 / (s(i) * s(i)); return b;
  }
  template <int R, int C, int O, int MR, int MC> static shared_ptr Variances(const Eigen::Matrix<double, R, C, O, MR, MC>& s) {
    const int n = s.size(); auto b = make(n); for (int i = 0; i < n; ++i) b->info[(size_t)i * n + i] = 1.0 / s(i); return b;
  }
};
struct Isotropic : public Base {
  typedef noiseModel::shared_ptr shared_ptr;
  static shared_ptr Sigma(int dim, double s) { auto b = make(dim); for (int i = 0; i < dim; ++i) b->info[(size_t)i * dim + i] = 1.0 / (s * s); return b; }
  static shared_ptr Variance(int dim, double v) { auto b = make(dim); for (int i = 0; i < dim; ++i) b->info[(size_t)i * dim + i] = 1.0 / v; return b; }
};
struct Gaussian : public Base {
  typedef noiseModel::shared_ptr shared_ptr;
  template <int R, int C, int O, int MR, int MC> static shared_ptr Information(const Eigen::Matrix<double, R, C, O, MR, MC>& m) {
    const int n = m.rows(); auto b = make(n); for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) b->info[(size_t)i * n + j] = m(i, j); return b;
  }
  template <int R, int C, int O, int MR, int MC> static shared_ptr Covariance(const Eigen::Matrix<double, R, C, O, MR, MC>& S) {
    return Information(Eigen::Matrix<double, R, C>(S.inverse()));
  }
};
}  // namespace noiseModel
typedef noiseModel::shared_ptr SharedNoiseModel;

// ------------------------------------------------------------------ Values
struct Value {
  int type = -1;             // FG_T_*
  double v[12] = {0};
};
class ValuesKeyDoesNotExist : public std::runtime_error { public: ValuesKeyDoesNotExist(Key k) : std::runtime_error("ValuesKeyDoesNotExist " + std::to_string(k)) {} };
class ValuesKeyAlreadyExists : public std::runtime_error { public: ValuesKeyAlreadyExists(Key k) : std::runtime_error("ValuesKeyAlreadyExists " + std::to_string(k)) {} };

template <class T> struct ValueTraits;
template <> struct ValueTraits<Pose3> { static const int type = FG_T_POSE; static void put(const Pose3& p, double* a) { p.toArray12(a); } static Pose3 get(const double* a) { return Pose3::FromArray12(a); } };
template <> struct ValueTraits<Vector3> { static const int type = FG_T_VEC3; static void put(const Vector3& p, double* a) { for (int i = 0; i < 3; ++i) a[i] = p(i); } static Vector3 get(const double* a) { return Vector3(a[0], a[1], a[2]); } };
template <> struct ValueTraits<Point3> { static const int type = FG_T_POINT; static void put(const Point3& p, double* a) { for (int i = 0; i < 3; ++i) a[i] = p(i); } static Point3 get(const double* a) { return Point3(a[0], a[1], a[2]); } };
template <> struct ValueTraits<imuBias::ConstantBias> {
  static const int type = FG_T_BIAS;
  static void put(const imuBias::ConstantBias& b, double* a) { for (int i = 0; i < 3; ++i) { a[i] = b.accelerometer()(i); a[3 + i] = b.gyroscope()(i); } }
  static imuBias::ConstantBias get(const double* a) { return imuBias::ConstantBias(Vector3(a[0], a[1], a[2]), Vector3(a[3], a[4], a[5])); }
};
template <> struct ValueTraits<OrientedPlane3> {
  static const int type = FG_T_PLANE;
  static void put(const OrientedPlane3& p, double* a) { const Vector4 c = p.planeCoefficients(); for (int i = 0; i < 4; ++i) a[i] = c(i); }
  static OrientedPlane3 get(const double* a) { return OrientedPlane3(a[0], a[1], a[2], a[3]); }
};

class Values {
 public:
  std::map<Key, Value> m;
  template <class T> void insert(Key k, const T& v) {
    if (m.count(k)) throw ValuesKeyAlreadyExists(k);
    Value x; x.type = ValueTraits<T>::type; ValueTraits<T>::put(v, x.v); m[k] = x;
  }
  template <class T> void update(Key k, const T& v) {
    auto it = m.find(k); if (it == m.end()) throw ValuesKeyDoesNotExist(k);
    ValueTraits<T>::put(v, it->second.v);
  }
  template <class T> T at(Key k) const { auto it = m.find(k); if (it == m.end()) throw ValuesKeyDoesNotExist(k); return ValueTraits<T>::get(it->second.v); }
  bool exists(Key k) const { return m.count(k) != 0; }
  size_t size() const { return m.size(); }
  bool empty() const { return m.empty(); }
  void clear() { m.clear(); }
  void erase(Key k) { m.erase(k); }
  void insert(const Values& o) { for (auto& kv : o.m) { if (m.count(kv.first)) throw ValuesKeyAlreadyExists(kv.first); m[kv.first] = kv.second; } }
  void print(const std::string& s = "") const { std::cout << s << "Values with " << m.size() << " values" << std::endl; }
};

// ------------------------------------------------------------------ factors: each one knows how to add itself through the C ABI
class NonlinearFactor {
 public:
  typedef boost::shared_ptr<NonlinearFactor> shared_ptr;
  virtual ~NonlinearFactor() {}
  virtual int emit(fg_ctx* c) const = 0;
  virtual void print(const std::string& s = "") const { std::cout << s << " factor" << std::endl; }
};
template <class T> class PriorFactor;
template <> class PriorFactor<Pose3> : public NonlinearFactor {
 public:
  Key k; Pose3 prior; noiseModel::shared_ptr nm;
  PriorFactor(Key key, const Pose3& p, const noiseModel::shared_ptr& n) : k(key), prior(p), nm(n) {}
  int emit(fg_ctx* c) const override { double T[12]; prior.toArray12(T); return fg_add_prior_pose(c, k, T, nm->info.data()); }
};
template <> class PriorFactor<Vector3> : public NonlinearFactor {
 public:
  Key k; Vector3 prior; noiseModel::shared_ptr nm;
  PriorFactor(Key key, const Vector3& p, const noiseModel::shared_ptr& n) : k(key), prior(p), nm(n) {}
  int emit(fg_ctx* c) const override { return fg_add_prior_vec3(c, k, prior.data(), nm->info.data()); }
};
template <> class PriorFactor<Point3> : public NonlinearFactor {        // PriorFactor<Point3>(Q(id), p, Isotropic::Sigma(3, s))  gtsam_graph.cpp:379,394
 public:
  Key k; Point3 prior; noiseModel::shared_ptr nm;
  PriorFactor(Key key, const Point3& p, const noiseModel::shared_ptr& n) : k(key), prior(p), nm(n) {}
  int emit(fg_ctx* c) const override { return fg_add_prior_point(c, k, prior.data(), 1.0 / std::sqrt(nm->info[0])); }
};
template <> class PriorFactor<imuBias::ConstantBias> : public NonlinearFactor {
 public:
  Key k; imuBias::ConstantBias prior; noiseModel::shared_ptr nm;
  PriorFactor(Key key, const imuBias::ConstantBias& p, const noiseModel::shared_ptr& n) : k(key), prior(p), nm(n) {}
  int emit(fg_ctx* c) const override { const Vector6 v = prior.vector(); return fg_add_prior_bias(c, k, v.data(), nm->info.data()); }
};
template <class T> class BetweenFactor;
template <> class BetweenFactor<Pose3> : public NonlinearFactor {
 public:
  Key k1, k2; Pose3 z; noiseModel::shared_ptr nm;
  BetweenFactor(Key a, Key b, const Pose3& m, const noiseModel::shared_ptr& n) : k1(a), k2(b), z(m), nm(n) {}
  int emit(fg_ctx* c) const override { double T[12]; z.toArray12(T); return fg_add_between(c, k1, k2, T, nm->info.data()); }
};
class OrientedPlane3Factor : public NonlinearFactor {
 public:
  Key kp, kl; Vector4 z; noiseModel::shared_ptr nm;
  OrientedPlane3Factor() : kp(0), kl(0) {}
  OrientedPlane3Factor(const Vector4& meas, const noiseModel::shared_ptr& n, Key pose, Key lm) : kp(pose), kl(lm), z(meas), nm(n) {}
  int emit(fg_ctx* c) const override {
    // the ABI takes the covariance (gtsam_graph.cpp:1265 passes Gaussian::Covariance); invert the stored information back
    Matrix3 I; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) I(i, j) = nm->info[3 * i + j];
    const Matrix3 cov = I.inverse();
    double cv[9]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) cv[3 * i + j] = cov(i, j);
    return fg_add_plane_factor(c, kp, kl, z.data(), cv);
  }
  // OrientedPlane3Factor::evaluateError / error(Values): 1/2 |transform(plane, pose).error(z)|^2_Sigma  (host arithmetic on two values)
  Vector3 evaluateError(const Pose3& pose, const OrientedPlane3& plane) const { return plane.transform(pose).error(OrientedPlane3(z)); }
  double error(const Values& v) const {
    const Vector3 e = evaluateError(v.at<Pose3>(kp), v.at<OrientedPlane3>(kl));
    double s = 0; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s += e(i) * nm->info[3 * i + j] * e(j);
    return 0.5 * s;
  }
};
// GenericProjectionFactor<Pose3, Point3, Cal3DS2>  (gtsam_graph.cpp:405-406).  Distinct (calibration, body_P_sensor) pairs of
// one graph get distinct ids, so that fg_finalize rejects a graph that mixes them instead of silently using the last one.
namespace detail {
struct CameraSlot { double K[9]; double T[12]; };
inline std::map<fg_ctx*, std::vector<CameraSlot>>& camera_slots() { static std::map<fg_ctx*, std::vector<CameraSlot>> m; return m; }   // erased when the context dies
inline int camera_slot(fg_ctx* c, const double K[9], const double T[12]) {
  std::vector<CameraSlot>& v = camera_slots()[c];
  for (size_t i = 0; i < v.size(); ++i) if (!std::memcmp(v[i].K, K, sizeof v[i].K) && !std::memcmp(v[i].T, T, sizeof v[i].T)) return (int)i;
  CameraSlot s; std::memcpy(s.K, K, sizeof s.K); std::memcpy(s.T, T, sizeof s.T);
  v.push_back(s);
  const int id = (int)v.size() - 1;
  if (fg_set_calibration(c, id, K) != FG_OK || fg_set_sensor(c, id, T) != FG_OK) return -1;
  return id;
}
}  // namespace detail
template <class POSE, class LANDMARK, class CALIBRATION = Cal3_S2>
class GenericProjectionFactor : public NonlinearFactor {
 public:
  Point2 uv; double sigma; Key kp, kq; boost::shared_ptr<CALIBRATION> K; Pose3 body_P_sensor;
  GenericProjectionFactor(const Point2& m, const noiseModel::shared_ptr& n, Key pose, Key point, const boost::shared_ptr<CALIBRATION>& k,
                          bool /*throwCheirality*/ = false, bool /*verbose*/ = false, const Pose3& bPs = Pose3())
      : uv(m), sigma(1.0 / std::sqrt(n->info[0])), kp(pose), kq(point), K(k), body_P_sensor(bPs) {}
  GenericProjectionFactor(const Point2& m, const noiseModel::shared_ptr& n, Key pose, Key point, const boost::shared_ptr<CALIBRATION>& k, const Pose3& bPs)
      : uv(m), sigma(1.0 / std::sqrt(n->info[0])), kp(pose), kq(point), K(k), body_P_sensor(bPs) {}
  int emit(fg_ctx* c) const override {
    double T[12]; body_P_sensor.toArray12(T);
    const int id = detail::camera_slot(c, K->K, T);
    if (id < 0) return FG_ERR_INVALID;
    return fg_add_projection(c, kp, kq, uv.data(), sigma, id, id);
  }
};

// ------------------------------------------------------------------ preintegration (imu_base.cpp:72-99, imu_vn100.cpp:24-67)
// PreintegrationType is GTSAM 4.0's TangentPreintegration (the default build); CImuBase keeps its integrator through a
// pointer to it (gtsam/imu_base.h:72) and the drivers dynamic_cast it to PreintegratedCombinedMeasurements for the factor.
struct PreintegrationCombinedParams {
  Matrix33 accelerometerCovariance, gyroscopeCovariance, integrationCovariance, biasAccCovariance, biasOmegaCovariance;
  Matrix66 biasAccOmegaInt;
  Vector3 n_gravity;
  PreintegrationCombinedParams() : accelerometerCovariance(Matrix33::Identity()), gyroscopeCovariance(Matrix33::Identity()), integrationCovariance(Matrix33::Identity()),
             biasAccCovariance(Matrix33::Identity()), biasOmegaCovariance(Matrix33::Identity()), biasAccOmegaInt(Matrix66::Identity()), n_gravity(0, 0, -9.81) {}
  static boost::shared_ptr<PreintegrationCombinedParams> MakeSharedD(double g = 9.81) { boost::shared_ptr<PreintegrationCombinedParams> p(new PreintegrationCombinedParams()); p->n_gravity = Vector3(0, 0, g); return p; }
  static boost::shared_ptr<PreintegrationCombinedParams> MakeSharedU(double g = 9.81) { boost::shared_ptr<PreintegrationCombinedParams> p(new PreintegrationCombinedParams()); p->n_gravity = Vector3(0, 0, -g); return p; }
  void print(const std::string& s = "") const { std::cout << s << " gravity " << n_gravity(0) << " " << n_gravity(1) << " " << n_gravity(2) << std::endl; }
};
class PreintegrationType {
 public:
  typedef PreintegrationCombinedParams Params;
 protected:
  boost::shared_ptr<Params> p_;
  imuBias::ConstantBias biasHat_;
  std::vector<double> samples_;       // [gx gy gz ax ay az] per integrateMeasurement call
  double dt_ = 0.0;
  mutable fg_pim pim_;
  mutable bool dirty_ = true;
 public:
  PreintegrationType(const boost::shared_ptr<Params>& p, const imuBias::ConstantBias& b = imuBias::ConstantBias()) : p_(p), biasHat_(b) { resetIntegration(); }
  virtual ~PreintegrationType() {}
  Params& p() const { return *p_; }
  const boost::shared_ptr<Params>& params() const { return p_; }
  const imuBias::ConstantBias& biasHat() const { return biasHat_; }
  void resetIntegration() { samples_.clear(); dirty_ = true; }
  void resetIntegrationAndSetBias(const imuBias::ConstantBias& b) { biasHat_ = b; resetIntegration(); }
  // integrateMeasurement(measuredAcc, measuredOmega, dt): samples are queued and integrated on the device on demand
  // (fg_preintegrate: the CImuBase::predictNext sample loop, imu_base.cpp:76-85)
  void integrateMeasurement(const Vector3& acc, const Vector3& omega, double dt) {
    const double s[6] = {omega(0), omega(1), omega(2), acc(0), acc(1), acc(2)};
    samples_.insert(samples_.end(), s, s + 6);
    dt_ = dt; dirty_ = true;
  }
  const fg_pim& pim() const {
    if (dirty_) {
      fg_imu_params ip;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          ip.acc_cov[3 * i + j] = p_->accelerometerCovariance(i, j); ip.gyro_cov[3 * i + j] = p_->gyroscopeCovariance(i, j);
          ip.int_cov[3 * i + j] = p_->integrationCovariance(i, j); ip.bias_acc_cov[3 * i + j] = p_->biasAccCovariance(i, j);
          ip.bias_gyro_cov[3 * i + j] = p_->biasOmegaCovariance(i, j);
        }
      for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) ip.bias_acc_omega_int[6 * i + j] = p_->biasAccOmegaInt(i, j);
      for (int i = 0; i < 3; ++i) ip.gravity[i] = p_->n_gravity(i);
      int off[2] = {0, (int)(samples_.size() / 6)};
      const Vector6 bh = biasHat_.vector();
      double dummy[6] = {0, 0, 0, 0, 0, 0};
      const int rc = fg_preintegrate(nullptr, 1, off, samples_.empty() ? dummy : samples_.data(), dt_ > 0 ? dt_ : 1.0, &ip, bh.data(), &pim_);
      if (rc != FG_OK) throw std::runtime_error("fg_preintegrate failed (no CUDA device? there is no CPU fallback)");
      dirty_ = false;
    }
    return pim_;
  }
  NavState predict(const NavState& s, const imuBias::ConstantBias& b) const {
    double Xi[12], Xj[12], vj[3]; s.pose().toArray12(Xi);
    const Vector6 bv = b.vector();
    const Vector3 vi = s.v();
    fg_pim_predict(&pim(), Xi, vi.data(), bv.data(), Xj, vj);
    return NavState(Pose3::FromArray12(Xj), Vector3(vj[0], vj[1], vj[2]));
  }
  double deltaTij() const { return pim().dt; }
  virtual void print(const std::string& s = "") const { std::cout << s << " preintegrated " << samples_.size() / 6 << " samples, deltaTij " << deltaTij() << std::endl; }
};
typedef PreintegrationType TangentPreintegration;
class PreintegratedCombinedMeasurements : public PreintegrationType {
 public:
  typedef PreintegrationCombinedParams Params;
  PreintegratedCombinedMeasurements(const boost::shared_ptr<Params>& p, const imuBias::ConstantBias& b = imuBias::ConstantBias()) : PreintegrationType(p, b) {}
  Matrix15 preintMeasCov() const { Matrix15 m; const fg_pim& q = pim(); for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) m(i, j) = q.cov[15 * i + j]; return m; }
};

class CombinedImuFactor : public NonlinearFactor {
 public:
  Key k[6]; fg_pim pim;
  CombinedImuFactor(Key pose_i, Key vel_i, Key pose_j, Key vel_j, Key bias_i, Key bias_j, const PreintegratedCombinedMeasurements& p) {
    k[0] = pose_i; k[1] = vel_i; k[2] = pose_j; k[3] = vel_j; k[4] = bias_i; k[5] = bias_j; pim = p.pim();
  }
  int emit(fg_ctx* c) const override { return fg_add_imu(c, k, &pim); }
};

// ------------------------------------------------------------------ graph + optimisers
class NonlinearFactorGraph {
 public:
  std::vector<boost::shared_ptr<NonlinearFactor>> f;
  template <class F> void add(const F& fac) { f.push_back(boost::shared_ptr<NonlinearFactor>(new F(fac))); }
  template <class F> void push_back(const F& fac) { add(fac); }
  template <class F> void push_back(const boost::shared_ptr<F>& fac) { f.push_back(fac); }
  template <class F, class... Args> void emplace_shared(Args&&... args) { f.push_back(boost::shared_ptr<NonlinearFactor>(new F(std::forward<Args>(args)...))); }
  void resize(size_t n) { f.resize(n); }
  size_t size() const { return f.size(); }
  double error(const Values& v) const;
  void print(const std::string& s = "") const { std::cout << s << "NonlinearFactorGraph with " << f.size() << " factors" << std::endl; }
  // graphviz dump (CGraphGT::writeGTSAM, gtsam_graph.cpp:160-171): one node per value, one dot per factor
  void saveGraph(std::ostream& os, const Values& v = Values()) const {
    os << "graph {\n";
    for (auto& kv : v.m) os << "  var" << kv.first << " [label=\"" << Symbol(kv.first).chr() << Symbol(kv.first).index() << "\"];\n";
    for (size_t i = 0; i < f.size(); ++i) os << "  factor" << i << " [shape=point];\n";
    os << "}\n";
  }
};

namespace detail {
inline void check(fg_ctx* c, int rc, const char* what) {
  if (rc != FG_OK) { std::string m = std::string(what) + ": " + fg_last_error(c); throw std::runtime_error(m); }
}
inline void add_value(fg_ctx* c, Key k, const Value& v, const char* what) {
  int rc = FG_OK;
  switch (v.type) {
    case FG_T_POSE: rc = fg_add_pose(c, k, v.v); break;
    case FG_T_VEC3: rc = fg_add_vec3(c, k, v.v); break;
    case FG_T_BIAS: rc = fg_add_bias(c, k, v.v); break;
    case FG_T_POINT: rc = fg_add_point(c, k, v.v); break;
    case FG_T_PLANE: rc = fg_add_plane(c, k, v.v); break;
    default: rc = FG_ERR_INVALID;
  }
  check(c, rc, what);
}
struct Ctx {          // owns a context for the duration of one call
  fg_ctx* c;
  Ctx() : c(fg_create(0, 0, 1)) {
    if (!c) throw std::runtime_error("fg_create failed: no CUDA device (this backend has no CPU fallback)");
    fg_set_pose_chart(c, FG_POSE3_CHART);
  }
  ~Ctx() { if (c) { camera_slots().erase(c); fg_destroy(c); } }
  Ctx(const Ctx&) = delete;
};
inline void build(fg_ctx* c, const NonlinearFactorGraph& g, const Values& v) {
  for (auto& kv : v.m) add_value(c, kv.first, kv.second, "Values -> ctx");
  for (auto& fac : g.f) if (fac) check(c, fac->emit(c), "factor -> ctx");
}
inline void readback(fg_ctx* c, Values& v) {
  for (auto& kv : v.m) { int n = 0; check(c, fg_get_value(c, kv.first, kv.second.v, &n), "fg_get_value"); }
}
}  // namespace detail

inline double NonlinearFactorGraph::error(const Values& v) const {
  detail::Ctx x;
  detail::build(x.c, *this, v);
  double e = 0;
  detail::check(x.c, fg_error(x.c, &e), "fg_error");
  return e;
}

struct LevenbergMarquardtParams {
  double lambdaInitial = 1e-5, lambdaFactor = 10.0, lambdaUpperBound = 1e5, lambdaLowerBound = 0.0, minModelFidelity = 1e-3;
  int maxIterations = 100; double relativeErrorTol = 1e-5, absoluteErrorTol = 1e-5, errorTol = 0.0;
  void setVerbosity(const std::string&) {}
  void setVerbosityLM(const std::string&) {}
  void setMaxIterations(int n) { maxIterations = n; }
  void setRelativeErrorTol(double v) { relativeErrorTol = v; }
  void setAbsoluteErrorTol(double v) { absoluteErrorTol = v; }
  void setlambdaInitial(double v) { lambdaInitial = v; }
};
class LevenbergMarquardtOptimizer {
  const NonlinearFactorGraph& g_; Values v_; LevenbergMarquardtParams p_; fg_lm_report report_;
 public:
  LevenbergMarquardtOptimizer(const NonlinearFactorGraph& graph, const Values& initial, const LevenbergMarquardtParams& p = LevenbergMarquardtParams()) : g_(graph), v_(initial), p_(p) { std::memset(&report_, 0, sizeof report_); }
  const Values& optimize() {
    detail::Ctx x;
    detail::build(x.c, g_, v_);
    fg_lm_params q; fg_lm_params_default(&q);
    q.lambda_initial = p_.lambdaInitial; q.lambda_factor = p_.lambdaFactor; q.lambda_upper = p_.lambdaUpperBound; q.lambda_lower = p_.lambdaLowerBound;
    q.min_model_fidelity = p_.minModelFidelity; q.max_iterations = p_.maxIterations; q.relative_error_tol = p_.relativeErrorTol;
    q.absolute_error_tol = p_.absoluteErrorTol; q.error_tol = p_.errorTol;
    detail::check(x.c, fg_optimize_lm(x.c, &q, &report_), "fg_optimize_lm");
    detail::readback(x.c, v_);
    return v_;
  }
  const Values& values() const { return v_; }
  double error() const { return report_.final_error; }
  int iterations() const { return report_.iterations; }
  const fg_lm_report& report() const { return report_; }
};

// Marginals(graph, values, Marginals::CHOLESKY).marginalCovariance(key)   gtsam/gtsam_graph.cpp:598-601, :1357
// (SURVEY 8 f2).  One undamped device factorisation per call; dimension 6 / 3 / 6 / 3 for a pose / velocity / bias / plane key.
class Marginals {
 public:
  enum Factorization { CHOLESKY, QR };
 private:
  const NonlinearFactorGraph& g_; Values v_;
 public:
  Marginals(const NonlinearFactorGraph& graph, const Values& solution, Factorization = CHOLESKY) : g_(graph), v_(solution) {}
  Matrix marginalCovariance(Key key) const {
    detail::Ctx x;
    detail::build(x.c, g_, v_);
    double cov[36]; int d = 0;
    detail::check(x.c, fg_marginal_cov(x.c, key, cov, &d), "fg_marginal_cov");
    Matrix m(d, d);
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) m(i, j) = cov[i * d + j];
    return m;
  }
  Matrix marginalInformation(Key key) const { return Matrix(marginalCovariance(key).inverse()); }
};

// ISAM2 (gtsam/gtsam_graph.cpp:93-99,1768-1776): one persistent device context.  update() hands the new
// values and factors to it and runs fg_update_incremental -- new variables enter at their initial value, variables whose
// delta reaches relinearizeThreshold move their linearisation point, one undamped Gauss-Newton system is solved on the
// device; calculateEstimate() reads theta (+) delta back.  (A full re-factorisation per update instead of the partial
// Bayes-tree re-elimination: see include/fg_abi.h.)
class ISAM2Params { public: double relinearizeThreshold = 0.1; int relinearizeSkip = 10; void print(const std::string& s = "") const { std::cout << s << " relinearizeThreshold " << relinearizeThreshold << " relinearizeSkip " << relinearizeSkip << std::endl; } };
struct ISAM2Result { size_t variablesRelinearized = 0, variablesReeliminated = 0; double errorBefore = 0, errorAfter = 0; };
class ISAM2 {
 public:
  ISAM2Params params; Values estimate; fg_inc_report report;
  ISAM2() { std::memset(&report, 0, sizeof report); }
  explicit ISAM2(const ISAM2Params& p) : params(p) { std::memset(&report, 0, sizeof report); }
  ISAM2(const ISAM2&) = delete;
  ISAM2& operator=(const ISAM2&) = delete;
  ~ISAM2() { if (c_) { detail::camera_slots().erase(c_); fg_destroy(c_); } }
  ISAM2Result update(const NonlinearFactorGraph& nf = NonlinearFactorGraph(), const Values& nv = Values()) {
    if (!c_) {
      c_ = fg_create(0, 0, 1);
      if (!c_) throw std::runtime_error("fg_create failed: no CUDA device (this backend has no CPU fallback)");
      fg_set_pose_chart(c_, FG_POSE3_CHART);
    }
    for (auto& kv : nv.m) detail::add_value(c_, kv.first, kv.second, "ISAM2::update (new value)");
    estimate.insert(nv);
    for (auto& fac : nf.f) if (fac) detail::check(c_, fac->emit(c_), "ISAM2::update (new factor)");
    fg_isam2_params p; p.relinearize_threshold = params.relinearizeThreshold; p.relinearize_skip = params.relinearizeSkip;
    detail::check(c_, fg_update_incremental(c_, &p, &report), "fg_update_incremental");
    fresh_ = false;
    ISAM2Result r; r.variablesRelinearized = (size_t)report.n_relinearized; r.variablesReeliminated = (size_t)report.n_variables;
    r.errorBefore = report.error_before; r.errorAfter = report.error_after;
    return r;
  }
  Values calculateEstimate() {
    if (!fresh_ && c_) { detail::readback(c_, estimate); fresh_ = true; }
    return estimate;
  }
  template <class T> T calculateEstimate(Key k) { return calculateEstimate().at<T>(k); }
 private:
  fg_ctx* c_ = nullptr;
  bool fresh_ = true;
};

// writeG2o(graph, values, file)  (CGraphGT::writeG2O, gtsam_graph.cpp:1941-1945): VERTEX_SE3:QUAT / EDGE_SE3:QUAT lines;
// g2o's tangent order is [trans, rot], GTSAM's [rot, trans], so the information blocks are swapped on the way out.
inline void writeG2o(const NonlinearFactorGraph& g, const Values& v, const std::string& file) {
  std::ofstream os(file.c_str());
  os.precision(17);
  for (auto& kv : v.m) {
    if (kv.second.type != FG_T_POSE) continue;
    const Pose3 p = Pose3::FromArray12(kv.second.v);
    const Eigen::Quaterniond q = p.rotation().toQuaternion();
    os << "VERTEX_SE3:QUAT " << Symbol(kv.first).index() << " " << p.x() << " " << p.y() << " " << p.z() << " " << q.x() << " " << q.y() << " " << q.z() << " " << q.w() << "\n";
  }
  for (auto& fac : g.f) {
    const BetweenFactor<Pose3>* b = dynamic_cast<const BetweenFactor<Pose3>*>(fac.get());
    if (!b) continue;
    const Eigen::Quaterniond q = b->z.rotation().toQuaternion();
    os << "EDGE_SE3:QUAT " << Symbol(b->k1).index() << " " << Symbol(b->k2).index() << " " << b->z.x() << " " << b->z.y() << " " << b->z.z() << " "
       << q.x() << " " << q.y() << " " << q.z() << " " << q.w();
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) { const int a = (i + 3) % 6, c = (j + 3) % 6; os << " " << b->nm->info[(size_t)a * 6 + c]; }
    os << "\n";
  }
}

}  // namespace gtsam
