// gtsam_lite.h -- header-only subset of the gtsam:: types the reference's wrapper and drivers use
// (SURVEY.md section 8b item 2), as host-side VALUE types that forward all arithmetic-heavy work to the C ABI
// (include/fg_abi.h).  Nothing here solves anything on the CPU: LevenbergMarquardtOptimizer::optimize,
// NonlinearFactorGraph::error and PreintegratedCombinedMeasurements run on the GPU through libfg_b200.so.
//
// Mirrors (reference file:line of the usage):
//   Pose3/Rot3/Point3/NavState/imuBias::ConstantBias   gtsam/gtsam_graph.cpp:320-368,613-695
//   Values insert/update/at/exists                      gtsam/gtsam_graph.cpp:333,619-623,632-666
//   NonlinearFactorGraph add/push_back/resize/error     gtsam/gtsam_graph.cpp:173-176,341,691-692,1773
//   PriorFactor / BetweenFactor / CombinedImuFactor / OrientedPlane3Factor / GenericProjectionFactor
//   LevenbergMarquardtOptimizer(graph, values).optimize()   gtsam/gtsam_graph.cpp:1786-1787
//   ISAM2::update + calculateEstimate                   gtsam/gtsam_graph.cpp:1770-1772  (batch stand-in, see below)
//   PreintegratedCombinedMeasurements (+Params::MakeSharedD)  gtsam/imu_base.cpp:83-98,258-263; imu_vn100.cpp:16,57-62
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/fg_abi.h"

namespace gtsam {

typedef uint64_t Key;

// ------------------------------------------------------------------ tiny fixed-size matrices (row-major)
template <int R, int C>
struct Mat {
  double d[R * C];
  Mat() { for (int i = 0; i < R * C; ++i) d[i] = 0.0; }
  double& operator()(int r, int c) { return d[r * C + c]; }
  double operator()(int r, int c) const { return d[r * C + c]; }
  double& operator()(int i) { return d[i]; }
  double operator()(int i) const { return d[i]; }
  double& operator[](int i) { return d[i]; }
  double operator[](int i) const { return d[i]; }
  static Mat Zero() { return Mat(); }
  static Mat Identity() { Mat m; for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0; return m; }
  Mat<C, R> transpose() const { Mat<C, R> t; for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c); return t; }
  Mat operator+(const Mat& o) const { Mat m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] + o.d[i]; return m; }
  Mat operator-(const Mat& o) const { Mat m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] - o.d[i]; return m; }
  Mat operator*(double s) const { Mat m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] * s; return m; }
  double norm() const { double s = 0; for (int i = 0; i < R * C; ++i) s += d[i] * d[i]; return std::sqrt(s); }
  const double* data() const { return d; }
  double* data() { return d; }
};
template <int R, int K, int C>
Mat<R, C> operator*(const Mat<R, K>& a, const Mat<K, C>& b) {
  Mat<R, C> m;
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) { double s = 0; for (int k = 0; k < K; ++k) s += a(r, k) * b(k, c); m(r, c) = s; }
  return m;
}
typedef Mat<3, 1> Vector3;
typedef Mat<4, 1> Vector4;
typedef Mat<6, 1> Vector6;
typedef Mat<9, 1> Vector9;
typedef Mat<3, 3> Matrix3;
typedef Mat<3, 3> Matrix33;
typedef Mat<4, 4> Matrix4;
typedef Mat<6, 6> Matrix6;
typedef Mat<6, 6> Matrix66;
typedef Mat<15, 15> Matrix15;
typedef Vector3 Point3;
typedef Mat<2, 1> Point2;
inline Vector3 vec3(double x, double y, double z) { Vector3 v; v[0] = x; v[1] = y; v[2] = z; return v; }

// ------------------------------------------------------------------ Symbol (gtsam_graph.cpp:50-54)
inline Key Symbol(unsigned char c, uint64_t j) { return (uint64_t(c) << 56) | j; }
namespace symbol_shorthand {
inline Key X(uint64_t j) { return Symbol('x', j); }
inline Key V(uint64_t j) { return Symbol('v', j); }
inline Key B(uint64_t j) { return Symbol('b', j); }
inline Key L(uint64_t j) { return Symbol('l', j); }
inline Key Q(uint64_t j) { return Symbol('q', j); }
}  // namespace symbol_shorthand

// ------------------------------------------------------------------ Rot3 / Pose3 (A.1)
class Rot3 {
 public:
  Matrix3 R;
  Rot3() : R(Matrix3::Identity()) {}
  explicit Rot3(const Matrix3& m) : R(m) {}
  static Rot3 RzRyRx(double x, double y, double z) {
    double cx = std::cos(x), sx = std::sin(x), cy = std::cos(y), sy = std::sin(y), cz = std::cos(z), sz = std::sin(z);
    Matrix3 m;
    m(0, 0) = cz * cy; m(0, 1) = cz * sy * sx - sz * cx; m(0, 2) = cz * sy * cx + sz * sx;
    m(1, 0) = sz * cy; m(1, 1) = sz * sy * sx + cz * cx; m(1, 2) = sz * sy * cx - cz * sx;
    m(2, 0) = -sy;     m(2, 1) = cy * sx;                m(2, 2) = cy * cx;
    return Rot3(m);
  }
  static Rot3 Ypr(double y, double p, double r) { return RzRyRx(r, p, y); }
  const Matrix3& matrix() const { return R; }
  Rot3 operator*(const Rot3& o) const { return Rot3(R * o.R); }
  Vector3 operator*(const Vector3& p) const { return R * p; }
  Rot3 inverse() const { return Rot3(R.transpose()); }
  Vector3 rpy() const { return vec3(std::atan2(R(2, 1), R(2, 2)), -std::asin(R(2, 0)), std::atan2(R(1, 0), R(0, 0))); }
};

class Pose3 {
 public:
  Rot3 r; Point3 t;
  Pose3() {}
  Pose3(const Rot3& R, const Point3& T) : r(R), t(T) {}
  explicit Pose3(const Matrix4& m) { for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) r.R(i, j) = m(i, j); t[i] = m(i, 3); } }
  static Pose3 Create(const Rot3& R, const Point3& T) { return Pose3(R, T); }
  static Pose3 FromArray12(const double* a) { Pose3 p; for (int i = 0; i < 9; ++i) p.r.R.d[i] = a[i]; for (int i = 0; i < 3; ++i) p.t[i] = a[9 + i]; return p; }
  void toArray12(double* a) const { for (int i = 0; i < 9; ++i) a[i] = r.R.d[i]; for (int i = 0; i < 3; ++i) a[9 + i] = t[i]; }
  const Rot3& rotation() const { return r; }
  const Point3& translation() const { return t; }
  double x() const { return t[0]; } double y() const { return t[1]; } double z() const { return t[2]; }
  Matrix4 matrix() const { Matrix4 m = Matrix4::Identity(); for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) m(i, j) = r.R(i, j); m(i, 3) = t[i]; } return m; }
  Pose3 operator*(const Pose3& o) const { return Pose3(r * o.r, r.R * o.t + t); }
  Point3 operator*(const Point3& p) const { return transform_from(p); }
  Pose3 inverse() const { Rot3 ri = r.inverse(); return Pose3(ri, (ri.R * t) * -1.0); }
  Point3 transform_from(const Point3& p) const { return r.R * p + t; }
  Point3 transform_to(const Point3& p) const { return r.R.transpose() * (p - t); }
  Pose3 transform_pose_to(const Pose3& b) const { return inverse() * b; }
  Pose3 between(const Pose3& b) const { return inverse() * b; }
  Matrix6 AdjointMap() const {
    Matrix6 A; Matrix3 S;
    S(0, 1) = -t[2]; S(0, 2) = t[1]; S(1, 0) = t[2]; S(1, 2) = -t[0]; S(2, 0) = -t[1]; S(2, 1) = t[0];
    Matrix3 SR = S * r.R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A(i, j) = r.R(i, j); A(3 + i, 3 + j) = r.R(i, j); A(3 + i, j) = SR(i, j); }
    return A;
  }
  void print(const std::string& s = "") const { printf("%s t = [%g %g %g]\n", s.c_str(), t[0], t[1], t[2]); }
  // Pose3::ChartAtOrigin::{Retract, Local}: full EXPMAP chart, tangent [rot, trans]
  struct ChartAtOrigin {
    static Pose3 Retract(const Vector6& xi);
    static Vector6 Local(const Pose3& p);
  };
};

namespace detail {
inline void so3_coeff(double th2, double& a, double& b, double& c) {
  if (th2 < 1e-10) { a = 1 - th2 / 6; b = 0.5 - th2 / 24; c = 1.0 / 6 - th2 / 120; }
  else { double t = std::sqrt(th2), s = std::sin(t), sh = std::sin(0.5 * t); a = s / t; b = 2 * sh * sh / th2; c = (t - s) / (th2 * t); }
}
inline Matrix3 skew(const Vector3& w) { Matrix3 S; S(0, 1) = -w[2]; S(0, 2) = w[1]; S(1, 0) = w[2]; S(1, 2) = -w[0]; S(2, 0) = -w[1]; S(2, 1) = w[0]; return S; }
}  // namespace detail

inline Pose3 Pose3::ChartAtOrigin::Retract(const Vector6& xi) {
  Vector3 w = vec3(xi[0], xi[1], xi[2]), v = vec3(xi[3], xi[4], xi[5]);
  double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], a, b, c;
  detail::so3_coeff(th2, a, b, c);
  Matrix3 W = detail::skew(w), W2 = W * W;
  Matrix3 R = Matrix3::Identity() + W * a + W2 * b;
  Matrix3 Vm = Matrix3::Identity() + W * b + W2 * c;
  return Pose3(Rot3(R), Vm * v);
}
inline Vector6 Pose3::ChartAtOrigin::Local(const Pose3& p) {
  const Matrix3& R = p.r.R;
  double tr = R(0, 0) + R(1, 1) + R(2, 2);
  Vector3 w;
  double tr3 = tr - 3.0, mag;
  if (tr3 < -1e-7) { double cth = std::min(1.0, std::max(-1.0, 0.5 * (tr - 1.0))), th = std::acos(cth); mag = th / (2 * std::sin(th)); }
  else mag = 0.5 - tr3 * tr3 / 12.0;
  w = vec3(mag * (R(2, 1) - R(1, 2)), mag * (R(0, 2) - R(2, 0)), mag * (R(1, 0) - R(0, 1)));
  double th = w.norm();
  Vector6 xi;
  for (int i = 0; i < 3; ++i) xi[i] = w[i];
  if (th < 1e-10) { for (int i = 0; i < 3; ++i) xi[3 + i] = p.t[i]; return xi; }
  Matrix3 W = detail::skew(w * (1.0 / th));
  Vector3 WT = W * p.t, WWT = W * WT;
  double coef = 1.0 - th / (2.0 * std::tan(0.5 * th));
  for (int i = 0; i < 3; ++i) xi[3 + i] = p.t[i] - 0.5 * th * WT[i] + coef * WWT[i];
  return xi;
}

// ------------------------------------------------------------------ NavState / ConstantBias / OrientedPlane3
namespace imuBias {
class ConstantBias {
 public:
  Vector3 acc, gyro;
  ConstantBias() {}
  ConstantBias(const Vector3& a, const Vector3& g) : acc(a), gyro(g) {}
  const Vector3& accelerometer() const { return acc; }
  const Vector3& gyroscope() const { return gyro; }
  Vector6 vector() const { Vector6 v; for (int i = 0; i < 3; ++i) { v[i] = acc[i]; v[3 + i] = gyro[i]; } return v; }
  void print(const std::string& s = "") const { printf("%s acc [%g %g %g] gyro [%g %g %g]\n", s.c_str(), acc[0], acc[1], acc[2], gyro[0], gyro[1], gyro[2]); }
};
}  // namespace imuBias

class NavState {
 public:
  Pose3 p; Vector3 vel;
  NavState() {}
  NavState(const Pose3& pose, const Vector3& v) : p(pose), vel(v) {}
  const Pose3& pose() const { return p; }
  const Vector3& v() const { return vel; }
  const Vector3& velocity() const { return vel; }
};

class OrientedPlane3 {
 public:
  Vector3 n; double d;
  OrientedPlane3() : d(0) { n[2] = 1; }
  OrientedPlane3(double a, double b, double c, double dd) : d(dd) { double s = std::sqrt(a * a + b * b + c * c); n = vec3(a / s, b / s, c / s); }
  explicit OrientedPlane3(const Vector4& v) : OrientedPlane3(v[0], v[1], v[2], v[3]) {}
  Vector4 planeCoefficients() const { Vector4 v; v[0] = n[0]; v[1] = n[1]; v[2] = n[2]; v[3] = d; return v; }
  double distance() const { return d; }
  // n' = R^T n, d' = n.t + d   (gtsam/test/testOrientedPlane3.cpp:61-70)
  OrientedPlane3 transform(const Pose3& xr) const {
    Vector3 q = xr.r.R.transpose() * n;
    return OrientedPlane3(q[0], q[1], q[2], n[0] * xr.t[0] + n[1] * xr.t[1] + n[2] * xr.t[2] + d);
  }
};

// ------------------------------------------------------------------ noise models (only Sigma^-1 matters, A.1)
namespace noiseModel {
struct Base { std::vector<double> info; int dim = 0; };   // information matrix, row-major dim x dim
typedef std::shared_ptr<Base> shared_ptr;
inline shared_ptr make(int dim) { auto b = std::make_shared<Base>(); b->dim = dim; b->info.assign(dim * dim, 0.0); return b; }
struct Diagonal {
  typedef noiseModel::shared_ptr shared_ptr;
  template <int N> static shared_ptr Sigmas(const Mat<N, 1>& s) { auto b = make(N); for (int i = 0; i < N; ++i) b->info[i * N + i] = 1.0 / (s[i] * s[i]); return b; }
};
struct Isotropic {
  typedef noiseModel::shared_ptr shared_ptr;
  static shared_ptr Sigma(int dim, double s) { auto b = make(dim); for (int i = 0; i < dim; ++i) b->info[i * dim + i] = 1.0 / (s * s); return b; }
};
struct Gaussian {
  typedef noiseModel::shared_ptr shared_ptr;
  template <int N> static shared_ptr Information(const Mat<N, N>& m) { auto b = make(N); for (int i = 0; i < N * N; ++i) b->info[i] = m.d[i]; return b; }
  // Covariance(S): the inverse is taken inside the C ABI where a covariance is what it accepts (plane factor),
  // otherwise here by Gauss-Jordan on the small matrix.
  template <int N> static shared_ptr Covariance(const Mat<N, N>& S) {
    double a[N][2 * N];
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) { a[i][j] = S(i, j); a[i][N + j] = (i == j); }
    for (int c = 0; c < N; ++c) {
      int piv = c; for (int r = c + 1; r < N; ++r) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
      for (int j = 0; j < 2 * N; ++j) std::swap(a[c][j], a[piv][j]);
      double inv = 1.0 / a[c][c];
      for (int j = 0; j < 2 * N; ++j) a[c][j] *= inv;
      for (int r = 0; r < N; ++r) if (r != c) { double f = a[r][c]; for (int j = 0; j < 2 * N; ++j) a[r][j] -= f * a[c][j]; }
    }
    auto b = make(N); for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) b->info[i * N + j] = a[i][N + j];
    return b;
  }
};
}  // namespace noiseModel

// ------------------------------------------------------------------ Values
struct Value {
  int type = -1;             // FG_T_*
  double v[12] = {0};
};
class ValuesKeyDoesNotExist : public std::runtime_error { public: ValuesKeyDoesNotExist(Key k) : std::runtime_error("ValuesKeyDoesNotExist " + std::to_string(k)) {} };
class ValuesKeyAlreadyExists : public std::runtime_error { public: ValuesKeyAlreadyExists(Key k) : std::runtime_error("ValuesKeyAlreadyExists " + std::to_string(k)) {} };

template <class T> struct ValueTraits;
template <> struct ValueTraits<Pose3> { static const int type = FG_T_POSE; static void put(const Pose3& p, double* a) { p.toArray12(a); } static Pose3 get(const double* a) { return Pose3::FromArray12(a); } };
template <> struct ValueTraits<Vector3> { static const int type = FG_T_VEC3; static void put(const Vector3& p, double* a) { for (int i = 0; i < 3; ++i) a[i] = p[i]; } static Vector3 get(const double* a) { return vec3(a[0], a[1], a[2]); } };
template <> struct ValueTraits<imuBias::ConstantBias> { static const int type = FG_T_BIAS; static void put(const imuBias::ConstantBias& b, double* a) { for (int i = 0; i < 3; ++i) { a[i] = b.acc[i]; a[3 + i] = b.gyro[i]; } } static imuBias::ConstantBias get(const double* a) { return imuBias::ConstantBias(vec3(a[0], a[1], a[2]), vec3(a[3], a[4], a[5])); } };
template <> struct ValueTraits<OrientedPlane3> { static const int type = FG_T_PLANE; static void put(const OrientedPlane3& p, double* a) { a[0] = p.n[0]; a[1] = p.n[1]; a[2] = p.n[2]; a[3] = p.d; } static OrientedPlane3 get(const double* a) { return OrientedPlane3(a[0], a[1], a[2], a[3]); } };
// Point3 is a typedef of Vector3 here; points are inserted with insertPoint / at via atPoint.

class Values {
 public:
  std::map<Key, Value> m;
  template <class T> void insert(Key k, const T& v) {
    if (m.count(k)) throw ValuesKeyAlreadyExists(k);
    Value x; x.type = ValueTraits<T>::type; ValueTraits<T>::put(v, x.v); m[k] = x;
  }
  void insertPoint(Key k, const Point3& p) { if (m.count(k)) throw ValuesKeyAlreadyExists(k); Value x; x.type = FG_T_POINT; for (int i = 0; i < 3; ++i) x.v[i] = p[i]; m[k] = x; }
  template <class T> void update(Key k, const T& v) {
    auto it = m.find(k); if (it == m.end()) throw ValuesKeyDoesNotExist(k);
    ValueTraits<T>::put(v, it->second.v);
  }
  template <class T> T at(Key k) const { auto it = m.find(k); if (it == m.end()) throw ValuesKeyDoesNotExist(k); return ValueTraits<T>::get(it->second.v); }
  Point3 atPoint(Key k) const { auto it = m.find(k); if (it == m.end()) throw ValuesKeyDoesNotExist(k); return vec3(it->second.v[0], it->second.v[1], it->second.v[2]); }
  bool exists(Key k) const { return m.count(k) != 0; }
  size_t size() const { return m.size(); }
  void clear() { m.clear(); }
  void insert(const Values& o) { for (auto& kv : o.m) { if (m.count(kv.first)) throw ValuesKeyAlreadyExists(kv.first); m[kv.first] = kv.second; } }
};

// ------------------------------------------------------------------ factors: each one knows how to add itself through the C ABI
class NonlinearFactor {
 public:
  virtual ~NonlinearFactor() {}
  virtual int emit(fg_ctx* c) const = 0;
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
  Key k; Vector3 prior; noiseModel::shared_ptr nm; bool is_point;
  PriorFactor(Key key, const Vector3& p, const noiseModel::shared_ptr& n, bool point = false) : k(key), prior(p), nm(n), is_point(point) {}
  int emit(fg_ctx* c) const override {
    if (is_point) return fg_add_prior_point(c, k, prior.d, 1.0 / std::sqrt(nm->info[0]));
    return fg_add_prior_vec3(c, k, prior.d, nm->info.data());
  }
};
template <> class PriorFactor<imuBias::ConstantBias> : public NonlinearFactor {
 public:
  Key k; imuBias::ConstantBias prior; noiseModel::shared_ptr nm;
  PriorFactor(Key key, const imuBias::ConstantBias& p, const noiseModel::shared_ptr& n) : k(key), prior(p), nm(n) {}
  int emit(fg_ctx* c) const override { Vector6 v = prior.vector(); return fg_add_prior_bias(c, k, v.d, nm->info.data()); }
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
  OrientedPlane3Factor(const Vector4& meas, const noiseModel::shared_ptr& n, Key pose, Key lm) : kp(pose), kl(lm), z(meas), nm(n) {}
  int emit(fg_ctx* c) const override {
    // the ABI takes the covariance (gtsam_graph.cpp:1265 passes Gaussian::Covariance); invert the stored information back
    Mat<3, 3> I; for (int i = 0; i < 9; ++i) I.d[i] = nm->info[i];
    auto cov = noiseModel::Gaussian::Covariance(I);
    return fg_add_plane_factor(c, kp, kl, z.d, cov->info.data());
  }
};
struct Cal3DS2 {
  double K[9];
  Cal3DS2(double fx, double fy, double s, double u0, double v0, double k1, double k2, double p1 = 0, double p2 = 0) {
    double k[9] = {fx, fy, s, u0, v0, k1, k2, p1, p2}; for (int i = 0; i < 9; ++i) K[i] = k[i];
  }
};
class GenericProjectionFactor : public NonlinearFactor {   // <Pose3, Point3, Cal3DS2>
 public:
  Point2 uv; double sigma; Key kp, kq; std::shared_ptr<Cal3DS2> K; Pose3 body_P_sensor;
  GenericProjectionFactor(const Point2& m, const noiseModel::shared_ptr& n, Key pose, Key point, const std::shared_ptr<Cal3DS2>& k,
                          bool /*throwCheirality*/, bool /*verbose*/, const Pose3& bPs)
      : uv(m), sigma(1.0 / std::sqrt(n->info[0])), kp(pose), kq(point), K(k), body_P_sensor(bPs) {}
  int emit(fg_ctx* c) const override {
    double T[12]; body_P_sensor.toArray12(T);
    int rc = fg_set_calibration(c, 0, K->K); if (rc) return rc;
    rc = fg_set_sensor(c, 0, T); if (rc) return rc;
    return fg_add_projection(c, kp, kq, uv.d, sigma, 0, 0);
  }
};

// ------------------------------------------------------------------ preintegration (imu_base.cpp:72-99, imu_vn100.cpp:24-67)
class PreintegrationType { public: virtual ~PreintegrationType() {} };
class PreintegratedCombinedMeasurements : public PreintegrationType {
 public:
  struct Params {
    Matrix33 accelerometerCovariance, gyroscopeCovariance, integrationCovariance, biasAccCovariance, biasOmegaCovariance;
    Matrix66 biasAccOmegaInt;
    Vector3 n_gravity;
    static std::shared_ptr<Params> MakeSharedD(double g = 9.81) { auto p = std::make_shared<Params>(); p->n_gravity = vec3(0, 0, g); return p; }
    static std::shared_ptr<Params> MakeSharedU(double g = 9.81) { auto p = std::make_shared<Params>(); p->n_gravity = vec3(0, 0, -g); return p; }
  };
  std::shared_ptr<Params> p_;
  imuBias::ConstantBias biasHat_;
  std::vector<double> samples_;       // [gx gy gz ax ay az] per integrateMeasurement call
  double dt_ = 0.0;
  mutable fg_pim pim_;
  mutable bool dirty_ = true;
  PreintegratedCombinedMeasurements(const std::shared_ptr<Params>& p, const imuBias::ConstantBias& b) : p_(p), biasHat_(b) { resetIntegration(); }
  void resetIntegration() { samples_.clear(); dirty_ = true; }
  void resetIntegrationAndSetBias(const imuBias::ConstantBias& b) { biasHat_ = b; resetIntegration(); }
  // integrateMeasurement(measuredAcc, measuredOmega, dt): samples are queued and integrated on the device on demand
  void integrateMeasurement(const Vector3& acc, const Vector3& omega, double dt) {
    const double s[6] = {omega[0], omega[1], omega[2], acc[0], acc[1], acc[2]};
    samples_.insert(samples_.end(), s, s + 6);
    dt_ = dt; dirty_ = true;
  }
  const fg_pim& pim() const {
    if (dirty_) {
      fg_imu_params ip;
      for (int i = 0; i < 9; ++i) { ip.acc_cov[i] = p_->accelerometerCovariance.d[i]; ip.gyro_cov[i] = p_->gyroscopeCovariance.d[i]; ip.int_cov[i] = p_->integrationCovariance.d[i]; ip.bias_acc_cov[i] = p_->biasAccCovariance.d[i]; ip.bias_gyro_cov[i] = p_->biasOmegaCovariance.d[i]; }
      for (int i = 0; i < 36; ++i) ip.bias_acc_omega_int[i] = p_->biasAccOmegaInt.d[i];
      for (int i = 0; i < 3; ++i) ip.gravity[i] = p_->n_gravity[i];
      int off[2] = {0, (int)(samples_.size() / 6)};
      Vector6 bh = biasHat_.vector();
      double dummy[6] = {0, 0, 0, 0, 0, 0};
      int rc = fg_preintegrate(nullptr, 1, off, samples_.empty() ? dummy : samples_.data(), dt_ > 0 ? dt_ : 1.0, &ip, bh.d, &pim_);
      if (rc != FG_OK) throw std::runtime_error("fg_preintegrate failed (no CUDA device? there is no CPU fallback)");
      dirty_ = false;
    }
    return pim_;
  }
  NavState predict(const NavState& s, const imuBias::ConstantBias& b) const {
    double Xi[12], Xj[12], vj[3]; s.pose().toArray12(Xi);
    Vector6 bv = b.vector();
    fg_pim_predict(&pim(), Xi, s.v().d, bv.d, Xj, vj);
    return NavState(Pose3::FromArray12(Xj), vec3(vj[0], vj[1], vj[2]));
  }
  Matrix15 preintMeasCov() const { Matrix15 m; const fg_pim& q = pim(); for (int i = 0; i < 225; ++i) m.d[i] = q.cov[i]; return m; }
  double deltaTij() const { return pim().dt; }
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
  std::vector<std::shared_ptr<NonlinearFactor>> f;
  template <class F> void add(const F& fac) { f.push_back(std::make_shared<F>(fac)); }
  template <class F> void push_back(const F& fac) { add(fac); }
  void resize(size_t n) { f.resize(n); }
  size_t size() const { return f.size(); }
  double error(const Values& v) const;
};

namespace detail {
inline void check(fg_ctx* c, int rc, const char* what) {
  if (rc != FG_OK) { std::string m = std::string(what) + ": " + fg_last_error(c); throw std::runtime_error(m); }
}
inline fg_ctx* build(const NonlinearFactorGraph& g, const Values& v) {
  fg_ctx* c = fg_create(0, 0, 1);
  if (!c) throw std::runtime_error("fg_create failed: no CUDA device (this backend has no CPU fallback)");
  for (auto& kv : v.m) {
    int rc = FG_OK;
    switch (kv.second.type) {
      case FG_T_POSE: rc = fg_add_pose(c, kv.first, kv.second.v); break;
      case FG_T_VEC3: rc = fg_add_vec3(c, kv.first, kv.second.v); break;
      case FG_T_BIAS: rc = fg_add_bias(c, kv.first, kv.second.v); break;
      case FG_T_POINT: rc = fg_add_point(c, kv.first, kv.second.v); break;
      case FG_T_PLANE: rc = fg_add_plane(c, kv.first, kv.second.v); break;
    }
    if (rc != FG_OK) { std::string m = fg_last_error(c); fg_destroy(c); throw std::runtime_error("Values -> ctx: " + m); }
  }
  for (auto& fac : g.f) if (fac) { int rc = fac->emit(c); if (rc != FG_OK) { std::string m = fg_last_error(c); fg_destroy(c); throw std::runtime_error("factor -> ctx: " + m); } }
  return c;
}
inline void readback(fg_ctx* c, Values& v) {
  for (auto& kv : v.m) { int n = 0; check(c, fg_get_value(c, kv.first, kv.second.v, &n), "fg_get_value"); }
}
}  // namespace detail

inline double NonlinearFactorGraph::error(const Values& v) const {
  fg_ctx* c = detail::build(*this, v);
  double e = 0; int rc = fg_error(c, &e);
  std::string m = rc ? fg_last_error(c) : "";
  fg_destroy(c);
  if (rc) throw std::runtime_error("fg_error: " + m);
  return e;
}

class LevenbergMarquardtOptimizer {
 public:
  const NonlinearFactorGraph& g; Values v; fg_lm_report report;
  LevenbergMarquardtOptimizer(const NonlinearFactorGraph& graph, const Values& initial) : g(graph), v(initial) {}
  Values optimize() {
    fg_ctx* c = detail::build(g, v);
    int rc = fg_optimize_lm(c, nullptr, &report);
    if (rc != FG_OK) { std::string m = fg_last_error(c); fg_destroy(c); throw std::runtime_error("fg_optimize_lm: " + m); }
    detail::readback(c, v);
    fg_destroy(c);
    return v;
  }
  double error() const { return report.final_error; }
  int iterations() const { return report.iterations; }
};

// Marginals(graph, values, Marginals::CHOLESKY).marginalCovariance(key)   gtsam/gtsam_graph.cpp:598-601, :1357
// (SURVEY 8 f2).  One undamped device factorisation per call of marginalCovariance; dimension 6 / 3 / 6 / 3 for a
// pose / velocity / bias / plane key, returned row-major.
class Marginals {
 public:
  enum Factorization { CHOLESKY, QR };
  const NonlinearFactorGraph& g; const Values& v;
  Marginals(const NonlinearFactorGraph& graph, const Values& solution, Factorization = CHOLESKY) : g(graph), v(solution) {}
  std::vector<double> marginalCovariance(Key key, int* dim = nullptr) const {
    fg_ctx* c = detail::build(g, v);
    double cov[36]; int d = 0;
    int rc = fg_marginal_cov(c, key, cov, &d);
    std::string m = rc ? fg_last_error(c) : "";
    fg_destroy(c);
    if (rc) throw std::runtime_error("fg_marginal_cov: " + m);
    if (dim) *dim = d;
    return std::vector<double>(cov, cov + d * d);
  }
};

// ISAM2 (gtsam/gtsam_graph.cpp:93-99,1768-1776): one persistent device context.  update() hands the new
// values and factors to it and runs fg_update_incremental -- new variables enter at their initial value, variables whose
// delta reaches relinearizeThreshold move their linearisation point, one undamped Gauss-Newton system is solved on the
// device; calculateEstimate() reads theta (+) delta back.  (A full re-factorisation per update instead of the partial
// Bayes-tree re-elimination: see include/fg_abi.h.)
struct ISAM2Params { double relinearizeThreshold = 0.1; int relinearizeSkip = 10; };
class ISAM2 {
 public:
  ISAM2Params params; Values estimate; fg_inc_report report;
  ISAM2() { std::memset(&report, 0, sizeof report); }
  explicit ISAM2(const ISAM2Params& p) : params(p) { std::memset(&report, 0, sizeof report); }
  ISAM2(const ISAM2&) = delete;
  ISAM2& operator=(const ISAM2&) = delete;
  ~ISAM2() { if (c_) fg_destroy(c_); }
  void update(const NonlinearFactorGraph& nf, const Values& nv) {
    if (!c_) {
      c_ = fg_create(0, 0, 1);
      if (!c_) throw std::runtime_error("fg_create failed: no CUDA device (this backend has no CPU fallback)");
    }
    for (auto& kv : nv.m) {
      int rc = FG_OK;
      switch (kv.second.type) {
        case FG_T_POSE: rc = fg_add_pose(c_, kv.first, kv.second.v); break;
        case FG_T_VEC3: rc = fg_add_vec3(c_, kv.first, kv.second.v); break;
        case FG_T_BIAS: rc = fg_add_bias(c_, kv.first, kv.second.v); break;
        case FG_T_POINT: rc = fg_add_point(c_, kv.first, kv.second.v); break;
        case FG_T_PLANE: rc = fg_add_plane(c_, kv.first, kv.second.v); break;
      }
      detail::check(c_, rc, "ISAM2::update (new value)");
    }
    estimate.insert(nv);
    for (auto& fac : nf.f) if (fac) detail::check(c_, fac->emit(c_), "ISAM2::update (new factor)");
    fg_isam2_params p; p.relinearize_threshold = params.relinearizeThreshold; p.relinearize_skip = params.relinearizeSkip;
    detail::check(c_, fg_update_incremental(c_, &p, &report), "fg_update_incremental");
    fresh_ = false;
  }
  void update() { NonlinearFactorGraph none; Values nov; update(none, nov); }
  Values calculateEstimate() {
    if (!fresh_ && c_) { detail::readback(c_, estimate); fresh_ = true; }
    return estimate;
  }
 private:
  fg_ctx* c_ = nullptr;
  bool fresh_ = true;
};

}  // namespace gtsam
