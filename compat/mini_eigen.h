// mini_eigen.h -- a small dense linear-algebra header with the Eigen 3 spellings the reference's wrapper sources use
// (SURVEY.md section 8b-3 / Appendix C).  Eigen itself is not installed in this image; this header exists only so that
// gtsam/gtsam_graph.cpp, imu_base.cpp, imu_vn100.cpp, g2o/g2o_graph.cpp and the two offline drivers of the reference
// compile UNCHANGED against the B200 backend (host-side glue: 3x3 ... 6x6 matrices, a handful per frame).  With a real
// Eigen on the include path this directory is simply not used.  No expression templates: every operation returns a value.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_DEFINE_STL_VECTOR_SPECIALIZATION(...)

namespace Eigen {

const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum TransformTraits { Isometry = 1, Affine = 2, AffineCompact = 3, Projective = 4 };
typedef std::ptrdiff_t Index;
template <typename T> using aligned_allocator = std::allocator<T>;

namespace internal {
template <typename T, int N> struct Storage {
  std::array<T, (N > 0 ? N : 1)> v{};
  void resize(size_t) {}
  T* data() { return v.data(); }
  const T* data() const { return v.data(); }
};
template <typename T> struct Storage<T, Dynamic> {
  std::vector<T> v;
  void resize(size_t n) { v.assign(n, T()); }
  T* data() { return v.data(); }
  const T* data() const { return v.data(); }
};
constexpr int pick(int a, int b) { return a == Dynamic ? b : a; }
}  // namespace internal

template <typename T, int R, int C, int Opt = 0, int MR = R, int MC = C> class Matrix;
template <typename T> class Quaternion;

// comma initialiser:  m << a, b, c, ...;   (row-major fill, like Eigen)
template <typename M> struct CommaInitializer {
  M& m; int k;
  CommaInitializer(M& mm, typename M::Scalar first) : m(mm), k(0) { put(first); }
  void put(typename M::Scalar v) { const int c = m.cols(); m(k / c, k % c) = v; ++k; }
  CommaInitializer& operator,(typename M::Scalar v) { put(v); return *this; }
};

template <typename T, int R, int C, int Opt, int MR, int MC>
class Matrix {
 public:
  typedef T Scalar;
  typedef Eigen::Index Index;
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C, IsRowMajor = (Opt & RowMajor) ? 1 : 0, SizeAtCompileTime = (R == Dynamic || C == Dynamic) ? Dynamic : R * C };

 protected:
  internal::Storage<T, (R == Dynamic || C == Dynamic) ? Dynamic : R * C> s_;
  int r_ = (R == Dynamic ? 0 : R), c_ = (C == Dynamic ? 0 : C);
  int idx(int i, int j) const { return IsRowMajor ? i * c_ + j : j * r_ + i; }

 public:
  Matrix() {}
  // (rows, cols) for dynamic matrices, (x, y) for 2-vectors
  Matrix(T a, T b) {
    if (R == Dynamic || C == Dynamic) { resize((int)a, (int)b); }
    else { assert(R * C == 2); s_.v[0] = a; s_.v[1] = b; }
  }
  explicit Matrix(int n) { if (R == Dynamic && C == 1) resize(n, 1); else if (C == Dynamic && R == 1) resize(1, n); else if (R == Dynamic || C == Dynamic) resize(n, n); else if (R * C == 1) s_.v[0] = (T)n; }
  Matrix(T a, T b, T c) { static_assert(R * C == 3 || R == Dynamic, "3 coefficients"); if (R == Dynamic) resize(3, 1); s_.data()[0] = a; s_.data()[1] = b; s_.data()[2] = c; }
  Matrix(T a, T b, T c, T d) { static_assert(R * C == 4 || R == Dynamic, "4 coefficients"); if (R == Dynamic) resize(4, 1); s_.data()[0] = a; s_.data()[1] = b; s_.data()[2] = c; s_.data()[3] = d; }
  template <int R2, int C2, int O2, int MR2, int MC2>
  Matrix(const Matrix<T, R2, C2, O2, MR2, MC2>& o) { assign(o); }
  template <int R2, int C2, int O2, int MR2, int MC2>
  Matrix& operator=(const Matrix<T, R2, C2, O2, MR2, MC2>& o) { assign(o); return *this; }
  template <int R2, int C2, int O2, int MR2, int MC2>
  void assign(const Matrix<T, R2, C2, O2, MR2, MC2>& o) {
    if (R == Dynamic || C == Dynamic) resize(o.rows(), o.cols());
    if (rows() != o.rows() || cols() != o.cols()) {
      if (rows() * cols() == o.rows() * o.cols() && (rows() == 1 || cols() == 1) && (o.rows() == 1 || o.cols() == 1)) {   // vector <-> row vector
        for (int k = 0; k < rows() * cols(); ++k) (*this)(k) = o(k);
        return;
      }
      throw std::runtime_error("mini_eigen: size mismatch in assignment");
    }
    for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) (*this)(i, j) = o(i, j);
  }

  void resize(int r, int c) {
    if (R != Dynamic && r != R) throw std::runtime_error("mini_eigen: resize of a fixed dimension");
    if (C != Dynamic && c != C) throw std::runtime_error("mini_eigen: resize of a fixed dimension");
    r_ = r; c_ = c; s_.resize((size_t)r * c);
  }
  void resize(int n) { if (C == 1) resize(n, 1); else resize(1, n); }
  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  // a 1 x 1 result (an inner product written as a^T * b) reads as a scalar
  template <int RR = R, int CC = C, typename = typename std::enable_if<RR == 1 && CC == 1>::type> operator T() const { return s_.data()[0]; }
  T* data() { return s_.data(); }
  const T* data() const { return s_.data(); }

  T& operator()(int i, int j) { return s_.data()[idx(i, j)]; }
  const T& operator()(int i, int j) const { return s_.data()[idx(i, j)]; }
  T& operator()(int i) { return s_.data()[i]; }
  const T& operator()(int i) const { return s_.data()[i]; }
  T& operator[](int i) { return s_.data()[i]; }
  const T& operator[](int i) const { return s_.data()[i]; }
  T& coeffRef(int i, int j) { return (*this)(i, j); }
  T coeff(int i, int j) const { return (*this)(i, j); }
  T& x() { return s_.data()[0]; } const T& x() const { return s_.data()[0]; }
  T& y() { return s_.data()[1]; } const T& y() const { return s_.data()[1]; }
  T& z() { return s_.data()[2]; } const T& z() const { return s_.data()[2]; }
  T& w() { return s_.data()[3]; } const T& w() const { return s_.data()[3]; }

  // ---- generators
  static Matrix Zero() { Matrix m; m.setZero(); return m; }
  static Matrix Zero(int r, int c) { Matrix m; m.resize(r, c); m.setZero(); return m; }
  static Matrix Zero(int n) { Matrix m; m.resize(n); m.setZero(); return m; }
  static Matrix Ones() { Matrix m; m.setConstant(T(1)); return m; }
  static Matrix Ones(int r, int c) { Matrix m; m.resize(r, c); m.setConstant(T(1)); return m; }
  static Matrix Constant(T v) { Matrix m; m.setConstant(v); return m; }
  static Matrix Constant(int r, int c, T v) { Matrix m; m.resize(r, c); m.setConstant(v); return m; }
  static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
  static Matrix Identity(int r, int c) { Matrix m; m.resize(r, c); m.setIdentity(); return m; }
  static Matrix UnitX() { Matrix m; m.setZero(); m(0) = 1; return m; }
  static Matrix UnitY() { Matrix m; m.setZero(); m(1) = 1; return m; }
  static Matrix UnitZ() { Matrix m; m.setZero(); m(2) = 1; return m; }
  Matrix& setZero() { std::fill(data(), data() + size(), T(0)); return *this; }
  Matrix& setZero(int r, int c) { resize(r, c); return setZero(); }
  Matrix& setOnes() { return setConstant(T(1)); }
  Matrix& setConstant(T v) { std::fill(data(), data() + size(), v); return *this; }
  Matrix& setIdentity() { setZero(); for (int i = 0; i < std::min(r_, c_); ++i) (*this)(i, i) = T(1); return *this; }
  Matrix& fill(T v) { return setConstant(v); }
  CommaInitializer<Matrix> operator<<(T v) { return CommaInitializer<Matrix>(*this, v); }

  // ---- blocks: a copy of the values that writes back to its parent on assignment
  template <int BR, int BC>
  struct BlockRef : public Matrix<T, BR, BC> {
    Matrix* p; int i0, j0;
    BlockRef(Matrix* parent, int i, int j, int br, int bc) : p(parent), i0(i), j0(j) {
      if (BR == Dynamic || BC == Dynamic) this->resize(br, bc);
      for (int b = 0; b < bc; ++b) for (int a = 0; a < br; ++a) (*static_cast<Matrix<T, BR, BC>*>(this))(a, b) = (*p)(i + a, j + b);
    }
    void push() { for (int b = 0; b < this->cols(); ++b) for (int a = 0; a < this->rows(); ++a) (*p)(i0 + a, j0 + b) = (*static_cast<Matrix<T, BR, BC>*>(this))(a, b); }
    template <int R2, int C2, int O2, int MR2, int MC2>
    BlockRef& operator=(const Matrix<T, R2, C2, O2, MR2, MC2>& o) { Matrix<T, BR, BC>::assign(o); push(); return *this; }
    BlockRef& operator=(const BlockRef& o) { Matrix<T, BR, BC>::assign(static_cast<const Matrix<T, BR, BC>&>(o)); push(); return *this; }
    template <int R2, int C2, int O2, int MR2, int MC2>
    BlockRef& operator+=(const Matrix<T, R2, C2, O2, MR2, MC2>& o) { Matrix<T, BR, BC>::operator+=(o); push(); return *this; }
    template <int R2, int C2, int O2, int MR2, int MC2>
    BlockRef& operator-=(const Matrix<T, R2, C2, O2, MR2, MC2>& o) { Matrix<T, BR, BC>::operator-=(o); push(); return *this; }
    BlockRef& operator*=(T v) { Matrix<T, BR, BC>::operator*=(v); push(); return *this; }
    BlockRef& setZero() { Matrix<T, BR, BC>::setZero(); push(); return *this; }
    BlockRef& setIdentity() { Matrix<T, BR, BC>::setIdentity(); push(); return *this; }
    CommaInitializer<BlockRef> operator<<(T v) { return CommaInitializer<BlockRef>(*this, v); }
    // coefficient writes through a block go straight to the parent
    struct Cell { BlockRef* b; int i, j; Cell& operator=(T v) { (*static_cast<Matrix<T, BR, BC>*>(b))(i, j) = v; (*b->p)(b->i0 + i, b->j0 + j) = v; return *this; } operator T() const { return (*static_cast<const Matrix<T, BR, BC>*>(b))(i, j); } };
  };
  template <int BR, int BC> BlockRef<BR, BC> block(int i, int j) { return BlockRef<BR, BC>(this, i, j, BR, BC); }
  template <int BR, int BC> Matrix<T, BR, BC> block(int i, int j) const { return const_cast<Matrix*>(this)->template block<BR, BC>(i, j); }
  BlockRef<Dynamic, Dynamic> block(int i, int j, int br, int bc) { return BlockRef<Dynamic, Dynamic>(this, i, j, br, bc); }
  Matrix<T, Dynamic, Dynamic> block(int i, int j, int br, int bc) const { return const_cast<Matrix*>(this)->block(i, j, br, bc); }
  template <int BR, int BC> BlockRef<BR, BC> topLeftCorner() { return block<BR, BC>(0, 0); }
  template <int BR, int BC> Matrix<T, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
  template <int BR, int BC> BlockRef<BR, BC> topRightCorner() { return block<BR, BC>(0, c_ - BC); }
  template <int BR, int BC> Matrix<T, BR, BC> topRightCorner() const { return block<BR, BC>(0, c_ - BC); }
  template <int BR, int BC> BlockRef<BR, BC> bottomRightCorner() { return block<BR, BC>(r_ - BR, c_ - BC); }
  template <int BR, int BC> BlockRef<BR, BC> bottomLeftCorner() { return block<BR, BC>(r_ - BR, 0); }
  template <int N> BlockRef<(C == 1 ? N : 1), (C == 1 ? 1 : N)> head() { return segment<N>(0); }
  template <int N> Matrix<T, (C == 1 ? N : 1), (C == 1 ? 1 : N)> head() const { return segment<N>(0); }
  template <int N> BlockRef<(C == 1 ? N : 1), (C == 1 ? 1 : N)> tail() { return segment<N>(size() - N); }
  template <int N> Matrix<T, (C == 1 ? N : 1), (C == 1 ? 1 : N)> tail() const { return segment<N>(size() - N); }
  template <int N> BlockRef<(C == 1 ? N : 1), (C == 1 ? 1 : N)> segment(int i) { return (C == 1) ? block<(C == 1 ? N : 1), (C == 1 ? 1 : N)>(i, 0) : block<(C == 1 ? N : 1), (C == 1 ? 1 : N)>(0, i); }
  template <int N> Matrix<T, (C == 1 ? N : 1), (C == 1 ? 1 : N)> segment(int i) const { return const_cast<Matrix*>(this)->template segment<N>(i); }
  BlockRef<Dynamic, Dynamic> head(int n) { return (C == 1) ? block(0, 0, n, 1) : block(0, 0, 1, n); }
  BlockRef<Dynamic, Dynamic> tail(int n) { return (C == 1) ? block(size() - n, 0, n, 1) : block(0, size() - n, 1, n); }
  BlockRef<R, 1> col(int j) { return BlockRef<R, 1>(this, 0, j, r_, 1); }
  Matrix<T, R, 1> col(int j) const { return const_cast<Matrix*>(this)->col(j); }
  BlockRef<1, C> row(int i) { return BlockRef<1, C>(this, i, 0, 1, c_); }
  Matrix<T, 1, C> row(int i) const { return const_cast<Matrix*>(this)->row(i); }
  Matrix<T, internal::pick(R, C), 1> diagonal() const { Matrix<T, internal::pick(R, C), 1> d; if (R == Dynamic || C == Dynamic) d.resize(std::min(r_, c_), 1); for (int i = 0; i < std::min(r_, c_); ++i) d(i) = (*this)(i, i); return d; }
  Matrix<T, internal::pick(R, C) == 1 ? internal::pick(C, R) : internal::pick(R, C), internal::pick(R, C) == 1 ? internal::pick(C, R) : internal::pick(R, C)> asDiagonal() const {
    const int n = size();
    Matrix<T, internal::pick(R, C) == 1 ? internal::pick(C, R) : internal::pick(R, C), internal::pick(R, C) == 1 ? internal::pick(C, R) : internal::pick(R, C)> m;
    m.resize(n, n); m.setZero();
    for (int i = 0; i < n; ++i) m(i, i) = (*this)(i);
    return m;
  }

  // ---- algebra
  Matrix<T, C, R> transpose() const { Matrix<T, C, R> t; if (R == Dynamic || C == Dynamic) t.resize(c_, r_); for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) t(j, i) = (*this)(i, j); return t; }
  void transposeInPlace() { *this = Matrix(transpose()); }
  T trace() const { T s = 0; for (int i = 0; i < std::min(r_, c_); ++i) s += (*this)(i, i); return s; }
  T sum() const { T s = 0; for (int k = 0; k < size(); ++k) s += data()[k]; return s; }
  T squaredNorm() const { T s = 0; for (int k = 0; k < size(); ++k) s += data()[k] * data()[k]; return s; }
  T norm() const { return std::sqrt(squaredNorm()); }
  Matrix normalized() const { Matrix m(*this); m.normalize(); return m; }
  void normalize() { const T n = norm(); if (n > 0) for (int k = 0; k < size(); ++k) data()[k] /= n; }
  T maxCoeff() const { T m = data()[0]; for (int k = 1; k < size(); ++k) m = std::max(m, data()[k]); return m; }
  T minCoeff() const { T m = data()[0]; for (int k = 1; k < size(); ++k) m = std::min(m, data()[k]); return m; }
  T mean() const { return sum() / T(size()); }
  Matrix cwiseAbs() const { Matrix m(*this); for (int k = 0; k < size(); ++k) m.data()[k] = std::abs(m.data()[k]); return m; }
  Matrix cwiseSqrt() const { Matrix m(*this); for (int k = 0; k < size(); ++k) m.data()[k] = std::sqrt(m.data()[k]); return m; }
  template <int R2, int C2> Matrix cwiseProduct(const Matrix<T, R2, C2>& o) const { Matrix m(*this); for (int k = 0; k < size(); ++k) m.data()[k] *= o.data()[k]; return m; }
  bool hasNaN() const { for (int k = 0; k < size(); ++k) if (data()[k] != data()[k]) return true; return false; }
  bool allFinite() const { for (int k = 0; k < size(); ++k) if (!std::isfinite((double)data()[k])) return false; return true; }
  bool isZero(T eps = T(1e-12)) const { for (int k = 0; k < size(); ++k) if (std::abs(data()[k]) > eps) return false; return true; }
  template <int R2, int C2, int O2, int MR2, int MC2> bool isApprox(const Matrix<T, R2, C2, O2, MR2, MC2>& o, T eps = T(1e-12)) const { return (*this - o).norm() <= eps * std::min(norm(), o.norm()); }
  template <int R2, int C2, int O2, int MR2, int MC2> T dot(const Matrix<T, R2, C2, O2, MR2, MC2>& o) const { T s = 0; for (int k = 0; k < size(); ++k) s += (*this)(k) * o(k); return s; }
  template <int R2, int C2, int O2, int MR2, int MC2> Matrix cross(const Matrix<T, R2, C2, O2, MR2, MC2>& o) const {
    Matrix m(*this);
    m(0) = (*this)(1) * o(2) - (*this)(2) * o(1); m(1) = (*this)(2) * o(0) - (*this)(0) * o(2); m(2) = (*this)(0) * o(1) - (*this)(1) * o(0);
    return m;
  }
  template <typename U> Matrix<U, R, C, Opt> cast() const { Matrix<U, R, C, Opt> m; if (R == Dynamic || C == Dynamic) m.resize(r_, c_); for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) m(i, j) = (U)(*this)(i, j); return m; }
  const Matrix& eval() const { return *this; }
  const Matrix& matrix() const { return *this; }
  const Matrix& array() const { return *this; }

  // Gauss-Jordan with partial pivoting (sizes here are <= 15)
  Matrix inverse() const {
    const int n = r_;
    if (n != c_) throw std::runtime_error("mini_eigen: inverse of a non-square matrix");
    Matrix a(*this), inv; inv.resize(n, n); inv.setIdentity();
    for (int c = 0; c < n; ++c) {
      int piv = c;
      for (int i = c + 1; i < n; ++i) if (std::abs(a(i, c)) > std::abs(a(piv, c))) piv = i;
      if (piv != c) for (int j = 0; j < n; ++j) { std::swap(a(c, j), a(piv, j)); std::swap(inv(c, j), inv(piv, j)); }
      const T d = a(c, c);      // a singular matrix yields inf / nan, as Eigen's inverse() does
      for (int j = 0; j < n; ++j) { a(c, j) /= d; inv(c, j) /= d; }
      for (int i = 0; i < n; ++i) if (i != c) { const T f = a(i, c); if (f != T(0)) for (int j = 0; j < n; ++j) { a(i, j) -= f * a(c, j); inv(i, j) -= f * inv(c, j); } }
    }
    return inv;
  }
  T determinant() const {
    const int n = r_;
    Matrix a(*this); T det = 1;
    for (int c = 0; c < n; ++c) {
      int piv = c;
      for (int i = c + 1; i < n; ++i) if (std::abs(a(i, c)) > std::abs(a(piv, c))) piv = i;
      if (a(piv, c) == T(0)) return T(0);
      if (piv != c) { for (int j = 0; j < n; ++j) std::swap(a(c, j), a(piv, j)); det = -det; }
      det *= a(c, c);
      for (int i = c + 1; i < n; ++i) { const T f = a(i, c) / a(c, c); for (int j = c; j < n; ++j) a(i, j) -= f * a(c, j); }
    }
    return det;
  }

  // ---- operators
  Matrix operator-() const { Matrix m(*this); for (int k = 0; k < size(); ++k) m.data()[k] = -m.data()[k]; return m; }
  template <int R2, int C2, int O2, int MR2, int MC2> Matrix& operator+=(const Matrix<T, R2, C2, O2, MR2, MC2>& o) { for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) (*this)(i, j) += o(i, j); return *this; }
  template <int R2, int C2, int O2, int MR2, int MC2> Matrix& operator-=(const Matrix<T, R2, C2, O2, MR2, MC2>& o) { for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) (*this)(i, j) -= o(i, j); return *this; }
  Matrix& operator*=(T v) { for (int k = 0; k < size(); ++k) data()[k] *= v; return *this; }
  Matrix& operator/=(T v) { for (int k = 0; k < size(); ++k) data()[k] /= v; return *this; }
  template <int R2, int C2, int O2, int MR2, int MC2> Matrix& operator*=(const Matrix<T, R2, C2, O2, MR2, MC2>& o) { *this = Matrix((*this) * o); return *this; }
  template <int R2, int C2, int O2, int MR2, int MC2> bool operator==(const Matrix<T, R2, C2, O2, MR2, MC2>& o) const { if (r_ != o.rows() || c_ != o.cols()) return false; for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) if ((*this)(i, j) != o(i, j)) return false; return true; }
  template <int R2, int C2, int O2, int MR2, int MC2> bool operator!=(const Matrix<T, R2, C2, O2, MR2, MC2>& o) const { return !(*this == o); }
};

template <typename T, int R1, int C1, int O1, int MR1, int MC1, int R2, int C2, int O2, int MR2, int MC2>
Matrix<T, internal::pick(R1, R2), internal::pick(C1, C2)> operator+(const Matrix<T, R1, C1, O1, MR1, MC1>& a, const Matrix<T, R2, C2, O2, MR2, MC2>& b) {
  Matrix<T, internal::pick(R1, R2), internal::pick(C1, C2)> m; m.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); ++j) for (int i = 0; i < a.rows(); ++i) m(i, j) = a(i, j) + b(i, j);
  return m;
}
template <typename T, int R1, int C1, int O1, int MR1, int MC1, int R2, int C2, int O2, int MR2, int MC2>
Matrix<T, internal::pick(R1, R2), internal::pick(C1, C2)> operator-(const Matrix<T, R1, C1, O1, MR1, MC1>& a, const Matrix<T, R2, C2, O2, MR2, MC2>& b) {
  Matrix<T, internal::pick(R1, R2), internal::pick(C1, C2)> m; m.resize(a.rows(), a.cols());
  for (int j = 0; j < a.cols(); ++j) for (int i = 0; i < a.rows(); ++i) m(i, j) = a(i, j) - b(i, j);
  return m;
}
template <typename T, int R1, int C1, int O1, int MR1, int MC1, int R2, int C2, int O2, int MR2, int MC2>
Matrix<T, R1, C2> operator*(const Matrix<T, R1, C1, O1, MR1, MC1>& a, const Matrix<T, R2, C2, O2, MR2, MC2>& b) {
  if (a.cols() != b.rows()) throw std::runtime_error("mini_eigen: size mismatch in product");
  Matrix<T, R1, C2> m; m.resize(a.rows(), b.cols());
  for (int i = 0; i < a.rows(); ++i)
    for (int j = 0; j < b.cols(); ++j) { T s = 0; for (int k = 0; k < a.cols(); ++k) s += a(i, k) * b(k, j); m(i, j) = s; }
  return m;
}
template <typename T, int R, int C, int O, int MR, int MC, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator*(const Matrix<T, R, C, O, MR, MC>& a, S v) { Matrix<T, R, C> m(a); m *= (T)v; return m; }
template <typename T, int R, int C, int O, int MR, int MC, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator*(S v, const Matrix<T, R, C, O, MR, MC>& a) { Matrix<T, R, C> m(a); m *= (T)v; return m; }
template <typename T, int R, int C, int O, int MR, int MC, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C> operator/(const Matrix<T, R, C, O, MR, MC>& a, S v) { Matrix<T, R, C> m(a); m /= (T)v; return m; }

// scalar (+|-) 1 x 1 matrix, as in  double s = a + v.transpose() * M * v;
template <typename T, int O, int MR, int MC, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
T operator+(S a, const Matrix<T, 1, 1, O, MR, MC>& b) { return (T)a + b(0); }
template <typename T, int O, int MR, int MC, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
T operator+(const Matrix<T, 1, 1, O, MR, MC>& a, S b) { return a(0) + (T)b; }
template <typename T, int O, int MR, int MC, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
T operator-(S a, const Matrix<T, 1, 1, O, MR, MC>& b) { return (T)a - b(0); }
template <typename T, int O, int MR, int MC, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
T operator-(const Matrix<T, 1, 1, O, MR, MC>& a, S b) { return a(0) - (T)b; }

template <typename T, int R, int C, int O, int MR, int MC>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C, O, MR, MC>& m) {
  for (int i = 0; i < m.rows(); ++i) { for (int j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j); if (i + 1 < m.rows()) os << "\n"; }
  return os;
}

#define MINI_EIGEN_TYPEDEFS(T, S)                                                                  \
  typedef Matrix<T, 2, 2> Matrix2##S; typedef Matrix<T, 3, 3> Matrix3##S; typedef Matrix<T, 4, 4> Matrix4##S; \
  typedef Matrix<T, Dynamic, Dynamic> MatrixX##S;                                                  \
  typedef Matrix<T, 2, 1> Vector2##S; typedef Matrix<T, 3, 1> Vector3##S; typedef Matrix<T, 4, 1> Vector4##S; \
  typedef Matrix<T, Dynamic, 1> VectorX##S; typedef Matrix<T, 1, 3> RowVector3##S; typedef Matrix<T, 1, Dynamic> RowVectorX##S;
MINI_EIGEN_TYPEDEFS(double, d)
MINI_EIGEN_TYPEDEFS(float, f)
MINI_EIGEN_TYPEDEFS(int, i)
#undef MINI_EIGEN_TYPEDEFS

// Map: a matrix view of caller-owned memory (a copy that is written back by the destructor and by assignment)
template <typename M> class Map : public M {
  typename M::Scalar* ptr_;
 public:
  Map(typename M::Scalar* p) : ptr_(p) { load(); }
  Map(const typename M::Scalar* p) : ptr_(nullptr) { ptr_ = const_cast<typename M::Scalar*>(p); load(); ptr_ = nullptr; }
  Map(typename M::Scalar* p, int r, int c) : ptr_(p) { this->resize(r, c); load(); }
  Map(typename M::Scalar* p, int n) : ptr_(p) { this->resize(n); load(); }
  ~Map() { store(); }
  template <int R2, int C2, int O2, int MR2, int MC2>
  Map& operator=(const Matrix<typename M::Scalar, R2, C2, O2, MR2, MC2>& o) { M::assign(o); store(); return *this; }
 private:
  void load() { for (int k = 0; k < this->size(); ++k) this->data()[k] = ptr_[k]; }
  void store() { if (ptr_) for (int k = 0; k < this->size(); ++k) ptr_[k] = this->data()[k]; }
};

// ------------------------------------------------------------------ geometry
template <typename T> class AngleAxis {
 public:
  T a; Matrix<T, 3, 1> ax;
  AngleAxis() : a(0), ax(1, 0, 0) {}
  AngleAxis(T angle, const Matrix<T, 3, 1>& axis) : a(angle), ax(axis) {}
  T angle() const { return a; }
  const Matrix<T, 3, 1>& axis() const { return ax; }
  Matrix<T, 3, 3> toRotationMatrix() const {
    const T c = std::cos(a), s = std::sin(a), x = ax(0), y = ax(1), z = ax(2);
    Matrix<T, 3, 3> R;
    R << c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s,
         y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s,
         z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c);
    return R;
  }
  Matrix<T, 3, 3> matrix() const { return toRotationMatrix(); }
};
typedef AngleAxis<double> AngleAxisd;
typedef AngleAxis<float> AngleAxisf;

template <typename T> class Quaternion {
  T w_, x_, y_, z_;
 public:
  typedef T Scalar;
  Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
  Quaternion(T w, T x, T y, T z) : w_(w), x_(x), y_(y), z_(z) {}
  template <int O, int MR, int MC> explicit Quaternion(const Matrix<T, 3, 3, O, MR, MC>& R) { *this = R; }
  template <int O, int MR, int MC> explicit Quaternion(const Matrix<T, 4, 1, O, MR, MC>& c) : w_(c(3)), x_(c(0)), y_(c(1)), z_(c(2)) {}   // coefficient order (x, y, z, w)
  explicit Quaternion(const AngleAxis<T>& aa) { const T h = aa.angle() / 2, s = std::sin(h); w_ = std::cos(h); x_ = s * aa.axis()(0); y_ = s * aa.axis()(1); z_ = s * aa.axis()(2); }
  template <int O, int MR, int MC> Quaternion& operator=(const Matrix<T, 3, 3, O, MR, MC>& M) {
    const T tr = M(0, 0) + M(1, 1) + M(2, 2);
    if (tr > 0) { T s = std::sqrt(tr + 1) * 2; w_ = s / 4; x_ = (M(2, 1) - M(1, 2)) / s; y_ = (M(0, 2) - M(2, 0)) / s; z_ = (M(1, 0) - M(0, 1)) / s; }
    else if (M(0, 0) > M(1, 1) && M(0, 0) > M(2, 2)) { T s = std::sqrt(1 + M(0, 0) - M(1, 1) - M(2, 2)) * 2; w_ = (M(2, 1) - M(1, 2)) / s; x_ = s / 4; y_ = (M(0, 1) + M(1, 0)) / s; z_ = (M(0, 2) + M(2, 0)) / s; }
    else if (M(1, 1) > M(2, 2)) { T s = std::sqrt(1 + M(1, 1) - M(0, 0) - M(2, 2)) * 2; w_ = (M(0, 2) - M(2, 0)) / s; x_ = (M(0, 1) + M(1, 0)) / s; y_ = s / 4; z_ = (M(1, 2) + M(2, 1)) / s; }
    else { T s = std::sqrt(1 + M(2, 2) - M(0, 0) - M(1, 1)) * 2; w_ = (M(1, 0) - M(0, 1)) / s; x_ = (M(0, 2) + M(2, 0)) / s; y_ = (M(1, 2) + M(2, 1)) / s; z_ = s / 4; }
    return *this;
  }
  static Quaternion Identity() { return Quaternion(); }
  T& w() { return w_; } T& x() { return x_; } T& y() { return y_; } T& z() { return z_; }
  T w() const { return w_; } T x() const { return x_; } T y() const { return y_; } T z() const { return z_; }
  Matrix<T, 4, 1> coeffs() const { return Matrix<T, 4, 1>(x_, y_, z_, w_); }
  Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(x_, y_, z_); }
  T norm() const { return std::sqrt(w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_); }
  void normalize() { const T n = norm(); w_ /= n; x_ /= n; y_ /= n; z_ /= n; }
  Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Quaternion inverse() const { const T n2 = w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_; return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2); }
  Quaternion operator*(const Quaternion& o) const {
    return Quaternion(w_ * o.w_ - x_ * o.x_ - y_ * o.y_ - z_ * o.z_, w_ * o.x_ + x_ * o.w_ + y_ * o.z_ - z_ * o.y_,
                      w_ * o.y_ - x_ * o.z_ + y_ * o.w_ + z_ * o.x_, w_ * o.z_ + x_ * o.y_ - y_ * o.x_ + z_ * o.w_);
  }
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& v) const { return toRotationMatrix() * v; }
  Matrix<T, 3, 3> toRotationMatrix() const {
    Matrix<T, 3, 3> R;
    R << 1 - 2 * (y_ * y_ + z_ * z_), 2 * (x_ * y_ - z_ * w_), 2 * (x_ * z_ + y_ * w_),
         2 * (x_ * y_ + z_ * w_), 1 - 2 * (x_ * x_ + z_ * z_), 2 * (y_ * z_ - x_ * w_),
         2 * (x_ * z_ - y_ * w_), 2 * (y_ * z_ + x_ * w_), 1 - 2 * (x_ * x_ + y_ * y_);
    return R;
  }
  Matrix<T, 3, 3> matrix() const { return toRotationMatrix(); }
  template <typename U> Quaternion<U> cast() const { return Quaternion<U>((U)w_, (U)x_, (U)y_, (U)z_); }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <typename T, int Dim> class Translation {
 public:
  Matrix<T, Dim, 1> v;
  Translation() {}
  Translation(T x, T y, T z) : v(x, y, z) {}
  explicit Translation(const Matrix<T, Dim, 1>& t) : v(t) {}
  const Matrix<T, Dim, 1>& vector() const { return v; }
};
typedef Translation<double, 3> Translation3d;
typedef Translation<float, 3> Translation3f;

// Transform<T, 3, Mode>: a 4x4 homogeneous matrix with the usual accessors
template <typename T, int Dim, int Mode> class Transform {
  static_assert(Dim == 3, "only 3-D transforms");
  Matrix<T, 4, 4> m_;
 public:
  typedef T Scalar;
  typedef Matrix<T, 4, 4> MatrixType;
  Transform() { m_.setIdentity(); }
  template <int O, int MR, int MC> Transform(const Matrix<T, 4, 4, O, MR, MC>& m) : m_(m) {}
  template <int M2> Transform(const Transform<T, Dim, M2>& o) : m_(o.matrix()) {}
  explicit Transform(const Quaternion<T>& q) { m_.setIdentity(); linear() = q.toRotationMatrix(); }
  explicit Transform(const Translation<T, 3>& t) { m_.setIdentity(); translation() = t.v; }
  template <int O, int MR, int MC> Transform& operator=(const Matrix<T, 4, 4, O, MR, MC>& m) { m_ = m; return *this; }
  Transform& operator=(const Quaternion<T>& q) { m_.setIdentity(); linear() = q.toRotationMatrix(); return *this; }
  static Transform Identity() { return Transform(); }
  void setIdentity() { m_.setIdentity(); }
  MatrixType& matrix() { return m_; }
  const MatrixType& matrix() const { return m_; }
  T& operator()(int i, int j) { return m_(i, j); }
  T operator()(int i, int j) const { return m_(i, j); }
  T* data() { return m_.data(); }
  const T* data() const { return m_.data(); }
  typename MatrixType::template BlockRef<3, 1> translation() { return m_.template block<3, 1>(0, 3); }
  Matrix<T, 3, 1> translation() const { return m_.template block<3, 1>(0, 3); }
  typename MatrixType::template BlockRef<3, 3> linear() { return m_.template block<3, 3>(0, 0); }
  Matrix<T, 3, 3> linear() const { return m_.template block<3, 3>(0, 0); }
  Matrix<T, 3, 3> rotation() const { return linear(); }
  Transform inverse() const {
    Transform r;
    const Matrix<T, 3, 3> Rt = linear().transpose();
    r.linear() = Rt;
    r.translation() = -(Rt * translation());
    return r;
  }
  Transform operator*(const Transform& o) const { return Transform(MatrixType(m_ * o.m_)); }
  template <int O, int MR, int MC> Transform operator*(const Matrix<T, 4, 4, O, MR, MC>& o) const { return Transform(MatrixType(m_ * o)); }
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1>& p) const { return Matrix<T, 3, 1>(linear() * p + translation()); }
  Transform& operator*=(const Transform& o) { m_ = MatrixType(m_ * o.m_); return *this; }
  Transform operator*(const Quaternion<T>& q) const { Transform r(q); return *this * r; }
  Transform operator*(const Translation<T, 3>& t) const { Transform r(t); return *this * r; }
  Transform& translate(const Matrix<T, 3, 1>& t) { translation() = Matrix<T, 3, 1>(translation() + linear() * t); return *this; }
  Transform& pretranslate(const Matrix<T, 3, 1>& t) { translation() = Matrix<T, 3, 1>(translation() + t); return *this; }
  template <typename Rot> Transform& rotate(const Rot& r) { linear() = Matrix<T, 3, 3>(linear() * r.toRotationMatrix()); return *this; }
  Transform& rotate(const Matrix<T, 3, 3>& r) { linear() = Matrix<T, 3, 3>(linear() * r); return *this; }
  template <typename U> Transform<U, Dim, Mode> cast() const { return Transform<U, Dim, Mode>(m_.template cast<U>()); }
};
template <typename T> Transform<T, 3, Isometry> operator*(const Translation<T, 3>& t, const Quaternion<T>& q) { Transform<T, 3, Isometry> r(q); r.translation() = t.v; return r; }
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<float, 3, Isometry> Isometry3f;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<float, 3, Affine> Affine3f;
template <typename T, int Dim, int Mode>
std::ostream& operator<<(std::ostream& os, const Transform<T, Dim, Mode>& t) { return os << t.matrix(); }

}  // namespace Eigen
