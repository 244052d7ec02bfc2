// stand-in for the sibling package `plane` (not part of the reference repository): CPlane, a fitted plane
// n.p + d = 0 with the 4x4 covariance of (n, d).  Plane SEGMENTATION (fitting from a depth image) is front end and is not
// reproduced; what the graph wrapper reads (gtsam/gtsam_graph.cpp:1118-1298, 1346-1503) is.
#pragma once
#include <cmath>
#include <memory>
#include <vector>
#include <Eigen/Core>
#include "opencv2/opencv.hpp"
// colours of the plane viewer (global, unlike gtsam/color.h's CG::COLOR) and its pixel marker (display only: a no-op)
typedef enum { RED = 0, GREEN, BLUE, PURPLE, WHITE, YELLOW, DARK } COLOR;
inline void markColor(cv::Mat&, std::vector<int>&, COLOR) {}
struct Point { float x = 0, y = 0, z = 0; };
struct Cloud { std::vector<Point> points; };
typedef std::shared_ptr<Cloud> CloudPtr;
class CPlane {
 public:
  double nx_ = 0, ny_ = 0, nz_ = 1, d1_ = 0;   // unit normal and offset
  double m_CP[4][4];                           // covariance of (nx, ny, nz, d), row-major
  double m_E_Sdi = 0;                          // estimated variance of d
  CPlane() { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m_CP[i][j] = (i == j) ? 1e-4 : 0.0; }
  double dis2plane(double px, double py, double pz) const { return std::fabs(nx_ * px + ny_ * py + nz_ * pz + d1_); }
  template <class M> void getNVCov(M& S) const { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S(i, j) = m_CP[i][j]; }
  double getTraceSVN() const { return m_CP[0][0] + m_CP[1][1] + m_CP[2][2]; }
  // make the covariance usable: symmetric, positive diagonal (the reference calls this when its conditioning check fails)
  void regularizeCOV() {
    for (int i = 0; i < 4; ++i) {
      for (int j = 0; j < i; ++j) m_CP[i][j] = m_CP[j][i] = 0.0;
      if (!(m_CP[i][i] > 1e-8) || m_CP[i][i] != m_CP[i][i]) m_CP[i][i] = 1e-4;
    }
  }
  template <class... A> bool computeCOVSparse(A&&...) { return false; }      // covariance from the supporting pixels: front end
  bool computeParameters(CloudPtr&) { return false; }
  template <class V> bool computeParameters(std::vector<V>&) { return false; }
  void print_m1() const {}
};

// Small-matrix checks the reference's addPlaneFactor calls (gtsam/gtsam_graph.cpp:1167,1251-1254).  Their definitions live
// in a sibling package that is not part of the reference repository; these follow what the call sites expect of them:
// MatrixCheck -- a usable covariance (finite, positive diagonal); DominateCheck -- diagonally dominant; TriangleMatrix --
// make it so by dropping the off-diagonal terms.
template <class M> bool MatrixCheck(const M& m) {
  for (int i = 0; i < m.rows(); ++i) {
    for (int j = 0; j < m.cols(); ++j) if (!std::isfinite((double)m(i, j))) return false;
    if (!(m(i, i) > 0)) return false;
  }
  return true;
}
template <class M> bool DominateCheck(const M& m) {
  for (int i = 0; i < m.rows(); ++i) {
    double off = 0;
    for (int j = 0; j < m.cols(); ++j) if (j != i) off += std::fabs((double)m(i, j));
    if (std::fabs((double)m(i, i)) < off) return false;
  }
  return true;
}
template <class M> void TriangleMatrix(M& m) {
  for (int i = 0; i < m.rows(); ++i) for (int j = 0; j < m.cols(); ++j) if (i != j) m(i, j) = 0;
}
template <class A, class B> bool MatrixEqual(const A& a, const B& b, double tol) {
  for (int i = 0; i < a.rows(); ++i) for (int j = 0; j < a.cols(); ++j) if (std::fabs((double)a(i, j) - (double)b(i, j)) > tol) return false;
  return true;
}
