// g2o_lite.h -- header-only subset of the g2o API that CGraphG2O uses (g2o/g2o_graph.cpp:30-31,65-134,241-258; SURVEY.md
// section 8 a14), as host-side value types that forward all numeric work to the C ABI (include/fg_abi.h): the optimiser
// holds the vertices and edges, `optimize(n)` runs n Levenberg iterations with g2o's own rule on the GPU
// (fg_optimize_g2o), `chi2()` is fg_g2o_chi2.  Its purpose is that the reference's g2o/g2o_graph.cpp compiles UNCHANGED:
// the headers under compat/g2o/... all include this file.  BlockSolver / LinearSolverCSparse / OptimizationAlgorithmLevenberg
// are tags here -- the linear solver is the device's sparse block Cholesky.
#pragma once
#include <cmath>
#include <iostream>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include "../../include/fg_abi.h"

namespace g2o {

class HyperGraph {
 public:
  class Edge;
  class Vertex {
    int id_ = -1;
   public:
    virtual ~Vertex() {}
    int id() const { return id_; }
    void setId(int i) { id_ = i; }
  };
  class Edge {
   protected:
    std::vector<Vertex*> v_;
   public:
    virtual ~Edge() {}
    std::vector<Vertex*>& vertices() { return v_; }
    const std::vector<Vertex*>& vertices() const { return v_; }
    void resize(size_t n) { v_.resize(n, nullptr); }
  };
  typedef std::set<Edge*> EdgeSet;
  typedef std::map<int, Vertex*> VertexIDMap;
};
namespace OptimizableGraph_ {}
class OptimizableGraph : public HyperGraph {
 public:
  typedef HyperGraph::Vertex Vertex;
  typedef HyperGraph::Edge Edge;
};

class VertexSE3 : public HyperGraph::Vertex {
  Eigen::Isometry3d est_;
  bool fixed_ = false;
 public:
  VertexSE3() { est_.setIdentity(); }
  const Eigen::Isometry3d& estimate() const { return est_; }
  void setEstimate(const Eigen::Isometry3d& e) { est_ = e; }
  bool fixed() const { return fixed_; }
  void setFixed(bool f) { fixed_ = f; }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

class EdgeSE3 : public HyperGraph::Edge {
  Eigen::Isometry3d z_;
  Eigen::Matrix<double, 6, 6> info_;
 public:
  EdgeSE3() : info_(Eigen::Matrix<double, 6, 6>::Identity()) { z_.setIdentity(); resize(2); }
  void setMeasurement(const Eigen::Isometry3d& m) { z_ = m; }
  const Eigen::Isometry3d& measurement() const { return z_; }
  template <int O, int MR, int MC> void setInformation(const Eigen::Matrix<double, 6, 6, O, MR, MC>& i) { info_ = i; }
  const Eigen::Matrix<double, 6, 6>& information() const { return info_; }
  void setRobustKernel(void*) {}
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

// solver stack of CGraphG2O::createOptimizer (g2o_graph.cpp:65-77): tags only
template <int P, int L> struct BlockSolverTraits { typedef Eigen::Matrix<double, P, P> PoseMatrixType; typedef Eigen::Matrix<double, L, L> LandmarkMatrixType; };
template <class MatrixType> class LinearSolverCSparse { public: void setBlockOrdering(bool) {} };
template <class MatrixType> class LinearSolverCholmod { public: void setBlockOrdering(bool) {} };
template <class Traits> class BlockSolver {
 public:
  typedef typename Traits::PoseMatrixType PoseMatrixType;
  typedef typename Traits::LandmarkMatrixType LandmarkMatrixType;
  template <class LS> explicit BlockSolver(LS* ls) { delete ls; }
};
class OptimizationAlgorithm { public: virtual ~OptimizationAlgorithm() {} };
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithm {
 public:
  template <class S> explicit OptimizationAlgorithmLevenberg(S* solver) { delete solver; }
  void setUserLambdaInit(double) {}
  void setMaxTrialsAfterFailure(int) {}
};

class SparseOptimizer : public OptimizableGraph {
  std::map<int, VertexSE3*> vertices_;
  std::vector<EdgeSE3*> edges_;
  OptimizationAlgorithm* algo_ = nullptr;
  fg_ctx* c_ = nullptr;          // the graph as last handed to the device; rebuilt when vertices or edges were added
  size_t built_v_ = 0, built_e_ = 0;
  double chi2_ = 0.0;
  bool verbose_ = false;
  static fg_key key(int id) { return (fg_key('x') << 56) | (fg_key)(unsigned)id; }
  static void to12(const Eigen::Isometry3d& T, double a[12]) { for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) a[3 * i + j] = T(i, j); a[9 + i] = T(i, 3); } }
  void check(int rc, const char* what) { if (rc != FG_OK) throw std::runtime_error(std::string(what) + ": " + fg_last_error(c_)); }
  // (re)build the device graph from the current vertices and edges
  void sync() {
    if (c_ && built_v_ == vertices_.size() && built_e_ == edges_.size()) {
      for (auto& kv : vertices_) { double a[12]; to12(kv.second->estimate(), a); check(fg_update_value(c_, key(kv.first), a), "fg_update_value"); }
      return;
    }
    if (c_) fg_destroy(c_);
    c_ = fg_create(0, 0, 1);
    if (!c_) throw std::runtime_error("fg_create failed: no CUDA device (this backend has no CPU fallback)");
    for (auto& kv : vertices_) {
      double a[12]; to12(kv.second->estimate(), a);
      check(fg_add_pose(c_, key(kv.first), a), "fg_add_pose");
      if (kv.second->fixed()) check(fg_set_fixed(c_, key(kv.first), 1), "fg_set_fixed");
    }
    for (EdgeSE3* e : edges_) {
      double a[12], info[36]; to12(e->measurement(), a);
      for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) info[6 * i + j] = e->information()(i, j);
      check(fg_add_g2o_edge(c_, key(e->vertices()[0]->id()), key(e->vertices()[1]->id()), a, info), "fg_add_g2o_edge");
    }
    built_v_ = vertices_.size(); built_e_ = edges_.size();
  }
  void readback() {
    for (auto& kv : vertices_) {
      double a[12]; int n = 0;
      check(fg_get_value(c_, key(kv.first), a, &n), "fg_get_value");
      Eigen::Isometry3d T; T.setIdentity();
      for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T(i, j) = a[3 * i + j]; T(i, 3) = a[9 + i]; }
      kv.second->setEstimate(T);
    }
  }
 public:
  SparseOptimizer() {}
  ~SparseOptimizer() { clear(); delete algo_; }
  SparseOptimizer(const SparseOptimizer&) = delete;
  void setVerbose(bool v) { verbose_ = v; }
  void setAlgorithm(OptimizationAlgorithm* a) { delete algo_; algo_ = a; }
  bool addVertex(VertexSE3* v) { if (vertices_.count(v->id())) return false; vertices_[v->id()] = v; return true; }
  bool addEdge(EdgeSE3* e) { if (!e->vertices()[0] || !e->vertices()[1]) return false; edges_.push_back(e); return true; }
  HyperGraph::Vertex* vertex(int id) { auto it = vertices_.find(id); return it == vertices_.end() ? nullptr : it->second; }
  const std::map<int, VertexSE3*>& vertices() const { return vertices_; }
  const std::vector<EdgeSE3*>& edges() const { return edges_; }
  void clear() {
    for (auto& kv : vertices_) delete kv.second;
    for (EdgeSE3* e : edges_) delete e;
    vertices_.clear(); edges_.clear();
    if (c_) { fg_destroy(c_); c_ = nullptr; }
    built_v_ = built_e_ = 0;
  }
  bool initializeOptimization(int = 0) { sync(); return true; }
  // n Levenberg iterations of OptimizationAlgorithmLevenberg on the device; lambda is initialised at the start of the call, as
  // g2o does.  Returns the iterations performed (0 when the graph has no edge to optimise).
  int optimize(int iterations, bool = false) {
    if (edges_.empty() || iterations <= 0) return 0;
    sync();
    fg_g2o_params p; fg_g2o_params_default(&p);
    p.iterations = iterations; p.iterations_per_call = iterations;
    fg_g2o_report rep;
    check(fg_optimize_g2o(c_, &p, &rep), "fg_optimize_g2o");
    readback();
    chi2_ = rep.final_chi2;
    if (verbose_) std::cerr << "g2o iterations " << rep.iterations << " chi2 " << rep.final_chi2 << " lambda " << rep.lambda << std::endl;
    return rep.iterations;
  }
  void computeActiveErrors() {
    if (edges_.empty()) { chi2_ = 0.0; return; }
    sync();
    check(fg_g2o_chi2(c_, &chi2_), "fg_g2o_chi2");
  }
  double chi2() const { return chi2_; }
  double activeChi2() const { return chi2_; }
  // SparseOptimizer::save: the .g2o text format (VERTEX_SE3:QUAT / EDGE_SE3:QUAT, upper-triangular information)
  bool save(std::ostream& os) const {
    os.precision(17);
    for (auto& kv : vertices_) {
      const Eigen::Isometry3d& T = kv.second->estimate();
      const Eigen::Quaterniond q(Eigen::Matrix3d(T.rotation()));
      os << "VERTEX_SE3:QUAT " << kv.first << " " << T(0, 3) << " " << T(1, 3) << " " << T(2, 3) << " " << q.x() << " " << q.y() << " " << q.z() << " " << q.w() << "\n";
      if (kv.second->fixed()) os << "FIX " << kv.first << "\n";
    }
    for (const EdgeSE3* e : edges_) {
      const Eigen::Isometry3d& T = e->measurement();
      const Eigen::Quaterniond q(Eigen::Matrix3d(T.rotation()));
      os << "EDGE_SE3:QUAT " << e->vertices()[0]->id() << " " << e->vertices()[1]->id() << " " << T(0, 3) << " " << T(1, 3) << " " << T(2, 3) << " "
         << q.x() << " " << q.y() << " " << q.z() << " " << q.w();
      for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) os << " " << e->information()(i, j);
      os << "\n";
    }
    return os.good();
  }
  bool save(const char* f) const;
};

}  // namespace g2o
