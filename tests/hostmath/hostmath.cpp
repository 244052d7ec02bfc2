// TEST-ONLY shim: compiles the product's device math headers (fg_math.cuh / fg_factors.cuh) for the host
// with g++ so that `-m "not gpu"` tests can compare the exact formulas the CUDA kernels execute against
// the numpy oracle in this GPU-less container.  It is never linked into libfg_b200.so and is not a
// CPU fallback: it exposes single-factor evaluations only, no assembly and no solver.
#include <cstring>
#include "../../graph_slam_b200/csrc/fg_factors.cuh"
using namespace fg;
extern "C" {
void hm_so3_exp(const double* w, double* R) { so3_exp(w, R); }
void hm_so3_log(const double* R, double* w) { so3_log(R, w); }
void hm_se3_exp(const double* xi, double* T) { se3_exp(xi, T, T + 9); }
void hm_se3_log(const double* T, double* xi) { se3_log(T, T + 9, xi); }
void hm_jr(const double* w, double* J) { so3_jr(w, J); }
void hm_jr_inv(const double* w, double* J) { so3_jr_inv(w, J); }
void hm_pose_retract(const double* T, const double* xi, double* To) { pose_retract(T, T + 9, xi, To, To + 9); }
void hm_plane_retract(const double* pl, const double* v, double* out) { plane_retract(pl, v, out); }
void hm_between(const double* X1, const double* X2, const double* Z, double* r, double* J1) { between_eval<true>(X1, X2, Z, r, J1); }
void hm_prior_pose(const double* X, const double* Pm, double* r) { prior_pose_eval(X, Pm, r); }
void hm_projection(const double* X, const double* p, const double* uv, const double* K, const double* S, double* r, double* Jp, double* Jl) {
  projection_eval<true>(X, p, uv, K, S, r, Jp, Jl);
}
void hm_plane(const double* X, const double* pl, const double* z, double* r, double* Hr, double* Hp) { plane_eval<true>(X, pl, z, r, Hr, Hp); }
void hm_imu(const double* Xi, const double* vi, const double* Xj, const double* vj, const double* bi, const double* bj,
            double dt, const double* preint, const double* Hba, const double* Hbg, const double* bias_hat, const double* gravity,
            double* r, double* J) {
  ImuRec f;
  f.dt = dt;
  std::memcpy(f.preint, preint, sizeof f.preint); std::memcpy(f.Hba, Hba, sizeof f.Hba); std::memcpy(f.Hbg, Hbg, sizeof f.Hbg);
  std::memcpy(f.bias_hat, bias_hat, sizeof f.bias_hat); std::memcpy(f.gravity, gravity, sizeof f.gravity);
  imu_eval<true>(Xi, vi, Xj, vj, bi, bj, &f, r, J);
}
void hm_d_jr_c(const double* th, const double* c, double* D) { d_jr_c(th, c, D); }
void hm_chart_retract(const double* T, const double* xi, int chart, double* To) { pose_chart_retract(T, T + 9, xi, chart, To, To + 9); }
void hm_chart_local0(const double* T, int chart, double* xi) { pose_chart_local0(T, T + 9, chart, xi); }
void hm_between_chart(const double* X1, const double* X2, const double* Z, int chart, double* r, double* J1) { between_eval<true>(X1, X2, Z, r, J1, chart); }
void hm_g2o_edge(const double* X1, const double* X2, const double* Z, double* e, double* J1, double* J2) { g2o_edge_eval<true>(X1, X2, Z, e, J1, J2); }
void hm_g2o_oplus(const double* T, const double* d, double* To) { g2o_oplus(T, T + 9, d, To, To + 9); }
}
