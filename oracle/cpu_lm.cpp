// oracle/cpu_lm.cpp -- TIMING BASELINE ONLY (test infrastructure, not part of the product, never linked into
// libfg_b200.so).  One Levenberg-Marquardt iteration of the BA + IMU graph CGraphGT builds
// (gtsam/gtsam_graph.cpp:370-448 projection factors + point priors, test_ba_imu_graph.cpp:239-244 CombinedImuFactor,
// gtsam_graph.cpp:320-368 firstNode priors) on the host cores, OpenMP over all of them: linearise, eliminate the
// landmarks (Schur complement), factor the reduced pose system (block-banded Cholesky in frame order, 15 x 15 blocks
// [X V B]), back-substitute, retract, evaluate the new error -- the same unit of work bench.py times on the GPU.
// It stands in for "the reference's GTSAM CPU path", which cannot be built here: GTSAM would run the same
// arithmetic through virtual linearize() calls and a multifrontal elimination, single threaded unless built with TBB.
//
// The per-factor formulas are the product's own math headers (fg_math.cuh / fg_factors.cuh) compiled for the host, the
// way tests/hostmath does; parity is NOT defined by this file -- the numpy oracle (oracle/*.py) is the checker, and
// tests/test_cpu_baseline.py checks this file against it.
//
//   g++ -O3 -mavx2 -mfma -fopenmp -shared -fPIC oracle/cpu_lm.cpp -o oracle/_build/libcpu_lm.so
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <omp.h>
#include "../graph_slam_b200/csrc/fg_factors.cuh"
using namespace fg;

namespace {
const int FD = 15;                       // frame dims: [X 6, V 3, B 6]
const int FB = FD * FD;
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Problem {
  int P; double *pose, *vel, *bias;
  const double *pp_T, *pp_info, *pv_mean, *pv_info, *pb_mean, *pb_info;
  int n_imu; const ImuRec* imu;
  int64_t L; double* pts; const double* pt_mean; double pt_w;
  int64_t M; const int *obs_pose, *obs_point; const double* obs_uv; double obs_w;
  const double *K, *sensor;
};

// chi2 of every factor at (pose, vel, bias, pts)
double total_chi2(const Problem& q, const double* pose, const double* vel, const double* bias, const double* pts, const std::vector<int64_t>& lm_ptr) {
  double e = 0.0;
  {
    double r[6], wr[6];
    prior_pose_eval(pose, q.pp_T, r);
    for (int i = 0; i < 6; ++i) { wr[i] = 0; for (int j = 0; j < 6; ++j) wr[i] += q.pp_info[6 * i + j] * r[j]; e += r[i] * wr[i]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) e += (vel[i] - q.pv_mean[i]) * q.pv_info[3 * i + j] * (vel[j] - q.pv_mean[j]);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) e += (bias[i] - q.pb_mean[i]) * q.pb_info[6 * i + j] * (bias[j] - q.pb_mean[j]);
  }
#pragma omp parallel for reduction(+ : e) schedule(static)
  for (int f = 0; f < q.n_imu; ++f) {
    double r[15], J[450];
    imu_eval<false>(pose + 12 * f, vel + 3 * f, pose + 12 * (f + 1), vel + 3 * (f + 1), bias + 6 * f, bias + 6 * (f + 1), q.imu + f, r, J);
    const double* Om = q.imu[f].info;
    for (int i = 0; i < 15; ++i) { double s = 0; for (int j = 0; j < 15; ++j) s += Om[15 * i + j] * r[j]; e += r[i] * s; }
  }
#pragma omp parallel for reduction(+ : e) schedule(static, 256)
  for (int64_t l = 0; l < q.L; ++l) {
    const double* p = pts + 3 * l;
    for (int i = 0; i < 3; ++i) { const double d = p[i] - q.pt_mean[3 * l + i]; e += q.pt_w * d * d; }
    for (int64_t o = lm_ptr[l]; o < lm_ptr[l + 1]; ++o) {
      double r[2], Jp[12], Jl[6];
      projection_eval<false>(pose + 12 * (int64_t)q.obs_pose[o], p, q.obs_uv + 2 * o, q.K, q.sensor, r, Jp, Jl);
      e += q.obs_w * (r[0] * r[0] + r[1] * r[1]);
    }
  }
  return e;
}
}  // namespace

// out[0] error before, [1] error after, [2..6] seconds: linearise, Schur, factor, solve + back-substitution + retract, error;
// out[7] = half bandwidth in frames.  Returns 0, or 1 when the damped reduced system is not positive definite (state unchanged).
extern "C" int cpu_lm_iteration(int P, double* pose, double* vel, double* bias, const double* pp_T, const double* pp_info,
                                const double* pv_mean, const double* pv_info, const double* pb_mean, const double* pb_info, int n_imu,
                                const void* imu_recs, int64_t L, double* pts, const double* pt_mean, double pt_sigma, int64_t M,
                                const int* obs_pose, const int* obs_point, const double* obs_uv, double obs_sigma, const double* K9,
                                const double* sensor12, double lambda, int nthreads, double* out) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  Problem q{P, pose, vel, bias, pp_T, pp_info, pv_mean, pv_info, pb_mean, pb_info, n_imu, (const ImuRec*)imu_recs,
            L, pts, pt_mean, 1.0 / (pt_sigma * pt_sigma), M, obs_pose, obs_point, obs_uv, 1.0 / (obs_sigma * obs_sigma), K9, sensor12};
  // ---- index structures (observations must be sorted by landmark)
  std::vector<int64_t> lm_ptr(L + 1, 0), pose_ptr(P + 1, 0);
  for (int64_t o = 0; o < M; ++o) { lm_ptr[obs_point[o] + 1]++; pose_ptr[obs_pose[o] + 1]++; }
  for (int64_t l = 0; l < L; ++l) lm_ptr[l + 1] += lm_ptr[l];
  for (int p = 0; p < P; ++p) pose_ptr[p + 1] += pose_ptr[p];
  std::vector<int64_t> pose_obs(M);
  { std::vector<int64_t> cur(pose_ptr.begin(), pose_ptr.end() - 1); for (int64_t o = 0; o < M; ++o) pose_obs[cur[obs_pose[o]]++] = o; }
  int w = 1;
  for (int64_t l = 0; l < L; ++l)
    if (lm_ptr[l + 1] > lm_ptr[l]) {
      int lo = P, hi = -1;
      for (int64_t o = lm_ptr[l]; o < lm_ptr[l + 1]; ++o) { lo = std::min(lo, obs_pose[o]); hi = std::max(hi, obs_pose[o]); }
      w = std::max(w, hi - lo);
    }
  out[7] = w;
  const size_t W1 = (size_t)w + 1;
  std::vector<double> S((size_t)P * W1 * FB, 0.0), rhs((size_t)P * FD, 0.0);
  auto blk = [&](int f, int d) -> double* { return S.data() + ((size_t)f * W1 + d) * FB; };      // block (frame f, frame f - d), rows of f
  std::vector<double> Wm((size_t)M * 18), V((size_t)L * 9), gl((size_t)L * 3), Vinv((size_t)L * 9);
  double t0 = now();
  // ---- linearise
  double chi2 = 0.0;
#pragma omp parallel for reduction(+ : chi2) schedule(static, 256)
  for (int64_t l = 0; l < L; ++l) {
    const double* p = pts + 3 * l;
    double Vl[9] = {q.pt_w, 0, 0, 0, q.pt_w, 0, 0, 0, q.pt_w}, g[3];
    for (int i = 0; i < 3; ++i) { const double d = p[i] - pt_mean[3 * l + i]; g[i] = q.pt_w * d; chi2 += q.pt_w * d * d; }
    for (int64_t o = lm_ptr[l]; o < lm_ptr[l + 1]; ++o) {
      double r[2], Jp[12], Jl[6];
      projection_eval<true>(pose + 12 * (int64_t)obs_pose[o], p, obs_uv + 2 * o, K9, sensor12, r, Jp, Jl);
      chi2 += q.obs_w * (r[0] * r[0] + r[1] * r[1]);
      double* Wo = Wm.data() + 18 * o;
      for (int i = 0; i < 6; ++i) for (int c = 0; c < 3; ++c) Wo[3 * i + c] = q.obs_w * (Jp[i] * Jl[c] + Jp[6 + i] * Jl[3 + c]);
      for (int a = 0; a < 3; ++a) {
        for (int c = 0; c < 3; ++c) Vl[3 * a + c] += q.obs_w * (Jl[a] * Jl[c] + Jl[3 + a] * Jl[3 + c]);
        g[a] += q.obs_w * (Jl[a] * r[0] + Jl[3 + a] * r[1]);
      }
    }
    std::memcpy(V.data() + 9 * l, Vl, sizeof Vl); std::memcpy(gl.data() + 3 * l, g, sizeof g);
  }
#pragma omp parallel for schedule(dynamic, 4)
  for (int p = 0; p < P; ++p) {
    double U[36] = {0}, g[6] = {0};
    for (int64_t k = pose_ptr[p]; k < pose_ptr[p + 1]; ++k) {
      const int64_t o = pose_obs[k];
      double r[2], Jp[12], Jl[6];
      projection_eval<true>(pose + 12 * (int64_t)p, pts + 3 * (int64_t)obs_point[o], obs_uv + 2 * o, K9, sensor12, r, Jp, Jl);
      for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) U[6 * i + j] += q.obs_w * (Jp[i] * Jp[j] + Jp[6 + i] * Jp[6 + j]);
        g[i] += q.obs_w * (Jp[i] * r[0] + Jp[6 + i] * r[1]);
      }
    }
    double* D = blk(p, 0);
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) D[FD * i + j] += U[6 * i + j]; rhs[(size_t)FD * p + i] += g[i]; }
  }
  // IMU factors couple frame f and f + 1: even f first, then odd f (no two threads touch the same blocks)
  for (int parity = 0; parity < 2; ++parity) {
#pragma omp parallel for reduction(+ : chi2) schedule(static)
    for (int f = parity; f < n_imu; f += 2) {
      double r[15], J[450], OJ[450], wr[15];
      imu_eval<true>(pose + 12 * f, vel + 3 * f, pose + 12 * (f + 1), vel + 3 * (f + 1), bias + 6 * f, bias + 6 * (f + 1), q.imu + f, r, J);
      const double* Om = q.imu[f].info;
      for (int i = 0; i < 15; ++i) { double s = 0; for (int j = 0; j < 15; ++j) s += Om[15 * i + j] * r[j]; wr[i] = s; chi2 += r[i] * s; }
      for (int i = 0; i < 15; ++i) for (int c = 0; c < 30; ++c) { double s = 0; for (int k = 0; k < 15; ++k) s += Om[15 * i + k] * J[30 * k + c]; OJ[30 * i + c] = s; }
      // factor columns: [Xi 0-5, vi 6-8, Xj 9-14, vj 15-17, bi 18-23, bj 24-29] -> (frame, offset in frame [X 0, V 6, B 9])
      int fr[30], of[30];
      for (int c = 0; c < 30; ++c) {
        const int seg = c < 6 ? 0 : c < 9 ? 1 : c < 15 ? 2 : c < 18 ? 3 : c < 24 ? 4 : 5;
        const int start[6] = {0, 6, 9, 15, 18, 24}, frame[6] = {0, 0, 1, 1, 0, 1}, off[6] = {0, 6, 0, 6, 9, 9};
        fr[c] = f + frame[seg]; of[c] = off[seg] + c - start[seg];
      }
      for (int a = 0; a < 30; ++a) {
        double ga = 0; for (int k = 0; k < 15; ++k) ga += J[30 * k + a] * wr[k];
        rhs[(size_t)FD * fr[a] + of[a]] += ga;
        for (int b = 0; b < 30; ++b) {
          if (fr[b] > fr[a]) continue;                 // block (fr[a], fr[b]) with fr[b] <= fr[a]
          double h = 0; for (int k = 0; k < 15; ++k) h += J[30 * k + a] * OJ[30 * k + b];
          blk(fr[a], fr[a] - fr[b])[FD * of[a] + of[b]] += h;
        }
      }
    }
  }
  {  // firstNode priors on frame 0
    double r[6], wr[6];
    prior_pose_eval(pose, pp_T, r);
    double* D = blk(0, 0);
    for (int i = 0; i < 6; ++i) { wr[i] = 0; for (int j = 0; j < 6; ++j) { wr[i] += pp_info[6 * i + j] * r[j]; D[FD * i + j] += pp_info[6 * i + j]; } chi2 += r[i] * wr[i]; rhs[i] += wr[i]; }
    for (int i = 0; i < 3; ++i) { double s = 0; for (int j = 0; j < 3; ++j) { s += pv_info[3 * i + j] * (vel[j] - pv_mean[j]); D[FD * (6 + i) + 6 + j] += pv_info[3 * i + j]; } chi2 += (vel[i] - pv_mean[i]) * s; rhs[6 + i] += s; }
    for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) { s += pb_info[6 * i + j] * (bias[j] - pb_mean[j]); D[FD * (9 + i) + 9 + j] += pb_info[6 * i + j]; } chi2 += (bias[i] - pb_mean[i]) * s; rhs[9 + i] += s; }
  }
  double t1 = now();
  // ---- Schur complement: thread-owned pose rows, S_pq -= W_a Vinv W_b^T, rhs_p -= W_a Vinv g_l
#pragma omp parallel for schedule(static, 256)
  for (int64_t l = 0; l < L; ++l) {
    double A[9]; std::memcpy(A, V.data() + 9 * l, sizeof A);
    A[0] += lambda; A[4] += lambda; A[8] += lambda;
    inv3(A, Vinv.data() + 9 * l);
  }
#pragma omp parallel for schedule(dynamic, 4)
  for (int p = 0; p < P; ++p) {
    for (int64_t k = pose_ptr[p]; k < pose_ptr[p + 1]; ++k) {
      const int64_t a = pose_obs[k];
      const int64_t l = obs_point[a];
      const double* Wa = Wm.data() + 18 * a; const double* Vi = Vinv.data() + 9 * l; const double* g = gl.data() + 3 * l;
      double Y[18];
      for (int i = 0; i < 6; ++i) for (int c = 0; c < 3; ++c) Y[3 * i + c] = Wa[3 * i] * Vi[c] + Wa[3 * i + 1] * Vi[3 + c] + Wa[3 * i + 2] * Vi[6 + c];
      for (int i = 0; i < 6; ++i) rhs[(size_t)FD * p + i] -= Y[3 * i] * g[0] + Y[3 * i + 1] * g[1] + Y[3 * i + 2] * g[2];
      for (int64_t b = lm_ptr[l]; b < lm_ptr[l + 1]; ++b) {
        const int qq = obs_pose[b];
        if (qq > p) continue;
        const double* Wb = Wm.data() + 18 * b;
        double* D = blk(p, p - qq);
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) D[FD * i + j] -= Y[3 * i] * Wb[3 * j] + Y[3 * i + 1] * Wb[3 * j + 1] + Y[3 * i + 2] * Wb[3 * j + 2];
      }
    }
  }
  for (int f = 0; f < P; ++f) for (int i = 0; i < FD; ++i) blk(f, 0)[FD * i + i] += lambda;
  double t2 = now();
  // ---- block-banded Cholesky (right looking): for frame j, L_jj, the w blocks below it, then the trailing window
  int bad = 0;
  for (int j = 0; j < P && !bad; ++j) {
    double* D = blk(j, 0);
    for (int c = 0; c < FD; ++c) {
      double s = D[FD * c + c]; for (int k = 0; k < c; ++k) s -= D[FD * c + k] * D[FD * c + k];
      if (!(s > 0)) { bad = 1; break; }
      const double d = std::sqrt(s); D[FD * c + c] = d;
      for (int r = c + 1; r < FD; ++r) { double t = D[FD * r + c]; for (int k = 0; k < c; ++k) t -= D[FD * r + k] * D[FD * c + k]; D[FD * r + c] = t / d; }
    }
    if (bad) break;
    const int imax = std::min(P - 1, j + w);
#pragma omp parallel
    {
#pragma omp for schedule(static)
      for (int i = j + 1; i <= imax; ++i) {              // L_ij = S_ij L_jj^-T
        double* B = blk(i, i - j);
        for (int r = 0; r < FD; ++r)
          for (int c = 0; c < FD; ++c) { double t = B[FD * r + c]; for (int k = 0; k < c; ++k) t -= B[FD * r + k] * D[FD * c + k]; B[FD * r + c] = t / D[FD * c + c]; }
      }
      const int n = imax - j;                            // trailing blocks (i, k), j < k <= i <= imax: S_ik -= L_ij L_kj^T
#pragma omp for schedule(dynamic, 8)
      for (int t = 0; t < n * (n + 1) / 2; ++t) {
        int ii = (int)((std::sqrt(8.0 * t + 1.0) - 1.0) / 2.0); while ((ii + 1) * (ii + 2) / 2 <= t) ++ii; while (ii * (ii + 1) / 2 > t) --ii;
        const int kk = t - ii * (ii + 1) / 2;
        const int i = j + 1 + ii, k = j + 1 + kk;
        const double* A = blk(i, i - j); const double* Bk = blk(k, k - j);
        double* C = blk(i, i - k);
        for (int r = 0; r < FD; ++r) for (int c = 0; c < FD; ++c) { double s = 0; for (int m = 0; m < FD; ++m) s += A[FD * r + m] * Bk[FD * c + m]; C[FD * r + c] -= s; }
      }
    }
  }
  double t3 = now();
  if (bad) { out[0] = 0.5 * chi2; out[1] = INFINITY; out[2] = t1 - t0; out[3] = t2 - t1; out[4] = t3 - t2; out[5] = out[6] = 0; return 1; }
  // ---- solve (L L^T) delta = -rhs; blocks above the diagonal of a diagonal block are ignored
  std::vector<double> x((size_t)P * FD);
  for (size_t i = 0; i < x.size(); ++i) x[i] = -rhs[i];
  for (int f = 0; f < P; ++f) {                           // forward
    double* xf = x.data() + (size_t)FD * f;
    for (int d = std::min(w, f); d >= 1; --d) { const double* B = blk(f, d); const double* xg = x.data() + (size_t)FD * (f - d); for (int r = 0; r < FD; ++r) { double s = 0; for (int c = 0; c < FD; ++c) s += B[FD * r + c] * xg[c]; xf[r] -= s; } }
    const double* D = blk(f, 0);
    for (int r = 0; r < FD; ++r) { double s = xf[r]; for (int c = 0; c < r; ++c) s -= D[FD * r + c] * xf[c]; xf[r] = s / D[FD * r + r]; }
  }
  for (int f = P - 1; f >= 0; --f) {                      // backward
    double* xf = x.data() + (size_t)FD * f;
    const double* D = blk(f, 0);
    for (int r = FD - 1; r >= 0; --r) { double s = xf[r]; for (int c = r + 1; c < FD; ++c) s -= D[FD * c + r] * xf[c]; xf[r] = s / D[FD * r + r]; }
    for (int d = 1; d <= std::min(w, f); ++d) { const double* B = blk(f, d); double* xg = x.data() + (size_t)FD * (f - d); for (int c = 0; c < FD; ++c) { double s = 0; for (int r = 0; r < FD; ++r) s += B[FD * r + c] * xf[r]; xg[c] -= s; } }
  }
  // ---- landmarks: delta_l = -Vinv (g_l + sum_o W_o^T delta_p), retraction
  std::vector<double> npose((size_t)P * 12), nvel((size_t)P * 3), nbias((size_t)P * 6), npts((size_t)L * 3);
#pragma omp parallel for schedule(static, 256)
  for (int64_t l = 0; l < L; ++l) {
    double s[3] = {gl[3 * l], gl[3 * l + 1], gl[3 * l + 2]};
    for (int64_t o = lm_ptr[l]; o < lm_ptr[l + 1]; ++o) {
      const double* Wo = Wm.data() + 18 * o; const double* dp = x.data() + (size_t)FD * obs_pose[o];
      for (int i = 0; i < 6; ++i) for (int c = 0; c < 3; ++c) s[c] += Wo[3 * i + c] * dp[i];
    }
    const double* Vi = Vinv.data() + 9 * l;
    for (int c = 0; c < 3; ++c) npts[3 * l + c] = pts[3 * l + c] - (Vi[3 * c] * s[0] + Vi[3 * c + 1] * s[1] + Vi[3 * c + 2] * s[2]);
  }
#pragma omp parallel for schedule(static)
  for (int f = 0; f < P; ++f) {
    const double* d = x.data() + (size_t)FD * f;
    pose_retract(pose + 12 * f, pose + 12 * f + 9, d, npose.data() + 12 * f, npose.data() + 12 * f + 9);
    for (int i = 0; i < 3; ++i) nvel[3 * f + i] = vel[3 * f + i] + d[6 + i];
    for (int i = 0; i < 6; ++i) nbias[6 * f + i] = bias[6 * f + i] + d[9 + i];
  }
  double t4 = now();
  const double chi2_new = total_chi2(q, npose.data(), nvel.data(), nbias.data(), npts.data(), lm_ptr);
  double t5 = now();
  std::memcpy(pose, npose.data(), sizeof(double) * npose.size()); std::memcpy(vel, nvel.data(), sizeof(double) * nvel.size());
  std::memcpy(bias, nbias.data(), sizeof(double) * nbias.size()); std::memcpy(pts, npts.data(), sizeof(double) * npts.size());
  out[0] = 0.5 * chi2; out[1] = 0.5 * chi2_new; out[2] = t1 - t0; out[3] = t2 - t1; out[4] = t3 - t2; out[5] = t4 - t3; out[6] = t5 - t4;
  return 0;
}

extern "C" int cpu_lm_max_threads(void) { return omp_get_max_threads(); }
