// fg_kernels.cu -- hand-written sm_100a CUDA for the Gauss-Newton / LM inner loop (fp64).
//
// Kernels (names follow SURVEY.md section 7):
//   K2  k_prior_pose / k_prior_vec<D>     priors                      (gtsam_graph.cpp:341,362-367)
//   K1  k_between                         BetweenFactor<Pose3>        (gtsam_graph.cpp:691-692)
//   K4  k_imu                             CombinedImuFactor           (test_vro_imu_graph.cpp:191-196)
//   K5  k_plane                           OrientedPlane3Factor        (gtsam_graph.cpp:1265)
//   K6  k_lm_prior / k_proj_obs / k_proj_pose    projection factors (gtsam_graph.cpp:370-448); the landmark
//       Schur complement itself is in fg_schur.cu
//   K9  k_lm_backsub_obs / k_lm_update    landmark back-substitution + retraction
//   K10 k_retract_reduced                 SE3 / vector / plane retraction (Values::retract)
//   K3  k_preintegrate                    PreintegratedCombinedMeasurements::integrateMeasurement loop
//                                                                     (imu_base.cpp:76-85)
// All kernels are templated on JAC: JAC=true linearises and assembles, JAC=false evaluates chi2 only
// (graph.error, gtsam_graph.cpp:173-176).
#include <cstdio>
#include "fg_internal.h"

namespace fg {

struct Vals { const double* v[T_COUNT]; };

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;
}
// Scalar sums (chi2, g^T delta, |delta|^2) are deterministic: every block of every kernel of a pass writes its own partial
// into a slot (no fp64 atomics, whose arrival order varies from run to run), and k_sum_partials adds the slots in a fixed order.
// Every thread of the block must call.
__device__ __forceinline__ double block_sum(double e) {
  __shared__ double bs_part[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  e = warp_sum(e);
  __syncthreads();                       // a second call in the same kernel reuses bs_part
  if (lane == 0) bs_part[w] = e;
  __syncthreads();
  double v = 0.0;
  if (w == 0) { v = lane < nw ? bs_part[lane] : 0.0; v = warp_sum(v); }
  return v;                              // valid in thread 0
}
__device__ __forceinline__ void chi2_accumulate(double e, double* part) {
  const double v = block_sum(e);
  if (threadIdx.x == 0) part[blockIdx.x] = v;
}
__global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ part, int n, double* target) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += part[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) { if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d]; __syncthreads(); }
  if (threadIdx.x == 0) *target = sh[0];
}

// 128-bit loads of a pose record (12 doubles, 96 B, 16 B aligned)
__device__ __forceinline__ void load_pose(const double* __restrict__ base, int idx, double* X) {
  const double2* p = reinterpret_cast<const double2*>(base + (int64_t)idx * 12);
#pragma unroll
  for (int i = 0; i < 6; ++i) { double2 v = __ldg(p + i); X[2 * i] = v.x; X[2 * i + 1] = v.y; }
}

// sym mat-vec helpers on row-major dense
template <int M, int N>
__device__ __forceinline__ void matvec(const double* A, const double* x, double* y) {
#pragma unroll
  for (int i = 0; i < M; ++i) { double s = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) s += A[i * N + j] * x[j]; y[i] = s; }
}

// ------------------------------------------------------------------ priors
template <bool JAC>
__global__ void k_prior_pose(int n, const int* __restrict__ var, const double* __restrict__ mean,
                             const double* __restrict__ info, Vals vals, const int* __restrict__ off,
                             SysView sys, double* g_r, double* chi2, int chart) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (f < n) {
    double X[12], Pm[12], r[6], wr[6];
    load_pose(vals.v[T_POSE], var[f], X);
#pragma unroll
    for (int i = 0; i < 12; ++i) Pm[i] = mean[12 * f + i];
    prior_pose_eval(X, Pm, r, chart);
    const double* Om = info + 36 * f;
    matvec<6, 6>(Om, r, wr);
#pragma unroll
    for (int i = 0; i < 6; ++i) e += r[i] * wr[i];
    if (JAC) {
      int o = off[var[f]];
      sys_add_block(sys, o, 6, o, 6, Om, 6);
#pragma unroll
      for (int i = 0; i < 6; ++i) atomicAdd(&g_r[o + i], wr[i]);
    }
  }
  chi2_accumulate(e, chi2);
}

template <bool JAC, int D, int TYPE>
__global__ void k_prior_vec(int n, const int* __restrict__ var, const double* __restrict__ mean,
                            const double* __restrict__ info, Vals vals, const int* __restrict__ off,
                            SysView sys, double* g_r, double* chi2) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (f < n) {
    double r[D], wr[D];
    const double* x = vals.v[TYPE] + (int64_t)var[f] * D;
#pragma unroll
    for (int i = 0; i < D; ++i) r[i] = x[i] - mean[D * f + i];
    const double* Om = info + D * D * f;
    matvec<D, D>(Om, r, wr);
#pragma unroll
    for (int i = 0; i < D; ++i) e += r[i] * wr[i];
    if (JAC) {
      int o = off[var[f]];
      sys_add_block(sys, o, D, o, D, Om, D);
#pragma unroll
      for (int i = 0; i < D; ++i) atomicAdd(&g_r[o + i], wr[i]);
    }
  }
  chi2_accumulate(e, chi2);
}

// ------------------------------------------------------------------ K1 between
template <bool JAC>
__global__ void k_between(int n, const int* __restrict__ vi, const int* __restrict__ vj,
                          const double* __restrict__ meas, const double* __restrict__ info, Vals vals,
                          const int* __restrict__ off, SysView sys, double* g_r, double* chi2, int chart) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (f < n) {
    double X1[12], X2[12], Z[12], r[6], J1[36], wr[6];
    load_pose(vals.v[T_POSE], vi[f], X1);
    load_pose(vals.v[T_POSE], vj[f], X2);
    load_pose(meas, f, Z);
    between_eval<JAC>(X1, X2, Z, r, J1, chart);
    const double* Om = info + 36 * (int64_t)f;
    double O[36];
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      double2 v = __ldg(reinterpret_cast<const double2*>(Om) + i);
      O[2 * i] = v.x; O[2 * i + 1] = v.y;
    }
    matvec<6, 6>(O, r, wr);
#pragma unroll
    for (int i = 0; i < 6; ++i) e += r[i] * wr[i];
    if (JAC) {
      int o1 = off[vi[f]], o2 = off[vj[f]];
      double M[36], H11[36], g1[6];
      // M = Omega J1 ; H11 = J1^T M ; H21 = M ; H22 = Omega
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          double s = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) s += O[6 * i + k] * J1[6 * k + j];
          M[6 * i + j] = s;
        }
#pragma unroll
      for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          double s = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) s += J1[6 * k + i] * M[6 * k + j];
          H11[6 * i + j] = s;
        }
        double s = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) s += J1[6 * k + i] * wr[k];
        g1[i] = s;
      }
      sys_add_block(sys, o1, 6, o1, 6, H11, 6);
      sys_add_block(sys, o2, 6, o1, 6, M, 6);
      sys_add_block(sys, o2, 6, o2, 6, O, 6);
#pragma unroll
      for (int i = 0; i < 6; ++i) { atomicAdd(&g_r[o1 + i], g1[i]); atomicAdd(&g_r[o2 + i], wr[i]); }
    }
  }
  chi2_accumulate(e, chi2);
}

// ------------------------------------------------------------------ g2o EdgeSE3 (config 1: CGraphG2O, g2o/g2o_graph.cpp:96-134)
// chi2 = sum e^T Omega e (no 1/2: g2o's chi2(), g2o_graph.cpp:254-258).  A fixed end (VertexSE3::setFixed, :90) takes no
// Hessian or gradient contribution: its delta stays 0 (k_fix_identity keeps its diagonal block non-singular).
template <bool JAC>
__global__ void k_g2o_edge(int n, const int* __restrict__ vi, const int* __restrict__ vj, const double* __restrict__ meas,
                           const double* __restrict__ info, Vals vals, const int* __restrict__ off, const char* __restrict__ fixed,
                           SysView sys, double* g_r, double* chi2) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (f < n) {
    double X1[12], X2[12], Z[12], r[6], J1[36], J2[36], wr[6], O[36];
    load_pose(vals.v[T_POSE], vi[f], X1);
    load_pose(vals.v[T_POSE], vj[f], X2);
    load_pose(meas, f, Z);
    g2o_edge_eval<JAC>(X1, X2, Z, r, J1, J2);
#pragma unroll
    for (int i = 0; i < 36; ++i) O[i] = info[36 * (int64_t)f + i];
    matvec<6, 6>(O, r, wr);
#pragma unroll
    for (int i = 0; i < 6; ++i) e += r[i] * wr[i];
    if (JAC) {
      const bool free1 = !fixed || !fixed[vi[f]], free2 = !fixed || !fixed[vj[f]];
      const int o1 = off[vi[f]], o2 = off[vj[f]];
      double M1[36], M2[36], H[36], g[6];
      auto mul = [](const double* A, const double* B, double* C, bool at) {     // C = A B or A^T B (6x6)
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += (at ? A[6 * k + i] : A[6 * i + k]) * B[6 * k + j];
            C[6 * i + j] = s;
          }
      };
      mul(O, J1, M1, false);
      mul(O, J2, M2, false);
      if (free1) {
        mul(J1, M1, H, true);
        sys_add_block(sys, o1, 6, o1, 6, H, 6);
        for (int i = 0; i < 6; ++i) { double s = 0; for (int k = 0; k < 6; ++k) s += J1[6 * k + i] * wr[k]; g[i] = s; }
        for (int i = 0; i < 6; ++i) atomicAdd(&g_r[o1 + i], g[i]);
      }
      if (free2) {
        mul(J2, M2, H, true);
        sys_add_block(sys, o2, 6, o2, 6, H, 6);
        for (int i = 0; i < 6; ++i) { double s = 0; for (int k = 0; k < 6; ++k) s += J2[6 * k + i] * wr[k]; g[i] = s; }
        for (int i = 0; i < 6; ++i) atomicAdd(&g_r[o2 + i], g[i]);
      }
      if (free1 && free2) {
        mul(J2, M1, H, true);                      // rows of vertex 2, columns of vertex 1
        sys_add_block(sys, o2, 6, o1, 6, H, 6);
      }
    }
  }
  chi2_accumulate(e, chi2);
}
// identity on the diagonal block of every fixed pose (the block has no other contribution)
__global__ void k_fix_identity(int n, const int* __restrict__ list, const int* __restrict__ off, SysView sys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 6 * n) return;
  const int C = off[list[i / 6]] + i % 6;
  int ld;
  sys.L[sys_find(sys, C, C, &ld)] += 1.0;
}
// max diagonal entry over the free columns (g2o: OptimizationAlgorithmLevenberg::computeLambdaInit); doubles >= 0 order
// like their bit patterns
__global__ void k_max_diag(SysView sys, const char* __restrict__ fixed_col, unsigned long long* out) {
  const int C = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (C < sys.n_r && !(fixed_col && fixed_col[C])) {
    const int sn = sys.col2sn[C];
    const int c = C - sys.sn_col0[sn];
    v = fabs(sys.L[sys.sn_valptr[sn] + c + (int64_t)c * sys.sn_nrows[sn]]);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, d));
  if ((threadIdx.x & 31) == 0 && v > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(v));
}

// ------------------------------------------------------------------ K5 plane
template <bool JAC>
__global__ void k_plane(int n, const int* __restrict__ vp, const int* __restrict__ vl,
                        const double* __restrict__ meas, const double* __restrict__ info, Vals vals,
                        const int* __restrict__ off_pose, const int* __restrict__ off_plane, SysView sys,
                        double* g_r, double* chi2) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (f < n) {
    double X[12], pl[4], z[4], r[3], Hr[18], Hp[9], wr[3];
    load_pose(vals.v[T_POSE], vp[f], X);
#pragma unroll
    for (int i = 0; i < 4; ++i) { pl[i] = vals.v[T_PLANE][4 * (int64_t)vl[f] + i]; z[i] = meas[4 * (int64_t)f + i]; }
    plane_eval<JAC>(X, pl, z, r, Hr, Hp);
    const double* Om = info + 9 * (int64_t)f;
    matvec<3, 3>(Om, r, wr);
    e = r[0] * wr[0] + r[1] * wr[1] + r[2] * wr[2];
    if (JAC) {
      int op = off_pose[vp[f]], ol = off_plane[vl[f]];
      double Mr[18], Mp[9];   // Omega Hr (3x6), Omega Hp (3x3)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) Mr[6 * i + j] = Om[3 * i] * Hr[j] + Om[3 * i + 1] * Hr[6 + j] + Om[3 * i + 2] * Hr[12 + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) Mp[3 * i + j] = Om[3 * i] * Hp[j] + Om[3 * i + 1] * Hp[3 + j] + Om[3 * i + 2] * Hp[6 + j];
      }
      double Hpp[36], Hlp[18], Hll[9];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) Hpp[6 * i + j] = Hr[i] * Mr[j] + Hr[6 + i] * Mr[6 + j] + Hr[12 + i] * Mr[12 + j];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) Hlp[6 * i + j] = Hp[i] * Mr[j] + Hp[3 + i] * Mr[6 + j] + Hp[6 + i] * Mr[12 + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) Hll[3 * i + j] = Hp[i] * Mp[j] + Hp[3 + i] * Mp[3 + j] + Hp[6 + i] * Mp[6 + j];
      }
      sys_add_block(sys, op, 6, op, 6, Hpp, 6);
      sys_add_block(sys, ol, 3, op, 6, Hlp, 6);
      sys_add_block(sys, ol, 3, ol, 3, Hll, 3);
#pragma unroll
      for (int i = 0; i < 6; ++i) atomicAdd(&g_r[op + i], Hr[i] * wr[0] + Hr[6 + i] * wr[1] + Hr[12 + i] * wr[2]);
#pragma unroll
      for (int i = 0; i < 3; ++i) atomicAdd(&g_r[ol + i], Hp[i] * wr[0] + Hp[3 + i] * wr[1] + Hp[6 + i] * wr[2]);
    }
  }
  chi2_accumulate(e, chi2);
}

// ------------------------------------------------------------------ K4 imu: one warp per factor
#define IMU_WPB 4
template <bool JAC>
__global__ void __launch_bounds__(32 * IMU_WPB) k_imu(int n, const int* __restrict__ var, const ImuRec* __restrict__ rec,
                                                      Vals vals, const int* __restrict__ off_pose,
                                                      const int* __restrict__ off_vel, const int* __restrict__ off_bias,
                                                      SysView sys, double* g_r, double* chi2) {
  __shared__ double sJ[IMU_WPB][450];
  __shared__ double sM[IMU_WPB][450];
  __shared__ double sr[IMU_WPB][16];
  __shared__ double swr[IMU_WPB][16];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int f = blockIdx.x * IMU_WPB + w;
  double e = 0.0;
  bool act = f < n;
  const ImuRec* F = rec + (act ? f : 0);
  const int* v = var + 6 * (int64_t)(act ? f : 0);
  if (act && lane == 0) {
    double Xi[12], Xj[12];
    load_pose(vals.v[T_POSE], v[0], Xi);
    load_pose(vals.v[T_POSE], v[2], Xj);
    const double* vi = vals.v[T_VEC3] + 3 * (int64_t)v[1];
    const double* vj = vals.v[T_VEC3] + 3 * (int64_t)v[3];
    const double* bi = vals.v[T_BIAS] + 6 * (int64_t)v[4];
    const double* bj = vals.v[T_BIAS] + 6 * (int64_t)v[5];
    imu_eval<JAC>(Xi, vi, Xj, vj, bi, bj, F, sr[w], sJ[w]);
  }
  __syncwarp();
  if (act) {
    // wr = Omega r (lanes 0..14)
    if (lane < 15) {
      double s = 0;
      for (int k = 0; k < 15; ++k) s += F->info[15 * lane + k] * sr[w][k];
      swr[w][lane] = s;
      e = s * sr[w][lane];
    }
  }
  if (JAC) {
    __syncwarp();
    if (act && lane < 30) {
      // column `lane` of M = Omega J
      double col[15];
      for (int k = 0; k < 15; ++k) col[k] = sJ[w][30 * k + lane];
      for (int i = 0; i < 15; ++i) {
        double s = 0;
        for (int k = 0; k < 15; ++k) s += F->info[15 * i + k] * col[k];
        sM[w][30 * i + lane] = s;
      }
    }
    __syncwarp();
    if (act && lane < 30) {
      // gradient entry and Hessian column `lane`: H[:, lane] = J^T M[:, lane]
      const int seg_start[6] = {0, 6, 9, 15, 18, 24};
      const int seg_dim[6] = {6, 3, 6, 3, 6, 6};
      int offs[6] = {off_pose[v[0]], off_vel[v[1]], off_pose[v[2]], off_vel[v[3]], off_bias[v[4]], off_bias[v[5]]};
      int myseg = 0;
      for (int s = 0; s < 6; ++s) if (lane >= seg_start[s]) myseg = s;
      int C = offs[myseg] + (lane - seg_start[myseg]);   // global reduced column of this lane
      double gsum = 0;
      for (int k = 0; k < 15; ++k) gsum += sJ[w][30 * k + lane] * swr[w][k];
      atomicAdd(&g_r[C], gsum);
      for (int s = 0; s < 6; ++s) {
        for (int i = 0; i < seg_dim[s]; ++i) {
          int row_local = seg_start[s] + i;
          int R = offs[s] + i;
          if (R < C) continue;              // lower triangle only (R >= C)
          double h = 0;
          for (int k = 0; k < 15; ++k) h += sJ[w][30 * k + row_local] * sM[w][30 * k + lane];
          int ld;
          int64_t idx = sys_find(sys, R, C, &ld);
          atomicAdd(&sys.L[idx], h);
        }
      }
    }
  }
  chi2_accumulate(e, chi2);
}

// ------------------------------------------------------------------ K6 projection + Schur
// per landmark: initialise V, gl with the point prior (PriorFactor<Point3>, gtsam_graph.cpp:394)
template <bool JAC>
__global__ void k_lm_prior(int64_t L, const double* __restrict__ pts, const double* __restrict__ mean,
                           const double* __restrict__ w, double* V, double* gl, double* chi2) {
  int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double e = 0.0;
  if (l < L) {
    double wl = w[l];
    double r0 = 0, r1 = 0, r2 = 0;
    if (wl > 0) {
      r0 = pts[3 * l] - mean[3 * l]; r1 = pts[3 * l + 1] - mean[3 * l + 1]; r2 = pts[3 * l + 2] - mean[3 * l + 2];
      e = wl * (r0 * r0 + r1 * r1 + r2 * r2);
    }
    if (JAC) {
      V[6 * l + 0] = wl; V[6 * l + 1] = 0; V[6 * l + 2] = 0; V[6 * l + 3] = wl; V[6 * l + 4] = 0; V[6 * l + 5] = wl;
      gl[3 * l] = wl * r0; gl[3 * l + 1] = wl * r1; gl[3 * l + 2] = wl * r2;
    }
  }
  chi2_accumulate(e, chi2);
}

// segmented (by key) inclusive-suffix reduction inside a warp: head lane of each run gets the run sum
__device__ __forceinline__ double seg_sum(double v, int key, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    double o = __shfl_down_sync(0xffffffffu, v, d);
    int k2 = __shfl_down_sync(0xffffffffu, key, d);
    if (lane + d < 32 && k2 == key) v += o;
  }
  return v;
}

// In-block ordered merge of per-landmark sums (deterministic: no atomics).  Observations are sorted by landmark and a block's
// range starts and ends on landmark boundaries (fg_finalize: oblk_ptr), so every landmark is summed inside ONE block: the
// warp-segmented reduction leaves a partial in the head lane of every (warp, landmark) run; the heads are numbered in
// observation order, parked in shared memory, and the first head of each landmark adds its pieces in that order.
template <int NV, class Write>
__device__ __forceinline__ void ordered_merge(const double (&v)[NV], int l, bool act, double* hbuf, int* hl, int* wcnt, Write write) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int prev = __shfl_up_sync(FULL, l, 1);
  const bool head = act && (lane == 0 || prev != l);
  const unsigned hm = __ballot_sync(FULL, head);
  if (lane == 0) wcnt[w] = __popc(hm);
  double sums[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) sums[i] = seg_sum(v[i], l, lane);
  __syncthreads();
  int base = 0, nh = 0;
  for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { if (k < w) base += wcnt[k]; nh += wcnt[k]; }
  if (head) {
    const int ord = base + __popc(hm & ((1u << lane) - 1u));
    hl[ord] = l;
#pragma unroll
    for (int i = 0; i < NV; ++i) hbuf[ord * (NV + 1) + i] = sums[i];
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < nh && (t == 0 || hl[t - 1] != hl[t])) {
    const int lm = hl[t];
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = hbuf[t * (NV + 1) + i];
    for (int k = t + 1; k < nh && hl[k] == lm; ++k)
#pragma unroll
      for (int i = 0; i < NV; ++i) acc[i] += hbuf[k * (NV + 1) + i];
    write(lm, acc);                                                     // single owner: plain read-modify-writes
  }
  __syncthreads();
}

// The same ordered merge for runs that are CONTIGUOUS across the block (observations sorted by landmark), with ONE barrier: a
// run that crosses a warp boundary continues as the first run of the following warps, so every warp parks the sum of its
// first run (and the key of its last lane); after the barrier the head of a run that does not continue one from the warp
// before adds the first-run sums of the following warps while their key matches -- observation order, no atomics.
// The head with the final sum calls write(key, sums).  `more` = the caller loops and calls again (block-uniform).
template <int NV>
struct RunMergeSmem { double first[8][NV]; int first_key[8]; int last_key[8]; };
template <int NV, class Write>
__device__ __forceinline__ void run_merge(const double (&v)[NV], int key, bool act, RunMergeSmem<NV>& sm, bool more, Write write) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int prev = __shfl_up_sync(FULL, key, 1);
  const bool head = lane == 0 || prev != key;
  double sums[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) sums[i] = seg_sum(v[i], key, lane);
  if (lane == 0) {
    sm.first_key[w] = act ? key : -1;
#pragma unroll
    for (int i = 0; i < NV; ++i) sm.first[w][i] = sums[i];
  }
  if (lane == 31) sm.last_key[w] = act ? key : -1;
  const unsigned hm = __ballot_sync(FULL, head);
  const bool last_run = (hm >> lane) <= 1u;               // no later head in this warp: the run reaches lane 31
  __syncthreads();
  if (head && act && !(lane == 0 && w > 0 && sm.last_key[w - 1] == key)) {
    if (last_run)
      for (int k = w + 1; k < nw && sm.first_key[k] == key; ++k)
#pragma unroll
        for (int i = 0; i < NV; ++i) sums[i] += sm.first[k][i];
    write(key, sums);
  }
  if (more) __syncthreads();
}

// Jacobian pass of the between factors without colours: one thread per factor END, ends sorted by (pose, factor), block ranges
// cut on pose boundaries (fg_finalize: bt_end, bt_eblk).  Every thread evaluates its factor; the j-end thread owns the factor's
// off-diagonal block and its chi2 (host guarantee: no two factors on one pose pair -- otherwise the coloured k_between runs);
// the diagonal block and the gradient of a pose are summed over its ends by the in-block ordered merge.  One launch instead of
// one per colour (a VIO graph with 5 look-back edges has 11).
#define BTE_T 128
__global__ void __launch_bounds__(BTE_T) k_between_ends(const int* __restrict__ eblk_ptr, const int* __restrict__ ends, const int* __restrict__ vi,
                                                        const int* __restrict__ vj, const double* __restrict__ meas, const double* __restrict__ info,
                                                        Vals vals, const int* __restrict__ off, SysView sys, double* g_r, double* chi2, int chart) {
  __shared__ double hbuf[BTE_T * 28];
  __shared__ int hl[BTE_T];
  __shared__ int wcnt[BTE_T / 32];
  const int s0 = eblk_ptr[blockIdx.x], s1 = eblk_ptr[blockIdx.x + 1];
  double e = 0.0;
  for (int base = s0; base < s1; base += BTE_T) {
    const int q = base + threadIdx.x;
    const bool act = q < s1;
    int pose = -1;
    double v27[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) v27[i] = 0.0;
    if (act) {
      const int rec = ends[q], f = rec >> 1, end = rec & 1;
      double X1[12], X2[12], Z[12], r[6], J1[36], wr[6], O[36];
      const int p1 = vi[f], p2 = vj[f];
      pose = end ? p2 : p1;
      load_pose(vals.v[T_POSE], p1, X1);
      load_pose(vals.v[T_POSE], p2, X2);
      load_pose(meas, f, Z);
      between_eval<true>(X1, X2, Z, r, J1, chart);
      const double* Om = info + 36 * (int64_t)f;
#pragma unroll
      for (int i = 0; i < 18; ++i) {
        double2 v = __ldg(reinterpret_cast<const double2*>(Om) + i);
        O[2 * i] = v.x; O[2 * i + 1] = v.y;
      }
      matvec<6, 6>(O, r, wr);
      if (end) {
        // H22 = Omega, g2 = Omega r; the off-diagonal H21 = Omega J1 and the factor's chi2 belong to this end
        double M[36];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          e += r[i] * wr[i];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            double sacc = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) sacc += O[6 * i + k] * J1[6 * k + j];
            M[6 * i + j] = sacc;
          }
        }
        sys_add_block(sys, off[p2], 6, off[p1], 6, M, 6);
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j <= i; ++j) v27[t++] = O[6 * i + j];
#pragma unroll
        for (int i = 0; i < 6; ++i) v27[21 + i] = wr[i];
      } else {
        // H11 = J1^T Omega J1, g1 = J1^T Omega r
        double M[36];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            double sacc = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) sacc += O[6 * i + k] * J1[6 * k + j];
            M[6 * i + j] = sacc;
          }
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
          for (int j = 0; j <= i; ++j) {
            double sacc = 0;
#pragma unroll
            for (int k = 0; k < 6; ++k) sacc += J1[6 * k + i] * M[6 * k + j];
            v27[t++] = sacc;
          }
          double sacc = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) sacc += J1[6 * k + i] * wr[k];
          v27[21 + i] = sacc;
        }
      }
    }
    ordered_merge<27>(v27, pose, act, hbuf, hl, wcnt, [&](int p, const double (&acc)[27]) {
      const int o = off[p];
      int ld;
      const int64_t b0 = sys_find(sys, o, o, &ld);
      int t = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) sys.L[b0 + i + (int64_t)j * ld] += acc[t++];
#pragma unroll
      for (int i = 0; i < 6; ++i) g_r[o + i] += acc[21 + i];
    });
  }
  chi2_accumulate(e, chi2);
}

// one thread per observation (observations sorted by landmark): residual, Jp, Jl; W = w Jp^T Jl stored AoS;
// V_l, g_l by the one-barrier run merge above.  Block b owns observations [oblk_ptr[b], oblk_ptr[b + 1]) (<= 256 unless a
// single landmark has more).  The W records (144 B each) of a warp's 32 consecutive observations are transposed through a
// warp-private piece of shared memory (no block barrier) so that the global store is coalesced: a thread writing its own
// record would cost 32 cache-line wavefronts per store instruction.
typedef DevGraph::ProjCal ProjCal;
// Calibration argument of the projection kernels.  One (Cal3DS2, body_P_sensor) pair in the graph -- what the reference builds:
// the pair travels by value in the constant bank.  Several pairs: a table in global memory and an index per observation.
template <bool MULTI> struct CalArg;
template <> struct CalArg<false> {
  ProjCal cal;
  __device__ __forceinline__ const ProjCal& at(int64_t) const { return cal; }
};
template <> struct CalArg<true> {
  const ProjCal* cals; const unsigned char* idx;
  __device__ __forceinline__ const ProjCal& at(int64_t o) const { return cals[idx[o]]; }
};
template <bool JAC, bool MULTI>
__global__ void __launch_bounds__(256, JAC ? 3 : 5) k_proj_obs(const int64_t* __restrict__ oblk_ptr, const int* __restrict__ obs_pose, const int* __restrict__ obs_point,
                                                  const double* __restrict__ obs_uv, const double* __restrict__ obs_w,
                                                  Vals vals, const __grid_constant__ CalArg<MULTI> ca,
                                                  double* __restrict__ W, double* V, double* gl, double* part) {
  __shared__ __align__(16) double wbuf[JAC ? 8 : 1][JAC ? 32 * 18 : 2];   // 144-byte records back to back: 16-byte accesses of 8 lanes hit bank groups (lane + piece) mod 8
  __shared__ RunMergeSmem<9> ms;
  const int64_t s0 = oblk_ptr[blockIdx.x], s1 = oblk_ptr[blockIdx.x + 1];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  double e = 0.0;
  for (int64_t base = s0; base < s1; base += 256) {
    const int64_t o = base + threadIdx.x;
    const bool act = o < s1;
    const int l = act ? obs_point[o] : -1;
    double r[2], Jp[12], Jl[6], w = 0.0;
    if (act) {
      double X[12], p[3];
      load_pose(vals.v[T_POSE], obs_pose[o], X);
      p[0] = vals.v[T_POINT][3 * (int64_t)l]; p[1] = vals.v[T_POINT][3 * (int64_t)l + 1]; p[2] = vals.v[T_POINT][3 * (int64_t)l + 2];
      double2 uvv = __ldg(reinterpret_cast<const double2*>(obs_uv) + o);
      double uv[2] = {uvv.x, uvv.y};
      w = obs_w[o];
      const ProjCal& pc = ca.at(o);
      projection_eval<JAC>(X, p, uv, pc.K, pc.S, r, Jp, Jl);
      e += w * (r[0] * r[0] + r[1] * r[1]);
    }
    if (JAC) {
      double2* wb = reinterpret_cast<double2*>(wbuf[wp]);
      if (act) {
        double wr[18];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            wr[3 * i + c] = w * (Jp[i] * Jl[c] + Jp[6 + i] * Jl[3 + c]);
#pragma unroll
        for (int q = 0; q < 9; ++q) wb[lane * 9 + q] = make_double2(wr[2 * q], wr[2 * q + 1]);
      }
      __syncwarp();
      {
        const int64_t o0 = base + 32 * wp;                                  // first observation of this warp
        const int n2 = (int)max((int64_t)0, min((int64_t)32, s1 - o0)) * 9;    // 16-byte pieces to store
        double2* dst = reinterpret_cast<double2*>(W + o0 * 18);
        for (int i = lane; i < n2; i += 32) dst[i] = wb[i];
      }
      __syncwarp();
      // V (upper: 00 01 02 11 12 22) and gl
      double c9[9];
      if (act) {
        c9[0] = w * (Jl[0] * Jl[0] + Jl[3] * Jl[3]);
        c9[1] = w * (Jl[0] * Jl[1] + Jl[3] * Jl[4]);
        c9[2] = w * (Jl[0] * Jl[2] + Jl[3] * Jl[5]);
        c9[3] = w * (Jl[1] * Jl[1] + Jl[4] * Jl[4]);
        c9[4] = w * (Jl[1] * Jl[2] + Jl[4] * Jl[5]);
        c9[5] = w * (Jl[2] * Jl[2] + Jl[5] * Jl[5]);
        c9[6] = w * (Jl[0] * r[0] + Jl[3] * r[1]);
        c9[7] = w * (Jl[1] * r[0] + Jl[4] * r[1]);
        c9[8] = w * (Jl[2] * r[0] + Jl[5] * r[1]);
      } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) c9[i] = 0.0;
      }
      run_merge<9>(c9, l, act, ms, base + 256 < s1, [&](int lm, const double (&acc)[9]) {
#pragma unroll
        for (int i = 0; i < 6; ++i) V[6 * (int64_t)lm + i] += acc[i];       // single owner: plain read-modify-writes
#pragma unroll
        for (int i = 0; i < 3; ++i) gl[3 * (int64_t)lm + i] += acc[6 + i];
      });
    }
  }
  chi2_accumulate(e, part);
}

// Several projection factors on one (pose, landmark) pair act as ONE coupling block W = sum of their W records: fold the
// secondaries into the primary and clear them (rare; a secondary chain on one pair is walked by the thread of its last link
// only when the list is ordered, which fg_finalize guarantees: the pairs of one primary are consecutive).
__global__ void k_merge_dup(int n, const int* __restrict__ prim, const int* __restrict__ sec, double* __restrict__ W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i > 0 && prim[i - 1] == prim[i]) return;           // the first pair of a primary does the whole group, in list order
  double* wp = W + 18 * (int64_t)prim[i];
  for (int j = i; j < n && prim[j] == prim[i]; ++j) {
    double* ws = W + 18 * (int64_t)sec[j];
#pragma unroll
    for (int e = 0; e < 18; ++e) { wp[e] += ws[e]; ws[e] = 0.0; }
  }
}

// U_pp = sum w Jp^T Jp, g_p = sum w Jp^T r over the pose's observations (recomputed).  PP_W warps per pose, each over a
// contiguous quarter of the pose's list (a pose has ~2000 observations at C5: one warp per pose is 5000 long warps, 2.1 waves of
// the resident 2368, and the last wave runs 11 % full); the quarters are added in order through shared memory.
#define PP_W 4
template <bool MULTI>
__global__ void __launch_bounds__(256, 2) k_proj_pose(int P, const int64_t* __restrict__ pose_obs_ptr, const int64_t* __restrict__ pose_obs,
                                                   const int* __restrict__ obs_point, const double* __restrict__ obs_uv,
                                                   const double* __restrict__ obs_w, Vals vals, const __grid_constant__ CalArg<MULTI> ca,
                                                   const int* __restrict__ off_pose, SysView sys, double* g_r) {
  __shared__ double part[256 / 32][28];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pib = warp / PP_W, q = warp % PP_W;                 // pose within the block, quarter of its list
  const int wid = blockIdx.x * (256 / 32 / PP_W) + pib;
  const bool live = wid < P;
  int64_t b = 0, e = 0;
  if (live) { b = pose_obs_ptr[wid]; e = pose_obs_ptr[wid + 1]; }
  const int64_t len = (e - b + PP_W - 1) / PP_W;
  const int64_t qb = b + q * len, qe = min(e, qb + len);
  double acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.0;
  if (qb < qe) {
    double X[12];
    load_pose(vals.v[T_POSE], wid, X);
    for (int64_t k = qb + lane; k < qe; k += 32) {
      int64_t o = pose_obs[k];
      int l = obs_point[o];
      double p[3] = {vals.v[T_POINT][3 * (int64_t)l], vals.v[T_POINT][3 * (int64_t)l + 1], vals.v[T_POINT][3 * (int64_t)l + 2]};
      double uv[2] = {obs_uv[2 * o], obs_uv[2 * o + 1]};
      double w = obs_w[o], r[2], Jp[12], Jl[6];
      const ProjCal& pc = ca.at(o);
      projection_eval<true>(X, p, uv, pc.K, pc.S, r, Jp, Jl);
      int t = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j) acc[t++] += w * (Jp[i] * Jp[j] + Jp[6 + i] * Jp[6 + j]);
#pragma unroll
      for (int i = 0; i < 6; ++i) acc[21 + i] += w * (Jp[i] * r[0] + Jp[6 + i] * r[1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 27; ++i) {
    const double v = warp_sum(acc[i]);
    if (lane == 0) part[warp][i] = v;
  }
  __syncthreads();
  if (q == 0 && live && b < e && lane < 27) {
    double v = part[warp][lane];
#pragma unroll
    for (int k = 1; k < PP_W; ++k) v += part[warp + k][lane];
    const int o = off_pose[wid];
    if (lane < 21) {
      int t = lane, i = 0;
      while (t > i) { t -= i + 1; ++i; }                       // lower-triangle index -> (i, j = t)
      int ld;
      const int64_t base = sys_find(sys, o, o, &ld);
      sys.L[base + i + (int64_t)t * ld] += v;                   // the only writer of this block in this kernel
    } else {
      g_r[o + lane - 21] += v;
    }
  }
}

// damp diagonal and write the rhs row: L[C,C] += lambda ; L[rhs, C] = -g_r[C]
__global__ void k_damp_rhs(SysView sys, const double* __restrict__ g_r, double lambda, int add_rhs) {
  int C = blockIdx.x * blockDim.x + threadIdx.x;
  if (C >= sys.n_r) return;
  int sn = sys.col2sn[C];
  int c = C - sys.sn_col0[sn], nr = sys.sn_nrows[sn];
  double* col = sys.L + sys.sn_valptr[sn] + (int64_t)c * nr;
  col[c] += lambda;
  if (add_rhs) col[nr - 1] -= g_r[C];
}

// ------------------------------------------------------------------ K9/K10 back-substitution and retraction
// t_l = sum_o W_o^T delta_p(o) (thread per observation, block ranges on landmark boundaries, one-barrier run merge); a warp's
// 32 consecutive W records arrive coalesced and are transposed through its private piece of shared memory
__global__ void __launch_bounds__(256) k_lm_backsub_obs(const int64_t* __restrict__ oblk_ptr, const int* __restrict__ obs_pose, const int* __restrict__ obs_point,
                                                        const double* __restrict__ W, const double* __restrict__ delta,
                                                        const int* __restrict__ off_pose, double* tl) {
  __shared__ __align__(16) double wbuf[8][32 * 18];
  __shared__ RunMergeSmem<3> ms;
  const int64_t s0 = oblk_ptr[blockIdx.x], s1 = oblk_ptr[blockIdx.x + 1];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int64_t base = s0; base < s1; base += 256) {
    const int64_t o = base + threadIdx.x;
    const bool act = o < s1;
    const int l = act ? obs_point[o] : -1;
    double t3[3] = {0, 0, 0};
    double2* wb = reinterpret_cast<double2*>(wbuf[wp]);
    {
      // coalesced load of the warp's consecutive W records, issued before the dependent gather below
      const int64_t o0 = base + 32 * wp;
      const int n2 = (int)max((int64_t)0, min((int64_t)32, s1 - o0)) * 9;
      const double2* src = reinterpret_cast<const double2*>(W + o0 * 18);
      double2 v[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) { const int i = lane + 32 * q; v[q] = i < n2 ? __ldg(src + i) : make_double2(0.0, 0.0); }
      // obs_pose -> off_pose -> delta
      double dl[6];
      const double* d = delta + (act ? off_pose[obs_pose[o]] : 0);
#pragma unroll
      for (int i = 0; i < 6; ++i) dl[i] = act ? d[i] : 0.0;
#pragma unroll
      for (int q = 0; q < 9; ++q) {
        const int i = lane + 32 * q;
        if (i < n2) wb[i] = v[q];
      }
      __syncwarp();
      if (act) {
        double wr[18];
#pragma unroll
        for (int q = 0; q < 9; ++q) { const double2 x = wb[lane * 9 + q]; wr[2 * q] = x.x; wr[2 * q + 1] = x.y; }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
          for (int c = 0; c < 3; ++c) t3[c] += wr[3 * i + c] * dl[i];
        }
      }
      __syncwarp();
    }
    run_merge<3>(t3, l, act, ms, base + 256 < s1, [&](int lm, const double (&acc)[3]) {
#pragma unroll
      for (int i = 0; i < 3; ++i) tl[3 * (int64_t)lm + i] += acc[i];
    });
  }
}

// delta_l = -Vinv (g_l + t_l) ; p_new = p + delta_l ; accumulates g^T delta and |delta|^2
__global__ void k_lm_update(int64_t L, const double* __restrict__ pts, const double* __restrict__ Vinv,
                            const double* __restrict__ gl, const double* __restrict__ tl, double* pts_new, double* part_gd, double* part_dd) {
  int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double gd = 0.0, dd = 0.0;
  if (l < L) {
    double s[3] = {gl[3 * l] + tl[3 * l], gl[3 * l + 1] + tl[3 * l + 1], gl[3 * l + 2] + tl[3 * l + 2]};
    const double* vi = Vinv + 6 * l;
    double d0 = -(vi[0] * s[0] + vi[1] * s[1] + vi[2] * s[2]);
    double d1 = -(vi[1] * s[0] + vi[3] * s[1] + vi[4] * s[2]);
    double d2 = -(vi[2] * s[0] + vi[4] * s[1] + vi[5] * s[2]);
    pts_new[3 * l] = pts[3 * l] + d0; pts_new[3 * l + 1] = pts[3 * l + 1] + d1; pts_new[3 * l + 2] = pts[3 * l + 2] + d2;
    gd = gl[3 * l] * d0 + gl[3 * l + 1] * d1 + gl[3 * l + 2] * d2;
    dd = d0 * d0 + d1 * d1 + d2 * d2;
  }
  gd = block_sum(gd); dd = block_sum(dd);
  if (threadIdx.x == 0) { part_gd[blockIdx.x] = gd; part_dd[blockIdx.x] = dd; }
}

// Values::retract for the reduced variables; accumulates g_r^T delta and |delta|^2 (once per scalar)
template <int TYPE>
__global__ void k_retract_reduced(int64_t n, const double* __restrict__ val, double* val_new, const int* __restrict__ off,
                                  const double* __restrict__ delta, const double* __restrict__ g_r, double* part_gd, double* part_dd, int count_scal, int chart) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double gd = 0.0, dd = 0.0;
  if (i < n) {
    const int D = (TYPE == T_POSE || TYPE == T_BIAS) ? 6 : 3;
    const double* d = delta + off[i];
    const double* g = g_r + off[i];
    double dl[6];
#pragma unroll
    for (int k = 0; k < D; ++k) { dl[k] = d[k]; gd += g[k] * dl[k]; dd += dl[k] * dl[k]; }
    if (TYPE == T_POSE) {
      double X[12], Y[12];
      load_pose(val, (int)i, X);
      if (chart == 1) g2o_oplus(X, X + 9, dl, Y, Y + 9);                 // VertexSE3::oplus
      else pose_chart_retract(X, X + 9, dl, chart, Y, Y + 9);           // Pose3::Retract under the context's chart
#pragma unroll
      for (int k = 0; k < 12; ++k) val_new[12 * i + k] = Y[k];
    } else if (TYPE == T_PLANE) {
      double out[4];
      plane_retract(val + 4 * i, dl, out);
#pragma unroll
      for (int k = 0; k < 4; ++k) val_new[4 * i + k] = out[k];
    } else {
#pragma unroll
      for (int k = 0; k < D; ++k) val_new[D * i + k] = val[D * i + k] + dl[k];
    }
  }
  // g_r is this rank's share of the gradient (always counted); |delta_r|^2 is replicated (rank 0 only)
  gd = block_sum(gd); dd = block_sum(dd);
  if (threadIdx.x == 0) { part_gd[blockIdx.x] = gd; part_dd[blockIdx.x] = count_scal ? dd : 0.0; }
}

// ------------------------------------------------------------------ K3 IMU preintegration
// PreintegratedCombinedMeasurements::integrateMeasurement over the samples of one inter-frame interval
// (CImuBase::predictNext, gtsam/imu_base.cpp:76-85; TangentPreintegration).  The samples of an interval are a
// strictly sequential recurrence; intervals are independent.  ONE WARP per interval: the 15 x 15 covariance update
// P <- F P F^T + Q (the dominant cost, ~7 kflop per sample) and the two 9 x 3 bias Jacobians are spread over the lanes with
// P, F and the intermediate product in shared memory; lane 0 carries the 9-vector and the 3 x 3 pieces of the sample
// (SO(3) exp / right Jacobian and its derivative) in registers.  (Round 1 ran one THREAD per interval with three 225-double
// local arrays: 1.5 ms for 5 k intervals; this kernel: ~0.1 ms.)
#define PI_WPB 4
struct PiWarp {
  double P[225], F[225], T[225];
  double H[54], nH[54];          // [0, 27): preintegrated_H_biasAcc, [27, 54): preintegrated_H_biasOmega (9 x 3 row-major each)
  double wH[9], aH[9], invH[9], R[9], Q[27];      // Q: vv block, theta-theta block, v-theta cross block of the noise term
};
__global__ void __launch_bounds__(32 * PI_WPB, 4) k_preintegrate(int n, const int* __restrict__ offsets, const double* __restrict__ imu6, double dt,
                                                              const ImuParamsDev* __restrict__ par, const double* __restrict__ bias_hat, fg_pim* out) {
  __shared__ PiWarp sm_all[PI_WPB];
  const int lane = threadIdx.x & 31, f = blockIdx.x * PI_WPB + (threadIdx.x >> 5);
  if (f >= n) return;                       // whole warps leave together
  PiWarp& sm = sm_all[threadIdx.x >> 5];
  for (int e = lane; e < 225; e += 32) sm.P[e] = 0.0;
  for (int e = lane; e < 54; e += 32) sm.H[e] = 0.0;
  double pre[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};          // lane 0
  double Ca[9], Cw[9], Cx[9];                              // sample-independent noise sums (lane 0)
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      Ca[3 * i + j] = par->acc_cov[3 * i + j] + par->bint[6 * i + j];
      Cw[3 * i + j] = par->gyro_cov[3 * i + j] + par->bint[6 * (3 + i) + 3 + j];
      Cx[3 * i + j] = par->bint[6 * (3 + i) + j];
    }
  const double* bh = bias_hat + 6 * (int64_t)f;
  const double dt22 = 0.5 * dt * dt;
  const int s0 = offsets[f], s1 = offsets[f + 1];
  __syncwarp();
  for (int s = s0; s < s1; ++s) {
    if (lane == 0) {
      const double* m = imu6 + 6 * (int64_t)s;
      double acc[3] = {m[3] - bh[0], m[4] - bh[1], m[5] - bh[2]};
      double om[3] = {m[0] - bh[3], m[1] - bh[4], m[2] - bh[5]};
      double Jr[9], invH[9], wt[3], D[9], wH[9], R[9], an[3];
      so3_jr(pre, Jr);
      inv3(Jr, invH);
      m3_vec(invH, om, wt);
      d_jr_c(pre, wt, D);
      m3_mul(invH, D, wH);             // w_tangent_H_theta = -invH * D
      so3_exp(pre, R);
      m3_vec(R, acc, an);
      double Sa[9], RS[9], aH[9];      // a_nav_H_theta = R * skew(-acc) * Jr
      double nacc[3] = {-acc[0], -acc[1], -acc[2]};
      skew3(nacc, Sa);
      m3_mul(R, Sa, RS);
      m3_mul(RS, Jr, aH);
#pragma unroll
      for (int i = 0; i < 9; ++i) { sm.wH[i] = wH[i]; sm.aH[i] = aH[i]; sm.invH[i] = invH[i]; sm.R[i] = R[i]; }
      // noise blocks: vv = vH Ca vH^T / dt, theta-theta = thH Cw thH^T / dt, v-theta = vH Cx thH^T  (thH = -dt invH, vH = -dt R)
      double thH[9], vH[9], t1[9], t2[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) { thH[i] = -dt * invH[i]; vH[i] = -dt * R[i]; }
      m3_mul(vH, Ca, t1); m3_mult(t1, vH, t2);
#pragma unroll
      for (int i = 0; i < 9; ++i) sm.Q[i] = t2[i] / dt;
      m3_mul(thH, Cw, t1); m3_mult(t1, thH, t2);
#pragma unroll
      for (int i = 0; i < 9; ++i) sm.Q[9 + i] = t2[i] / dt;
      m3_mul(vH, Cx, t1); m3_mult(t1, thH, t2);
#pragma unroll
      for (int i = 0; i < 9; ++i) sm.Q[18 + i] = t2[i];
      // the 9-vector
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double th = pre[i] + wt[i] * dt, pp = pre[3 + i] + pre[6 + i] * dt + an[i] * dt22, vv = pre[6 + i] + an[i] * dt;
        pre[i] = th; pre[3 + i] = pp; pre[6 + i] = vv;
      }
    }
    // F = [[A, G_b], [0, I]] with A = d(next 9-vector)/d(9-vector), G_b the bias columns (A.5)
    for (int e = lane; e < 225; e += 32) sm.F[e] = (e / 15 == e % 15) ? 1.0 : 0.0;
    __syncwarp();
    if (lane < 9) {
      const int i = lane / 3, j = lane % 3;
      sm.F[15 * i + j] -= dt * sm.wH[3 * i + j];                 // A theta-theta
      sm.F[15 * (3 + i) + j] = dt22 * sm.aH[3 * i + j];         // A p-theta
      sm.F[15 * (6 + i) + j] = dt * sm.aH[3 * i + j];           // A v-theta
      sm.F[15 * i + 12 + j] = -dt * sm.invH[3 * i + j];         // theta_H_biasOmega = -C[0:3]
      sm.F[15 * (6 + i) + 9 + j] = -dt * sm.R[3 * i + j];       // vel_H_biasAcc = -B[6:9]
      if (j == 0) sm.F[15 * (3 + i) + 6 + i] = dt;              // A p-v
    }
    __syncwarp();
    // bias Jacobians: H_ba <- A H_ba - B, H_bg <- A H_bg - C
    for (int e = lane; e < 54; e += 32) {
      const int which = e / 27, idx = e % 27, i = idx / 3, c = idx % 3;
      const double* H = sm.H + 27 * which;
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fma(sm.F[15 * i + k], H[3 * k + c], a);
      if (which == 0) { if (i >= 6) a -= dt * sm.R[3 * (i - 6) + c]; else if (i >= 3) a -= dt22 * sm.R[3 * (i - 3) + c]; }
      else if (i < 3) a -= dt * sm.invH[3 * i + c];
      sm.nH[e] = a;
    }
    // T = F P (rows 9..14 of F are identity rows)
    for (int e = lane; e < 225; e += 32) {
      const int i = e / 15, j = e % 15;
      double a;
      if (i >= 9) a = sm.P[e];
      else { a = 0.0;
#pragma unroll
        for (int k = 0; k < 15; ++k) a = fma(sm.F[15 * i + k], sm.P[15 * k + j], a); }
      sm.T[e] = a;
    }
    __syncwarp();
    for (int e = lane; e < 54; e += 32) sm.H[e] = sm.nH[e];
    // P = T F^T + Q
    for (int e = lane; e < 225; e += 32) {
      const int i = e / 15, j = e % 15;
      double a;
      if (j >= 9) a = sm.T[e];
      else { a = 0.0;
#pragma unroll
        for (int k = 0; k < 15; ++k) a = fma(sm.T[15 * i + k], sm.F[15 * j + k], a); }
      const int bi = i / 3, bj = j / 3, ii = i % 3, jj = j % 3;
      if (bi == 2 && bj == 2) a += sm.Q[3 * ii + jj];                                   // vv
      else if (bi == 0 && bj == 0) a += sm.Q[9 + 3 * ii + jj];                          // theta-theta
      else if (bi == 2 && bj == 0) a += sm.Q[18 + 3 * ii + jj];                         // v-theta
      else if (bi == 0 && bj == 2) a += sm.Q[18 + 3 * jj + ii];                         // theta-v (transpose)
      else if (bi == 1 && bj == 1) a += dt * par->int_cov[3 * ii + jj];
      else if (bi == 3 && bj == 3) a += dt * par->bias_acc_cov[3 * ii + jj];
      else if (bi == 4 && bj == 4) a += dt * par->bias_gyro_cov[3 * ii + jj];
      sm.P[e] = a;
    }
    __syncwarp();
  }
  fg_pim* o = out + f;
  if (lane == 0) {
    o->dt = (s1 - s0) * dt;
    for (int i = 0; i < 9; ++i) o->preint[i] = pre[i];
    for (int i = 0; i < 6; ++i) o->bias_hat[i] = bh[i];
    for (int i = 0; i < 3; ++i) o->gravity[i] = par->gravity[i];
  }
  for (int e = lane; e < 27; e += 32) { o->H_ba[e] = sm.H[e]; o->H_bg[e] = sm.H[27 + e]; }
  for (int e = lane; e < 225; e += 32) o->cov[e] = sm.P[e];
}

void launch_preintegrate(int n, const int* d_off, const double* d_imu, double dt, const ImuParamsDev* d_par,
                         const double* d_bias, fg_pim* d_out, cudaStream_t st) {
  if (n <= 0) return;
  k_preintegrate<<<cdiv(n, PI_WPB), 32 * PI_WPB, 0, FGS(st)>>>(n, d_off, d_imu, dt, d_par, d_bias, d_out);
}

// ------------------------------------------------------------------ launch wrappers
static SysView make_view(fg_ctx* c, double* Lbuf) {
  SysView s;
  s.L = Lbuf; s.col2sn = c->d.col2sn; s.sn_col0 = c->d.sn_col0; s.sn_ncols = c->d.sn_ncols; s.sn_nrows = c->d.sn_nrows;
  s.sn_rowptr = c->d.sn_rowptr; s.sn_valptr = c->d.sn_valptr; s.rowidx = c->d.rowidx; s.n_r = c->sym.n_r;
  return s;
}

// All factors of the graph at `val` (or the trial state): chi2 -> *target, and with JAC the linearisation.
// Pose-side factors are launched COLOUR BY COLOUR (fg_finalize: no two factors of a colour share a variable), so the
// read-modify-writes of sys_add_block / g_r never meet inside a launch and the assembled system is bitwise repeatable;
// the launches of a kind walk its colour-sorted arrays through [cp[k], cp[k + 1]).
template <bool JAC>
static void run_factors(fg_ctx* c, bool trial, double* target) {
  DevGraph& d = c->d;
  cudaStream_t st = c->stream;
  Vals v;
  for (int t = 0; t < T_COUNT; ++t) v.v[t] = trial ? d.val_new[t] : d.val[t];
  SysView sys = make_view(c, d.U0);
  // Pose-side factors are known to every rank but each is evaluated by ONE of them: rank r takes the r-th slice of
  // every launch; the all-reduce of the packed system (and of chi2) adds the slices up.  One rank: the slice is everything.
  const int rk = c->rank, nrk = c->nranks;
  auto slice = [&](int b, int n, int& b2, int& n2) {
    const int lo = (int)((int64_t)n * rk / nrk), hi = (int)((int64_t)n * (rk + 1) / nrk);
    b2 = b + lo; n2 = hi - lo;
  };
  const bool pose_side = true;
  const int T = 128;
  int np = 0;                                // partial-sum slots handed out so far
  auto slots = [&](int grid) { double* p = d.part + np; np += grid; return p; };
  // Jacobian pass: one launch per colour class (no two factors of a class touch the same variable: plain read-modify-writes,
  // deterministic).  Error pass: nothing is assembled, the whole (colour-sorted, contiguous) kind goes in one launch.
  auto classes = [&](int kind, auto&& launch) {
    const std::vector<int>& cp = c->color_ptr[kind];
    if (cp.size() < 2) return;
    int b2, n2;
    if (!JAC) { slice(cp.front(), cp.back() - cp.front(), b2, n2); if (n2 > 0) launch(b2, n2); return; }
    for (size_t k = 0; k + 1 < cp.size(); ++k) { slice(cp[k], cp[k + 1] - cp[k], b2, n2); if (n2 > 0) launch(b2, n2); }
  };
  if (pose_side) {
    classes(K_PP, [&](int b, int n) {
      k_prior_pose<JAC><<<cdiv(n, T), T, 0, FGS(st)>>>(n, d.pp_var + b, d.pp_mean + 12 * (size_t)b, d.pp_info + 36 * (size_t)b, v, d.off[T_POSE], sys, d.g_r, slots(cdiv(n, T)), d.pose_chart);
    });
    classes(K_PV, [&](int b, int n) {
      k_prior_vec<JAC, 3, T_VEC3><<<cdiv(n, T), T, 0, FGS(st)>>>(n, d.pv_var + b, d.pv_mean + 3 * (size_t)b, d.pv_info + 9 * (size_t)b, v, d.off[T_VEC3], sys, d.g_r, slots(cdiv(n, T)));
    });
    classes(K_PB, [&](int b, int n) {
      k_prior_vec<JAC, 6, T_BIAS><<<cdiv(n, T), T, 0, FGS(st)>>>(n, d.pb_var + b, d.pb_mean + 6 * (size_t)b, d.pb_info + 36 * (size_t)b, v, d.off[T_BIAS], sys, d.g_r, slots(cdiv(n, T)));
    });
    if (JAC && d.n_bt_eblk) {
      int b2, n2;
      slice(0, d.n_bt_eblk, b2, n2);           // whole blocks: a block holds all the ends of its poses
      if (n2 > 0) k_between_ends<<<n2, BTE_T, 0, FGS(st)>>>(d.bt_eblk + b2, d.bt_end, d.bt_i, d.bt_j, d.bt_meas, d.bt_info, v, d.off[T_POSE], sys, d.g_r, slots(n2), d.pose_chart);
    } else classes(K_BT, [&](int b, int n) {
      k_between<JAC><<<cdiv(n, T), T, 0, FGS(st)>>>(n, d.bt_i + b, d.bt_j + b, d.bt_meas + 12 * (size_t)b, d.bt_info + 36 * (size_t)b, v, d.off[T_POSE], sys, d.g_r, slots(cdiv(n, T)), d.pose_chart);
    });
    classes(K_GE, [&](int b, int n) {
      k_g2o_edge<JAC><<<cdiv(n, 64), 64, 0, FGS(st)>>>(n, d.ge_i + b, d.ge_j + b, d.ge_meas + 12 * (size_t)b, d.ge_info + 36 * (size_t)b, v, d.off[T_POSE], d.fixed_pose, sys, d.g_r, slots(cdiv(n, 64)));
    });
    if (JAC && d.n_fixed && rk == 0) k_fix_identity<<<cdiv(6 * d.n_fixed, 64), 64, 0, FGS(st)>>>(d.n_fixed, d.fixed_list, d.off[T_POSE], sys);
    classes(K_IMU, [&](int b, int n) {
      k_imu<JAC><<<cdiv(n, IMU_WPB), 32 * IMU_WPB, 0, FGS(st)>>>(n, d.imu_var + 6 * (size_t)b, d.imu_rec + b, v, d.off[T_POSE], d.off[T_VEC3], d.off[T_BIAS], sys, d.g_r, slots(cdiv(n, IMU_WPB)));
    });
    classes(K_PL, [&](int b, int n) {
      k_plane<JAC><<<cdiv(n, T), T, 0, FGS(st)>>>(n, d.pl_pose + b, d.pl_plane + b, d.pl_meas + 4 * (size_t)b, d.pl_info + 9 * (size_t)b, v, d.off[T_POSE], d.off[T_PLANE], sys, d.g_r, slots(cdiv(n, T)));
    });
  }
  int64_t L = d.n[T_POINT];
  if (L) {
    k_lm_prior<JAC><<<cdiv(L, 256), 256, 0, FGS(st)>>>(L, v.v[T_POINT], d.lm_prior_mean, d.lm_prior_w, d.V, d.gl, slots(cdiv(L, 256)));
    if (d.n_obs) {
      if (JAC && c->kev[0]) cudaEventRecord(c->kev[0], st);
      if (d.n_cal > 1) {
        const CalArg<true> ca{d.cals, d.obs_cal};
        k_proj_obs<JAC, true><<<d.n_oblk, 256, 0, FGS(st)>>>(d.oblk_ptr, d.obs_pose, d.obs_point, d.obs_uv, d.obs_w, v, ca, d.W, d.V, d.gl, slots(d.n_oblk));
      } else {
        const CalArg<false> ca{d.cal};
        k_proj_obs<JAC, false><<<d.n_oblk, 256, 0, FGS(st)>>>(d.oblk_ptr, d.obs_pose, d.obs_point, d.obs_uv, d.obs_w, v, ca, d.W, d.V, d.gl, slots(d.n_oblk));
      }
      if (JAC && c->kev[1]) cudaEventRecord(c->kev[1], st);
      if (JAC && d.n_dup) k_merge_dup<<<cdiv(d.n_dup, 128), 128, 0, FGS(st)>>>(d.n_dup, d.dup_prim, d.dup_sec, d.W);
      if (JAC) {
        int P = (int)d.n[T_POSE];
        if (d.n_cal > 1) {
          const CalArg<true> ca{d.cals, d.obs_cal};
          k_proj_pose<true><<<cdiv((int64_t)P * 32 * PP_W, 256), 256, 0, FGS(st)>>>(P, d.pose_obs_ptr, d.pose_obs, d.obs_point, d.obs_uv, d.obs_w, v, ca, d.off[T_POSE], sys, d.g_r);
        } else {
          const CalArg<false> ca{d.cal};
          k_proj_pose<false><<<cdiv((int64_t)P * 32 * PP_W, 256), 256, 0, FGS(st)>>>(P, d.pose_obs_ptr, d.pose_obs, d.obs_point, d.obs_uv, d.obs_w, v, ca, d.off[T_POSE], sys, d.g_r);
        }
      }
    }
  }
  k_sum_partials<<<1, 256, 0, FGS(st)>>>(d.part, np, target);
}

void launch_linearize(fg_ctx* c) {
  DevGraph& d = c->d;
  cudaMemsetAsync(d.U0, 0, sizeof(double) * (c->sym.nnz + 8), c->stream);
  cudaMemsetAsync(d.g_r, 0, sizeof(double) * c->sym.n_r, c->stream);
  cudaMemsetAsync(d.scal, 0, sizeof(double) * 8, c->stream);
  run_factors<true>(c, false, d.scal + 0);
}

void launch_error_only(fg_ctx* c, bool trial) {
  DevGraph& d = c->d;
  double* target = d.scal + (trial ? 3 : 0);
  run_factors<false>(c, trial, target);
}

void launch_build_and_schur(fg_ctx* c, double lambda) {
  DevGraph& d = c->d;
  cudaStream_t st = c->stream;
  cudaMemcpyAsync(d.L, d.U0, sizeof(double) * (c->sym.nnz + 8), cudaMemcpyDeviceToDevice, st);
  SysView sys = make_view(c, d.L);
  // damping and the pose-side gradient are replicated terms: added by rank 0 only
  // damping is a replicated term (rank 0 only); every rank contributes its own share of the gradient
  k_damp_rhs<<<cdiv(c->sym.n_r, 256), 256, 0, FGS(st)>>>(sys, d.g_r, c->rank == 0 ? lambda : 0.0, 1);
  int64_t L = d.n[T_POINT];
  if (L) launch_schur(c, lambda);
}

void launch_max_diag(fg_ctx* c, double* d_out) {
  cudaMemsetAsync(d_out, 0, sizeof(double), c->stream);
  SysView sys = make_view(c, c->d.U0);
  k_max_diag<<<cdiv(c->sym.n_r, 256), 256, 0, FGS(c->stream)>>>(sys, c->d.fixed_col, reinterpret_cast<unsigned long long*>(d_out));
}

// ------------------------------------------------------------------ incremental update: relinearisation gating
// ISAM2's fluid relinearisation (gtsam_graph.cpp:93-99: relinearizeThreshold 0.1): delta_j = local(theta_j, estimate_j);
// where a component reaches the threshold the linearisation point moves to the estimate.
template <int TYPE>
__global__ void k_inc_gate(int64_t n, double* __restrict__ theta, const double* __restrict__ est, double thr, int* count, int chart) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int S = (TYPE == T_POSE) ? 12 : (TYPE == T_PLANE ? 4 : (TYPE == T_BIAS ? 6 : 3));
  const int D = (TYPE == T_POSE || TYPE == T_BIAS) ? 6 : 3;
  double a[12], b[12], dl[6];
#pragma unroll
  for (int k = 0; k < S; ++k) { a[k] = theta[S * i + k]; b[k] = est[S * i + k]; }
  if (TYPE == T_POSE) {
    double R[9], t[3];
    pose_between(a, a + 9, b, b + 9, R, t);
    pose_chart_local0(R, t, chart, dl);
  } else if (TYPE == T_PLANE) {
    unit3_local(a, b, dl);
    dl[2] = b[3] - a[3];
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) dl[k] = b[k] - a[k];
  }
  double m = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) m = fmax(m, fabs(dl[k]));
  if (m >= thr) {
#pragma unroll
    for (int k = 0; k < S; ++k) theta[S * i + k] = b[k];
    atomicAdd(count, 1);
  }
}
void launch_inc_gate(fg_ctx* c, double thr, int* d_count) {
  DevGraph& d = c->d;
  cudaStream_t st = c->stream;
  const int T = 128;
  if (d.n[T_POSE]) k_inc_gate<T_POSE><<<cdiv(d.n[T_POSE], T), T, 0, FGS(st)>>>(d.n[T_POSE], d.val[T_POSE], d.val_new[T_POSE], thr, d_count, d.pose_chart);
  if (d.n[T_VEC3]) k_inc_gate<T_VEC3><<<cdiv(d.n[T_VEC3], T), T, 0, FGS(st)>>>(d.n[T_VEC3], d.val[T_VEC3], d.val_new[T_VEC3], thr, d_count, d.pose_chart);
  if (d.n[T_BIAS]) k_inc_gate<T_BIAS><<<cdiv(d.n[T_BIAS], T), T, 0, FGS(st)>>>(d.n[T_BIAS], d.val[T_BIAS], d.val_new[T_BIAS], thr, d_count, d.pose_chart);
  if (d.n[T_POINT]) k_inc_gate<T_POINT><<<cdiv(d.n[T_POINT], T), T, 0, FGS(st)>>>(d.n[T_POINT], d.val[T_POINT], d.val_new[T_POINT], thr, d_count, d.pose_chart);
  if (d.n[T_PLANE]) k_inc_gate<T_PLANE><<<cdiv(d.n[T_PLANE], T), T, 0, FGS(st)>>>(d.n[T_PLANE], d.val[T_PLANE], d.val_new[T_PLANE], thr, d_count, d.pose_chart);
}

// ------------------------------------------------------------------ packed exchange (multi-GPU)
__global__ void k_pack(int64_t n, const int64_t* __restrict__ idx, const double* __restrict__ L, double* __restrict__ buf,
                       const double* __restrict__ scal, int with_chi2) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) buf[i] = L[idx[i]];
  if (i == n) buf[n] = with_chi2 ? scal[0] : 0.0;
}
__global__ void k_unpack(int64_t n, const int64_t* __restrict__ idx, double* __restrict__ L, const double* __restrict__ buf,
                         double* __restrict__ scal, int with_chi2) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) L[idx[i]] = buf[i];
  if (i == n && with_chi2) scal[0] = buf[n];
}
void launch_pack(fg_ctx* c, bool with_chi2) {
  DevGraph& d = c->d;
  k_pack<<<cdiv(d.n_pk + 1, 256), 256, 0, FGS(c->stream)>>>(d.n_pk, d.pk_idx, d.L, d.pk_buf, d.scal, with_chi2 ? 1 : 0);
}
void launch_unpack(fg_ctx* c, bool with_chi2) {
  DevGraph& d = c->d;
  k_unpack<<<cdiv(d.n_pk + 1, 256), 256, 0, FGS(c->stream)>>>(d.n_pk, d.pk_idx, d.L, d.pk_buf, d.scal, with_chi2 ? 1 : 0);
}

void launch_retract_error(fg_ctx* c, double lambda) {
  DevGraph& d = c->d;
  cudaStream_t st = c->stream;
  const int T = 128;
  int cnt = (c->rank == 0) ? 1 : 0;
  int np = 0;                                // slots of g^T delta (d.part) and |delta|^2 (d.part2)
  auto at = [&](int grid) { int o = np; np += grid; return o; };
  if (d.n[T_POSE]) { const int g = cdiv(d.n[T_POSE], T), o = at(g); k_retract_reduced<T_POSE><<<g, T, 0, FGS(st)>>>(d.n[T_POSE], d.val[T_POSE], d.val_new[T_POSE], d.off[T_POSE], d.delta, d.g_r, d.part + o, d.part2 + o, cnt, d.pose_chart); }
  if (d.n[T_VEC3]) { const int g = cdiv(d.n[T_VEC3], T), o = at(g); k_retract_reduced<T_VEC3><<<g, T, 0, FGS(st)>>>(d.n[T_VEC3], d.val[T_VEC3], d.val_new[T_VEC3], d.off[T_VEC3], d.delta, d.g_r, d.part + o, d.part2 + o, cnt, d.pose_chart); }
  if (d.n[T_BIAS]) { const int g = cdiv(d.n[T_BIAS], T), o = at(g); k_retract_reduced<T_BIAS><<<g, T, 0, FGS(st)>>>(d.n[T_BIAS], d.val[T_BIAS], d.val_new[T_BIAS], d.off[T_BIAS], d.delta, d.g_r, d.part + o, d.part2 + o, cnt, d.pose_chart); }
  if (d.n[T_PLANE]) { const int g = cdiv(d.n[T_PLANE], T), o = at(g); k_retract_reduced<T_PLANE><<<g, T, 0, FGS(st)>>>(d.n[T_PLANE], d.val[T_PLANE], d.val_new[T_PLANE], d.off[T_PLANE], d.delta, d.g_r, d.part + o, d.part2 + o, cnt, d.pose_chart); }
  int64_t L = d.n[T_POINT];
  if (L) {
    cudaMemsetAsync(d.tl, 0, sizeof(double) * 3 * L, st);
    if (d.n_obs) k_lm_backsub_obs<<<d.n_oblk, 256, 0, FGS(st)>>>(d.oblk_ptr, d.obs_pose, d.obs_point, d.W, d.delta, d.off[T_POSE], d.tl);
    const int g = cdiv(L, 256), o = at(g);
    k_lm_update<<<g, 256, 0, FGS(st)>>>(L, d.val[T_POINT], d.Vinv, d.gl, d.tl, d.val_new[T_POINT], d.part + o, d.part2 + o);
  }
  k_sum_partials<<<1, 256, 0, FGS(st)>>>(d.part, np, d.scal + 1);
  k_sum_partials<<<1, 256, 0, FGS(st)>>>(d.part2, np, d.scal + 2);
  run_factors<false>(c, true, d.scal + 3);
}

}  // namespace fg
