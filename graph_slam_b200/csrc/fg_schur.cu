// fg_schur.cu -- K6 (second half): landmark Schur complement onto the reduced pose system (fp64).
//
//   S_pq -= sum_l W_pl (V_l + lambda I)^-1 W_ql^T ,   rhs_p += sum_l W_pl (V_l + lambda I)^-1 g_l
//
// for the GenericProjectionFactor / PriorFactor<Point3> part of the graph CGraphGT::addToGTSAM builds
// (gtsam/gtsam_graph.cpp:370-448); in GTSAM this is the elimination of every Point3 before the poses.
//
// Output-stationary and deterministic: the host groups, once per graph, the (observation a, observation b)
// pairs of every landmark by the reduced-Hessian block (pose of a, pose of b) they fall in (fg_api.cu,
// fg_finalize).  One warp owns one 6x6 block: lanes stride over the block's pair list, accumulate
// Y_a W_b^T in registers (Y_a = W_a Vinv_l), a shuffle butterfly sums the 36 values and the block is written
// with plain stores -- no atomics, fixed summation order.  W is stored AoS (144 B per observation) so that a
// pair costs two contiguous 144 B reads plus 48 B of Vinv.
#include "fg_internal.h"

namespace fg {

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// per landmark: Vinv = (V + lambda I)^-1 (6 upper), yl = Vinv g_l
__global__ void k_vinv(int64_t L, const double* __restrict__ V, const double* __restrict__ gl, double lambda,
                       double* __restrict__ Vinv, double* __restrict__ yl) {
  int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (l >= L) return;
  double A[9] = {V[6 * l] + lambda, V[6 * l + 1], V[6 * l + 2],
                 V[6 * l + 1], V[6 * l + 3] + lambda, V[6 * l + 4],
                 V[6 * l + 2], V[6 * l + 4], V[6 * l + 5] + lambda};
  double Ai[9];
  inv3(A, Ai);
  Vinv[6 * l] = Ai[0]; Vinv[6 * l + 1] = Ai[1]; Vinv[6 * l + 2] = Ai[2];
  Vinv[6 * l + 3] = Ai[4]; Vinv[6 * l + 4] = Ai[5]; Vinv[6 * l + 5] = Ai[8];
  double g3[3] = {gl[3 * l], gl[3 * l + 1], gl[3 * l + 2]}, y[3];
  m3_vec(Ai, g3, y);
  yl[3 * l] = y[0]; yl[3 * l + 1] = y[1]; yl[3 * l + 2] = y[2];
}

__device__ __forceinline__ void load18(const double* __restrict__ W, int o, double* w) {
  const double2* p = reinterpret_cast<const double2*>(W + (int64_t)o * 18);
#pragma unroll
  for (int i = 0; i < 9; ++i) { double2 v = __ldg(p + i); w[2 * i] = v.x; w[2 * i + 1] = v.y; }
}

// per observation: Y_o = W_o Vinv_l  (6x3), so that the pair loop below is two 144 B reads and 108 DFMA
__global__ void __launch_bounds__(256) k_ymat(int64_t M, const int* __restrict__ obs_point, const double* __restrict__ W,
                                              const double* __restrict__ Vinv, double* __restrict__ Y) {
  const int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (o >= M) return;
  const int l = obs_point[o];
  double w[18], vi[6];
  load18(W, (int)o, w);
  {
    const double2* p = reinterpret_cast<const double2*>(Vinv + (int64_t)l * 6);
    double2 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
    vi[0] = v0.x; vi[1] = v0.y; vi[2] = v1.x; vi[3] = v1.y; vi[4] = v2.x; vi[5] = v2.y;
  }
  double2* out = reinterpret_cast<double2*>(Y + o * 18);
  double y[18];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    y[3 * i + 0] = w[3 * i] * vi[0] + w[3 * i + 1] * vi[1] + w[3 * i + 2] * vi[2];
    y[3 * i + 1] = w[3 * i] * vi[1] + w[3 * i + 1] * vi[3] + w[3 * i + 2] * vi[4];
    y[3 * i + 2] = w[3 * i] * vi[2] + w[3 * i + 1] * vi[4] + w[3 * i + 2] * vi[5];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) out[i] = make_double2(y[2 * i], y[2 * i + 1]);
}

// one warp per block (p, q)
__global__ void __launch_bounds__(512, 1) k_schur_blocks(int64_t n_blk, const int* __restrict__ blk_order, const int* __restrict__ blk_p, const int* __restrict__ blk_q,
                                                         const int64_t* __restrict__ blk_ptr, const int* __restrict__ pair_a,
                                                         const int* __restrict__ pair_b, const double* __restrict__ Y,
                                                         const double* __restrict__ W, const int* __restrict__ off_pose, SysView sys) {
  const int64_t wg = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wg >= n_blk) return;
  const int64_t blk = blk_order[wg];
  double acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.0;
  const int64_t s = blk_ptr[blk], e = blk_ptr[blk + 1];
  int64_t k = s + lane;
  int oa = -1, ob = -1;
  if (k < e) { oa = __ldg(pair_a + k); ob = __ldg(pair_b + k); }
  while (oa >= 0) {
    double wb[18];
    load18(W, ob, wb);
    const double* ya = Y + (int64_t)oa * 18;
    // prefetch the next pair's indices before the arithmetic
    k += 32;
    int na = -1, nb2 = -1;
    if (k < e) { na = __ldg(pair_a + k); nb2 = __ldg(pair_b + k); }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double y0 = __ldg(ya + 3 * i), y1 = __ldg(ya + 3 * i + 1), y2 = __ldg(ya + 3 * i + 2);
#pragma unroll
      for (int j = 0; j < 6; ++j) acc[6 * i + j] += y0 * wb[3 * j] + y1 * wb[3 * j + 1] + y2 * wb[3 * j + 2];
    }
    oa = na; ob = nb2;
  }
  // butterfly: every lane ends with the full sums
#pragma unroll
  for (int i = 0; i < 36; ++i) {
    double v = acc[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    acc[i] = v;
  }
  const int p = blk_p[blk], q = blk_q[blk];
  const int op = off_pose[p], oq = off_pose[q];
  int ld;
  const int64_t base = sys_find(sys, op, oq, &ld);      // rows of p, columns of q (order(q) <= order(p))
  // lane handles entries lane and lane + 32
#pragma unroll
  for (int i = 0; i < 36; ++i) {
    if ((i & 31) == lane && (i < 32 || lane < 4)) {
      const int r = i / 6, cc = i % 6;
      if (p != q || r >= cc) sys.L[base + r + (int64_t)cc * ld] -= acc[i];
    }
  }
}

// one warp per pose: rhs_p += sum_a W_a yl(a)
__global__ void __launch_bounds__(256) k_schur_rhs(int P, const int64_t* __restrict__ pose_obs_ptr, const int64_t* __restrict__ pose_obs,
                                                   const int* __restrict__ obs_point, const double* __restrict__ W,
                                                   const double* __restrict__ yl, const int* __restrict__ off_pose, SysView sys) {
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= P) return;
  const int64_t b = pose_obs_ptr[p], e = pose_obs_ptr[p + 1];
  if (b == e) return;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t k = b + lane; k < e; k += 32) {
    const int o = (int)pose_obs[k];
    const int l = obs_point[o];
    double w[18];
    load18(W, o, w);
    const double y0 = yl[3 * (int64_t)l], y1 = yl[3 * (int64_t)l + 1], y2 = yl[3 * (int64_t)l + 2];
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[i] += w[3 * i] * y0 + w[3 * i + 1] * y1 + w[3 * i + 2] * y2;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], d);
  if (lane < 6) {
    const int C = off_pose[p] + lane;
    const int sn = sys.col2sn[C];
    const int nr = sys.sn_nrows[sn];
    double v = acc[0];
    if (lane == 1) v = acc[1]; else if (lane == 2) v = acc[2]; else if (lane == 3) v = acc[3];
    else if (lane == 4) v = acc[4]; else if (lane == 5) v = acc[5];
    sys.L[sys.sn_valptr[sn] + (int64_t)(C - sys.sn_col0[sn]) * nr + nr - 1] += v;
  }
}

void launch_schur(fg_ctx* c, double lambda) {
  DevGraph& d = c->d;
  cudaStream_t st = c->stream;
  SysView sys;
  sys.L = d.L; sys.col2sn = d.col2sn; sys.sn_col0 = d.sn_col0; sys.sn_ncols = d.sn_ncols; sys.sn_nrows = d.sn_nrows;
  sys.sn_rowptr = d.sn_rowptr; sys.sn_valptr = d.sn_valptr; sys.rowidx = d.rowidx; sys.n_r = c->sym.n_r;
  const int64_t L = d.n[T_POINT];
  k_vinv<<<cdiv(L, 256), 256, 0, st>>>(L, d.V, d.gl, lambda, d.Vinv, d.yl);
  if (d.n_obs) k_ymat<<<cdiv(d.n_obs, 256), 256, 0, st>>>(d.n_obs, d.obs_point, d.W, d.Vinv, d.Y);
  if (c->kev[2]) cudaEventRecord(c->kev[2], st);
  if (d.n_blk) k_schur_blocks<<<cdiv(d.n_blk * 32, 512), 512, 0, st>>>(d.n_blk, d.blk_order, d.blk_p, d.blk_q, d.blk_ptr, d.pair_a, d.pair_b, d.Y,
                                                                        d.W, d.off[T_POSE], sys);
  if (c->kev[3]) cudaEventRecord(c->kev[3], st);
  const int P = (int)d.n[T_POSE];
  if (d.n_obs) k_schur_rhs<<<cdiv((int64_t)P * 32, 256), 256, 0, st>>>(P, d.pose_obs_ptr, d.pose_obs, d.obs_point, d.W, d.yl, d.off[T_POSE], sys);
}

}  // namespace fg
