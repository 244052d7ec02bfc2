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

#define REC 18            // doubles per W / Y record
#define RST 19            // padded record stride in shared memory (odd: conflict-free 8-byte lane accesses)

// per observation: Y_o = W_o Vinv_l (6x3).  256 consecutive records are moved through shared memory so that the global
// loads and stores are fully coalesced (a thread touching its own 144-byte record costs 32 cache-line wavefronts
// per instruction).
__global__ void __launch_bounds__(256) k_ymat(int64_t M, const int* __restrict__ obs_point, const double* __restrict__ W,
                                              const double* __restrict__ Vinv, double* __restrict__ Y) {
  __shared__ double buf[256 * RST];
  const int64_t o0 = blockIdx.x * (int64_t)256;
  const int n = (int)min((int64_t)256, M - o0);
  const int tid = threadIdx.x;
  for (int i = tid; i < n * REC; i += 256) buf[(i / REC) * RST + (i % REC)] = __ldg(W + o0 * REC + i);
  __syncthreads();
  if (tid < n) {
    const int l = obs_point[o0 + tid];
    double vi[6];
    {
      const double2* p = reinterpret_cast<const double2*>(Vinv + (int64_t)l * 6);
      double2 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
      vi[0] = v0.x; vi[1] = v0.y; vi[2] = v1.x; vi[3] = v1.y; vi[4] = v2.x; vi[5] = v2.y;
    }
    double* w = buf + tid * RST;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double w0 = w[3 * i], w1 = w[3 * i + 1], w2 = w[3 * i + 2];
      w[3 * i + 0] = w0 * vi[0] + w1 * vi[1] + w2 * vi[2];
      w[3 * i + 1] = w0 * vi[1] + w1 * vi[3] + w2 * vi[4];
      w[3 * i + 2] = w0 * vi[2] + w1 * vi[4] + w2 * vi[5];
    }
  }
  __syncthreads();
  for (int i = tid; i < n * REC; i += 256) Y[o0 * REC + i] = buf[(i / REC) * RST + (i % REC)];
}

// one warp per block (p, q).  Each iteration handles 32 pairs: the 64 records (Y of the 32 a-observations, W of the
// 32 b-observations) are fetched cooperatively -- 9 lanes x 16 B per record, so one load instruction touches a few
// cache lines instead of 32 -- and parked in the warp's shared-memory slab; then every lane multiplies its own pair.
#define SB_WARPS 4
__global__ void __launch_bounds__(32 * SB_WARPS, 4) k_schur_blocks(int64_t n_blk, const int* __restrict__ blk_p, const int* __restrict__ blk_q,
                                                                    const int64_t* __restrict__ blk_ptr, const int* __restrict__ pair_a,
                                                                    const int* __restrict__ pair_b, const double* __restrict__ Y,
                                                                    const double* __restrict__ W, const int* __restrict__ off_pose, SysView sys) {
  __shared__ double slab[SB_WARPS][2][32 * RST];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t blk = blockIdx.x * (int64_t)SB_WARPS + warp;
  if (blk >= n_blk) return;
  double* Ys = slab[warp][0];
  double* Ws = slab[warp][1];
  double acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.0;
  const int64_t s = blk_ptr[blk], e = blk_ptr[blk + 1];
  for (int64_t k0 = s; k0 < e; k0 += 32) {
    const int64_t k = k0 + lane;
    const bool valid = k < e;
    const int oa = valid ? __ldg(pair_a + k) : -1, ob = valid ? __ldg(pair_b + k) : -1;
    // 32 records x 9 chunks of 16 B = 288 chunks per operand: 9 rounds of 32 lanes
#pragma unroll
    for (int rnd = 0; rnd < 9; ++rnd) {
      const int c = rnd * 32 + lane, j = c / 9, part = c - 9 * j;
      const int ja = __shfl_sync(0xffffffffu, oa, j), jb = __shfl_sync(0xffffffffu, ob, j);
      if (ja >= 0) {
        const double2 y = __ldg(reinterpret_cast<const double2*>(Y + (int64_t)ja * REC) + part);
        const double2 w = __ldg(reinterpret_cast<const double2*>(W + (int64_t)jb * REC) + part);
        Ys[j * RST + 2 * part] = y.x; Ys[j * RST + 2 * part + 1] = y.y;
        Ws[j * RST + 2 * part] = w.x; Ws[j * RST + 2 * part + 1] = w.y;
      }
    }
    __syncwarp();
    if (valid) {
      const double* ya = Ys + lane * RST;
      const double* wb = Ws + lane * RST;
      double wbr[18];
#pragma unroll
      for (int i = 0; i < 18; ++i) wbr[i] = wb[i];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const double y0 = ya[3 * i], y1 = ya[3 * i + 1], y2 = ya[3 * i + 2];
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[6 * i + j] += y0 * wbr[3 * j] + y1 * wbr[3 * j + 1] + y2 * wbr[3 * j + 2];
      }
    }
    __syncwarp();
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
  if (d.n_blk) k_schur_blocks<<<cdiv(d.n_blk, SB_WARPS), 32 * SB_WARPS, 0, st>>>(d.n_blk, d.blk_p, d.blk_q, d.blk_ptr, d.pair_a, d.pair_b, d.Y,
                                                                        d.W, d.off[T_POSE], sys);
  if (c->kev[3]) cudaEventRecord(c->kev[3], st);
  const int P = (int)d.n[T_POSE];
  if (d.n_obs) k_schur_rhs<<<cdiv((int64_t)P * 32, 256), 256, 0, st>>>(P, d.pose_obs_ptr, d.pose_obs, d.obs_point, d.W, d.yl, d.off[T_POSE], sys);
}

}  // namespace fg
