// fg_chol.cu -- K8: backward substitution and marginal covariances on the factored reduced system (fp64).
//
// Replaces the back-substitution of gtsam::LevenbergMarquardtOptimizer's linear solve (multifrontal Cholesky,
// SURVEY.md section 3A / A.7) and Marginals::marginalCovariance (gtsam/gtsam_graph.cpp:598-601,1357).
//
// Layout: supernodal panels, column-major, ld = nrows (diag rows, below rows, then ONE extra row that carries the
// right-hand side).  Because the rhs rides as an extra matrix row, the factorisation (fg_chol_rs.cu) also performs the
// forward substitution: afterwards the rhs row holds y = L^-1 b, and k_backsolve_w solves L^T x = y top-down.
#include <cstdlib>
#include "fg_internal.h"

namespace fg {

#define CH_KMAX 32        // max supernode width (must match kMaxSnCols)

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Marginal covariance block of one reduced variable: Sigma_jj = E_j^T (L L^T)^-1 E_j = Z^T Z with L Z = E_j.
// Z is non-zero only on the path from the variable's supernode to the root of the elimination tree, so one CTA walks
// that path: solve the diagonal block for the (<= 6) right-hand sides, add t^T t to Sigma, push L_below t to the
// ancestors' rows of the work vector (n_r x 6, zeroed by the launcher), move to the parent.
__global__ void __launch_bounds__(256) k_marginal(SysView s, int col0, int dim, double* work, double* out36) {
  __shared__ double t[CH_KMAX][6];
  __shared__ double sig[36];
  const int tid = threadIdx.x;
  const int ldw = s.n_r + 1;
  if (tid < 36) sig[tid] = 0.0;
  if (tid < dim) work[col0 + tid + (int64_t)tid * ldw] = 1.0;
  __threadfence_block();
  __syncthreads();
  int sn = s.col2sn[col0];
  while (true) {
    const int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
    const double* Lp = s.L + s.sn_valptr[sn];
    const int* rows = s.rowidx + s.sn_rowptr[sn];
    // forward substitution on the diagonal block, one right-hand side per thread
    if (tid < dim) {
      double v[CH_KMAX];
      for (int c = 0; c < nc; ++c) {
        double a = work[c0 + c + (int64_t)tid * ldw];
        for (int k = 0; k < c; ++k) a -= Lp[c + (int64_t)k * nr] * v[k];
        v[c] = a / Lp[c + (int64_t)c * nr];
        t[c][tid] = v[c];
      }
    }
    __syncthreads();
    if (tid < dim * dim) {
      const int a = tid / dim, b = tid % dim;
      double acc = 0.0;
      for (int c = 0; c < nc; ++c) acc += t[c][a] * t[c][b];
      sig[a * 6 + b] += acc;
    }
    // below rows (the right-hand-side row nr - 1 is not part of the matrix)
    const int nbelow = nr - 1 - nc;
    for (int i = tid; i < nbelow * dim; i += blockDim.x) {
      const int r = nc + i % nbelow, k = i / nbelow;
      double a = 0.0;
      for (int c = 0; c < nc; ++c) a += Lp[r + (int64_t)c * nr] * t[c][k];
      work[rows[r] + (int64_t)k * ldw] -= a;
    }
    __threadfence_block();
    __syncthreads();
    if (nbelow <= 0) break;
    sn = s.col2sn[rows[nc]];          // parent in the elimination tree
  }
  if (tid < 36) out36[tid] = sig[tid];
}

void launch_marginal(fg_ctx* c, int col0, int dim, double* work, double* out36);

// Warp-per-supernode variant of the backward substitution (default).  The solve is a dependency chain of ~n_levels
// steps, so what counts is the time from "nearest ancestor published" to "own solution published": one warp does the
// whole supernode without block barriers -- far ancestors' rows are folded in while the nearest ancestor is still
// pending, the partial sums are reduced by shuffles, the diagonal block (one column per lane, loaded up front)
// is back-substituted with static register indices.
#define BW_WARPS 4
__device__ __forceinline__ int bw_ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(32 * BW_WARPS, 3) k_backsolve_w(SysView s, const int* __restrict__ sched, const int* __restrict__ anc_ptr,
                                                              const int* __restrict__ anc_t, const int* __restrict__ anc_b,
                                                              int* flags2, int* counters, int epoch, int n_sn, double* x) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  while (true) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(&counters[1], 1);
    slot = __shfl_sync(FULL, slot, 0);
    if (slot >= n_sn) break;
    const int sn = sched[slot];                            // processing order: reverse level order (this rank's supernodes)
    const int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
    const double* Lp = s.L + s.sn_valptr[sn];
    const int* rows = s.rowidx + s.sn_rowptr[sn];
    const int a0 = anc_ptr[sn], a1 = anc_ptr[sn + 1];
    // lane k keeps column k of the diagonal block (rows >= k), its reciprocal pivot and the right-hand side y_k
    double Lcol[CH_KMAX];
#pragma unroll
    for (int c = 0; c < CH_KMAX; ++c) Lcol[c] = (lane < nc && c < nc && c >= lane) ? Lp[c + (int64_t)lane * nr] : 0.0;
    const double rdg = (lane < nc) ? 1.0 / Lp[lane + (int64_t)lane * nr] : 0.0;
    const double yk = (lane < nc) ? Lp[(nr - 1) + (int64_t)lane * nr] : 0.0;
    double acc[CH_KMAX];
#pragma unroll
    for (int c = 0; c < CH_KMAX; ++c) acc[c] = 0.0;
    // ancestors finish from the far end of the row list towards the nearest one: fold in the rows of every ready suffix
    // of the ancestor list as soon as it is ready, so that only the nearest ancestor's rows are left on the critical path
    int hi = a1;                                           // entries [hi, a1) are folded in
    while (hi > a0) {
      int lo;
      while (true) {
        int nr_max = a0 - 1;                               // highest entry that is not ready
        for (int base = a0; base < hi; base += 32) {
          const int e = base + lane;
          const bool notready = (e < hi) && bw_ld_relaxed(&flags2[anc_t[e]]) != epoch;
          const unsigned m = __ballot_sync(FULL, notready);
          if (m) nr_max = base + 31 - __clz(m);
        }
        lo = nr_max + 1;
        if (lo < hi) break;
      }
      __threadfence();
      const int r_lo = (lo == a0) ? nc : anc_b[lo - 1], r_hi = anc_b[hi - 1];
      for (int r = r_lo + lane; r < r_hi; r += 32) {
        const double xr = __ldcg(&x[rows[r]]);
#pragma unroll
        for (int c = 0; c < CH_KMAX; ++c) if (c < nc) acc[c] = fma(Lp[r + (int64_t)c * nr], xr, acc[c]);
      }
      hi = lo;
    }
    // sums over the lanes; lane k ends with t_k = y_k - sum_k
    double t = 0.0;
#pragma unroll
    for (int c = 0; c < CH_KMAX; ++c) {
      double v = acc[c];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
      if (lane == c) t = yk - v;
    }
    // L_dd^T x = t, backwards
    double xk = 0.0;
#pragma unroll
    for (int c = CH_KMAX - 1; c >= 0; --c) {
      if (c < nc) {                                        // warp uniform
        const double xc = __shfl_sync(FULL, t * rdg, c);
        if (lane == c) xk = xc;
        t = fma(-Lcol[c], xc, t);                          // lanes k < c: t_k -= L[c][k] x_c (Lcol[c] is 0 for k > c; lane c is done)
      }
    }
    if (lane < nc) x[c0 + lane] = xk;
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(&flags2[sn], epoch);
  }
}

static SysView chol_view(fg_ctx* c) {
  DevGraph& d = c->d;
  SysView s;
  s.L = d.L; s.col2sn = d.col2sn; s.sn_col0 = d.sn_col0; s.sn_ncols = d.sn_ncols; s.sn_nrows = d.sn_nrows;
  s.sn_rowptr = d.sn_rowptr; s.sn_valptr = d.sn_valptr; s.rowidx = d.rowidx; s.n_r = c->sym.n_r;
  return s;
}

void launch_marginal(fg_ctx* c, int col0, int dim, double* work, double* out36) {
  SysView s = chol_view(c);
  cudaMemsetAsync(work, 0, sizeof(double) * 6 * ((size_t)c->sym.n_r + 1), c->stream);
  k_marginal<<<1, 256, 0, FGS(c->stream)>>>(s, col0, dim, work, out36);
}

void launch_backsolve(fg_ctx* c) {
  DevGraph& d = c->d;
  SysView s = chol_view(c);
  const int n = c->sym.n_sn;
  int gridw = c->num_sms;                                // measured: 1.19 / 1.26 / 1.32 ms at 1 / 2 / 4 CTAs per SM (fewer pollers)
  if (gridw * BW_WARPS > n) gridw = (n + BW_WARPS - 1) / BW_WARPS;
  k_backsolve_w<<<gridw, 32 * BW_WARPS, 0, FGS(c->stream)>>>(s, d.bs_order, d.anc_ptr, d.anc_t, d.anc_b, d.flags2, d.counters, c->epoch, n, d.delta);
}

}  // namespace fg
