// fg_chol.cu -- K7/K8: reduced-pose sparse block Cholesky and triangular solves (fp64).
//
// Replaces the linear-solver call inside gtsam::LevenbergMarquardtOptimizer (multifrontal Cholesky,
// SURVEY.md section 3A / A.7) for the damped reduced system built by fg_kernels.cu.
//
// Layout: supernodal panels, column-major, ld = nrows (diag rows, below rows, then ONE extra row that
// carries the right-hand side).  Because the rhs rides as an extra matrix row, the factorisation also
// performs the forward substitution: after k_chol the rhs row holds y = L^-1 b.
//
// k_chol is a persistent left-looking supernodal factorisation.  CTAs take supernodes from a
// level-sorted schedule (a topological order of the nested-dissection elimination tree, fg_symbolic.cpp)
// through an atomic counter, so every independent chain of the tree is worked on at once; a supernode
// pulls the updates of its descendants in list order as soon as their epoch flags are published (flags
// are polled 256 at a time), i.e. everything except the immediate predecessor's update is applied while
// waiting.  k_backsolve runs the backward substitution the same way, top-down.
// There is no fp64 kind of tcgen05.mma, so the dense tiles use DFMA on CUDA cores.
#include <cstdlib>
#include "fg_internal.h"

namespace fg {

#define CH_T 256
#define CH_KMAX 32        // max supernode width (must match kMaxSnCols)
#define CH_DP 33
#define CH_PS 4096        // doubles staged per chunk of a descendant panel
#define CH_RMAX 512       // max rows per chunk
#define CH_ROWCAP 2048    // own row list cached in shared memory up to this length

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct CholSmem {
  double Ps[CH_PS];                 // [k][i] chunk of descendant rows
  double Bs[CH_KMAX * CH_KMAX];     // [k][j] descendant rows that fall in this supernode's columns
  double Ds[CH_KMAX * CH_DP];       // diagonal block
  int rows_s[CH_ROWCAP];
  int rel[CH_RMAX];
  int rowg[CH_RMAX];
  int colj[CH_KMAX];
  int slot;
  int first_not_ready;
};

__global__ void __launch_bounds__(CH_T) k_chol(SysView s, const int* __restrict__ sched, const int* __restrict__ upd_ptr,
                                               const int* __restrict__ upd_d, const int* __restrict__ upd_a,
                                               const int* __restrict__ upd_b, int* flags, int* counters, int epoch,
                                               int n_sn, int* status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CholSmem& sm = *reinterpret_cast<CholSmem*>(smem_raw);
  const int tid = threadIdx.x;

  while (true) {
    if (tid == 0) sm.slot = atomicAdd(&counters[0], 1);
    __syncthreads();
    const int slot = sm.slot;
    __syncthreads();
    if (slot >= n_sn) break;
    const int sn = sched[slot];
    const int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
    double* Lp = s.L + s.sn_valptr[sn];
    const int* rows_g = s.rowidx + s.sn_rowptr[sn];
    const bool cached = nr <= CH_ROWCAP;
    if (cached) for (int i = tid; i < nr; i += CH_T) sm.rows_s[i] = rows_g[i];
    const int* rows_s = cached ? sm.rows_s : rows_g;
    __syncthreads();

    int u = upd_ptr[sn];
    const int u1 = upd_ptr[sn + 1];
    while (u < u1) {
      // ---- poll up to 256 pending descendants at once; process the ready prefix in list order
      const int win = min(CH_T, u1 - u);
      if (tid == 0) sm.first_not_ready = win;
      __syncthreads();
      if (tid < win && ld_acquire(&flags[upd_d[u + tid]]) != epoch) atomicMin(&sm.first_not_ready, tid);
      __syncthreads();
      const int nready = sm.first_not_ready;
      __syncthreads();
      if (nready == 0) { __nanosleep(100); continue; }
      for (int uu = u; uu < u + nready; ++uu) {
        const int d = upd_d[uu], a = upd_a[uu], b = upd_b[uu];
        const int K = s.sn_ncols[d], nrd = s.sn_nrows[d];
        const double* Ld = s.L + s.sn_valptr[d];
        const int* rows_d = s.rowidx + s.sn_rowptr[d];
        const int nb = b - a;
        for (int i = tid; i < nb * K; i += CH_T) {
          int j = i % nb, k = i / nb;
          sm.Bs[k * CH_KMAX + j] = __ldcg(&Ld[a + j + (int64_t)k * nrd]);
        }
        if (tid < nb) sm.colj[tid] = rows_d[a + tid] - c0;
        int rch = (CH_PS / K) & ~3;
        if (rch > CH_RMAX) rch = CH_RMAX;
        for (int r0 = a; r0 < nrd; r0 += rch) {
          const int nrc = min(rch, nrd - r0);
          for (int i = tid; i < nrc * K; i += CH_T) {
            int ii = i % nrc, k = i / nrc;
            sm.Ps[k * rch + ii] = __ldcg(&Ld[r0 + ii + (int64_t)k * nrd]);
          }
          for (int i = tid; i < nrc; i += CH_T) {
            int R = rows_d[r0 + i];
            sm.rowg[i] = R;
            int r;
            if (R < c0 + nc) r = R - c0;
            else {
              int lo = nc, hi = nr - 1;
              while (lo < hi) { int mid = (lo + hi) >> 1; if (rows_s[mid] < R) lo = mid + 1; else hi = mid; }
              r = lo;
            }
            sm.rel[i] = r;
          }
          __syncthreads();
          // micro tiles: 4 rows x 4 cols
          const int ntr = (nrc + 3) >> 2, ntc = (nb + 3) >> 2;
          for (int t = tid; t < ntr * ntc; t += CH_T) {
            const int ti = t % ntr, tj = t / ntr;
            const int i0 = ti << 2, j0 = tj << 2;
            double acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
            for (int k = 0; k < K; ++k) {
              double p[4], q[4];
#pragma unroll
              for (int x = 0; x < 4; ++x) p[x] = (i0 + x < nrc) ? sm.Ps[k * rch + i0 + x] : 0.0;
#pragma unroll
              for (int y = 0; y < 4; ++y) q[y] = (j0 + y < nb) ? sm.Bs[k * CH_KMAX + j0 + y] : 0.0;
#pragma unroll
              for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] += p[x] * q[y];
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              if (i0 + x >= nrc) continue;
              const int rr = sm.rel[i0 + x];
              const int Rg = sm.rowg[i0 + x];
#pragma unroll
              for (int y = 0; y < 4; ++y) {
                if (j0 + y >= nb) continue;
                const int cj = sm.colj[j0 + y];
                if (Rg < c0 + cj) continue;     // strictly upper part of the diagonal block: not stored
                Lp[rr + (int64_t)cj * nr] -= acc[x][y];
              }
            }
          }
          __syncthreads();
        }
      }
      u += nready;
    }

    // ---- dense Cholesky of the diagonal block (warp 0, shared memory)
    for (int i = tid; i < nc * nc; i += CH_T) {
      int r = i % nc, c = i / nc;
      sm.Ds[r * CH_DP + c] = (r >= c) ? Lp[r + (int64_t)c * nr] : 0.0;
    }
    __syncthreads();
    if (tid < 32) {
      const int lane = tid;
      for (int c = 0; c < nc; ++c) {
        double dcc = sm.Ds[c * CH_DP + c];
        if (!(dcc > 0.0)) {          // not positive definite (or NaN): flag and keep going with a safe pivot
          if (lane == 0) atomicExch(status, 1);
          dcc = 1.0;
        }
        double inv = rsqrt(dcc);
        double l = dcc * inv;
        __syncwarp();
        if (lane == c) sm.Ds[c * CH_DP + c] = l;
        if (lane > c && lane < nc) sm.Ds[lane * CH_DP + c] *= inv;
        __syncwarp();
        if (lane > c && lane < nc) {
          double li = sm.Ds[lane * CH_DP + c];
          for (int j = c + 1; j <= lane; ++j) sm.Ds[lane * CH_DP + j] -= li * sm.Ds[j * CH_DP + c];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = tid; i < nc * nc; i += CH_T) {
      int r = i % nc, c = i / nc;
      if (r >= c) Lp[r + (int64_t)c * nr] = sm.Ds[r * CH_DP + c];
    }
    // ---- panel solve: X L_dd^T = A  (one row per thread, registers)
    for (int r = nc + tid; r < nr; r += CH_T) {
      double x[CH_KMAX];
#pragma unroll
      for (int c = 0; c < CH_KMAX; ++c) {
        if (c < nc) {
          double v = Lp[r + (int64_t)c * nr];
#pragma unroll
          for (int k = 0; k < c; ++k) v -= x[k] * sm.Ds[c * CH_DP + k];
          x[c] = v / sm.Ds[c * CH_DP + c];
          Lp[r + (int64_t)c * nr] = x[c];
        }
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(&flags[sn], epoch);
  }
}

// Backward substitution x = L^-T y (y = rhs rows).  Supernodes are taken from the schedule in reverse;
// x_s needs the solutions of the supernodes that own its below-diagonal rows (ancestor list).
__global__ void __launch_bounds__(CH_T) k_backsolve(SysView s, const int* __restrict__ sched, const int* __restrict__ anc_ptr,
                                                    const int* __restrict__ anc_t, const int* __restrict__ anc_b,
                                                    int* flags2, int* counters, int epoch, int n_sn, double* x) {
  __shared__ double Ds[CH_KMAX * CH_DP];
  __shared__ double part[8][CH_KMAX];
  __shared__ double xs[CH_KMAX];
  __shared__ int s_slot, s_pending;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  while (true) {
    if (tid == 0) s_slot = atomicAdd(&counters[1], 1);
    __syncthreads();
    const int slot = s_slot;
    __syncthreads();
    if (slot >= n_sn) break;
    const int sn = sched[n_sn - 1 - slot];
    const int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
    const double* Lp = s.L + s.sn_valptr[sn];
    const int* rows = s.rowidx + s.sn_rowptr[sn];
    for (int i = tid; i < nc * nc; i += CH_T) {
      int r = i % nc, c = i / nc;
      Ds[r * CH_DP + c] = Lp[r + (int64_t)c * nr];
    }
    double acc[CH_KMAX / 8];
#pragma unroll
    for (int q = 0; q < CH_KMAX / 8; ++q) acc[q] = 0.0;
    // phase 1: all ancestors except the nearest one; phase 2: the nearest (it finishes last)
    const int a0 = anc_ptr[sn], a1 = anc_ptr[sn + 1];
    for (int phase = 0; phase < 2; ++phase) {
      int lo_e, hi_e, r_lo, r_hi;
      if (phase == 0) { lo_e = a0 + 1; hi_e = a1; r_lo = (a1 > a0) ? anc_b[a0] : nr - 1; r_hi = nr - 1; }
      else { lo_e = a0; hi_e = min(a0 + 1, a1); r_lo = nc; r_hi = (a1 > a0) ? anc_b[a0] : nc; }
      // wait for every ancestor of this phase (polled in parallel)
      while (true) {
        if (tid == 0) s_pending = 0;
        __syncthreads();
        int pend = 0;
        for (int e = lo_e + tid; e < hi_e; e += CH_T)
          if (ld_acquire(&flags2[anc_t[e]]) != epoch) pend = 1;
        if (pend) atomicOr(&s_pending, 1);
        __syncthreads();
        const int p = s_pending;
        __syncthreads();
        if (!p) break;
        __nanosleep(100);
      }
      // warp w owns columns w, w+8, ... ; lanes stride over the rows of this phase
#pragma unroll
      for (int q = 0; q < CH_KMAX / 8; ++q) {
        const int c = w + 8 * q;
        if (c < nc) {
          const double* col = Lp + (int64_t)c * nr;
          double a = 0.0;
          for (int r = r_lo + lane; r < r_hi; r += 32) a += col[r] * __ldcg(&x[rows[r]]);
          acc[q] += a;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < CH_KMAX / 8; ++q) {
      double a = acc[q];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) a += __shfl_down_sync(0xffffffffu, a, d);
      const int c = w + 8 * q;
      if (lane == 0 && c < nc) part[w][q] = Lp[(nr - 1) + (int64_t)c * nr] - a;
    }
    __syncthreads();
    if (tid < 32) {
      // L_dd^T x = t : backward, lane c owns t_c
      double t = (lane < nc) ? part[lane & 7][lane >> 3] : 0.0;
      for (int c = nc - 1; c >= 0; --c) {
        double xc = __shfl_sync(0xffffffffu, t, c) / Ds[c * CH_DP + c];
        if (lane == c) xs[c] = xc;
        if (lane < c) t -= Ds[c * CH_DP + lane] * xc;
      }
    }
    __syncthreads();
    if (tid < nc) x[c0 + tid] = xs[tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(&flags2[sn], epoch);
  }
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
// pending, the 16 partial sums are reduced by shuffles, the diagonal block (one column per lane, loaded up front)
// is back-substituted with static register indices.
#define BW_WARPS 4
__device__ __forceinline__ int bw_ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(32 * BW_WARPS, 4) k_backsolve_w(SysView s, const int* __restrict__ sched, const int* __restrict__ anc_ptr,
                                                              const int* __restrict__ anc_t, const int* __restrict__ anc_b,
                                                              int* flags2, int* counters, int epoch, int n_sn, double* x) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  while (true) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(&counters[1], 1);
    slot = __shfl_sync(FULL, slot, 0);
    if (slot >= n_sn) break;
    const int sn = sched[n_sn - 1 - slot];
    const int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
    const double* Lp = s.L + s.sn_valptr[sn];
    const int* rows = s.rowidx + s.sn_rowptr[sn];
    const int a0 = anc_ptr[sn], a1 = anc_ptr[sn + 1];
    // lane k keeps column k of the diagonal block (rows >= k), its reciprocal pivot and the right-hand side y_k
    double Lcol[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) Lcol[c] = (lane < nc && c < nc && c >= lane) ? Lp[c + (int64_t)lane * nr] : 0.0;
    const double rdg = (lane < nc) ? 1.0 / Lp[lane + (int64_t)lane * nr] : 0.0;
    const double yk = (lane < nc) ? Lp[(nr - 1) + (int64_t)lane * nr] : 0.0;
    double acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.0;
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
        for (int c = 0; c < 16; ++c) if (c < nc) acc[c] = fma(Lp[r + (int64_t)c * nr], xr, acc[c]);
      }
      hi = lo;
    }
    // sums over the lanes; lane k ends with t_k = y_k - sum_k
    double t = 0.0;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      double v = acc[c];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
      if (lane == c) t = yk - v;
    }
    // L_dd^T x = t, backwards
    double xk = 0.0;
#pragma unroll
    for (int c = 15; c >= 0; --c) {
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

void launch_factor(fg_ctx* c) {
  DevGraph& d = c->d;
  SysView s = chol_view(c);
  static int max_blocks_per_sm = 0;
  const size_t smem = sizeof(CholSmem);
  if (!max_blocks_per_sm) {
    cudaFuncSetAttribute(k_chol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, k_chol, CH_T, smem);
    if (max_blocks_per_sm < 1) max_blocks_per_sm = 1;
  }
  int grid = c->num_sms * max_blocks_per_sm;
  if (grid > c->sym.n_sn) grid = c->sym.n_sn;
  c->epoch += 1;
  cudaMemsetAsync(d.status, 0, sizeof(int), c->stream);
  cudaMemsetAsync(d.counters, 0, sizeof(int) * 4, c->stream);
  k_chol<<<grid, CH_T, smem, c->stream>>>(s, d.sched, d.upd_ptr, d.upd_d, d.upd_a, d.upd_b, d.flags, d.counters, c->epoch,
                                          c->sym.n_sn, d.status);
}

void launch_marginal(fg_ctx* c, int col0, int dim, double* work, double* out36) {
  SysView s = chol_view(c);
  cudaMemsetAsync(work, 0, sizeof(double) * 6 * ((size_t)c->sym.n_r + 1), c->stream);
  k_marginal<<<1, 256, 0, c->stream>>>(s, col0, dim, work, out36);
}

void launch_backsolve(fg_ctx* c) {
  DevGraph& d = c->d;
  SysView s = chol_view(c);
  const char* cta = getenv("FG_BACKSOLVE_CTA");          // tests keep the CTA-per-supernode kernel covered
  if (c->sym.max_ncols <= 16 && !(cta && cta[0] == '1')) {
    const char* gm = getenv("FG_BW_GRID");
    int gridw = c->num_sms * (gm ? atoi(gm) : 1);      // measured: 1.19 / 1.26 / 1.32 ms at 1 / 2 / 4 CTAs per SM (fewer pollers)
    if (gridw * BW_WARPS > c->sym.n_sn) gridw = (c->sym.n_sn + BW_WARPS - 1) / BW_WARPS;
    k_backsolve_w<<<gridw, 32 * BW_WARPS, 0, c->stream>>>(s, d.sched, d.anc_ptr, d.anc_t, d.anc_b, d.flags2, d.counters, c->epoch,
                                                         c->sym.n_sn, d.delta);
    return;
  }
  int grid = c->num_sms * 2;
  if (grid > c->sym.n_sn) grid = c->sym.n_sn;
  k_backsolve<<<grid, CH_T, 0, c->stream>>>(s, d.sched, d.anc_ptr, d.anc_t, d.anc_b, d.flags2, d.counters, c->epoch,
                                            c->sym.n_sn, d.delta);
}

}  // namespace fg
