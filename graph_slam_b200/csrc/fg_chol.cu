// fg_chol.cu -- K7/K8: reduced-pose sparse block Cholesky and triangular solves (fp64).
//
// Replaces the linear-solver call inside gtsam::LevenbergMarquardtOptimizer (multifrontal Cholesky,
// SURVEY.md section 3A / A.7) for the damped reduced system built by fg_kernels.cu.
//
// Layout: supernodal panels, column-major, ld = nrows (diag rows, below rows, then ONE extra row that
// carries the right-hand side).  Because the rhs rides as an extra matrix row, the factorisation also
// performs the forward substitution: after k_chol the rhs row holds y = L^-1 b.
//
// k_chol is a persistent left-looking supernodal factorisation: CTA b owns supernodes b, b+G, ...;
// a supernode pulls the updates of its descendants (host-precomputed list, ascending) as soon as each
// descendant's epoch flag is published, so on the chain-like elimination trees of VIO/BA graphs the
// updates from all but the immediate predecessor are applied while waiting.  There is no fp64 kind of
// tcgen05.mma, so the dense tiles use DFMA on CUDA cores (SURVEY.md section 7, K7 note).
#include "fg_internal.h"

namespace fg {

#define CH_T 256
#define CH_RCH 96      // rows per staged chunk of a descendant panel
#define CH_KMAX 32     // max supernode width (must match kMaxSnCols)
#define CH_DP 33

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(CH_T) k_chol(SysView s, const int* __restrict__ upd_ptr, const int* __restrict__ upd_d,
                                               const int* __restrict__ upd_a, const int* __restrict__ upd_b,
                                               int* flags, int epoch, int n_sn, int* status) {
  __shared__ double Ps[CH_KMAX * CH_RCH];   // [k][i]  chunk of descendant rows
  __shared__ double Bs[CH_KMAX * CH_KMAX];  // [k][j]  descendant rows that fall in this supernode's columns
  __shared__ double Ds[CH_KMAX * CH_DP];    // diagonal block
  __shared__ int rel[CH_RCH];
  __shared__ int colj[CH_KMAX];
  __shared__ int rowg[CH_RCH];
  const int tid = threadIdx.x;

  for (int sn = blockIdx.x; sn < n_sn; sn += gridDim.x) {
    const int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
    double* Lp = s.L + s.sn_valptr[sn];
    const int* rows_s = s.rowidx + s.sn_rowptr[sn];

    for (int u = upd_ptr[sn]; u < upd_ptr[sn + 1]; ++u) {
      const int d = upd_d[u], a = upd_a[u], b = upd_b[u];
      if (tid == 0) {
        while (ld_acquire(&flags[d]) != epoch) __nanosleep(40);
      }
      __syncthreads();
      const int K = s.sn_ncols[d], nrd = s.sn_nrows[d];
      const double* Ld = s.L + s.sn_valptr[d];
      const int* rows_d = s.rowidx + s.sn_rowptr[d];
      const int nb = b - a;
      for (int i = tid; i < nb * K; i += CH_T) {
        int j = i % nb, k = i / nb;
        Bs[k * CH_KMAX + j] = __ldcg(&Ld[a + j + (int64_t)k * nrd]);
      }
      if (tid < nb) colj[tid] = rows_d[a + tid] - c0;
      for (int r0 = a; r0 < nrd; r0 += CH_RCH) {
        const int nrc = min(CH_RCH, nrd - r0);
        for (int i = tid; i < nrc * K; i += CH_T) {
          int ii = i % nrc, k = i / nrc;
          Ps[k * CH_RCH + ii] = __ldcg(&Ld[r0 + ii + (int64_t)k * nrd]);
        }
        for (int i = tid; i < nrc; i += CH_T) {
          int R = rows_d[r0 + i];
          rowg[i] = R;
          int r;
          if (R < c0 + nc) r = R - c0;
          else {
            int lo = nc, hi = nr - 1;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (rows_s[mid] < R) lo = mid + 1; else hi = mid; }
            r = lo;
          }
          rel[i] = r;
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
            for (int x = 0; x < 4; ++x) p[x] = (i0 + x < nrc) ? Ps[k * CH_RCH + i0 + x] : 0.0;
#pragma unroll
            for (int y = 0; y < 4; ++y) q[y] = (j0 + y < nb) ? Bs[k * CH_KMAX + j0 + y] : 0.0;
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 4; ++y) acc[x][y] += p[x] * q[y];
          }
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            if (i0 + x >= nrc) continue;
            const int rr = rel[i0 + x];
            const int Rg = rowg[i0 + x];
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              if (j0 + y >= nb) continue;
              const int cj = colj[j0 + y];
              if (Rg < c0 + cj) continue;     // strictly upper part of the diagonal block: not stored
              Lp[rr + (int64_t)cj * nr] -= acc[x][y];
            }
          }
        }
        __syncthreads();
      }
    }

    // ---- dense Cholesky of the diagonal block (warp 0, shared memory)
    for (int i = tid; i < nc * nc; i += CH_T) {
      int r = i % nc, c = i / nc;
      Ds[r * CH_DP + c] = (r >= c) ? Lp[r + (int64_t)c * nr] : 0.0;
    }
    __syncthreads();
    if (tid < 32) {
      const int lane = tid;
      for (int c = 0; c < nc; ++c) {
        double dcc = Ds[c * CH_DP + c];
        if (!(dcc > 0.0)) {          // not positive definite (or NaN): flag and keep going with a safe pivot
          if (lane == 0) atomicExch(status, 1);
          dcc = 1.0;
        }
        double l = sqrt(dcc), inv = 1.0 / l;
        __syncwarp();
        if (lane == c) Ds[c * CH_DP + c] = l;
        if (lane > c && lane < nc) Ds[lane * CH_DP + c] *= inv;
        __syncwarp();
        if (lane > c && lane < nc) {
          double li = Ds[lane * CH_DP + c];
          for (int j = c + 1; j <= lane; ++j) Ds[lane * CH_DP + j] -= li * Ds[j * CH_DP + c];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    for (int i = tid; i < nc * nc; i += CH_T) {
      int r = i % nc, c = i / nc;
      if (r >= c) Lp[r + (int64_t)c * nr] = Ds[r * CH_DP + c];
    }
    // ---- panel solve: X L_dd^T = A  (one row per thread, registers)
    for (int r = nc + tid; r < nr; r += CH_T) {
      double x[CH_KMAX];
#pragma unroll
      for (int c = 0; c < CH_KMAX; ++c) {
        if (c < nc) {
          double v = Lp[r + (int64_t)c * nr];
#pragma unroll
          for (int k = 0; k < c; ++k) v -= x[k] * Ds[c * CH_DP + k];
          x[c] = v / Ds[c * CH_DP + c];
          Lp[r + (int64_t)c * nr] = x[c];
        }
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release(&flags[sn], epoch);
  }
}

// Backward substitution x = L^-T y (y = rhs rows), single CTA walking the supernodes in reverse.
__global__ void __launch_bounds__(256) k_backsolve(SysView s, int n_sn, double* x) {
  __shared__ double tsum[CH_KMAX];
  __shared__ double xs[CH_KMAX];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int sn = n_sn - 1; sn >= 0; --sn) {
    const int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
    const double* Lp = s.L + s.sn_valptr[sn];
    const int* rows = s.rowidx + s.sn_rowptr[sn];
    // t_c = y_c - sum_{r in below rows} L[r,c] x[rows[r]]
    for (int c = w; c < nc; c += 8) {
      double acc = 0.0;
      const double* col = Lp + (int64_t)c * nr;
      for (int r = nc + lane; r < nr - 1; r += 32) acc += col[r] * x[rows[r]];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
      if (lane == 0) tsum[c] = col[nr - 1] - acc;
    }
    __syncthreads();
    if (tid < 32) {
      // L_dd^T x = t : backward, lane c owns t_c
      double t = (lane < nc) ? tsum[lane] : 0.0;
      for (int c = nc - 1; c >= 0; --c) {
        double xc = __shfl_sync(0xffffffffu, t, c) / Lp[c + (int64_t)c * nr];
        if (lane == c) xs[c] = xc;
        if (lane < c) t -= Lp[c + (int64_t)lane * nr] * xc;
      }
    }
    __syncthreads();
    if (tid < nc) x[c0 + tid] = xs[tid];
    __syncthreads();
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
  if (!max_blocks_per_sm) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks_per_sm, k_chol, CH_T, 0);
    if (max_blocks_per_sm < 1) max_blocks_per_sm = 1;
  }
  int grid = c->num_sms * max_blocks_per_sm;      // all CTAs must be co-resident (flag waits)
  if (grid > c->sym.n_sn) grid = c->sym.n_sn;
  c->epoch += 1;
  cudaMemsetAsync(d.status, 0, sizeof(int), c->stream);
  k_chol<<<grid, CH_T, 0, c->stream>>>(s, d.upd_ptr, d.upd_d, d.upd_a, d.upd_b, d.flags, c->epoch, c->sym.n_sn, d.status);
}

void launch_backsolve(fg_ctx* c) {
  SysView s = chol_view(c);
  k_backsolve<<<1, 256, 0, c->stream>>>(s, c->sym.n_sn, c->d.delta);
}

}  // namespace fg
