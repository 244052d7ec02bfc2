// fg_chol_reg.cu -- K7, fast path: register-tiled, output-stationary supernodal Cholesky (fp64).
//
// Same algorithm, schedule and flag protocol as k_chol (fg_chol.cu) but the TARGET panel of a supernode lives
// in registers for the whole time it is being built: thread t owns rows t and t+512 of the panel and all
// (<= 16) of its columns.  A descendant's update L_d[rows, :] * L_d[a:b, :]^T is accumulated straight into
// those registers -- the descendant's rows are read once from L2 with coalesced loads, its (b-a) x K block is
// broadcast from shared memory, and nothing is written until the supernode is finished.  The diagonal block is
// factored by one warp in shared memory, the panel rows are solved against it in registers, and the finished
// panel goes to global memory with one coalesced store per column.
// Limits: supernode width <= 16 columns (fg_symbolic.cpp caps it), panel height <= 1024 rows; graphs beyond
// that use the generic kernel.
#include "fg_internal.h"

namespace fg {

#define CR_T 512
#define CR_RPT 2          // rows per thread
#define CR_NC 16          // max columns
#define CR_DP 17
#define CR_ROWS (CR_T * CR_RPT)

__device__ __forceinline__ int cr_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void cr_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(CR_T, 1) k_chol_reg(SysView s, const int* __restrict__ sched, const int* __restrict__ upd_ptr,
                                                      const int* __restrict__ upd_d, const int* __restrict__ upd_a,
                                                      const int* __restrict__ upd_b, int* flags, int* counters, int epoch,
                                                      int n_sn, int* status) {
  __shared__ __align__(16) double Bf[CR_NC * CR_NC];   // [k][c]: descendant block scattered to target columns, zero padded
  __shared__ double Ds[CR_NC * CR_DP];
  __shared__ int rows_d[CR_ROWS];
  __shared__ int s_slot, s_first_not_ready, s_groups;
  const int tid = threadIdx.x;

  while (true) {
    if (tid == 0) s_slot = atomicAdd(&counters[0], 1);
    __syncthreads();
    const int slot = s_slot;
    __syncthreads();
    if (slot >= n_sn) break;
    const int sn = sched[slot];
    const int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
    double* Lp = s.L + s.sn_valptr[sn];
    const int* rows_g = s.rowidx + s.sn_rowptr[sn];

    // ---- own rows and panel entries into registers
    int myR[CR_RPT];
    double acc[CR_RPT][CR_NC];
#pragma unroll
    for (int m = 0; m < CR_RPT; ++m) {
      const int r = tid + m * CR_T;
      myR[m] = (r < nr) ? rows_g[r] : -1;
#pragma unroll
      for (int c = 0; c < CR_NC; ++c) acc[m][c] = (r < nr && c < nc) ? Lp[r + (int64_t)c * nr] : 0.0;
    }

    int u = upd_ptr[sn];
    const int u1 = upd_ptr[sn + 1];
    while (u < u1) {
      const int win = min(CR_T, u1 - u);
      if (tid == 0) s_first_not_ready = win;
      __syncthreads();
      if (tid < win && cr_ld_acquire(&flags[upd_d[u + tid]]) != epoch) atomicMin(&s_first_not_ready, tid);
      __syncthreads();
      const int nready = s_first_not_ready;
      __syncthreads();
      if (nready == 0) { __nanosleep(64); continue; }
      for (int uu = u; uu < u + nready; ++uu) {
        const int d = upd_d[uu], a = upd_a[uu], b = upd_b[uu];
        const int K = s.sn_ncols[d], nrd = s.sn_nrows[d];
        const double* Ld = s.L + s.sn_valptr[d];
        const int* rd = s.rowidx + s.sn_rowptr[d];
        const int nrows_u = nrd - a;                       // descendant rows that take part (sorted)
        // stage: row list segment, scattered block
        for (int i = tid; i < nrows_u; i += CR_T) rows_d[i] = rd[a + i];
        for (int i = tid; i < CR_NC * CR_NC; i += CR_T) Bf[i] = 0.0;
        if (tid == 0) s_groups = 0;
        __syncthreads();
        {
          const int nb = b - a;
          for (int i = tid; i < nb * K; i += CR_T) {
            const int j = i % nb, k = i / nb;
            const int cj = rows_d[j] - c0;
            Bf[k * CR_NC + cj] = __ldcg(&Ld[a + j + (int64_t)k * nrd]);
            if (k == 0) atomicOr(&s_groups, 1 << (cj >> 2));
          }
        }
        __syncthreads();
        const int groups = s_groups;
        // map my rows into the descendant's row segment
        int pos[CR_RPT];
#pragma unroll
        for (int m = 0; m < CR_RPT; ++m) {
          pos[m] = -1;
          const int R = myR[m];
          if (R >= 0 && nrows_u > 0 && R >= rows_d[0] && R <= rows_d[nrows_u - 1]) {
            int lo = 0, hi = nrows_u - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (rows_d[mid] < R) lo = mid + 1; else hi = mid; }
            if (rows_d[lo] == R) pos[m] = a + lo;
          }
        }
        // accumulate: acc[m][c] -= L_d[pos[m], k] * Bf[k][c]
        if (pos[0] >= 0 || pos[1] >= 0) {
#pragma unroll 3
          for (int k = 0; k < K; ++k) {
            double x[CR_RPT];
#pragma unroll
            for (int m = 0; m < CR_RPT; ++m) x[m] = (pos[m] >= 0) ? __ldcg(&Ld[pos[m] + (int64_t)k * nrd]) : 0.0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (groups & (1 << g)) {
                const double2 b01 = *reinterpret_cast<const double2*>(&Bf[k * CR_NC + 4 * g]);
                const double2 b23 = *reinterpret_cast<const double2*>(&Bf[k * CR_NC + 4 * g + 2]);
#pragma unroll
                for (int m = 0; m < CR_RPT; ++m) {
                  acc[m][4 * g + 0] -= x[m] * b01.x;
                  acc[m][4 * g + 1] -= x[m] * b01.y;
                  acc[m][4 * g + 2] -= x[m] * b23.x;
                  acc[m][4 * g + 3] -= x[m] * b23.y;
                }
              }
            }
          }
        }
        __syncthreads();
      }
      u += nready;
    }

    // ---- diagonal block: rows 0..nc-1 are owned by threads 0..nc-1 (m = 0)
    if (tid < nc) {
#pragma unroll
      for (int c = 0; c < CR_NC; ++c) Ds[tid * CR_DP + c] = (c <= tid) ? acc[0][c] : 0.0;
    }
    __syncthreads();
    if (tid < 32) {
      const int lane = tid;
      for (int c = 0; c < nc; ++c) {
        double dcc = Ds[c * CR_DP + c];
        if (!(dcc > 0.0)) {
          if (lane == 0) atomicExch(status, 1);
          dcc = 1.0;
        }
        const double inv = rsqrt(dcc);
        const double l = dcc * inv;
        __syncwarp();
        if (lane == c) Ds[c * CR_DP + c] = l;
        if (lane > c && lane < nc) Ds[lane * CR_DP + c] *= inv;
        __syncwarp();
        if (lane > c && lane < nc) {
          const double li = Ds[lane * CR_DP + c];
          for (int j = c + 1; j <= lane; ++j) Ds[lane * CR_DP + j] -= li * Ds[j * CR_DP + c];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- panel solve in registers and store
#pragma unroll
    for (int m = 0; m < CR_RPT; ++m) {
      const int r = tid + m * CR_T;
      if (r >= nr) continue;
      if (r < nc) {
#pragma unroll
        for (int c = 0; c < CR_NC; ++c)
          if (c <= r && c < nc) Lp[r + (int64_t)c * nr] = Ds[r * CR_DP + c];
      } else {
        double x[CR_NC];
#pragma unroll
        for (int c = 0; c < CR_NC; ++c) {
          if (c < nc) {
            double v = acc[m][c];
#pragma unroll
            for (int k = 0; k < c; ++k) v -= x[k] * Ds[c * CR_DP + k];
            x[c] = v / Ds[c * CR_DP + c];
            Lp[r + (int64_t)c * nr] = x[c];
          }
        }
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) cr_st_release(&flags[sn], epoch);
  }
}

bool chol_reg_supported(const fg_ctx* c) { return c->sym.max_ncols <= CR_NC && c->sym.max_nrows <= CR_ROWS; }

void launch_factor_reg(fg_ctx* c) {
  DevGraph& d = c->d;
  SysView s;
  s.L = d.L; s.col2sn = d.col2sn; s.sn_col0 = d.sn_col0; s.sn_ncols = d.sn_ncols; s.sn_nrows = d.sn_nrows;
  s.sn_rowptr = d.sn_rowptr; s.sn_valptr = d.sn_valptr; s.rowidx = d.rowidx; s.n_r = c->sym.n_r;
  int grid = c->num_sms;
  if (grid > c->sym.n_sn) grid = c->sym.n_sn;
  c->epoch += 1;
  cudaMemsetAsync(d.status, 0, sizeof(int), c->stream);
  cudaMemsetAsync(d.counters, 0, sizeof(int) * 4, c->stream);
  k_chol_reg<<<grid, CR_T, 0, c->stream>>>(s, d.sched, d.upd_ptr, d.upd_d, d.upd_a, d.upd_b, d.flags, d.counters, c->epoch,
                                           c->sym.n_sn, d.status);
}

}  // namespace fg
