// fg_chol_reg.cu -- K7, fast path: supernodal Cholesky with the target panel resident on chip (fp64).
//
// Same algorithm, schedule and flag protocol as k_chol (fg_chol.cu).  What changes is the data flow of one
// descendant update  P[rows, cols] -= L_d[rows, :] * L_d[a:b, :]^T :
//   * the TARGET panel P (<= 1024 rows x <= 16 columns, 128 KB) and the target's row list stay in shared memory
//     for the whole time the supernode is being built (one CTA per SM);
//   * each thread takes rows of the DESCENDANT (thread i: rows a+i and a+i+512): the row index and the K values
//     of that row are loaded with independent, coalesced L2 reads, so an update costs ONE L2 latency instead of
//     a chain of them; the (b-a) x K block is broadcast from a double-buffered shared tile;
//   * the row is located in the target's sorted row list by a binary search in shared memory and the products
//     are subtracted in place (distinct rows per thread: no conflicts, fixed order: deterministic).
// The diagonal block is factored by one warp with compact rolled loops in shared memory, the panel rows are solved
// against it in place in shared memory and the finished panel is written to global memory once, coalesced.
// Limits: supernode width <= 16 columns (fg_symbolic.cpp caps it), panel height <= 1024 rows; graphs beyond that
// use the generic kernel in fg_chol.cu.  No fp64 tcgen05 kind exists, hence DFMA.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fg_internal.h"

namespace fg {

#define CR_T 512
#define CR_RPT 2          // descendant rows per thread
#define CR_NC 16          // max columns
#define CR_DP 17
#define CR_ROWS (CR_T * CR_RPT)

__device__ __forceinline__ int cr_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int cr_ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void cr_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// acc[j] = sum_k x[k] * B[k][j] for j < 2*JP: straight-line (x is zero for k >= K, the tile is zero padded).
// Only JP = 8 is instantiated: the kernel's instruction footprint decides the latency of the single-warp sections
// (measured: ~35 cycles/instruction when the hot path did not fit the instruction cache), so one compact variant
// with some padded DFMAs beats eight exact ones.
template <int JP>
__device__ __forceinline__ void cr_row_times_block(const double* __restrict__ x, const double* __restrict__ Bt, double* acc) {
#pragma unroll
  for (int j = 0; j < 2 * JP; ++j) acc[j] = 0.0;
#pragma unroll
  for (int k = 0; k < CR_NC; ++k) {
    const double xk = x[k];
    const double2* brow = reinterpret_cast<const double2*>(Bt + k * CR_NC);
#pragma unroll
    for (int jp = 0; jp < JP; ++jp) {
      const double2 bb = brow[jp];
      acc[2 * jp] += xk * bb.x;
      acc[2 * jp + 1] += xk * bb.y;
    }
  }
}

struct CrSmem {
  double P[CR_ROWS * CR_NC];            // target panel, column-major, ld = nr
  double Bs[2][CR_NC * CR_NC];          // [buf][k][j] descendant block, j < nb
  double Ds[CR_NC * CR_DP];
  double colbuf[CR_NC];                 // current column of the diagonal factor (keeps the potrf inner loop alias-free)
  double dinv[CR_NC];                   // 1 / L_cc
  int rows_s[CR_ROWS];
  int rl[CR_ROWS];                      // front row list of the leaf being subtracted (phase C prologue)
  int colidx[CR_NC];
  int colj[2][CR_NC];
  int slot, first_not_ready;
};

__global__ void __launch_bounds__(CR_T, 1) k_chol_reg(SysView s, const int* __restrict__ sched, const int* __restrict__ upd_ptr,
                                                      const int* __restrict__ upd_d, const int* __restrict__ upd_a,
                                                      const int* __restrict__ upd_b, const UpdRec* __restrict__ upd_rec, int* flags, int* counters,
                                                      int epoch, int n_sn, int* status, long long* dbg, FrontView fv) {
  extern __shared__ __align__(16) unsigned char cr_raw[];
  CrSmem& sm = *reinterpret_cast<CrSmem*>(cr_raw);
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
    if (dbg && tid == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[8 * sn] = t; dbg[8 * sn + 7] = blockIdx.x; }
    for (int i = tid; i < nr; i += CR_T) sm.rows_s[i] = rows_g[i];
    for (int i = tid; i < nr * nc; i += CR_T) sm.P[i] = Lp[i];
    __syncthreads();
    // ---- phase C prologue: subtract the dense leaf fronts that reach this supernode (fg_front.cu)
    if (fv.tf_ptr) {
      for (int e = fv.tf_ptr[sn]; e < fv.tf_ptr[sn + 1]; ++e) {
        const int l = fv.tf_leaf[e];
        const int* Rl = fv.fr_rows + fv.fr_rowptr[l];
        const int nR = fv.fr_rowptr[l + 1] - fv.fr_rowptr[l];
        const double* Ul = fv.U + fv.fr_uptr[l];
        for (int i = tid; i < nR; i += CR_T) sm.rl[i] = Rl[i];
        __syncthreads();
        if (tid < nc) {
          const int g = c0 + tid;
          int lo = 0, hi = nR - 1;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.rl[mid] < g) lo = mid + 1; else hi = mid; }
          sm.colidx[tid] = (nR > 0 && sm.rl[lo] == g) ? lo : -1;
        }
        __syncthreads();
        for (int r = tid; r < nr; r += CR_T) {
          const int g = sm.rows_s[r];
          int lo = 0, hi = nR - 1;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.rl[mid] < g) lo = mid + 1; else hi = mid; }
          if (nR > 0 && sm.rl[lo] == g) {
            const double* urow = Ul + (int64_t)lo * nR;
            for (int c = 0; c < nc; ++c) {
              const int jc = sm.colidx[c];
              if (jc >= 0 && g >= c0 + c) sm.P[r + c * nr] -= urow[jc];
            }
          }
        }
        __syncthreads();
      }
    }

    int u = upd_ptr[sn];
    const int u1 = upd_ptr[sn + 1];
    int buf = 0;
    while (u < u1) {
      // warp 0 spins on the flags of the next (up to 32) descendants with relaxed loads and publishes the length
      // of the ready prefix; one acquire fence once something is ready, one block barrier per batch
      if (tid < 32) {
        const int win = min(32, u1 - u);
        const int* fp = flags + upd_d[u + min(tid, win - 1)];
        int n;
        while (true) {
          const int f = (tid < win) ? cr_ld_relaxed(fp) : epoch;
          const unsigned notready = ~__ballot_sync(0xffffffffu, f == epoch);
          n = notready ? (__ffs(notready) - 1) : 32;
          if (n > win) n = win;
          if (n > 0) break;
        }
        __threadfence();
        if (tid == 0) sm.first_not_ready = n;
      }
      __syncthreads();
      const int nready = sm.first_not_ready;
      if (dbg && tid == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[8 * sn + 1] = t; }
      UpdRec rec = upd_rec[u];
      for (int uu = u; uu < u + nready; ++uu, buf ^= 1) {
        const int K = rec.K, nrd = rec.nrd, nrows_u = rec.nrows_u, nb = rec.nb;
        const double* Ld = s.L + rec.val_off;          // points at descendant row a
        const int* rd = s.rowidx + rec.row_off;        // row list from row a on
        const int a = 0;
        if (uu + 1 < u + nready) rec = upd_rec[uu + 1];   // next record rides under this update's latency
        // ---- issue every load of this update up front (all independent)
        int R[CR_RPT];
        double x[CR_RPT][CR_NC];
#pragma unroll
        for (int m = 0; m < CR_RPT; ++m) {
          const int i = tid + m * CR_T;
          const bool act = i < nrows_u;
          R[m] = act ? __ldg(rd + a + i) : -1;
#pragma unroll
          for (int k = 0; k < CR_NC; ++k) x[m][k] = (act && k < K) ? __ldcg(&Ld[a + i + (int64_t)k * nrd]) : 0.0;
        }
        if (tid < CR_NC * CR_NC) {
          const int j = tid % CR_NC, k = tid / CR_NC;
          sm.Bs[buf][tid] = (j < nb && k < K) ? __ldcg(&Ld[a + j + (int64_t)k * nrd]) : 0.0;
        }
        if (tid < nb) sm.colj[buf][tid] = __ldg(rd + a + tid) - c0;
        __syncthreads();
        // ---- per descendant row: products against the block, subtract into the target panel
#pragma unroll
        for (int m = 0; m < CR_RPT; ++m) {
          if (R[m] < 0) continue;
          int r;
          if (R[m] < c0 + nc) r = R[m] - c0;
          else {
            int lo = nc, hi = nr - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.rows_s[mid] < R[m]) lo = mid + 1; else hi = mid; }
            r = lo;
          }
          double acc[CR_NC];
          const double* Bt = sm.Bs[buf];
          cr_row_times_block<CR_NC / 2>(x[m], Bt, acc);     // one compact variant: instruction-cache footprint matters more than DFMA count here
#pragma unroll
          for (int j = 0; j < CR_NC; ++j) {
            if (j < nb) {
              const int cj = sm.colj[buf][j];
              if (R[m] >= c0 + cj) sm.P[r + cj * nr] -= acc[j];     // strictly-upper part of the diagonal block is not stored
            }
          }
        }
        // no barrier here: the next update writes the other Bs/colj buffer, and distinct descendant rows map to
        // distinct target rows within one update; the barrier after the next update's loads orders the rest
      }
      u += nready;
      __syncthreads();
    }
    __syncthreads();
    if (dbg && tid == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[8 * sn + 2] = t; }

    // ---- diagonal block: compact rolled loops in shared memory (warp 0); code size is what counts on this path
    for (int i = tid; i < nc * nc; i += CR_T) {
      const int r = i % nc, c = i / nc;
      sm.Ds[r * CR_DP + c] = (r >= c) ? sm.P[r + c * nr] : 0.0;
    }
    __syncthreads();
    if (tid < 32) {
      const int lane = tid;
      for (int c = 0; c < nc; ++c) {
        double dcc = sm.Ds[c * CR_DP + c];
        if (!(dcc > 0.0)) {          // not positive definite (or NaN): flag and keep going with a safe pivot
          if (lane == 0) atomicExch(status, 1);
          dcc = 1.0;
        }
        const double inv = rsqrt(dcc);
        __syncwarp();
        double li = 0.0;
        if (lane == c) { sm.Ds[c * CR_DP + c] = dcc * inv; sm.dinv[c] = inv; }
        if (lane > c && lane < nc) { li = sm.Ds[lane * CR_DP + c] * inv; sm.Ds[lane * CR_DP + c] = li; sm.colbuf[lane] = li; }
        __syncwarp();
        if (lane > c && lane < nc) {
          double* row = &sm.Ds[lane * CR_DP];
          const double* col = sm.colbuf;
#pragma unroll 4
          for (int j = c + 1; j <= lane; ++j) row[j] -= li * col[j];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (dbg && tid == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[8 * sn + 3] = t; }
    // ---- panel solve in place in shared memory (thread per row, rolled), then one coalesced store per column
    for (int r = tid; r < nr; r += CR_T) {
      if (r < nc) {
        for (int c = 0; c <= r; ++c) sm.P[r + c * nr] = sm.Ds[r * CR_DP + c];
      } else {
        // four columns at a time: the k loop carries four independent accumulators, the 4x4 triangle is solved in registers
        for (int cc = 0; cc < nc; cc += 4) {
          double v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = (cc + i < nc) ? sm.P[r + (cc + i) * nr] : 0.0;
          for (int k = 0; k < cc; ++k) {
            const double xk = sm.P[r + k * nr];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] -= xk * sm.Ds[(cc + i) * CR_DP + k];     // rows >= nc of Ds are never read with cc+i >= nc: guarded below
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (cc + i < nc) {
#pragma unroll
              for (int q = 0; q < i; ++q) v[i] -= v[q] * sm.Ds[(cc + i) * CR_DP + cc + q];
              v[i] *= sm.dinv[cc + i];
              sm.P[r + (cc + i) * nr] = v[i];
            }
          }
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < nr * nc; i += CR_T) {
      const int r = i % nr, c = i / nr;
      if (r >= nc || c <= r) Lp[i] = sm.P[i];
    }
    if (dbg && tid == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[8 * sn + 4] = t; }
    __threadfence();
    __syncthreads();
    if (tid == 0) cr_st_release(&flags[sn], epoch);
    if (dbg && tid == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[8 * sn + 5] = t; }
  }
}

bool chol_reg_supported(const fg_ctx* c) {
  const char* force = getenv("FG_CHOL_GENERIC");      // tests force the generic kernel to keep it covered
  if (force && force[0] == '1') return false;
  return c->sym.max_ncols <= CR_NC && c->sym.max_nrows <= CR_ROWS;
}

void launch_factor_reg(fg_ctx* c) {
  DevGraph& d = c->d;
  SysView s;
  s.L = d.L; s.col2sn = d.col2sn; s.sn_col0 = d.sn_col0; s.sn_ncols = d.sn_ncols; s.sn_nrows = d.sn_nrows;
  s.sn_rowptr = d.sn_rowptr; s.sn_valptr = d.sn_valptr; s.rowidx = d.rowidx; s.n_r = c->sym.n_r;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_chol_reg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CrSmem));
    attr_set = true;
  }
  c->epoch += 1;
  cudaMemsetAsync(d.status, 0, sizeof(int), c->stream);
  // FG_CHOL_TRACE=<file>: dump per-supernode globaltimer stamps of the 3rd factorisation
  static int n_calls = 0;
  long long* dbg = nullptr;
  const char* trace = getenv("FG_CHOL_TRACE");
  if (trace && ++n_calls == 3) { cudaMalloc((void**)&dbg, sizeof(long long) * 8 * c->sym.n_sn); cudaMemset(dbg, 0, sizeof(long long) * 8 * c->sym.n_sn); }
  const char* nofront = getenv("FG_NO_FRONTS");
  const bool fronts = c->sym.use_fronts && !(nofront && nofront[0] == '1');
  FrontView none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (!fronts) {
    int grid = std::min(c->num_sms, c->sym.n_sn);
    cudaMemsetAsync(d.counters, 0, sizeof(int) * 4, c->stream);
    k_chol_reg<<<grid, CR_T, sizeof(CrSmem), c->stream>>>(s, d.sched, d.upd_ptr, d.upd_d, d.upd_a, d.upd_b, d.upd_rec, d.flags, d.counters,
                                                           c->epoch, c->sym.n_sn, d.status, dbg, none);
  } else {
    // phase A: the leaves (reduced update lists); phase B: one dense update matrix per leaf; phase C: the rest
    const int na = (int)c->sym.sched_a.size(), nc = (int)c->sym.sched_c.size();
    cudaMemsetAsync(d.counters, 0, sizeof(int) * 4, c->stream);
    if (na) k_chol_reg<<<std::min(c->num_sms, na), CR_T, sizeof(CrSmem), c->stream>>>(s, d.sched_a, d.updr_ptr, d.updr_d, nullptr, nullptr, d.updr_rec,
                                                                                    d.flags, d.counters, c->epoch, na, d.status, dbg, none);
    launch_front_syrk(c);
    cudaMemsetAsync(d.counters, 0, sizeof(int) * 4, c->stream);
    FrontView fv = {d.tf_ptr, d.tf_leaf, d.fr_rowptr, d.fr_rows, d.fr_uptr, d.U};
    if (nc) k_chol_reg<<<std::min(c->num_sms, nc), CR_T, sizeof(CrSmem), c->stream>>>(s, d.sched_c, d.updr_ptr, d.updr_d, nullptr, nullptr, d.updr_rec,
                                                                                    d.flags, d.counters, c->epoch, nc, d.status, dbg, fv);
  }
  if (dbg) {
    std::vector<long long> h(8 * (size_t)c->sym.n_sn);
    cudaStreamSynchronize(c->stream);
    cudaMemcpy(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
    cudaFree(dbg);
    FILE* f = fopen(trace, "w");
    if (f) {
      for (int i = 0; i < c->sym.n_sn; ++i)
        fprintf(f, "%d %d %d %d %d %lld %lld %lld %lld %lld %lld %lld\n", i, c->sym.level[i], c->sym.sn_ncols[i], c->sym.sn_nrows[i],
                c->sym.upd_ptr[i + 1] - c->sym.upd_ptr[i], h[8 * i], h[8 * i + 1], h[8 * i + 2], h[8 * i + 3], h[8 * i + 4], h[8 * i + 5], h[8 * i + 7]);
      fclose(f);
    }
  }
}

}  // namespace fg
