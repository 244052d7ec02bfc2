// fg_chol_rs.cu -- K7, row-split supernodal Cholesky (fp64): the default factorisation kernel.
//
// Same algorithm, panel layout and schedule idea as k_chol_reg (fg_chol_reg.cu), but the unit of work is a ROW BLOCK
// of a supernode instead of a whole supernode.  A nested-dissection ordering of a 50-frame covisibility band only has
// 32 independent leaf chains at C5 (fg_symbolic.cpp); with one CTA per supernode 32 of the 148 SMs carried three
// quarters of the flops.  Here a supernode of nr rows x nc columns (nc <= 16) is cut into blocks of <= RS_RB
// below-diagonal rows.  A unit (supernode, block)
//   * is output stationary: thread t owns local row t (the nc diagonal rows, then the block's own rows) and keeps its
//     nc values in registers from the first load to the final store -- no shared-memory panel, no write conflicts,
//   * pulls every descendant update restricted to those rows: a host-built map (int16 per unit, update and row) names
//     the descendant row that lands on each local row, so an update is one coalesced index load, K coalesced value
//     loads and K x nc DFMA per thread against the descendant's (rows in the target's columns) block, staged in shared
//     memory already scattered to target columns; the part that lands on the diagonal block is recomputed by every
//     block of the supernode (15 x 15 x K flops),
//   * factors the diagonal block (every block redundantly: no intra-supernode synchronisation), solves its own rows
//     against it and stores them,
//   * bumps the supernode's arrival counter; the block that arrives LAST stores the factored diagonal block (the others
//     read the assembled one when they start, so it must not be overwritten earlier) and publishes the supernode's
//     done flag (release); consumers poll the flag (acquire).
// The next update's loads are issued before the current one is multiplied.
// Deterministic: every panel entry is owned by one unit and updated in list order.  DFMA on CUDA cores: tcgen05 has
// no fp64 kind.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fg_internal.h"

namespace fg {

#define RS_T 256
#define RS_NC 16
#define RS_DP 17
#define RS_RB 240                      // below-diagonal rows per block; 16 + RS_RB <= RS_T: one descendant row per thread
#define RS_LR (RS_NC + RS_RB)          // local rows held by a unit: diagonal rows, then own rows

__device__ __forceinline__ int rs_ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct RsSmem {
  double Bs[2][RS_NC * RS_NC];          // [buf][k][c]: descendant rows that fall in the target's columns, scattered to TARGET columns
  double Ds[RS_NC * RS_DP];
  double dinv[RS_NC];
  int rl[1024];                         // row list of the leaf front being subtracted
  int colidx[RS_NC];
  int slot, first_not_ready;
};

// Cholesky of the nc x nc (nc <= 16) diagonal block by one warp, the matrix in registers: lane r holds row r, the pivot
// and the scaled column travel by shuffles, every index is static (fully unrolled: ~600 instructions, ~120 cycles per
// column).  Rows / columns beyond nc are padded with the identity.  Entries above the diagonal hold garbage that never
// feeds a used value.
__device__ __forceinline__ void rs_potrf_warp(RsSmem& sm, int nc, int lane, int* status) {
  const unsigned FULL = 0xffffffffu;
  const int r = lane & 15;
  double a[RS_NC];
#pragma unroll
  for (int c = 0; c < RS_NC; ++c) a[c] = (r < nc && c < nc) ? sm.Ds[r * RS_DP + c] : (r == c ? 1.0 : 0.0);
  bool bad = false;
#pragma unroll
  for (int c = 0; c < RS_NC; ++c) {
    double d = __shfl_sync(FULL, a[c], c);
    if (!(d > 0.0)) { bad = true; d = 1.0; }       // not positive definite (or NaN): flag and keep going with a safe pivot
    const double inv = rsqrt(d);
    const double l = (r == c) ? d * inv : a[c] * inv;
    a[c] = l;
    if (lane == c) sm.dinv[c] = inv;
#pragma unroll
    for (int j = c + 1; j < RS_NC; ++j) {
      const double lj = __shfl_sync(FULL, l, j);
      a[j] = fma(-l, lj, a[j]);
    }
  }
  if (bad && lane == 0) atomicExch(status, 1);
  if (lane < nc) {
#pragma unroll
    for (int c = 0; c < RS_NC; ++c)
      if (c <= lane) sm.Ds[lane * RS_DP + c] = a[c];
  }
}

__global__ void __launch_bounds__(RS_T, 2)
k_chol_rs(SysView s, const int4* __restrict__ units, const int64_t* __restrict__ unit_moff, const short* __restrict__ rowmap,
          const int* __restrict__ upd_ptr, const int* __restrict__ upd_d, const UpdRec* __restrict__ upd_rec,
          const signed char* __restrict__ colinv, int* arrived, int* done, int* counter, int n_units, int* status, FrontView fv,
          long long* dbg) {
  extern __shared__ __align__(16) unsigned char rs_raw[];
  RsSmem& sm = *reinterpret_cast<RsSmem*>(rs_raw);
  const int tid = threadIdx.x;

  while (true) {
    if (tid == 0) sm.slot = atomicAdd(counter, 1);
    __syncthreads();
    const int slot = sm.slot;
    __syncthreads();
    if (slot >= n_units) break;
#define RS_STAMP(k) if (dbg && tid == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[8 * slot + (k)] = t_; }
    RS_STAMP(0)
    const int4 un = units[slot];
    const int sn = un.x, r0 = un.y, r1 = un.z, nblk = un.w;   // own rows [r0, r1) of the panel, r0 >= nc; blocks of this supernode
    const int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
    const int nloc = nc + (r1 - r0);
    double* Lp = s.L + s.sn_valptr[sn];
    // thread tid owns local row tid of the unit: the diagonal rows first, then the own rows; its nc values live in registers
    const bool has_row = tid < nloc;
    const int prow = tid < nc ? tid : r0 + tid - nc;
    const short* umap = rowmap + unit_moff[slot] + tid;
    double acc[RS_NC];
#pragma unroll
    for (int c = 0; c < RS_NC; ++c) acc[c] = (has_row && c < nc) ? Lp[prow + (int64_t)c * nr] : 0.0;
    // ---- subtract the dense leaf fronts that reach this supernode (fg_front.cu)
    if (fv.tf_ptr) {
      const int g = has_row ? s.rowidx[s.sn_rowptr[sn] + prow] : -1;
      for (int e = fv.tf_ptr[sn]; e < fv.tf_ptr[sn + 1]; ++e) {
        const int l = fv.tf_leaf[e];
        const int* Rl = fv.fr_rows + fv.fr_rowptr[l];
        const int nR = fv.fr_rowptr[l + 1] - fv.fr_rowptr[l];
        const double* Ul = fv.U + fv.fr_uptr[l];
        __syncthreads();
        for (int i = tid; i < nR; i += RS_T) sm.rl[i] = Rl[i];
        __syncthreads();
        if (tid < RS_NC) {
          const int gc = c0 + tid;
          int lo = 0, hi = nR - 1;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.rl[mid] < gc) lo = mid + 1; else hi = mid; }
          sm.colidx[tid] = (tid < nc && nR > 0 && sm.rl[lo] == gc) ? lo : -1;
        }
        __syncthreads();
        if (has_row) {
          int lo = 0, hi = nR - 1;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (sm.rl[mid] < g) lo = mid + 1; else hi = mid; }
          if (nR > 0 && sm.rl[lo] == g) {
            const double* urow = Ul + (int64_t)lo * nR;
#pragma unroll
            for (int c = 0; c < RS_NC; ++c) {
              const int jc = sm.colidx[c];
              if (jc >= 0 && jc <= lo) acc[c] -= urow[jc];      // lower triangle of U only (above it: the unused upper part of the diagonal block)
            }
          }
        }
      }
    }

    int u = upd_ptr[sn];
    const int ubase = u;
    const int u1 = upd_ptr[sn + 1];
    int buf = 0;
    while (u < u1) {
      // warp 0 spins on the done flags of the next (up to 32) descendants and publishes the ready prefix
      if (tid < 32) {
        const int win = min(32, u1 - u);
        const int ui = u + min(tid, win - 1);
        const int* fp = done + upd_d[ui];
        int n;
        while (true) {
          const int f = (tid < win) ? rs_ld_relaxed(fp) : 1;
          const unsigned notready = ~__ballot_sync(0xffffffffu, f != 0);
          n = notready ? (__ffs(notready) - 1) : 32;
          if (n > win) n = win;
          if (n > 0) break;
        }
        __threadfence();
        if (tid == 0) sm.first_not_ready = n;
      }
      __syncthreads();
      const int nready = sm.first_not_ready;
      RS_STAMP(1)                                          // last time a batch of descendants was seen ready
      // ---- software pipeline over the ready updates, two stages deep: the record / row-map / column-map loads of
      //      update uu + 2 (static data) and the value loads of update uu + 1 are in flight while uu is multiplied.
      //      (Measured alternatives: parking the values in shared memory to prefetch two updates deep costs more
      //      shared-memory bandwidth than the latency it hides; cp.async is not usable because its 8-byte form allocates
      //      in L1, which may hold lines of panels that were not final when they were cached.)
      const int uend = u + nready;
      const double* Ld1 = nullptr; int K1 = 0, nrd1 = 0, half1 = 0, mi1 = -1, j1 = -1;      // stage A results (indices) of the next update
      int min_ = -1, halfn = 0;
      double xn[RS_NC], bn = 0.0;
      auto stage_a = [&](int uu) {
        const UpdRec rec = upd_rec[uu];
        Ld1 = s.L + rec.val_off; K1 = rec.K; nrd1 = rec.nrd; half1 = rec.pad[0];               // half: only target columns < 8 are touched
        mi1 = has_row ? (int)__ldg(umap + (int64_t)(uu - ubase) * nloc) : -1;                 // descendant row (from row a) landing on this thread's row
        j1 = colinv[(int64_t)uu * RS_NC + (tid % RS_NC)];                                     // descendant row (from a) holding target column tid % 16, or -1
      };
      auto stage_b = [&]() {
        const int k = tid / RS_NC;                        // RS_T == RS_NC * RS_NC
#pragma unroll
        for (int q = 0; q < RS_NC; ++q) xn[q] = (mi1 >= 0 && q < K1) ? __ldcg(&Ld1[mi1 + (int64_t)q * nrd1]) : 0.0;
        bn = (j1 >= 0 && k < K1) ? __ldcg(&Ld1[j1 + (int64_t)k * nrd1]) : 0.0;
        min_ = mi1; halfn = half1;
      };
      stage_a(u);
      stage_b();
      if (u + 1 < uend) stage_a(u + 1);
      for (int uu = u; uu < uend; ++uu, buf ^= 1) {
        const int mi = min_, half = halfn;
        double x[RS_NC];
#pragma unroll
        for (int k = 0; k < RS_NC; ++k) x[k] = xn[k];
        sm.Bs[buf][tid] = bn;                             // last read two updates ago: every thread is past that barrier
        __syncthreads();
        if (uu + 1 < uend) stage_b();
        if (uu + 2 < uend) stage_a(uu + 2);
        if (mi >= 0) {
          const double* Bt = sm.Bs[buf];
          if (half) {
#pragma unroll
            for (int k = 0; k < RS_NC; ++k) {
              const double xk = -x[k];
              const double2* brow = reinterpret_cast<const double2*>(Bt + k * RS_NC);
#pragma unroll
              for (int jp = 0; jp < RS_NC / 4; ++jp) {
                const double2 bb = brow[jp];
                acc[2 * jp] = fma(xk, bb.x, acc[2 * jp]);
                acc[2 * jp + 1] = fma(xk, bb.y, acc[2 * jp + 1]);
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < RS_NC; ++k) {
              const double xk = -x[k];
              const double2* brow = reinterpret_cast<const double2*>(Bt + k * RS_NC);
#pragma unroll
              for (int jp = 0; jp < RS_NC / 2; ++jp) {
                const double2 bb = brow[jp];
                acc[2 * jp] = fma(xk, bb.x, acc[2 * jp]);
                acc[2 * jp + 1] = fma(xk, bb.y, acc[2 * jp + 1]);
              }
            }
          }
        }
      }
      u += nready;
    }
    __syncthreads();
    RS_STAMP(2)

    // ---- diagonal block (every block of the supernode factors its own copy)
    if (tid < nc) {
#pragma unroll
      for (int c = 0; c < RS_NC; ++c) if (c <= tid) sm.Ds[tid * RS_DP + c] = acc[c];
    }
    __syncthreads();
    if (tid < 32) rs_potrf_warp(sm, nc, tid, status);
    __syncthreads();
    RS_STAMP(3)
    // ---- solve the own row against the diagonal factor in registers and store it (coalesced per column)
    if (has_row && tid >= nc) {
#pragma unroll
      for (int c = 0; c < RS_NC; ++c) {
        if (c < nc) {
          double v = acc[c];
#pragma unroll
          for (int k = 0; k < c; ++k) v = fma(-acc[k], sm.Ds[c * RS_DP + k], v);
          acc[c] = v * sm.dinv[c];
          Lp[prow + (int64_t)c * nr] = acc[c];
        }
      }
    }
    RS_STAMP(4)
    __threadfence();
    __syncthreads();
    if (tid == 0) sm.slot = atomicAdd(&arrived[sn], 1);
    __syncthreads();
    const bool last = sm.slot == nblk - 1;
    __syncthreads();
    RS_STAMP(5)
    if (last) {
      // every block of this supernode has read the assembled diagonal block and stored its rows
      for (int i = tid; i < nc * nc; i += RS_T) {
        const int r = i % nc, c = i / nc;
        if (c <= r) Lp[r + (int64_t)c * nr] = sm.Ds[r * RS_DP + c];
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(done + sn), "r"(1) : "memory");
      RS_STAMP(6)
    }
  }
}

bool chol_rs_supported(const fg_ctx* c) {
  const char* off = getenv("FG_CHOL_RS");
  if (off && off[0] == '0') return false;
  const char* gen = getenv("FG_CHOL_GENERIC");
  if (gen && gen[0] == '1') return false;
  return c->sym.rs_ok;
}

void launch_factor_rs(fg_ctx* c) {
  DevGraph& d = c->d;
  const Symbolic& S = c->sym;
  SysView s;
  s.L = d.L; s.col2sn = d.col2sn; s.sn_col0 = d.sn_col0; s.sn_ncols = d.sn_ncols; s.sn_nrows = d.sn_nrows;
  s.sn_rowptr = d.sn_rowptr; s.sn_valptr = d.sn_valptr; s.rowidx = d.rowidx; s.n_r = S.n_r;
  static int per_sm = 0;
  if (!per_sm) {
    cudaFuncSetAttribute(k_chol_rs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chol_rs, RS_T, sizeof(RsSmem));
    if (per_sm < 1) per_sm = 1;
  }
  cudaStream_t st = c->stream;
  c->epoch += 1;                                   // k_backsolve's flags are epoch stamped
  cudaMemsetAsync(d.status, 0, sizeof(int), st);
  cudaMemsetAsync(d.rs_done, 0, sizeof(int) * 2 * S.n_sn, st);     // done flags, then arrival counters
  cudaMemsetAsync(d.counters, 0, sizeof(int) * 4, st);
  const int na = S.rs_units_a, nc = (int)S.rs_units.size() - S.rs_units_a;
  // FG_CHOL_TRACE=<file>: per-unit %globaltimer stamps of the 3rd factorisation (dev tool, profiles/tools/chol_trace.py)
  static int n_calls = 0;
  long long* dbg = nullptr;
  const char* trace = getenv("FG_CHOL_TRACE");
  const size_t n_all = S.rs_units.size();
  if (trace && ++n_calls == 3) { cudaMalloc((void**)&dbg, sizeof(long long) * 8 * n_all); cudaMemset(dbg, 0, sizeof(long long) * 8 * n_all); }
  struct TraceDump {
    long long* dbg; const char* path; const Symbolic& S; cudaStream_t st;
    ~TraceDump() {
      if (!dbg) return;
      std::vector<long long> h(8 * S.rs_units.size());
      cudaStreamSynchronize(st);
      cudaMemcpy(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
      cudaFree(dbg);
      FILE* f = fopen(path, "w");
      if (!f) return;
      for (size_t i = 0; i < S.rs_units.size(); ++i) {
        const int4 u = S.rs_units[i];
        fprintf(f, "%zu %d %d %d %d %d %d %d", i, (int)(i >= (size_t)S.rs_units_a), u.x, u.y, u.z, u.w, S.sn_ncols[u.x], S.sn_leaf[u.x]);
        for (int k = 0; k < 7; ++k) fprintf(f, " %lld", h[8 * i + k]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  } dump{dbg, trace, S, c->stream};
  FrontView none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const int cap = c->num_sms * per_sm;
  if (!S.use_fronts) {
    k_chol_rs<<<std::min(cap, na), RS_T, sizeof(RsSmem), st>>>(s, d.rs_units, d.rs_moff, d.rs_map, d.upd_ptr, d.upd_d, d.upd_rec, d.rs_colinv,
                                                                d.rs_done + S.n_sn, d.rs_done, d.counters, na, d.status, none, dbg);
    return;
  }
  // phase A: the leaves; phase B: one dense update matrix per leaf; phase C: the separators
  if (na) k_chol_rs<<<std::min(cap, na), RS_T, sizeof(RsSmem), st>>>(s, d.rs_units, d.rs_moff, d.rs_map, d.updr_ptr, d.updr_d, d.updr_rec, d.rs_colinv,
                                                                      d.rs_done + S.n_sn, d.rs_done, d.counters, na, d.status, none, dbg);
  launch_front_syrk(c);
  FrontView fv = {d.tf_ptr, d.tf_leaf, d.fr_rowptr, d.fr_rows, d.fr_uptr, d.U};
  if (nc) k_chol_rs<<<std::min(cap, nc), RS_T, sizeof(RsSmem), st>>>(s, d.rs_units + na, d.rs_moff + na, d.rs_map, d.updr_ptr, d.updr_d,
                                                                      d.updr_rec, d.rs_colinv, d.rs_done + S.n_sn, d.rs_done, d.counters + 2, nc, d.status, fv, dbg ? dbg + 8 * (size_t)na : nullptr);
}

}  // namespace fg
