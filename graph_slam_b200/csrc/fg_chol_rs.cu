// fg_chol_rs.cu -- K7, row-split supernodal Cholesky (fp64): the default factorisation kernel.
//
// Persistent left-looking supernodal Cholesky whose unit of work is a ROW BLOCK of a supernode.  A nested-dissection
// ordering of a 50-frame covisibility band only has 32 independent leaf chains at C5 (fg_symbolic.cpp); with one CTA
// per supernode 32 of the 148 SMs carried three quarters of the flops.  Here a supernode of nr rows x nc columns
// (nc <= 32: two [X V B] frames or five poses) is cut into blocks of <= RS_RB below-diagonal rows.  A unit (supernode, block)
//   * is output stationary: the unit's (<= 256 rows) x (<= 32 columns) values live in registers from the first load
//     to the final store -- no shared-memory panel, no write conflicts,
//   * pulls every descendant update restricted to those rows: a host-built map (int16 per unit, update and row) names
//     the descendant row that lands on each local row; the rank-K update  P -= X B^T  (X: gathered descendant rows,
//     B: the descendant's rows in the target's columns, scattered to target columns) is a dense contraction and runs
//     on the fp64 tensor cores (DMMA m8n8k4); a descendant wider than 16 columns arrives as two column slices (the host
//     lists them as two updates, fg_symbolic.cpp); X and B arrive by cp.async through a 3-stage shared-memory ring, two
//     updates ahead of the multiplication; the part that lands on the diagonal block is recomputed by every block of the
//     supernode (nc x nc x K flops); 8-column groups of the target that the descendant does not reach are skipped,
//   * the first unit of a supernode owns the diagonal block: it factors it (one warp, matrix in registers) and stores it;
//     the units of below-diagonal rows finish their updates, wait for that unit's flag, solve their rows against the
//     factor and store them,
//   * publishes its own done flag (release); a consumer polls the flags of all units of a descendant (acquire).
// Deterministic: every panel entry is owned by one unit and updated in list order.  tcgen05 has no fp64 kind; DMMA
// measured at the DFMA peak on this part (profiles/tools/fp64_peak.cu): its gain is instruction and operand traffic.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fg_internal.h"

namespace fg {

#define RS_T 128                       // threads per unit: one panel row each
#define RS_NC 32                       // columns of a target supernode (fg_symbolic.cpp: kMaxSnCols)
#define RS_KC 16                       // columns of one update step (fg_symbolic.cpp: kUpdK)
#define RS_NT (RS_NC / 8)              // 8-column MMA tiles of the target
#define RS_DP 34                       // column stride of the diagonal factor in shared memory (even: 16-byte column starts)
#define RS_SP 33                       // row stride of the layout-conversion slabs

__device__ __forceinline__ int rs_ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

#define RS_ST 3                         // cp.async stages of the update pipeline
#define RS_XLD (RS_T + 4)               // leading dimension of a stage's [k][row] tile (doubles): 4 (mod 16), conflict-free MMA fragment loads
#define RS_BS 36                        // same for the [k][c] tile of the descendant's rows in the target's columns
struct RsSmem {
  double Xs[RS_ST][RS_KC * RS_XLD];     // [stage][k][row]: the descendant row that lands on each local row (gathered by the host row map)
  double Bs[RS_ST][RS_KC * RS_BS];      // [stage][k][c]: descendant rows that fall in the target's columns, scattered to TARGET columns
  double Dt[RS_NC * RS_DP];              // the diagonal factor, transposed: Dt[c * RS_DP + r] = L[r][c], r >= c (a column of L is contiguous)
  double dinv[RS_NC];
  int colidx[RS_NC];
  int slot, first_not_ready;
};
// aliases inside Xs, used outside the update pipeline: the row list of a leaf front (prologue) and the per-warp slabs
// (32 rows x 33) that convert between the thread-per-row layout and the MMA fragment layout
static_assert(sizeof(double) * RS_ST * RS_KC * RS_XLD >= sizeof(int) * 1024 + sizeof(double) * RS_T * RS_SP, "aliases must fit");

// D (8x8) += A (8x4, row major) * B (4x8, column major), fp64 tensor-core path (see fg_front.cu: dmma884)
__device__ __forceinline__ void rs_dmma(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
// 8-byte asynchronous global -> shared copy.  The 8-byte form exists only as .ca (allocates in L1), and L1 is not coherent:
// it is safe here because (i) panels start on 128-byte lines (fg_symbolic.cpp pads sn_valptr), so a line never mixes two
// supernodes, (ii) a supernode's lines are only ever read through L1 after its done flags, when they are final, and
// (iii) the earlier reads -- a unit loading its own assembled rows, the diagonal factor of its own supernode -- bypass L1 (ld.cg).
__device__ __forceinline__ void rs_cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}

// Right-looking Cholesky of the first 32 rows of a head unit by its first warp, in the registers that already hold them: lane r
// has panel row r.  Rows [0, nc) are the diagonal block; the lanes beyond it hold the first rows below the block, and the same
// recurrence (scale by 1 / L_cc, subtract the scaled column times L_jc) is their triangular solve.  The pivot of column c + 1 is
// updated, broadcast and its reciprocal square root started BEFORE the rest of column c's update is issued, so that the
// rsqrt / shuffle latency chain (the critical path of the whole factorisation) hides behind those independent updates.
// Entries above the diagonal hold garbage that never feeds a used value.  Returns 1 / L[lane][lane].
__device__ __forceinline__ double rs_potrf_rows(double (&a)[RS_NC], int nc, int lane, bool& bad) {
  const unsigned FULL = 0xffffffffu;
  // every index and every trip count below is static: the columns beyond nc run too, with pivot 1, on entries that are never
  // stored and never feed a column below nc
  double myinv = 1.0;
  double d = __shfl_sync(FULL, a[0], 0);
  if (!(d > 0.0)) { bad = true; d = 1.0; }          // not positive definite (or NaN): flag and keep going with a safe pivot
  double inv = rsqrt(d);
#pragma unroll
  for (int c = 0; c < RS_NC; ++c) {
    const double l = (lane == c) ? d * inv : a[c] * inv;
    a[c] = l;
    if (lane == c) myinv = inv;
    if (c + 1 < RS_NC) {
      const double l1 = __shfl_sync(FULL, l, (c + 1) & 31);
      a[(c + 1) & 31] = fma(-l, l1, a[(c + 1) & 31]);
      d = __shfl_sync(FULL, a[(c + 1) & 31], (c + 1) & 31);
      if (c + 1 >= nc) d = 1.0;
      else if (!(d > 0.0)) { bad = true; d = 1.0; }
      inv = rsqrt(d);
    }
#pragma unroll
    for (int j = 0; j < RS_NC; ++j) {
      if (j >= c + 2) {
        const double lj = __shfl_sync(FULL, l, j);
        a[j] = fma(-l, lj, a[j]);
      }
    }
  }
  return myinv;
}

__global__ void __launch_bounds__(RS_T, 3)
k_chol_rs(SysView s, const int4* __restrict__ units, const int64_t* __restrict__ unit_moff, const short* __restrict__ rowmap,
          const int* __restrict__ upd_ptr, const int* __restrict__ upd_d, const UpdRec* __restrict__ upd_rec,
          const signed char* __restrict__ colinv, const int2* __restrict__ sn_units, int* done, int unit_base,
          int* counter, int n_units, int* status, FrontView fv, long long* dbg, int* diag_done) {
  extern __shared__ __align__(16) unsigned char rs_raw[];
  RsSmem& sm = *reinterpret_cast<RsSmem*>(rs_raw);
  const int tid = threadIdx.x;
  int* const rl_s = reinterpret_cast<int*>(&sm.Xs[0][0]);                          // alias (prologue only)
  double* const slab_s = reinterpret_cast<double*>(rl_s + 1024);                  // alias (layout conversion only)

  while (true) {
    if (tid == 0) sm.slot = atomicAdd(counter, 1);
    __syncthreads();
    const int slot = sm.slot;
    __syncthreads();
    if (slot >= n_units) break;
// (the __syncwarp matters: a warp left diverged by the one-lane branch would take the slow divergent path of every following shuffle)
#define RS_STAMP(k) if (dbg) { if (tid == 0) { long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); dbg[8 * slot + (k)] = t_; } __syncwarp(); }
    RS_STAMP(0)
    const int uidx = slot;
    const int4 un = units[uidx];
    const int sn = un.x, r0 = un.y, r1 = un.z;               // rows [r0, r1) of the panel; r0 == 0: the head unit (diagonal block [0, nc) + rows below it)
    const bool is_diag = (r0 == 0);
    const int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
    const int nloc = r1 - r0;
    double* Lp = s.L + s.sn_valptr[sn];
    // thread tid owns panel row r0 + tid; its nc values live in registers from here to the final store
    const bool has_row = tid < nloc;
    const int prow = r0 + tid;
    double acc[RS_NC];
#pragma unroll
    for (int c = 0; c < RS_NC; ++c) acc[c] = (has_row && c < nc) ? __ldcg(&Lp[prow + (int64_t)c * nr]) : 0.0;
    // ---- subtract the dense leaf fronts that reach this supernode (fg_front.cu)
    if (fv.tf_ptr) {
      const int g = has_row ? s.rowidx[s.sn_rowptr[sn] + prow] : -1;
      for (int e = fv.tf_ptr[sn]; e < fv.tf_ptr[sn + 1]; ++e) {
        const int l = fv.tf_leaf[e];
        const int* Rl = fv.fr_rows + fv.fr_rowptr[l];
        const int nR = fv.fr_rowptr[l + 1] - fv.fr_rowptr[l];
        const double* Ul = fv.U + fv.fr_uptr[l];
        __syncthreads();
        for (int i = tid; i < nR; i += RS_T) rl_s[i] = Rl[i];
        __syncthreads();
        if (tid < RS_NC) {
          const int gc = c0 + tid;
          int lo = 0, hi = nR - 1;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (rl_s[mid] < gc) lo = mid + 1; else hi = mid; }
          sm.colidx[tid] = (tid < nc && nR > 0 && rl_s[lo] == gc) ? lo : -1;
        }
        __syncthreads();
        if (has_row) {
          int lo = 0, hi = nR - 1;
          while (lo < hi) { const int mid = (lo + hi) >> 1; if (rl_s[mid] < g) lo = mid + 1; else hi = mid; }
          if (nR > 0 && rl_s[lo] == g) {
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

    // ---- the update phase runs on the fp64 tensor cores (DMMA m8n8k4): warp w owns local rows [32 w, 32 w + 32) as four
    //      8-row tiles and the 32 target columns as four 8-column tiles; lane = 4 g + t holds the sums
    //      (8 m + g, 8 n + 2 t + {0, 1}).  The thread-per-row registers are converted through the warp's slab.
    const unsigned FULL = 0xffffffffu;
    const int lane = tid & 31, wrp = tid >> 5, fgi = lane >> 2, fti = lane & 3;
    double* slab = slab_s + wrp * 32 * RS_SP;
    double fr[4][RS_NT][2];
    __syncthreads();                                      // the front prologue is done with its alias
#pragma unroll
    for (int c = 0; c < RS_NC; ++c) slab[lane * RS_SP + c] = acc[c];
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int n = 0; n < RS_NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) fr[m][n][e] = slab[(8 * m + fgi) * RS_SP + 8 * n + 2 * fti + e];
    const short* umap = rowmap + unit_moff[uidx] + tid;
    const int bk = tid / RS_NC, bc = tid % RS_NC;          // this thread's elements of the B tile: (bk + 4 h, bc), h = 0..3

    int u = upd_ptr[sn];
    const int ubase = u;
    const int u1 = upd_ptr[sn + 1];
    const double* Ld1 = nullptr; int K1 = 0, nrd1 = 0, mask1 = 0, mi1 = -1, j1 = -1;        // indices of the next update to be issued
    auto stage_a = [&](int uu) {
      const UpdRec rec = upd_rec[uu];
      Ld1 = s.L + rec.val_off; K1 = rec.K; nrd1 = rec.nrd; mask1 = rec.pad[0];                 // mask: 8-column groups of the target that are touched
      mi1 = has_row ? (int)__ldg(umap + (int64_t)(uu - ubase) * nloc) : -1;                   // descendant row (from row a) landing on this thread's row
      j1 = colinv[(int64_t)uu * RS_NC + bc];                                                  // descendant row (from a) holding target column bc, or -1
    };
    while (u < u1) {
      stage_a(u);                                          // static data: its latency hides behind the wait for the descendant
      // warp 0 spins on the done flags of the next (up to 32) descendants and publishes the ready prefix
      if (tid < 32) {
        const int win = min(32, u1 - u);
        const int ui = u + min(tid, win - 1);
        const int2 du = sn_units[upd_d[ui]];               // (first unit, number of units) of the descendant
        const int* fp = done + du.x;
        int n;
        while (true) {
          int f = 1;
          if (tid < win) for (int b = 0; b < du.y; ++b) f &= rs_ld_relaxed(fp + b);     // every unit of the descendant is stored
          const unsigned notready = ~__ballot_sync(0xffffffffu, f != 0);
          n = notready ? (__ffs(notready) - 1) : 32;
          if (n > win) n = win;
          if (n > 0) break;
        }
        __threadfence();
        if (tid == 0) sm.first_not_ready = n;
      }
      __syncthreads();                                     // also: every product of the previous batch is done with the stages
      const int nready = sm.first_not_ready;
      RS_STAMP(1)                                          // last time a batch of descendants was seen ready
      // ---- pipeline over the ready updates: the descendant values of update uu + 2 travel by cp.async into a ring of
      //      RS_ST shared-memory stages while uu is multiplied; the record / row-map / column-map loads (static data) run
      //      one step further ahead.
      const int uend = u + nready;
      // issue the copies of the update described by stage_a into stage st; returns (any row of this warp touched) | mask << 1
      auto stage_b = [&](int st) -> int {
        const int any = __any_sync(FULL, mi1 >= 0) ? 1 : 0;
        if (any) {
          double* xs = sm.Xs[st] + tid;
#pragma unroll
          for (int q = 0; q < RS_KC; ++q) {
            if (mi1 >= 0 && q < K1) rs_cp_async8(xs + q * RS_XLD, &Ld1[mi1 + (int64_t)q * nrd1]);
            else xs[q * RS_XLD] = 0.0;
          }
        }
#pragma unroll
        for (int h = 0; h < RS_KC * RS_NC / RS_T; ++h) {
          const int k = bk + (RS_T / RS_NC) * h;
          if (j1 >= 0 && k < K1) rs_cp_async8(&sm.Bs[st][k * RS_BS + bc], &Ld1[j1 + (int64_t)k * nrd1]);
          else sm.Bs[st][k * RS_BS + bc] = 0.0;
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        return any | (mask1 << 1);
      };
      int meta[RS_ST] = {0, 0, 0};                         // per stage: stage_b's return value
      meta[0] = stage_b(0);
      if (u + 1 < uend) { stage_a(u + 1); meta[1] = stage_b(1); } else asm volatile("cp.async.commit_group;\n" ::: "memory");
      if (u + 2 < uend) stage_a(u + 2);
      int st = 0;
      for (int uu = u; uu < uend; ++uu) {
        asm volatile("cp.async.wait_group 1;\n" ::: "memory");       // this thread's copies of update uu have landed (uu + 1 may be pending)
        __syncthreads();                                               // ... and everybody's; stage st + 2 (read by the previous product) is free
        const int st2 = (st + 2 >= RS_ST) ? st + 2 - RS_ST : st + 2;
        const int mt = (st == 0) ? meta[0] : (st == 1 ? meta[1] : meta[2]);
        if (uu + 2 < uend) {
          const int r = stage_b(st2);
          if (st2 == 0) meta[0] = r; else if (st2 == 1) meta[1] = r; else meta[2] = r;
          if (uu + 3 < uend) stage_a(uu + 3);
        } else {
          asm volatile("cp.async.commit_group;\n" ::: "memory");     // keeps the group count in step
        }
        if (mt & 1) {
          const double* Xt = sm.Xs[st] + 32 * wrp + fgi;
          const double* Bt = sm.Bs[st] + fgi;
#pragma unroll
          for (int ks = 0; ks < RS_KC / 4; ++ks) {
            const int kk = 4 * ks + fti;
            double a[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) a[m] = Xt[kk * RS_XLD + 8 * m];
#pragma unroll
            for (int n = 0; n < RS_NT; ++n) {
              if (mt & (2 << n)) {                        // warp uniform
                const double bn = -Bt[kk * RS_BS + 8 * n];
#pragma unroll
                for (int m = 0; m < 4; ++m) rs_dmma(fr[m][n], a[m], bn);
              }
            }
          }
        }
        st = (st + 1 == RS_ST) ? 0 : st + 1;
      }
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      u += nready;
    }
    __syncthreads();                                      // the stages are free: back to thread-per-row registers through the slabs
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int n = 0; n < RS_NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) slab[(8 * m + fgi) * RS_SP + 8 * n + 2 * fti + e] = fr[m][n][e];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < RS_NC; ++c) acc[c] = slab[lane * RS_SP + c];
    __syncthreads();
    RS_STAMP(2)

    if (is_diag) {
      // ---- the head unit: its first warp factors the diagonal block (and solves the rows it holds beyond it) in registers, hands
      //      the factor to the other warps through shared memory, stores its rows and raises the supernode's diagonal flag -- the
      //      further row blocks of this supernode wait for it
      if (wrp == 0) {
        bool bad = false;
        const double myinv = rs_potrf_rows(acc, nc, lane, bad);
        if (bad) atomicExch(status, 1);
#pragma unroll
        for (int c = 0; c < RS_NC; ++c) sm.Dt[c * RS_DP + lane] = (lane < nc && c <= lane) ? acc[c] : 0.0;      // padded: no bounds in the solve
        sm.dinv[lane] = myinv;
      }
      __syncthreads();
      RS_STAMP(3)
      if (wrp == 0) {
        if (has_row) {
#pragma unroll
          for (int c = 0; c < RS_NC; ++c) if (c < nc && (prow >= nc || c <= prow)) Lp[prow + (int64_t)c * nr] = acc[c];
        }
        if (r1 < nr) {                                   // somebody is waiting
          __threadfence();
          __syncwarp();
          if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(diag_done + sn), "r"(1) : "memory");
        }
      }
    } else {
      // ---- further rows: wait for the diagonal factor of this supernode
      if (tid == 0) {
        const int* fp = diag_done + sn;
        while (rs_ld_relaxed(fp) == 0) { }
        __threadfence();
      }
      __syncthreads();
      for (int i = tid; i < RS_NC * RS_NC; i += RS_T) {
        const int r = i % RS_NC, c = i / RS_NC;
        sm.Dt[c * RS_DP + r] = (r < nc && c <= r) ? __ldcg(&Lp[r + (int64_t)c * nr]) : 0.0;       // padded: no bounds in the solve
      }
      __syncthreads();
      if (tid < RS_NC) sm.dinv[tid] = tid < nc ? 1.0 / sm.Dt[tid * RS_DP + tid] : 1.0;
      __syncthreads();
      RS_STAMP(3)
    }
    // ---- rows below the diagonal block: right-looking triangular solve of the own row in registers (the updates of one column are
    //      independent: the dependency chain is one multiply and one FMA per column), then the store (coalesced per column)
    if (has_row && !(is_diag && wrp == 0)) {
#pragma unroll
      for (int c = 0; c < RS_NC; ++c) {
        const double x = acc[c] * sm.dinv[c];
        acc[c] = x;
        if (c < nc) Lp[prow + (int64_t)c * nr] = x;
        const double* dc = &sm.Dt[c * RS_DP];
#pragma unroll
        for (int jj = 0; jj < RS_NC / 2; ++jj) {          // columns (j, j + 1) of the row, one 16-byte read of the factor's column c
          const int j = 2 * jj;
          if (j > c) {
            const double2 dd = *reinterpret_cast<const double2*>(dc + j);
            acc[j] = fma(-x, dd.x, acc[j]);
            acc[j + 1] = fma(-x, dd.y, acc[j + 1]);
          } else if (j + 1 > c) {
            acc[j + 1] = fma(-x, dc[j + 1], acc[j + 1]);
          }
        }
      }
    }
    RS_STAMP(4)
    __threadfence();
    __syncthreads();
    if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(done + unit_base + uidx), "r"(1) : "memory");
    RS_STAMP(6)
  }
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
  cudaMemsetAsync(d.rs_done, 0, sizeof(int) * (S.rs_units.size() + S.n_sn), st);     // one done flag per unit, then one diagonal flag per supernode
  int* const diag_done = d.rs_done + S.rs_units.size();
  cudaMemsetAsync(d.counters, 0, sizeof(int) * 4, st);
  const int na = S.rs_units_a, nc = (int)S.rs_units.size() - S.rs_units_a;
  // FG_CHOL_TRACE=<file>: per-unit %globaltimer stamps of the 3rd (FG_CHOL_TRACE_CALL-th) factorisation (dev tool, profiles/tools/chol_trace.py)
  static int n_calls = 0;
  long long* dbg = nullptr;
  const char* trace = getenv("FG_CHOL_TRACE");
  const size_t n_all = S.rs_units.size();
  static const int trace_call = getenv("FG_CHOL_TRACE_CALL") ? atoi(getenv("FG_CHOL_TRACE_CALL")) : 3;
  if (trace && ++n_calls == trace_call) { cudaMalloc((void**)&dbg, sizeof(long long) * 8 * n_all); cudaMemset(dbg, 0, sizeof(long long) * 8 * n_all); }
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
    k_chol_rs<<<std::min(cap, na), RS_T, sizeof(RsSmem), FGS(st)>>>(s, d.rs_units, d.rs_moff, d.rs_map, d.rsu_ptr, d.rsu_d, d.rsu_rec, d.rs_colinv,
                                                                d.rs_sn_units, d.rs_done, 0, d.counters, na, d.status, none, dbg, diag_done);
    return;
  }
  // phase A: the leaves; phase B: one dense update matrix per leaf; phase C: the separators
  if (na) k_chol_rs<<<std::min(cap, na), RS_T, sizeof(RsSmem), FGS(st)>>>(s, d.rs_units, d.rs_moff, d.rs_map, d.rsu_ptr, d.rsu_d, d.rsu_rec, d.rs_colinv,
                                                                      d.rs_sn_units, d.rs_done, 0, d.counters, na, d.status, none, dbg, diag_done);
  launch_front_syrk(c);
  FrontView fv = {d.tf_ptr, d.tf_leaf, d.fr_rowptr, d.fr_rows, d.fr_uptr, d.U};
  if (nc) k_chol_rs<<<std::min(cap, nc), RS_T, sizeof(RsSmem), FGS(st)>>>(s, d.rs_units + na, d.rs_moff + na, d.rs_map, d.rsu_ptr, d.rsu_d,
                                                                      d.rsu_rec, d.rs_colinv, d.rs_sn_units, d.rs_done, na, d.counters + 2, nc, d.status, fv, dbg ? dbg + 8 * (size_t)na : nullptr, diag_done);
}

}  // namespace fg
