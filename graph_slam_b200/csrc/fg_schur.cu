// fg_schur.cu -- K6 (second half): landmark Schur complement onto the reduced pose system (fp64).
//
//   S_pq -= sum_l W_pl (V_l + lambda I)^-1 W_ql^T ,   rhs_p += sum_l W_pl (V_l + lambda I)^-1 g_l
//
// for the GenericProjectionFactor / PriorFactor<Point3> part of the graph CGraphGT::addToGTSAM builds
// (gtsam/gtsam_graph.cpp:370-448); in GTSAM this is the elimination of every Point3 before the poses.
//
// Symmetric form.  With V_l + lambda I = G G^T (3x3 Cholesky) and C = G^-T,  W Vinv W^T = (W C)(W C)^T, so one array
// Z_o = W_o C_l (6x3 per observation, 144 B) serves both sides of the product.  k_zmat writes Z in POSE-major order (the
// observations of one pose are contiguous and sorted by landmark) as three planes, one per column of the 6 x 3 record.
//
// k_schur_tiles: output-stationary, deterministic, no atomics, no pair lists.  One CTA owns a 16 x 16 tile of 6x6
// blocks (16 consecutive row poses x 16 consecutive column poses); ONE LANE owns one block and keeps its 36 sums in
// registers.  The landmarks are walked in super-chunks of 96 consecutive ids (three 32-landmark table words).  Per
// super-chunk every lane intersects the landmark bit masks of its two poses once and lists its hits; then, one column k of
// Z at a time, the CTA stages that column of the records its 32 poses have in common with the other side of the tile
// (48 B per record) and every lane adds z_p^k (z_q^k)^T for its hits (36 DFMA per hit and column, operands by 16-byte
// shared-memory reads).  Shared-memory layout [slot][pose][48 B]: the 8 lanes of a quarter-warp sit on a wrapped diagonal
// of an 8 x 8 sub-tile (distinct row pose mod 8, distinct column pose mod 8), so every operand read is bank-conflict free
// whatever records the lanes are at.  HBM/L2 traffic is one record per (pose, landmark, tile) instead of two per pair.
#include "fg_internal.h"

namespace fg {

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// per landmark: Vinv = (V + lambda I)^-1 (6 upper) for the back-substitution; C = G^-T and u = G^-1 g_l for the
// symmetric Schur product
__global__ void k_vinv(int64_t L, const double* __restrict__ V, const double* __restrict__ gl, double lambda,
                       double* __restrict__ Vinv, double* __restrict__ Cf, double* __restrict__ ul) {
  int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (l >= L) return;
  const double a00 = V[6 * l] + lambda, a01 = V[6 * l + 1], a02 = V[6 * l + 2], a11 = V[6 * l + 3] + lambda, a12 = V[6 * l + 4],
               a22 = V[6 * l + 5] + lambda;
  double A[9] = {a00, a01, a02, a01, a11, a12, a02, a12, a22};
  double Ai[9];
  inv3(A, Ai);
  Vinv[6 * l] = Ai[0]; Vinv[6 * l + 1] = Ai[1]; Vinv[6 * l + 2] = Ai[2];
  Vinv[6 * l + 3] = Ai[4]; Vinv[6 * l + 4] = Ai[5]; Vinv[6 * l + 5] = Ai[8];
  // G (lower) and its inverse
  const double g00 = sqrt(a00), i00 = 1.0 / g00;
  const double g10 = a01 * i00, g20 = a02 * i00;
  const double g11 = sqrt(a11 - g10 * g10), i11 = 1.0 / g11;
  const double g21 = (a12 - g20 * g10) * i11;
  const double g22 = sqrt(a22 - g20 * g20 - g21 * g21), i22 = 1.0 / g22;
  const double i10 = -g10 * i00 * i11;
  const double i21 = -g21 * i11 * i22;
  const double i20 = -(g20 * i00 + g21 * i10) * i22;
  Cf[6 * l] = i00; Cf[6 * l + 1] = i10; Cf[6 * l + 2] = i20; Cf[6 * l + 3] = i11; Cf[6 * l + 4] = i21; Cf[6 * l + 5] = i22;
  const double g0 = gl[3 * l], g1 = gl[3 * l + 1], g2 = gl[3 * l + 2];
  ul[3 * l] = i00 * g0; ul[3 * l + 1] = i10 * g0 + i11 * g1; ul[3 * l + 2] = i20 * g0 + i21 * g1 + i22 * g2;
}

#define REC 18            // doubles per W / Z record
#define RST 19            // padded record stride in shared memory (odd: conflict-free 8-byte lane accesses)

// per observation: Z_o = W_o C_l (6x3), written at the observation's pose-major position.  256 consecutive W records
// are moved through shared memory so that the global loads are coalesced; the stores go out as 9 x 16 B per record.
__global__ void __launch_bounds__(256) k_zmat(int64_t M, const int* __restrict__ obs_point, const int* __restrict__ obs_ppos,
                                              const double* __restrict__ W, const double* __restrict__ Cf, double* __restrict__ Zp) {
  __shared__ double buf[256 * RST];
  __shared__ int ppos[256];
  const int64_t o0 = blockIdx.x * (int64_t)256;
  const int n = (int)min((int64_t)256, M - o0);
  const int tid = threadIdx.x;
  // the dependent gather obs_point -> Cf (and the scatter position) is issued first so that it overlaps the streaming W loads
  double c[6] = {0, 0, 0, 0, 0, 0};
  if (tid < n) {
    const int l = obs_point[o0 + tid];
    ppos[tid] = obs_ppos[o0 + tid];
    const double2* p = reinterpret_cast<const double2*>(Cf + (int64_t)l * 6);
    double2 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
    c[0] = v0.x; c[1] = v0.y; c[2] = v1.x; c[3] = v1.y; c[4] = v2.x; c[5] = v2.y;
  }
  for (int i = tid; i < n * REC; i += 256) buf[(i / REC) * RST + (i % REC)] = __ldg(W + o0 * REC + i);
  __syncthreads();
  if (tid < n) {
    double* w = buf + tid * RST;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const double w0 = w[3 * i], w1 = w[3 * i + 1], w2 = w[3 * i + 2];
      w[3 * i + 0] = w0 * c[0];
      w[3 * i + 1] = w0 * c[1] + w1 * c[3];
      w[3 * i + 2] = w0 * c[2] + w1 * c[4] + w2 * c[5];
    }
  }
  __syncthreads();
  // three planes Zk[k][M][6]: column k of every record, 48 contiguous bytes per (record, plane)
  for (int i = tid; i < n * 9; i += 256) {
    const int r = i / 9, part = i - 9 * r, k = part / 3, pr = part - 3 * k;
    const double* s = buf + r * RST + 6 * pr + k;
    reinterpret_cast<double2*>(Zp + ((int64_t)k * M + ppos[r]) * 6)[pr] = make_double2(s[0], s[3]);
  }
}

// ------------------------------------------------------------------ tile kernel
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}

#define ST_THREADS 256
#define ST_NW 3            // 32-landmark table chunks per super-chunk (96 landmarks between two staging rounds)
#define ST_KMAX 56         // record slots per pose and super-chunk
#define ST_MAXH 32         // hits per lane and super-chunk the hit list has room for
struct SchurSmem {
  static constexpr int NC = 3 * ST_KMAX * 16;        // 16-byte cells per side
  double2 rows[NC];                                  // [slot][pose][element pair]: ONE column of Z (6 doubles, 48 contiguous bytes) per record;
  double2 cols[NC];                                  // cell index = 3 (16 slot + pose) + pair, so the bank group of a read is (3 pose + pair) mod 8
  unsigned short hits[ST_MAXH * ST_THREADS];         // per lane: (row slot | column slot << 8) of every hit of the super-chunk
  int src[ST_THREADS / 32][4 * ST_KMAX];             // per warp: pose-major record index of every record it stages
};

// Z is stored as three planes Zk[k][M][6] (column k of every 6 x 3 record; pose-major).  S_pq = sum_k sum_l z_pl^k (z_ql^k)^T, so
// the tile can be built one column at a time: a staging round holds 48 bytes per record instead of 144, which lets a round
// cover 96 landmarks in the shared memory that held 32 -- the lanes of a warp run until the busiest one is out of hits, and
// the spread of the hit counts shrinks with the length of the round.  Per super-chunk: the hit list of every lane is
// built once (bit-mask walk), then three rounds (stage column k, barrier, 36 DFMA per hit from two 3 x LDS.128 operands).
// A super-chunk that does not fit (a pose with more than ST_KMAX records, a lane with more than ST_MAXH hits) is redone as
// three single-word rounds.
__global__ void __launch_bounds__(ST_THREADS, 2)
k_schur_tiles(const int4* __restrict__ tiles, int P, const int* __restrict__ pc_lo, const int* __restrict__ pc_n,
              const int64_t* __restrict__ pc_ptr, const uint2* __restrict__ pc_ent, const double* __restrict__ Zk, int64_t M,
              const int* __restrict__ off_pose, SysView sys) {
  extern __shared__ __align__(16) unsigned char st_raw[];
  SchurSmem& sm = *reinterpret_cast<SchurSmem*>(st_raw);
  const unsigned FULL = 0xffffffffu;
  const int4 td = tiles[blockIdx.x];
  const int gi = td.x, gj = td.y, cb = td.z, ce = td.w;
  const bool diag = gi == gj;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // metadata role: lane s holds the table entries of pose s (0..15 row poses, 16..31 column poses)
  const int mp = ((lane >> 4) ? gj : gi) * 16 + (lane & 15);
  int m_lo = 0, m_n = 0;
  int64_t m_ptr = 0;
  if (mp < P) { m_lo = pc_lo[mp]; m_n = pc_n[mp]; m_ptr = pc_ptr[mp]; }
  // compute role: two warps share an 8 x 8 sub-tile; the 8 lanes of a quarter-warp sit on one wrapped diagonal of it (distinct
  // row pose mod 8, distinct column pose mod 8), which is what makes the 16-byte operand reads bank-conflict free.  A landmark's
  // window covers ~50 consecutive poses, so at any time a tile is only partly inside it: with 8 x 8 sub-tiles the warps of the
  // part outside run no steps at all (wrapped diagonals of the whole 16 x 16 tile put every warp half in, half out).
  // (warps 0-1 and 4-5 issue from schedulers 0-1, warps 2-3 and 6-7 from 2-3: sub-tiles that share a row or a column of the
  // tile -- the usual shape of the part inside the window -- sit on different schedulers)
  const int sub = (warp >> 1) < 2 ? (warp >> 1) : 5 - (warp >> 1), ii = lane & 7, dgn = 4 * (warp & 1) + (lane >> 3);
  const int i = 8 * (sub >> 1) + ii, j = 8 * (sub & 1) + ((ii + dgn) & 7);
  bool active = (gi * 16 + i < P) && (gj * 16 + j < P);
  // In a diagonal tile the lanes (i, j) and (j, i) hold the same unordered pose pair: they split its landmarks by bit parity
  // and their sums are joined at the end; the p == q blocks are left to k_schur_rhs.
  const bool primary = !diag || i > j;
  unsigned hit_sel = 0xffffffffu;
  if (diag) { active = active && i != j; hit_sel = primary ? 0x55555555u : 0xaaaaaaaau; }
  double acc[36];
#pragma unroll
  for (int q = 0; q < 36; ++q) acc[q] = 0.0;
  bool any = false;
  const double2* cbase = diag ? sm.rows : sm.cols;
  int* srcl = sm.src[warp];                         // (M < 2^31 / 6 is checked by build_schur_tables)
  unsigned short* hl = sm.hits + threadIdx.x;
  // staging role: warp w copies the records of poses npw w .. npw w + npw - 1 (npw = 4, or 2 in a diagonal tile); a copy step
  // moves the three 16-byte parts of NR consecutive records of each of these poses
  const int npw = diag ? 2 : 4;
  const int st_t = lane % npw, st_part = (lane / npw) % 3, st_r = lane / (3 * npw), st_nr = diag ? 5 : 2;
  const int st_s = npw * warp + st_t;
  double2* st_dst = (st_s >> 4 ? sm.cols : sm.rows) + 3 * (st_s & 15) + st_part;

  auto fetch = [&](int c) -> uint2 {
    const int r = c - m_lo;
    return (c < ce && r >= 0 && r < m_n) ? __ldg(&pc_ent[m_ptr + r]) : make_uint2(0u, 0u);
  };
  uint2 nxt[ST_NW];
#pragma unroll
  for (int wd = 0; wd < ST_NW; ++wd) nxt[wd] = fetch(cb + wd);
  int c0 = cb, single = -1;                         // single >= 0: the super-chunk at c0 is being redone word by word
  while (c0 < ce) {
    uint2 cur[ST_NW];
    if (single < 0) {
#pragma unroll
      for (int wd = 0; wd < ST_NW; ++wd) { cur[wd] = nxt[wd]; nxt[wd] = fetch(c0 + ST_NW + wd); }   // entries ride one super-chunk ahead
    } else {
      cur[0] = fetch(c0 + single);
#pragma unroll
      for (int wd = 1; wd < ST_NW; ++wd) cur[wd] = make_uint2(0u, 0u);
    }
    // landmarks seen from both sides of the tile; a pose stages only its records of those
    unsigned f[ST_NW];
    int tot = 0;
#pragma unroll
    for (int wd = 0; wd < ST_NW; ++wd) {
      unsigned any16 = cur[wd].y;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) any16 |= __shfl_xor_sync(FULL, any16, o);
      f[wd] = cur[wd].y & __shfl_xor_sync(FULL, any16, 16);
      tot += __popc(f[wd]);
    }
    bool redo = false;
    if (__any_sync(FULL, tot != 0)) {               // identical decision in every warp: all hold the same 32 entries
      redo = __any_sync(FULL, tot > ST_KMAX);
      int nh = 0;
      if (!redo) {
        // ---- list the records of this warp's poses (slot order = landmark order)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (t < npw) {
            const int s = npw * warp + t;
            int base = t * ST_KMAX;
#pragma unroll
            for (int wd = 0; wd < ST_NW; ++wd) {
              const unsigned fm = __shfl_sync(FULL, f[wd], s), mm = __shfl_sync(FULL, cur[wd].y, s);
              const int st = (int)__shfl_sync(FULL, cur[wd].x, s);
              if ((fm >> lane) & 1u) {
                const unsigned low = (1u << lane) - 1u;
                srcl[base + __popc(fm & low)] = 6 * (st + __popc(mm & low));      // element offset of the record in a plane
              }
              base += __popc(fm);
            }
          }
        }
        int my_cnt = __shfl_sync(FULL, tot, npw * warp + st_t);
        if (st_r >= st_nr) my_cnt = 0;
        unsigned fi[ST_NW], fj[ST_NW];
#pragma unroll
        for (int wd = 0; wd < ST_NW; ++wd) {
          fi[wd] = __shfl_sync(FULL, f[wd], i); fj[wd] = __shfl_sync(FULL, f[wd], 16 + j);
          if (active) nh += __popc(fi[wd] & fj[wd] & hit_sel);
        }
        // ---- hit list of this lane's block: (row slot, column slot) of every landmark both poses see
        if (active) {
            int pi = 0, pj = 0, n = 0;
#pragma unroll
            for (int wd = 0; wd < ST_NW; ++wd) {
              unsigned hit = fi[wd] & fj[wd] & hit_sel;
              while (hit) {
                const int b = __ffs(hit) - 1;
                hit &= hit - 1u;
                const unsigned low = (1u << b) - 1u;
                const int a = pi + __popc(fi[wd] & low), bq = pj + __popc(fj[wd] & low);
                if (n < ST_MAXH) hl[n * ST_THREADS] = (unsigned short)(a | (bq << 8));
                ++n;
              }
              pi += __popc(fi[wd]); pj += __popc(fj[wd]);
            }
        }
        __syncwarp();
        // ---- three rounds, one column of Z each
        auto issue = [&](int k, double2* to) {           // cp.async of column k of this warp's records, one commit group
          const double* zk = Zk + (int64_t)k * M * 6 + 2 * st_part;
          const int* sl = srcl + st_t * ST_KMAX;
          double2* dst = to + st_r * 48;
#pragma unroll 2
          for (int kk = st_r; kk < my_cnt; kk += st_nr, dst += st_nr * 48) cp_async16(dst, zk + sl[kk]);
          asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
        for (int k = 0; k < 3; ++k) {
          const double2* rbase = sm.rows;
          const double2* qbase = cbase;
          if (!diag) {
            // the previous round's products are done with the staging area (k == 0: also the vote on the hit-list overflow)
            if (k == 0) {
              if (__syncthreads_or(nh > ST_MAXH)) { redo = true; break; }
            } else {
              __syncthreads();
            }
            issue(k, st_dst);
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();
          } else {
            // a diagonal tile stages 16 poses, so the column area is a second buffer: column 1 is on its way while column 0 is
            // multiplied, column 2 (back in the row area) while column 1 is
            if (k == 0) {
              if (__syncthreads_or(nh > ST_MAXH)) { redo = true; break; }
              issue(0, st_dst);
              issue(1, st_dst + SchurSmem::NC);
              asm volatile("cp.async.wait_group 1;\n" ::: "memory");
            } else if (k == 1) {
              __syncthreads();                       // column 0's products are done with the row area
              issue(2, st_dst);
              asm volatile("cp.async.wait_group 1;\n" ::: "memory");
              rbase = qbase = sm.cols;
            } else {
              asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            }
            __syncthreads();
          }
          for (int h = 0; h < nh; ++h) {
            const unsigned pk = hl[h * ST_THREADS];
            const double2* ra = rbase + ((pk & 0xffu) * 16 + i) * 3;
            const double2* cq = qbase + ((pk >> 8) * 16 + j) * 3;
            const double2 y01 = ra[0], y23 = ra[1], y45 = ra[2];
            const double2 w01 = cq[0], w23 = cq[1], w45 = cq[2];
            const double y[6] = {y01.x, y01.y, y23.x, y23.y, y45.x, y45.y};
            const double w[6] = {w01.x, w01.y, w23.x, w23.y, w45.x, w45.y};
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
              for (int cc = 0; cc < 6; ++cc) acc[6 * r + cc] = fma(y[r], w[cc], acc[6 * r + cc]);
          }
        }
        if (!redo && nh) any = true;
      }
    }
    // ---- advance
    if (redo) {
      if (single < 0) { single = 0; continue; }     // redo this super-chunk one word (<= 32 records, <= 32 hits) at a time
      // a single word always fits: not reached
    }
    if (single >= 0) {
      if (++single == ST_NW) { single = -1; c0 += ST_NW; }
    } else {
      c0 += ST_NW;
    }
  }
  if (diag) {
    // join the two halves of every pose pair: the secondary lane (j, i) parks its sums, transposed, in the staging area
    __syncthreads();
    double* xch = reinterpret_cast<double*>(sm.rows) + (primary ? i * 16 + j : j * 16 + i) * 37;
    if (active && !primary) {
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) xch[6 * r + cc] = acc[6 * cc + r];
      xch[36] = any ? 1.0 : 0.0;
    }
    __syncthreads();
    if (active && primary) {
#pragma unroll
      for (int q = 0; q < 36; ++q) acc[q] += xch[q];
      any = any || xch[36] != 0.0;
    }
    if (!primary) return;
  }
  if (!active || !any) return;
  // acc = sum Z_p Z_q^T with p = row pose of the tile, q = column pose; stored below the diagonal of the reduced system
  const int p = gi * 16 + i, q = gj * 16 + j;
  const int op = off_pose[p], oq = off_pose[q];
  int ld;
  if (op > oq) {
    const int64_t base = sys_find(sys, op, oq, &ld);
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) sys.L[base + r + (int64_t)cc * ld] -= acc[6 * r + cc];
  } else {
    const int64_t base = sys_find(sys, oq, op, &ld);     // rows of q, columns of p: the transpose
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) sys.L[base + r + (int64_t)cc * ld] -= acc[6 * cc + r];
  }
}

// one warp per pose over its (contiguous, pose-major) records: rhs_p += sum_k Z_k u_l(k), and the diagonal block
// S_pp -= sum_k Z_k Z_k^T (lower triangle) that k_schur_tiles leaves out
__global__ void __launch_bounds__(256) k_schur_rhs(int P, int64_t M, const int64_t* __restrict__ pose_obs_ptr, const int* __restrict__ pz_point,
                                                   const double* __restrict__ Zp, const double* __restrict__ ul,
                                                   const int* __restrict__ off_pose, SysView sys) {
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= P) return;
  const int64_t b = pose_obs_ptr[p], e = pose_obs_ptr[p + 1];
  if (b == e) return;
  double acc[27];
#pragma unroll
  for (int q = 0; q < 27; ++q) acc[q] = 0.0;
  for (int64_t k = b + lane; k < e; k += 32) {
    const int l = pz_point[k];
    double w[REC];
#pragma unroll
    for (int kc = 0; kc < 3; ++kc) {
      const double2* zp = reinterpret_cast<const double2*>(Zp + ((int64_t)kc * M + k) * 6);
#pragma unroll
      for (int q = 0; q < 3; ++q) { const double2 v = __ldg(zp + q); w[6 * q + kc] = v.x; w[6 * q + 3 + kc] = v.y; }
    }
    const double y0 = ul[3 * (int64_t)l], y1 = ul[3 * (int64_t)l + 1], y2 = ul[3 * (int64_t)l + 2];
#pragma unroll
    for (int q = 0; q < 6; ++q) acc[q] += w[3 * q] * y0 + w[3 * q + 1] * y1 + w[3 * q + 2] * y2;
    int t = 6;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc <= r; ++cc, ++t)
        acc[t] = fma(w[3 * r + 2], w[3 * cc + 2], fma(w[3 * r + 1], w[3 * cc + 1], fma(w[3 * r], w[3 * cc], acc[t])));
  }
#pragma unroll
  for (int q = 0; q < 27; ++q)
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], dd);
  const int C0 = off_pose[p];
  const int sn = sys.col2sn[C0];
  const int nr = sys.sn_nrows[sn];
  double* col0 = sys.L + sys.sn_valptr[sn] + (int64_t)(C0 - sys.sn_col0[sn]) * nr;     // column C0 of the panel; a pose never straddles supernodes
  const int r0 = C0 - sys.sn_col0[sn];                                               // row of the pose's first scalar inside the panel
  if (lane < 27) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 27; ++q) if (q == lane) v = acc[q];
    if (lane < 6) {
      col0[(int64_t)lane * nr + nr - 1] += v;                 // rhs row
    } else {
      int t = lane - 6, r = 0;
      while (t > r) { t -= r + 1; ++r; }                      // lower-triangle index -> (r, cc = t)
      col0[(int64_t)t * nr + r0 + r] -= v;
    }
  }
}

static void launch_tiles(fg_ctx* c, const SysView& sys) {
  DevGraph& d = c->d;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_schur_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SchurSmem));
    attr_set = true;
  }
  k_schur_tiles<<<d.n_tiles, ST_THREADS, sizeof(SchurSmem), FGS(c->stream)>>>(d.tile_desc, (int)d.n[T_POSE], d.pc_lo, d.pc_n, d.pc_ptr, d.pc_ent,
                                                                             d.Zp, d.n_obs, d.off[T_POSE], sys);
}

void launch_schur(fg_ctx* c, double lambda) {
  DevGraph& d = c->d;
  cudaStream_t st = c->stream;
  SysView sys;
  sys.L = d.L; sys.col2sn = d.col2sn; sys.sn_col0 = d.sn_col0; sys.sn_ncols = d.sn_ncols; sys.sn_nrows = d.sn_nrows;
  sys.sn_rowptr = d.sn_rowptr; sys.sn_valptr = d.sn_valptr; sys.rowidx = d.rowidx; sys.n_r = c->sym.n_r;
  const int64_t L = d.n[T_POINT];
  k_vinv<<<cdiv(L, 256), 256, 0, FGS(st)>>>(L, d.V, d.gl, lambda, d.Vinv, d.Cf, d.ul);
  if (d.n_obs) k_zmat<<<cdiv(d.n_obs, 256), 256, 0, FGS(st)>>>(d.n_obs, d.obs_point, d.obs_ppos, d.W, d.Cf, d.Zp);
  if (c->kev[2]) cudaEventRecord(c->kev[2], st);
  if (d.n_tiles) {
    launch_tiles(c, sys);                                // table chunks of 32 landmarks (d.schur_ch, build_schur_tables)
  }
  if (c->kev[3]) cudaEventRecord(c->kev[3], st);
  const int P = (int)d.n[T_POSE];
  if (d.n_obs) k_schur_rhs<<<cdiv((int64_t)P * 32, 256), 256, 0, FGS(st)>>>(P, d.n_obs, d.pose_obs_ptr, d.pz_point, d.Zp, d.ul, d.off[T_POSE], sys);
}

}  // namespace fg
