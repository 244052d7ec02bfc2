// fg_schur.cu -- K6 (second half): landmark Schur complement onto the reduced pose system (fp64).
//
//   S_pq -= sum_l W_pl (V_l + lambda I)^-1 W_ql^T ,   rhs_p += sum_l W_pl (V_l + lambda I)^-1 g_l
//
// for the GenericProjectionFactor / PriorFactor<Point3> part of the graph CGraphGT::addToGTSAM builds
// (gtsam/gtsam_graph.cpp:370-448); in GTSAM this is the elimination of every Point3 before the poses.
//
// Symmetric form.  With V_l + lambda I = G G^T (3x3 Cholesky) and C = G^-T,  W Vinv W^T = (W C)(W C)^T, so one array
// Z_o = W_o C_l (6x3 per observation, 144 B) serves both sides of the product.  k_zmat writes Z in POSE-major order:
// the observations of one pose are contiguous and sorted by landmark.
//
// k_schur_tiles: output-stationary, deterministic, no atomics, no pair lists.  One CTA owns a 16 x 16 tile of 6x6
// blocks (16 consecutive row poses x 16 consecutive column poses); ONE LANE owns one block and keeps its 36 sums in
// registers.  The landmarks are cut into chunks of CH consecutive ids.  Per chunk the CTA stages the Z records of its
// 32 poses in shared memory, each lane intersects the landmark bit masks of its two poses and walks the set bits:
//     S(p_i, q_j) += Z_(p_i, l) Z_(q_j, l)^T       (108 DFMA per hit, operands from shared memory).
// Shared-memory layout [element e][record k][pose i] with an odd plane stride: the 16 lanes of a half-warp sit on a
// wrapped diagonal of the tile (distinct i, distinct j), so every operand read is bank-conflict free whatever records
// the lanes are at.  HBM/L2 traffic is one record per (pose, landmark, tile) instead of two per pair.
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
  for (int i = tid; i < n * 9; i += 256) {
    const int r = i / 9, part = i - 9 * r;
    const double* s = buf + r * RST + 2 * part;
    reinterpret_cast<double2*>(Zp + (int64_t)ppos[r] * REC)[part] = make_double2(s[0], s[1]);
  }
}

// ------------------------------------------------------------------ tile kernel
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}

#define ST_THREADS 256
template <int CH, int KMAX>
struct SchurSmem {
  static constexpr int PS = KMAX * 16 + 1;        // plane stride (doubles): odd, so bank = (e + i) mod 16
  double rows[REC * PS];
  double cols[REC * PS];
  int src[ST_THREADS / 32][4 * CH];               // per warp: pose-major record index of every record it stages
};

// CH landmarks per chunk, room for KMAX records per pose and chunk.  KMAX < CH (32 / 24) bets that no pose sees more than
// KMAX of a chunk's landmarks together with the other side of the tile; a chunk that loses the bet is done as two
// half-chunks of 16 landmarks (<= 16 records per pose).
template <int CH, int KMAX>
__global__ void __launch_bounds__(ST_THREADS, 2)
k_schur_tiles(const int4* __restrict__ tiles, int P, const int* __restrict__ pc_lo, const int* __restrict__ pc_n,
              const int64_t* __restrict__ pc_ptr, const uint2* __restrict__ pc_ent, const double* __restrict__ Zp,
              const int* __restrict__ off_pose, SysView sys) {
  extern __shared__ __align__(16) unsigned char st_raw[];
  SchurSmem<CH, KMAX>& sm = *reinterpret_cast<SchurSmem<CH, KMAX>*>(st_raw);
  constexpr int PS = SchurSmem<CH, KMAX>::PS;
  const unsigned FULL = 0xffffffffu;
  const int4 td = tiles[blockIdx.x];
  const int gi = td.x, gj = td.y, cb = td.z, ce = td.w;
  const bool diag = gi == gj;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // metadata role: lane s holds the chunk entry of pose s (0..15 row poses, 16..31 column poses)
  const int mp = ((lane >> 4) ? gj : gi) * 16 + (lane & 15);
  int m_lo = 0, m_n = 0;
  int64_t m_ptr = 0;
  if (mp < P) { m_lo = pc_lo[mp]; m_n = pc_n[mp]; m_ptr = pc_ptr[mp]; }
  // compute role: half-warp h of warp w works on the wrapped diagonal d = 2 w + h of the tile: pose distance, hence the number
  // of shared landmarks, changes along d, so the two halves of a warp do about the same work and the lanes stay busy together
  // (pairing diagonals w and 15 - w balanced the warps instead and left half the lanes of every instruction idle:
  // 16.3 of 32 threads per instruction in profiles/r1_ncu_full_summary.md); an idle warp gives its issue slots to the other CTA.
  // In a diagonal tile the diagonals d and 16 - d hold the same unordered pose pairs: the two lanes of a pair split its
  // landmarks by bit parity and their sums are joined at the end; the p == q blocks are left to k_schur_rhs.
  const int i = lane & 15, d = 2 * warp + (lane >> 4), j = (i + d) & 15;
  bool active = (gi * 16 + i < P) && (gj * 16 + j < P);
  const bool primary = !diag || d < 8 || (d == 8 && i < 8);
  unsigned hit_sel = 0xffffffffu;
  if (diag) { active = active && d != 0; hit_sel = primary ? 0x55555555u : 0xaaaaaaaau; }
  double acc[36];
#pragma unroll
  for (int q = 0; q < 36; ++q) acc[q] = 0.0;
  bool any = false;
  const double* cbase = diag ? sm.rows : sm.cols;
  int* srcl = sm.src[warp];
  double* stage = sm.rows;                // rows, then cols: one staging area of 2 x REC planes
  const int n_st = diag ? 16 : 32;        // poses to stage
  const int spw = n_st / 8;               // poses per warp: s = warp + 8 t

  auto fetch = [&](int c) -> uint2 {
    const int r = c - m_lo;
    return (c < ce && r >= 0 && r < m_n) ? __ldg(&pc_ent[m_ptr + r]) : make_uint2(0u, 0u);
  };
  uint2 ent1 = fetch(cb), ent2 = fetch(cb + 1);
  for (int c = cb; c < ce; ++c) {
    const uint2 cur = ent1;
    ent1 = ent2;
    ent2 = fetch(c + 2);                          // entries ride two chunks ahead of their use
    // landmarks seen from both sides of the tile; a pose stages only its records of those
    unsigned any16 = cur.y;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) any16 |= __shfl_xor_sync(FULL, any16, o);
    const unsigned other = __shfl_xor_sync(FULL, any16, 16);
    const unsigned fall = cur.y & other;
    if (!__any_sync(FULL, fall != 0u)) continue;  // identical decision in every warp: all hold the same 32 entries
    int nparts = 1;
    if (KMAX < CH && __any_sync(FULL, __popc(fall) > KMAX)) nparts = 2;
    for (int part_i = 0; part_i < nparts; ++part_i) {
    const unsigned f = nparts == 1 ? fall : (part_i == 0 ? (fall & 0x0000ffffu) : (fall & 0xffff0000u));
    if (nparts == 2 && !__any_sync(FULL, f != 0u)) continue;
    __syncthreads();                              // the previous chunk's products are done with the staging area
    // ---- stage: list the records of this warp's poses, then copy them 3 records (27 lanes x 16 B) per step
    int base_t[5];
    base_t[0] = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int cnt = 0;
      if (t < spw) {
        const int s = warp + 8 * t;
        const unsigned fm = __shfl_sync(FULL, f, s), mm = __shfl_sync(FULL, cur.y, s);
        const int st = (int)__shfl_sync(FULL, cur.x, s);
        cnt = __popc(fm);
        if ((fm >> lane) & 1u) {
          const unsigned low = (1u << lane) - 1u;
          srcl[base_t[t] + __popc(fm & low)] = st + __popc(mm & low);
        }
      }
      base_t[t + 1] = base_t[t] + cnt;
    }
    __syncwarp();
    {
      // three records per step: lane (r, part) moves elements 2 part and 2 part + 1 of record pos0 + r with two
      // 8-byte cp.async (global side: 144 contiguous bytes per record; shared side: the transposed layout)
      const int total = base_t[4];
      const int r = lane / 9, part = lane - 9 * r;
      if (r < 3)
        for (int pos = r; pos < total; pos += 3) {
          const int t = (pos >= base_t[1]) + (pos >= base_t[2]) + (pos >= base_t[3]);
          const int k = pos - (t == 0 ? 0 : (t == 1 ? base_t[1] : (t == 2 ? base_t[2] : base_t[3])));
          const int s = warp + 8 * t;
          const int e0 = 2 * part + (r & 1), e1 = 2 * part + 1 - (r & 1);      // neighbouring records start on different banks
          const double* src = Zp + (int64_t)srcl[pos] * REC;
          double* dst = stage + (s >> 4) * (REC * PS) + k * 16 + (s & 15);
          cp_async8(dst + e0 * PS, src + e0);
          cp_async8(dst + e1 * PS, src + e1);
        }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    // ---- products: walk the landmarks both poses of this lane's block see
    const unsigned fi = __shfl_sync(FULL, f, i), fj = __shfl_sync(FULL, f, 16 + j);
    unsigned hit = active ? (fi & fj & hit_sel) : 0u;
    while (hit) {
      const int b = __ffs(hit) - 1;
      hit &= hit - 1u;
      const unsigned low = (1u << b) - 1u;
      const double* ra = sm.rows + __popc(fi & low) * 16 + i;
      const double* cq = cbase + __popc(fj & low) * 16 + j;
      double y[REC];
#pragma unroll
      for (int e = 0; e < REC; ++e) y[e] = ra[e * PS];
#pragma unroll
      for (int jj = 0; jj < 6; ++jj) {
        const double w0 = cq[(3 * jj) * PS], w1 = cq[(3 * jj + 1) * PS], w2 = cq[(3 * jj + 2) * PS];
#pragma unroll
        for (int ii = 0; ii < 6; ++ii) acc[6 * ii + jj] = fma(y[3 * ii + 2], w2, fma(y[3 * ii + 1], w1, fma(y[3 * ii], w0, acc[6 * ii + jj])));
      }
      any = true;
    }
    }
  }
  if (diag) {
    // join the two halves of every pose pair: the secondary lane (j, i) parks its sums, transposed, in the staging area
    __syncthreads();
    double* xch = stage + (primary ? i * 16 + j : j * 16 + i) * 37;
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
__global__ void __launch_bounds__(256) k_schur_rhs(int P, const int64_t* __restrict__ pose_obs_ptr, const int* __restrict__ pz_point,
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
    const double2* zp = reinterpret_cast<const double2*>(Zp + k * REC);
#pragma unroll
    for (int q = 0; q < 9; ++q) { const double2 v = __ldg(zp + q); w[2 * q] = v.x; w[2 * q + 1] = v.y; }
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

template <int CH, int KMAX>
static void launch_tiles(fg_ctx* c, const SysView& sys) {
  DevGraph& d = c->d;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_schur_tiles<CH, KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SchurSmem<CH, KMAX>));
    attr_set = true;
  }
  k_schur_tiles<CH, KMAX><<<d.n_tiles, ST_THREADS, sizeof(SchurSmem<CH, KMAX>), FGS(c->stream)>>>(d.tile_desc, (int)d.n[T_POSE], d.pc_lo, d.pc_n, d.pc_ptr, d.pc_ent,
                                                                                 d.Zp, d.off[T_POSE], sys);
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
    launch_tiles<32, 24>(c, sys);                        // d.schur_ch == 32 (build_schur_tables)
  }
  if (c->kev[3]) cudaEventRecord(c->kev[3], st);
  const int P = (int)d.n[T_POSE];
  if (d.n_obs) k_schur_rhs<<<cdiv((int64_t)P * 32, 256), 256, 0, FGS(st)>>>(P, d.pose_obs_ptr, d.pz_point, d.Zp, d.ul, d.off[T_POSE], sys);
}

}  // namespace fg
