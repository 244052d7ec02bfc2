// fg_front.cu -- K7, leaf fronts: the dense update matrix of every nested-dissection leaf onto its front (fp64).
//
// In the left-looking factorisation every separator supernode used to pull one rank-K update from each of the
// ~250 leaf supernodes below it (K = 15), re-streaming the same descendant panels for each of its ~25 siblings:
// half of the kernel time at C5 and 4.8x the algorithmic DRAM traffic (profiles/r1_ncu_full_summary.md).
// Here a leaf contributes once: after its own supernodes are factored (phase A of k_chol_reg),
//     U_leaf = sum_{d in leaf} A_d A_d^T ,   A_d = rows of L_d that lie in the leaf's front (gathered by posmap)
// is formed as a dense SYRK, one CTA per 64 x 64 tile of the lower triangle, and the supernodes outside the leaf
// subtract the entries of U_leaf that fall into their panels (prologue of phase C of k_chol_reg).
// This is the multifrontal "update matrix" of the leaf; GTSAM's multifrontal elimination forms the same quantity
// (SURVEY.md section 3A).  DFMA on CUDA cores: tcgen05 has no fp64 kind.
#include "fg_internal.h"

namespace fg {

#define FT 64          // tile edge
#define FK 16          // max supernode width

__global__ void __launch_bounds__(256) k_front_syrk(int n_tiles, const int* __restrict__ tile_leaf, const int* __restrict__ tile_i,
                                                    const int* __restrict__ tile_j, const int* __restrict__ leaf_sn_lo,
                                                    const int* __restrict__ leaf_sn_hi, const int* __restrict__ fr_rowptr,
                                                    const int64_t* __restrict__ fr_uptr, const int64_t* __restrict__ pm_ptr,
                                                    const int* __restrict__ posmap, const int64_t* __restrict__ pmne_ptr,
                                                    const unsigned char* __restrict__ pm_nonempty, SysView s, double* __restrict__ U) {
  __shared__ __align__(16) double Ai[FK][FT];
  __shared__ __align__(16) double Aj[FK][FT];
  const int t = blockIdx.x;
  if (t >= n_tiles) return;
  const int l = tile_leaf[t], ti = tile_i[t], tj = tile_j[t];
  const int nR = fr_rowptr[l + 1] - fr_rowptr[l];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int r = tid & 63, kq = tid >> 6;               // staging: this thread always loads front row r of the tile
  const int gi = ti * FT + r, gj = tj * FT + r;
  double acc[4][4];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
  const bool diag = (ti == tj);
  for (int d = leaf_sn_lo[l]; d < leaf_sn_hi[l]; ++d) {
    const unsigned char* ne = pm_nonempty + pmne_ptr[d];
    if (!ne[ti] || !ne[tj]) continue;                   // uniform: this supernode has no row in one of the two blocks
    const int K = s.sn_ncols[d], nrd = s.sn_nrows[d];
    const double* Ld = s.L + s.sn_valptr[d];
    const int* pm = posmap + pm_ptr[d];
    const int pi = (gi < nR) ? pm[gi] : -1;
    const int pj = (gj < nR) ? pm[gj] : -1;
    for (int k = kq; k < FK; k += 4) {
      Ai[k][r] = (pi >= 0 && k < K) ? __ldcg(&Ld[pi + (int64_t)k * nrd]) : 0.0;
      if (!diag) Aj[k][r] = (pj >= 0 && k < K) ? __ldcg(&Ld[pj + (int64_t)k * nrd]) : 0.0;
    }
    __syncthreads();
    const double (*Bj)[FT] = diag ? Ai : Aj;
#pragma unroll
    for (int k = 0; k < FK; ++k) {
      const double2 a01 = *reinterpret_cast<const double2*>(&Ai[k][4 * ty]);
      const double2 a23 = *reinterpret_cast<const double2*>(&Ai[k][4 * ty + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&Bj[k][4 * tx]);
      const double2 b23 = *reinterpret_cast<const double2*>(&Bj[k][4 * tx + 2]);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] += a[x] * b[y];
    }
    __syncthreads();
  }
  double* Ul = U + fr_uptr[l];
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int i = ti * FT + 4 * ty + x;
    if (i >= nR) continue;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int j = tj * FT + 4 * tx + y;
      if (j < nR && j <= i) Ul[(int64_t)i * nR + j] = acc[x][y];
    }
  }
}

void launch_front_syrk(fg_ctx* c) {
  DevGraph& d = c->d;
  const int n_tiles = (int)c->sym.tile_leaf.size();
  if (!n_tiles) return;
  SysView s;
  s.L = d.L; s.col2sn = d.col2sn; s.sn_col0 = d.sn_col0; s.sn_ncols = d.sn_ncols; s.sn_nrows = d.sn_nrows;
  s.sn_rowptr = d.sn_rowptr; s.sn_valptr = d.sn_valptr; s.rowidx = d.rowidx; s.n_r = c->sym.n_r;
  k_front_syrk<<<n_tiles, 256, 0, c->stream>>>(n_tiles, d.tile_leaf, d.tile_i, d.tile_j, d.leaf_sn_lo, d.leaf_sn_hi, d.fr_rowptr,
                                               d.fr_uptr, d.pm_ptr, d.posmap, d.pmne_ptr, d.pm_nonempty, s, d.U);
}

}  // namespace fg
