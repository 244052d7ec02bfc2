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
#define FM 2           // members staged per barrier pair (2 x 2 x 16 KB of static shared memory)

__global__ void __launch_bounds__(256) k_front_syrk(int n_tiles, const int* __restrict__ tile_leaf, const int* __restrict__ tile_i,
                                                    const int* __restrict__ tile_j, const int* __restrict__ tile_mptr,
                                                    const FrontRec* __restrict__ tile_mrec, const int* __restrict__ fr_rowptr,
                                                    const int64_t* __restrict__ fr_uptr, const int* __restrict__ posmap,
                                                    const double* __restrict__ L, double* __restrict__ U) {
  __shared__ __align__(16) double Ai[FM][FK][FT];
  __shared__ __align__(16) double Aj[FM][FK][FT];
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
  const int m1 = tile_mptr[t + 1];
  for (int m0 = tile_mptr[t]; m0 < m1; m0 += FM) {
    // ---- stage up to FM members with all their loads in flight together
#pragma unroll
    for (int q = 0; q < FM; ++q) {
      if (m0 + q < m1) {
        const FrontRec rec = tile_mrec[m0 + q];
        const double* Ld = L + rec.val_off;
        const int* pm = posmap + rec.pm_off;
        const int pi = (gi < nR) ? __ldg(pm + gi) : -1;
        const int pj = (!diag && gj < nR) ? __ldg(pm + gj) : -1;
#pragma unroll
        for (int k = kq; k < FK; k += 4) {
          Ai[q][k][r] = (pi >= 0 && k < rec.K) ? __ldcg(&Ld[pi + (int64_t)k * rec.nrd]) : 0.0;
          if (!diag) Aj[q][k][r] = (pj >= 0 && k < rec.K) ? __ldcg(&Ld[pj + (int64_t)k * rec.nrd]) : 0.0;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < FM; ++q) {
      if (m0 + q < m1) {
        const double (*Bi)[FT] = Ai[q];
        const double (*Bj)[FT] = diag ? Ai[q] : Aj[q];
#pragma unroll
        for (int k = 0; k < FK; ++k) {
          const double2 a01 = *reinterpret_cast<const double2*>(&Bi[k][4 * ty]);
          const double2 a23 = *reinterpret_cast<const double2*>(&Bi[k][4 * ty + 2]);
          const double2 b01 = *reinterpret_cast<const double2*>(&Bj[k][4 * tx]);
          const double2 b23 = *reinterpret_cast<const double2*>(&Bj[k][4 * tx + 2]);
          const double a[4] = {a01.x, a01.y, a23.x, a23.y}, b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] += a[x] * b[y];
        }
      }
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
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_front_syrk, cudaFuncAttributePreferredSharedMemoryCarveout, 100); attr = true; }
  k_front_syrk<<<n_tiles, 256, 0, c->stream>>>(n_tiles, d.tile_leaf, d.tile_i, d.tile_j, d.tile_mptr, d.tile_mrec, d.fr_rowptr, d.fr_uptr,
                                               d.posmap, d.L, d.U);
}

}  // namespace fg
