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
// (SURVEY.md section 3A).  The tiles are a true dense contraction and run on the fp64 tensor-core path (DMMA,
// mma.sync.m8n8k4.f64); tcgen05 has no fp64 kind.
#include "fg_internal.h"

namespace fg {

#define FT 64          // tile edge
#define FK 16          // max supernode width
#define FM 2           // members staged per barrier pair
#define FS 68          // shared-memory row stride (doubles): 4 (mod 16), so the 8 x 4 MMA fragment loads are conflict free

// D (8x8) += A (8x4, row major) * B (4x8, column major) in fp64 on the tensor cores (DMMA).  Fragments: lane = 4 g + t
// holds A[g][t], B[t][g] and D[g][2t], D[g][2t + 1].  Measured on this B200: 37.1 TFLOP/s against 36.2 TFLOP/s for DFMA
// (profiles/tools/fp64_peak.cu) -- the same arithmetic peak, but one instruction carries 256 FMAs and its operands
// come from registers, which is what a shared-memory-bound DFMA tile kernel lacks.  (tcgen05 has no fp64 kind.)
__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) k_front_syrk(int n_tiles, const int* __restrict__ tile_leaf, const int* __restrict__ tile_i,
                                                    const int* __restrict__ tile_j, const int* __restrict__ tile_mptr,
                                                    const FrontRec* __restrict__ tile_mrec, const int* __restrict__ fr_rowptr,
                                                    const int64_t* __restrict__ fr_uptr, const int* __restrict__ posmap,
                                                    const double* __restrict__ L, double* __restrict__ U) {
  __shared__ __align__(16) double Ai[FM][FK][FS];
  __shared__ __align__(16) double Aj[FM][FK][FS];
  if ((int)blockIdx.x >= n_tiles) return;
  const int t = blockIdx.x;
  const int l = tile_leaf[t], ti = tile_i[t], tj = tile_j[t];
  const int nR = fr_rowptr[l + 1] - fr_rowptr[l];
  const int tid = threadIdx.x;
  const int w = tid >> 5, g = (tid & 31) >> 2, tq = tid & 3;     // warp w owns rows [8w, 8w + 8) of the tile; MMA lane coordinates
  const int r = tid & 63, kq = tid >> 6;               // staging: this thread always loads front row r of the tile
  const int gi = ti * FT + r, gj = tj * FT + r;
  double acc[8][2];
#pragma unroll
  for (int x = 0; x < 8; ++x) { acc[x][0] = 0.0; acc[x][1] = 0.0; }
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
        const double (*Bi)[FS] = Ai[q];
        const double (*Bj)[FS] = diag ? Ai[q] : Aj[q];
#pragma unroll
        for (int ks = 0; ks < FK / 4; ++ks) {
          const double a = Bi[4 * ks + tq][8 * w + g];
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            if (diag && nt > w) continue;               // blocks above the diagonal of a diagonal tile are never stored
            dmma884(acc[nt], a, Bj[4 * ks + tq][8 * nt + g]);
          }
        }
      }
    }
    __syncthreads();
  }
  double* Ul = U + fr_uptr[l];
  const int i = ti * FT + 8 * w + g;
  if (i < nR) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = tj * FT + 8 * nt + 2 * tq + e;
        if (j < nR && j <= i) Ul[(int64_t)i * nR + j] = acc[nt][e];
      }
  }
}

void launch_front_syrk(fg_ctx* c) {
  DevGraph& d = c->d;
  const int n_tiles = (int)c->sym.tile_leaf.size();
  if (!n_tiles) return;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_front_syrk, cudaFuncAttributePreferredSharedMemoryCarveout, 100); attr = true; }
  k_front_syrk<<<n_tiles, 256, 0, FGS(c->stream)>>>(n_tiles, d.tile_leaf, d.tile_i, d.tile_j, d.tile_mptr, d.tile_mrec, d.fr_rowptr, d.fr_uptr,
                                               d.posmap, d.L, d.U);
}

}  // namespace fg
