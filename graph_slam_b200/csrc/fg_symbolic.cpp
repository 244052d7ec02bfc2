// fg_symbolic.cpp -- host-side symbolic analysis of the reduced (pose-side) system.
//
// Replaces the part of gtsam::LevenbergMarquardtOptimizer that chooses an elimination ordering and
// builds the elimination tree (COLAMD + multifrontal) for the graph
// CGraphGT builds (gtsam/gtsam_graph.cpp:1784-1788).  Point landmarks never enter this structure:
// they are eliminated first by the Schur complement (north_star); their co-visibility only
// contributes pose-pose couplings here.
//
// Ordering: nested dissection over the frame sequence.  Frames ([X_i V_i B_i] by key index) of a
// sequential VIO/BA graph form a band; a segment [lo,hi) is split at its middle, the separator is the
// set of variables of the right half adjacent to the left half (for a covisibility band of w frames:
// the w poses after the cut plus one velocity/bias pair), and the two halves recurse.  This turns the
// 5000-step dependency chain of a banded Cholesky into ~log2 levels of short independent chains, which
// is what lets the persistent kernel in fg_chol.cu use the whole GPU.  Plane landmarks are ordered last.
// Supernodes are runs of consecutive variables along one elimination-tree chain, capped at kMaxSnCols columns.  The
// amalgamation is relaxed: a child is merged with its parent when the parent brings few new rows (one more frame of a
// banded VIO/BA graph adds 15 rows to ~600), at the price of explicit zeros in the child's columns -- half as many
// dependency levels and half as many passes over the descendant panels in the left-looking factorisation.
#include <algorithm>
#include <cstdio>
#include <functional>
#include <numeric>
#include <tuple>
#include "fg_internal.h"

namespace fg {

static const int kMaxSnCols = 32;     // two [X V B] frames (30) or five poses (30): the target width of k_chol_rs (RS_NC)
static const int kUpdK = 16;          // a descendant wider than this is applied as two rank-<=16 updates (one pipeline stage each)

int build_symbolic(fg_ctx* c) {
  HostGraph& h = c->h;
  Symbolic& S = c->sym;
  S = Symbolic();
  // ---- reduced variables in frame-major base order
  struct RV { int type, idx; uint64_t kidx; int cls; };
  std::vector<RV> rv;
  const uint64_t mask = (1ull << 56) - 1;
  for (int t : {T_POSE, T_VEC3, T_BIAS, T_PLANE})
    for (int i = 0; i < (int)h.keys[t].size(); ++i)
      rv.push_back({t, i, h.keys[t][i] & mask, t == T_PLANE ? 1 : 0});
  std::stable_sort(rv.begin(), rv.end(), [](const RV& a, const RV& b) {
    if (a.cls != b.cls) return a.cls < b.cls;
    if (a.kidx != b.kidx) return a.kidx < b.kidx;
    return a.type < b.type;
  });
  const int nv = (int)rv.size();
  if (nv == 0) return FG_ERR_STATE;
  std::vector<int> base[T_COUNT];           // (type, idx) -> base index
  for (int t = 0; t < T_COUNT; ++t) base[t].assign(h.keys[t].size(), -1);
  std::vector<int> bdim(nv);
  for (int p = 0; p < nv; ++p) { base[rv[p].type][rv[p].idx] = p; bdim[p] = kDim[rv[p].type]; }

  // ---- symmetric adjacency on base indices
  std::vector<std::vector<int>> adj(nv);
  auto edge = [&](int a, int b) { if (a != b) { adj[a].push_back(b); adj[b].push_back(a); } };
  for (size_t f = 0; f < h.bt_i.size(); ++f) edge(base[T_POSE][h.bt_i[f]], base[T_POSE][h.bt_j[f]]);
  for (size_t f = 0; f < h.ge_i.size(); ++f) edge(base[T_POSE][h.ge_i[f]], base[T_POSE][h.ge_j[f]]);
  for (size_t f = 0; f < h.imu_rec.size(); ++f) {
    const int* v = &h.imu_var[6 * f];
    int p[6] = {base[T_POSE][v[0]], base[T_VEC3][v[1]], base[T_POSE][v[2]], base[T_VEC3][v[3]], base[T_BIAS][v[4]], base[T_BIAS][v[5]]};
    for (int a = 0; a < 6; ++a) for (int b = a + 1; b < 6; ++b) edge(p[a], p[b]);
  }
  for (size_t f = 0; f < h.pl_pose.size(); ++f) edge(base[T_POSE][h.pl_pose[f]], base[T_PLANE][h.pl_plane[f]]);
  for (size_t f = 0; f < h.se_a.size(); ++f) edge(base[T_POSE][h.se_a[f]], base[T_POSE][h.se_b[f]]);
  {
    // landmark co-visibility: every pair of poses observing one landmark is coupled through the Schur complement
    const int64_t L = h.count(T_POINT), M = (int64_t)h.pj_pose.size(), P = h.count(T_POSE);
    if (L && M) {
      std::vector<int64_t> lp(L + 1, 0), pp(P + 1, 0);
      for (int64_t o = 0; o < M; ++o) { lp[h.pj_point[o] + 1]++; pp[h.pj_pose[o] + 1]++; }
      for (int64_t l = 0; l < L; ++l) lp[l + 1] += lp[l];
      for (int64_t p = 0; p < P; ++p) pp[p + 1] += pp[p];
      std::vector<int> lm_pose(M), pose_lm(M);
      {
        std::vector<int64_t> c1(lp.begin(), lp.end() - 1), c2(pp.begin(), pp.end() - 1);
        for (int64_t o = 0; o < M; ++o) { lm_pose[c1[h.pj_point[o]]++] = h.pj_pose[o]; pose_lm[c2[h.pj_pose[o]]++] = h.pj_point[o]; }
      }
      std::vector<int> stamp(P, -1);
      S.cov_ptr.assign(P + 1, 0);
      for (int p = 0; p < (int)P; ++p) {
        stamp[p] = p;
        int bp = base[T_POSE][p];
        if (pp[p + 1] > pp[p]) S.cov_pose.push_back(p);      // self (diagonal block)
        for (int64_t k = pp[p]; k < pp[p + 1]; ++k) {
          int l = pose_lm[k];
          for (int64_t k2 = lp[l]; k2 < lp[l + 1]; ++k2) {
            int q = lm_pose[k2];
            if (stamp[q] != p) { stamp[q] = p; adj[bp].push_back(base[T_POSE][q]); S.cov_pose.push_back(q); }
          }
        }
        S.cov_ptr[p + 1] = (int64_t)S.cov_pose.size();
      }
    }
  }
  for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }

  // ---- nested dissection over frames -> elimination order
  std::vector<int> elim; elim.reserve(nv);
  std::vector<int> leaf_of;      // base index -> nested-dissection leaf (-1: separator / plane)
  int n_leaves = 0;
  {
    // frames: maximal runs of class-0 variables with the same key index (base order is frame-major)
    std::vector<std::vector<int>> frames;
    int n0 = 0;
    while (n0 < nv && rv[n0].cls == 0) ++n0;
    for (int p = 0; p < n0;) {
      int q = p;
      std::vector<int> f;
      while (q < n0 && rv[q].kidx == rv[p].kidx) f.push_back(q++);
      frames.push_back(f);
      p = q;
    }
    std::vector<int> region(nv, -1);
    int next_region = 0;
    leaf_of.assign(nv, -1);
    std::function<void(std::vector<std::vector<int>>&)> nd = [&](std::vector<std::vector<int>>& F) {
      auto emit = [&]() { for (auto& f : F) for (int v : f) { elim.push_back(v); leaf_of[v] = n_leaves; } ++n_leaves; };
      if (F.size() < 8) { emit(); return; }
      size_t mid = F.size() / 2;
      int ra = next_region++, rb = next_region++;
      for (size_t i = 0; i < F.size(); ++i) for (int v : F[i]) region[v] = i < mid ? ra : rb;
      std::vector<char> inS;
      std::vector<std::vector<int>> A(F.begin(), F.begin() + mid), B, Sep;
      int dS = 0;
      for (size_t i = mid; i < F.size(); ++i) {
        std::vector<int> keep, sep;
        for (int v : F[i]) {
          bool touch = false;
          for (int u : adj[v]) if (region[u] == ra) { touch = true; break; }
          if (touch) { sep.push_back(v); dS += bdim[v]; } else { keep.push_back(v); }
        }
        if (!keep.empty()) B.push_back(keep);
        if (!sep.empty()) Sep.push_back(sep);
      }
      // dissect all the way down: measured at C5 (profiles/r2_nd_depth.md) stopping early at a separator-to-part ratio left 32 long
      // leaf chains with 1.5x the fill and 1.5x the dependency levels of the full recursion
      if (dS == 0 || B.size() < 4) { emit(); return; }
      nd(A);
      nd(B);
      for (auto& f : Sep) for (int v : f) elim.push_back(v);
    };
    if (!frames.empty()) nd(frames);
    for (int p = n0; p < nv; ++p) elim.push_back(p);     // plane landmarks last
  }
  // ---- positions / offsets in elimination order
  std::vector<int> posb(nv);                 // base index -> position
  for (int p = 0; p < nv; ++p) posb[elim[p]] = p;
  std::vector<int> voff(nv + 1, 0), vdim(nv);
  for (int p = 0; p < nv; ++p) { vdim[p] = bdim[elim[p]]; voff[p + 1] = voff[p] + vdim[p]; }
  S.n_r = voff[nv];
  for (int t : {T_POSE, T_VEC3, T_BIAS, T_PLANE}) {
    S.off[t].resize(h.keys[t].size());
    for (size_t i = 0; i < h.keys[t].size(); ++i) S.off[t][i] = voff[posb[base[t][i]]];
  }
  std::vector<std::vector<int>> hadj(nv);
  for (int b = 0; b < nv; ++b)
    for (int u : adj[b]) if (posb[u] > posb[b]) hadj[posb[b]].push_back(posb[u]);
  adj.clear(); adj.shrink_to_fit();
  for (auto& a : hadj) std::sort(a.begin(), a.end());
  const std::vector<std::vector<int>> hadj_keep = hadj;      // later neighbours of every variable: nnz(S) and the packed-exchange index

  // ---- symbolic elimination (variable level)
  std::vector<std::vector<int>> st(nv);
  std::vector<int> parent(nv, -1);
  std::vector<std::vector<int>> children(nv);
  std::vector<int> tmp;
  for (int v = 0; v < nv; ++v) {
    std::vector<int>& s = st[v];
    s = hadj[v];
    for (int ch : children[v]) {
      tmp.clear();
      const std::vector<int>& cs = st[ch];
      size_t i = 0, j = 0;
      while (i < s.size() || j < cs.size()) {
        int a = i < s.size() ? s[i] : INT32_MAX;
        int b = j < cs.size() ? cs[j] : INT32_MAX;
        if (b == v) { ++j; continue; }
        if (a < b) { tmp.push_back(a); ++i; }
        else if (b < a) { tmp.push_back(b); ++j; }
        else { tmp.push_back(a); ++i; ++j; }
      }
      s.swap(tmp);
    }
    if (!s.empty()) { parent[v] = s[0]; children[s[0]].push_back(v); }
    hadj[v].clear(); hadj[v].shrink_to_fit();
  }
  // ---- supernodes
  std::vector<int> sn_first, sn_last;
  {
    std::vector<int> sdim(nv, 0);            // scalar rows below variable v
    for (int v = 0; v < nv; ++v) for (int u : st[v]) sdim[v] += vdim[u];
    int v = 0;
    while (v < nv) {
      int first = v, cols = vdim[v];
      while (v + 1 < nv && parent[v] == v + 1 && cols + vdim[v + 1] <= kMaxSnCols && leaf_of[elim[v]] == leaf_of[elim[v + 1]]) {
        // struct(v) \ {v+1} is a subset of struct(v+1): the merged panel's rows are struct(v+1); `extra` of them are
        // explicit zeros in the columns gathered so far
        const int extra = sdim[v + 1] + vdim[v + 1] - sdim[v];
        if (extra > std::max(16, sdim[v] / 16)) break;
        ++v; cols += vdim[v];
      }
      sn_first.push_back(first); sn_last.push_back(v);
      ++v;
    }
  }
  S.n_sn = (int)sn_first.size();
  S.sn_col0.resize(S.n_sn); S.sn_ncols.resize(S.n_sn); S.sn_nrows.resize(S.n_sn);
  S.sn_rowptr.assign(S.n_sn + 1, 0); S.sn_valptr.assign(S.n_sn + 1, 0);
  S.col2sn.resize(S.n_r);
  for (int s = 0; s < S.n_sn; ++s) {
    int c0 = voff[sn_first[s]], nc = voff[sn_last[s] + 1] - c0;
    S.sn_col0[s] = c0; S.sn_ncols[s] = nc;
    int below = 0;
    for (int u : st[sn_last[s]]) below += vdim[u];
    S.sn_nrows[s] = nc + below + 1;
    S.sn_rowptr[s + 1] = S.sn_rowptr[s] + S.sn_nrows[s];
    // panels start on 128-byte lines: a cache line never holds values of two supernodes (k_chol_rs reads finished panels
    // through the non-coherent L1)
    S.sn_valptr[s + 1] = (S.sn_valptr[s] + (int64_t)S.sn_nrows[s] * nc + 15) / 16 * 16;
    for (int k = 0; k < nc; ++k) S.col2sn[c0 + k] = s;
    S.max_nrows = std::max(S.max_nrows, S.sn_nrows[s]);
    S.max_ncols = std::max(S.max_ncols, nc);
    double m = S.sn_nrows[s] - 1, k = nc;
    S.flops_factor += k * k * k / 3.0 + (m - k) * k * k + (m - k) * (m - k) * k;
  }
  S.nnz = S.sn_valptr[S.n_sn];
  S.rowidx.resize(S.sn_rowptr[S.n_sn]);
  for (int s = 0; s < S.n_sn; ++s) {
    int* r = &S.rowidx[S.sn_rowptr[s]];
    int k = 0;
    for (int j = 0; j < S.sn_ncols[s]; ++j) r[k++] = S.sn_col0[s] + j;
    for (int u : st[sn_last[s]]) for (int j = 0; j < vdim[u]; ++j) r[k++] = voff[u] + j;
    r[k++] = S.n_r;
  }
  // ---- packed exchange index (multi-GPU): the panel entries that can be non-zero BEFORE the factorisation on
  //      any rank -- the blocks of coupled variable pairs, the diagonal blocks and the right-hand-side row.  Only these
  //      (about nnz(S), not nnz(L) with its fill) travel in the allreduce of the reduced pose Hessian.
  S.nnz_S = 0;                                         // structural non-zeros of the assembled reduced system (lower triangle + rhs row)
  for (int v = 0; v < nv; ++v) {
    int64_t below = 0;
    for (int u : hadj_keep[v]) below += vdim[u];
    S.nnz_S += (int64_t)vdim[v] * (vdim[v] + 1) / 2 + vdim[v] + (int64_t)vdim[v] * below;
  }
  if (c->nranks > 1) {
    const std::vector<std::vector<int>>& hadj0 = hadj_keep;
    auto find = [&](int R, int C) -> int64_t {          // panel offset of entry (R, C), R >= C (host twin of sys_find)
      const int sn = S.col2sn[C];
      const int c0 = S.sn_col0[sn], nc = S.sn_ncols[sn], nr = S.sn_nrows[sn];
      int r;
      if (R < c0 + nc) r = R - c0;
      else { const int* rows = &S.rowidx[S.sn_rowptr[sn]]; r = (int)(std::lower_bound(rows + nc, rows + nr, R) - rows); }
      return S.sn_valptr[sn] + r + (int64_t)(C - c0) * nr;
    };
    for (int v = 0; v < nv; ++v) {
      const int cv = voff[v], dv = vdim[v];
      for (int j = 0; j < dv; ++j) {
        const int64_t b0 = find(cv + j, cv + j);
        for (int i = j; i < dv; ++i) S.pk_idx.push_back(b0 + (i - j));           // diagonal block, lower triangle
        S.pk_idx.push_back(find(S.n_r, cv + j));                                 // right-hand-side row
        for (int u : hadj0[v]) {                                                 // later neighbours: rows of u, this column
          const int64_t bu = find(voff[u], cv + j);
          for (int i = 0; i < vdim[u]; ++i) S.pk_idx.push_back(bu + i);          // a variable's rows are consecutive in a row list
        }
      }
    }
    std::sort(S.pk_idx.begin(), S.pk_idx.end());
  }
  // ---- update lists (target <- descendants) and ancestor lists (descendant -> targets)
  std::vector<std::vector<int>> ul(S.n_sn);   // triples (d, a, b)
  S.anc_ptr.assign(S.n_sn + 1, 0);
  for (int d = 0; d < S.n_sn; ++d) {
    const int* r = &S.rowidx[S.sn_rowptr[d]];
    int nr = S.sn_nrows[d] - 1;   // exclude rhs row
    int a = S.sn_ncols[d];
    while (a < nr) {
      int t = S.col2sn[r[a]];
      int b = a + 1;
      while (b < nr && S.col2sn[r[b]] == t) ++b;
      ul[t].push_back(d); ul[t].push_back(a); ul[t].push_back(b);
      S.anc_t.push_back(t); S.anc_a.push_back(a); S.anc_b.push_back(b);
      a = b;
    }
    S.anc_ptr[d + 1] = (int)S.anc_t.size();
  }
  S.upd_ptr.assign(S.n_sn + 1, 0);
  for (int s = 0; s < S.n_sn; ++s) S.upd_ptr[s + 1] = S.upd_ptr[s] + (int)ul[s].size() / 3;
  S.upd_d.resize(S.upd_ptr[S.n_sn]); S.upd_a.resize(S.upd_d.size()); S.upd_b.resize(S.upd_d.size());
  for (int s = 0; s < S.n_sn; ++s)
    for (size_t k = 0; k < ul[s].size() / 3; ++k) {
      S.upd_d[S.upd_ptr[s] + k] = ul[s][3 * k];
      S.upd_a[S.upd_ptr[s] + k] = ul[s][3 * k + 1];
      S.upd_b[S.upd_ptr[s] + k] = ul[s][3 * k + 2];
    }
  // ---- leaf fronts: everything a leaf contributes to the supernodes OUTSIDE it is gathered in one dense update
  //      matrix per leaf (fg_front.cu) instead of one rank-K update per (leaf supernode, outside target) pair
  S.n_leaves = n_leaves;
  S.sn_leaf.assign(S.n_sn, -1);
  for (int s = 0; s < S.n_sn; ++s) S.sn_leaf[s] = leaf_of[elim[sn_first[s]]];
  S.use_fronts = false;
  if (n_leaves >= 2) {
    S.leaf_sn_lo.assign(n_leaves, S.n_sn); S.leaf_sn_hi.assign(n_leaves, 0);
    for (int s = 0; s < S.n_sn; ++s) {
      int l = S.sn_leaf[s];
      if (l >= 0) { S.leaf_sn_lo[l] = std::min(S.leaf_sn_lo[l], s); S.leaf_sn_hi[l] = std::max(S.leaf_sn_hi[l], s + 1); }
    }
    bool contiguous = true;
    for (int l = 0; l < n_leaves && contiguous; ++l)
      for (int s = S.leaf_sn_lo[l]; s < S.leaf_sn_hi[l]; ++s) if (S.sn_leaf[s] != l) { contiguous = false; break; }
    if (contiguous) {
      S.use_fronts = true;
      S.fr_rowptr.assign(n_leaves + 1, 0); S.fr_uptr.assign(n_leaves + 1, 0);
      S.pm_ptr.assign(S.n_sn + 1, 0); S.pmne_ptr.assign(S.n_sn + 1, 0);
      std::vector<std::vector<int>> tf(S.n_sn);
      for (int l = 0; l < n_leaves; ++l) {
        const int lo = S.leaf_sn_lo[l], hi = S.leaf_sn_hi[l];
        const int lc1 = (hi > lo) ? S.sn_col0[hi - 1] + S.sn_ncols[hi - 1] : 0;   // first column after the leaf
        std::vector<int> Rl;
        for (int s = lo; s < hi; ++s) {
          const int* r = &S.rowidx[S.sn_rowptr[s]];
          for (int i = S.sn_ncols[s]; i < S.sn_nrows[s]; ++i) if (r[i] >= lc1) Rl.push_back(r[i]);
        }
        std::sort(Rl.begin(), Rl.end()); Rl.erase(std::unique(Rl.begin(), Rl.end()), Rl.end());
        S.fr_rows.insert(S.fr_rows.end(), Rl.begin(), Rl.end());
        S.fr_rowptr[l + 1] = (int)S.fr_rows.size();
        const int64_t nR = (int64_t)Rl.size();
        S.fr_uptr[l + 1] = S.fr_uptr[l] + nR * nR;
        // targets that receive this front
        int last_t = -1;
        for (int g : Rl) if (g < S.n_r) { int t = S.col2sn[g]; if (t != last_t) { tf[t].push_back(l); last_t = t; } }
        // tiles of the lower triangle
        const int nt = (int)((nR + 63) / 64);
        for (int ti = 0; ti < nt; ++ti) for (int tj = 0; tj <= ti; ++tj) { S.tile_leaf.push_back(l); S.tile_i.push_back(ti); S.tile_j.push_back(tj); }
      }
      // position maps: for every member supernode, where each front row sits in its row list
      for (int s = 0; s < S.n_sn; ++s) {
        const int l = S.sn_leaf[s];
        int64_t n = 0, nb = 0;
        if (l >= 0) {
          const int* Rl = &S.fr_rows[S.fr_rowptr[l]];
          const int nR = S.fr_rowptr[l + 1] - S.fr_rowptr[l];
          const int* r = &S.rowidx[S.sn_rowptr[s]];
          const int nr = S.sn_nrows[s];
          n = nR; nb = (nR + 63) / 64;
          size_t base = S.posmap.size(), nbase = S.pm_nonempty.size();
          S.posmap.resize(base + nR, -1); S.pm_nonempty.resize(nbase + nb, 0);
          int i = S.sn_ncols[s];
          for (int k = 0; k < nR; ++k) {
            while (i < nr && r[i] < Rl[k]) ++i;
            if (i < nr && r[i] == Rl[k]) { S.posmap[base + k] = i; S.pm_nonempty[nbase + k / 64] = 1; }
          }
        }
        S.pm_ptr[s + 1] = S.pm_ptr[s] + n; S.pmne_ptr[s + 1] = S.pmne_ptr[s] + nb;
      }
      // per-tile work lists for k_front_syrk
      S.tile_mptr.assign(S.tile_leaf.size() + 1, 0);
      for (size_t t = 0; t < S.tile_leaf.size(); ++t) {
        const int l = S.tile_leaf[t], ti = S.tile_i[t], tj = S.tile_j[t];
        for (int m = S.leaf_sn_lo[l]; m < S.leaf_sn_hi[l]; ++m) {
          const unsigned char* ne = &S.pm_nonempty[S.pmne_ptr[m]];
          if (!ne[ti] || !ne[tj]) continue;
          for (int k0 = 0; k0 < S.sn_ncols[m]; k0 += kUpdK) {      // k_front_syrk stages <= 16 columns of a member per step
            FrontRec fr; fr.val_off = S.sn_valptr[m] + (int64_t)k0 * S.sn_nrows[m]; fr.pm_off = S.pm_ptr[m]; fr.nrd = S.sn_nrows[m];
            fr.K = std::min(kUpdK, S.sn_ncols[m] - k0); fr.pad[0] = fr.pad[1] = 0;
            S.tile_mrec.push_back(fr);
          }
        }
        S.tile_mptr[t + 1] = (int)S.tile_mrec.size();
      }
      // reduced update lists: drop (leaf member -> target outside that leaf)
      S.updr_ptr.assign(S.n_sn + 1, 0);
      for (int t = 0; t < S.n_sn; ++t) {
        for (int u = S.upd_ptr[t]; u < S.upd_ptr[t + 1]; ++u) {
          const int d = S.upd_d[u];
          if (S.sn_leaf[d] >= 0 && S.sn_leaf[d] != S.sn_leaf[t]) continue;
          S.updr_d.push_back(d); S.updr_a.push_back(S.upd_a[u]); S.updr_b.push_back(S.upd_b[u]);
        }
        S.updr_ptr[t + 1] = (int)S.updr_d.size();
      }
      S.tf_ptr.assign(S.n_sn + 1, 0);
      for (int t = 0; t < S.n_sn; ++t) {
        for (int l : tf[t]) if (S.sn_leaf[t] != l) { S.tf_leaf.push_back(l); if (S.sn_leaf[t] >= 0) S.use_fronts = false; }   // a leaf never feeds another leaf
        S.tf_ptr[t + 1] = (int)S.tf_leaf.size();
      }
      // levels and the two phase schedules
      std::vector<int> lv(S.n_sn, 0);
      for (int s = 0; s < S.n_sn; ++s) {
        for (int u = S.updr_ptr[s]; u < S.updr_ptr[s + 1]; ++u) lv[s] = std::max(lv[s], lv[S.updr_d[u]] + 1);
        for (int e = S.tf_ptr[s]; e < S.tf_ptr[s + 1]; ++e) {
          const int l = S.tf_leaf[e];
          for (int m = S.leaf_sn_lo[l]; m < S.leaf_sn_hi[l]; ++m) lv[s] = std::max(lv[s], lv[m] + 1);
        }
      }
      for (int s = 0; s < S.n_sn; ++s) (S.sn_leaf[s] >= 0 ? S.sched_a : S.sched_c).push_back(s);
      auto bylevel = [&](int a, int b) { return lv[a] < lv[b]; };
      std::stable_sort(S.sched_a.begin(), S.sched_a.end(), bylevel);
      std::stable_sort(S.sched_c.begin(), S.sched_c.end(), bylevel);
      S.n_levels_fronts = S.n_sn ? *std::max_element(lv.begin(), lv.end()) + 1 : 0;
      for (int l = 0; l < n_leaves; ++l) if (S.fr_rowptr[l + 1] - S.fr_rowptr[l] > 1024) S.use_fronts = false;   // k_chol_reg stages a front's row list in shared memory
    }
  }
  // ---- row-split work units for k_chol_rs: blocks of <= kRsRows panel rows per supernode
  {
    const int kRsRows = 128;                  // RS_T: one panel row per thread
    const bool fr = S.use_fronts;
    const std::vector<int>& UP = fr ? S.updr_ptr : S.upd_ptr;
    const std::vector<int>& UD = fr ? S.updr_d : S.upd_d;
    const std::vector<int>& UA = fr ? S.updr_a : S.upd_a;
    const std::vector<int>& UB = fr ? S.updr_b : S.upd_b;
    // levels of the supernodes under the update lists in use
    std::vector<int> lv(S.n_sn, 0);
    for (int s = 0; s < S.n_sn; ++s) {
      for (int u = UP[s]; u < UP[s + 1]; ++u) lv[s] = std::max(lv[s], lv[UD[u]] + 1);
      if (fr)
        for (int e = S.tf_ptr[s]; e < S.tf_ptr[s + 1]; ++e) {
          const int l = S.tf_leaf[e];
          for (int m = S.leaf_sn_lo[l]; m < S.leaf_sn_hi[l]; ++m) lv[s] = std::max(lv[s], lv[m] + 1);
        }
    }
    // the kernel's update list: the list in use with every descendant wider than kUpdK cut into column slices, in the order the
    // descendants are handed out (phase, level, index) -- a unit pulls its updates in list order and waits at the first one that is
    // not there yet, so a list in index order would park the whole second subtree of a separator behind the top of the first
    auto rank_of = [&](int d) { return std::make_tuple((fr && S.sn_leaf[d] < 0) ? 1 : 0, lv[d], d); };
    S.rsu_ptr.assign(S.n_sn + 1, 0);
    std::vector<int> us;
    for (int s = 0; s < S.n_sn; ++s) {
      us.resize(UP[s + 1] - UP[s]);
      std::iota(us.begin(), us.end(), UP[s]);
      std::sort(us.begin(), us.end(), [&](int x, int y) { return rank_of(UD[x]) < rank_of(UD[y]); });
      for (int u : us) {
        const int d = UD[u], a = UA[u], b = UB[u];
        const int* rd = &S.rowidx[S.sn_rowptr[d]];
        // which 8-column groups of the target the descendant's rows [a, b) reach, and where each target column comes from
        signed char inv[kMaxSnCols];
        for (int c = 0; c < kMaxSnCols; ++c) inv[c] = -1;
        int mask = 0;
        for (int i = a; i < b; ++i) { const int c = rd[i] - S.sn_col0[s]; inv[c] = (signed char)(i - a); mask |= 1 << (c >> 3); }
        for (int k0 = 0; k0 < S.sn_ncols[d]; k0 += kUpdK) {
          UpdRec r;
          r.val_off = S.sn_valptr[d] + a + (int64_t)k0 * S.sn_nrows[d]; r.row_off = S.sn_rowptr[d] + a; r.nrd = S.sn_nrows[d];
          r.nrows_u = S.sn_nrows[d] - a; r.K = (short)std::min(kUpdK, S.sn_ncols[d] - k0); r.nb = (short)(b - a);
          r.pad[0] = mask; r.pad[1] = k0;
          S.rsu_rec.push_back(r); S.rsu_d.push_back(d); S.rsu_src.push_back(u);
          S.rs_colinv.insert(S.rs_colinv.end(), inv, inv + kMaxSnCols);
        }
      }
      S.rsu_ptr[s + 1] = (int)S.rsu_d.size();
    }
    std::vector<int> nblk(S.n_sn);
    // the head unit takes the diagonal block and the rows right below it (rows [0, kRsRows) of the panel), the rest is cut evenly
    for (int s = 0; s < S.n_sn; ++s) nblk[s] = (std::max(0, S.sn_nrows[s] - kRsRows) + kRsRows - 1) / kRsRows;
    S.n_levels_rs = S.n_sn ? *std::max_element(lv.begin(), lv.end()) + 1 : 0;
    std::vector<int> order_a, order_c;
    for (int s = 0; s < S.n_sn; ++s) ((fr && S.sn_leaf[s] < 0) ? order_c : order_a).push_back(s);
    auto bylevel = [&](int a, int b) { return lv[a] < lv[b]; };
    std::stable_sort(order_a.begin(), order_a.end(), bylevel);
    std::stable_sort(order_c.begin(), order_c.end(), bylevel);
    std::vector<short> one;
    // units of a supernode: the head unit first -- the diagonal block and the rows right below it, i.e. the rows the next supernodes of
    // the chain need: it factors the block, publishes it (diagonal flag), then solves its other rows -- then the blocks of further
    // rows (they wait for the diagonal flag before their triangular solve)
    auto emit = [&](const std::vector<int>& order) {
      for (int s : order) {
        const int nr = S.sn_nrows[s], nb = nblk[s];
        const int head = std::min(nr, kRsRows);
        const int per = nb ? (nr - head + nb - 1) / nb : 0;
        const int* rows = &S.rowidx[S.sn_rowptr[s]];
        for (int b = -1; b < nb; ++b) {
          const int r0 = b < 0 ? 0 : head + b * per, r1 = b < 0 ? head : std::min(nr, r0 + per);
          S.rs_units.push_back(make_int4(s, r0, r1, nb + 1));
          S.rs_moff.push_back((int64_t)S.rs_map.size());
          const int nloc = r1 - r0;
          int last_src = -1;
          for (int q = S.rsu_ptr[s]; q < S.rsu_ptr[s + 1]; ++q) {
            const int u = S.rsu_src[q];
            if (u != last_src) {
              last_src = u;
              const int d = UD[u], a = UA[u];
              const int* rd = &S.rowidx[S.sn_rowptr[d]];
              const int nrd = S.sn_nrows[d];
              one.assign(nloc, (short)-1);
              // both row lists are sorted (a panel's rows: its own columns, then the structure below), and the descendant's rows from
              // a on are a subset of the target's rows
              int i = (int)(std::lower_bound(rd + a, rd + nrd, rows[r0]) - rd);
              for (int lr = 0; lr < nloc && i < nrd; ++lr) {
                const int g = rows[r0 + lr];
                while (i < nrd && rd[i] < g) { ++i; }
                if (i < nrd && rd[i] == g) { one[lr] = (short)(i - a); ++i; }
              }
            }
            S.rs_map.insert(S.rs_map.end(), one.begin(), one.end());      // a column slice of the same descendant: same rows
          }
        }
      }
    };
    emit(order_a);
    S.rs_units_a = (int)S.rs_units.size();
    emit(order_c);
    S.rs_sn_units.assign(S.n_sn, make_int2(0, 0));
    for (int k = (int)S.rs_units.size() - 1; k >= 0; --k) { int2& e = S.rs_sn_units[S.rs_units[k].x]; e.x = k; e.y += 1; }
    S.rs_ok = S.max_ncols <= kMaxSnCols && S.max_nrows <= 32767;      // row maps are int16
    for (const int4& un : S.rs_units) if (un.z - un.y > kRsRows || un.z <= un.y) S.rs_ok = false;
    static_assert(kMaxSnCols <= 32, "the head unit factors the diagonal block in its first warp");
  }
  // ---- schedule: supernodes by dependency level (longest path), a topological order that interleaves
  //      the independent chains so that the persistent kernel works on all of them at once
  S.level.assign(S.n_sn, 0);
  for (int s = 0; s < S.n_sn; ++s)
    for (int u = S.upd_ptr[s]; u < S.upd_ptr[s + 1]; ++u) S.level[s] = std::max(S.level[s], S.level[S.upd_d[u]] + 1);
  S.sched.resize(S.n_sn);
  std::iota(S.sched.begin(), S.sched.end(), 0);
  std::stable_sort(S.sched.begin(), S.sched.end(), [&](int a, int b) { return S.level[a] < S.level[b]; });
  S.n_levels = S.n_sn ? *std::max_element(S.level.begin(), S.level.end()) + 1 : 0;
  return FG_OK;
}

}  // namespace fg
