// fg_symbolic.cpp -- host-side symbolic analysis of the reduced (pose-side) system.
//
// Replaces the part of gtsam::LevenbergMarquardtOptimizer that chooses an elimination ordering and
// builds the elimination tree (COLAMD + multifrontal) for the graph
// CGraphGT builds (gtsam/gtsam_graph.cpp:1784-1788).  Point landmarks never enter this structure:
// they are eliminated first by the Schur complement (north_star); their co-visibility only
// contributes pose-pose couplings here.
//
// Ordering: frames in key-index order, [X_i V_i B_i] per frame, plane landmarks last.  For the
// sequential VIO/BA graphs the reference produces this is a banded ordering whose fill equals the
// covisibility band; loop closures and planes add a bordered block.  Supernodes are maximal runs of
// consecutive variables with nested structure, capped at kMaxSnCols scalar columns.
#include <algorithm>
#include <cstdio>
#include <numeric>
#include "fg_internal.h"

namespace fg {

static const int kMaxSnCols = 32;

int build_symbolic(fg_ctx* c) {
  HostGraph& h = c->h;
  Symbolic& S = c->sym;
  S = Symbolic();
  // ---- reduced variables and ordering
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
  std::vector<int> pos[T_COUNT];
  for (int t = 0; t < T_COUNT; ++t) pos[t].assign(h.keys[t].size(), -1);
  std::vector<int> voff(nv + 1, 0), vdim(nv);
  for (int p = 0; p < nv; ++p) {
    pos[rv[p].type][rv[p].idx] = p;
    vdim[p] = kDim[rv[p].type];
    voff[p + 1] = voff[p] + vdim[p];
  }
  S.n_r = voff[nv];
  for (int t : {T_POSE, T_VEC3, T_BIAS, T_PLANE}) {
    S.off[t].resize(h.keys[t].size());
    for (size_t i = 0; i < h.keys[t].size(); ++i) S.off[t][i] = voff[pos[t][i]];
  }
  // ---- adjacency (higher-ordered neighbours)
  std::vector<std::vector<int>> hadj(nv);
  auto edge = [&](int a, int b) {
    if (a == b) return;
    if (a > b) std::swap(a, b);
    hadj[a].push_back(b);
  };
  for (size_t f = 0; f < h.bt_i.size(); ++f) edge(pos[T_POSE][h.bt_i[f]], pos[T_POSE][h.bt_j[f]]);
  for (size_t f = 0; f < h.imu_rec.size(); ++f) {
    const int* v = &h.imu_var[6 * f];
    int p[6] = {pos[T_POSE][v[0]], pos[T_VEC3][v[1]], pos[T_POSE][v[2]], pos[T_VEC3][v[3]], pos[T_BIAS][v[4]], pos[T_BIAS][v[5]]};
    for (int a = 0; a < 6; ++a) for (int b = a + 1; b < 6; ++b) edge(p[a], p[b]);
  }
  for (size_t f = 0; f < h.pl_pose.size(); ++f) edge(pos[T_POSE][h.pl_pose[f]], pos[T_PLANE][h.pl_plane[f]]);
  // landmarks: clique over observing poses == star from the lowest-ordered pose (same filled graph)
  {
    const int64_t L = h.count(T_POINT);
    std::vector<int> lowest(L, INT32_MAX);
    for (size_t o = 0; o < h.pj_pose.size(); ++o) {
      int p = pos[T_POSE][h.pj_pose[o]];
      int& lo = lowest[h.pj_point[o]];
      if (p < lo) lo = p;
    }
    for (size_t o = 0; o < h.pj_pose.size(); ++o) {
      int p = pos[T_POSE][h.pj_pose[o]];
      int lo = lowest[h.pj_point[o]];
      if (p != lo) hadj[lo].push_back(p);
    }
  }
  for (auto& a : hadj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
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
      // merge s and cs \ {v}
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
    int v = 0;
    while (v < nv) {
      int first = v, cols = vdim[v];
      while (v + 1 < nv && parent[v] == v + 1 && st[v].size() == st[v + 1].size() + 1 &&
             cols + vdim[v + 1] <= kMaxSnCols) {
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
  std::vector<int> var2sn(nv);
  for (int s = 0; s < S.n_sn; ++s) {
    int c0 = voff[sn_first[s]], nc = voff[sn_last[s] + 1] - c0;
    S.sn_col0[s] = c0; S.sn_ncols[s] = nc;
    int below = 0;
    for (int u : st[sn_last[s]]) below += vdim[u];
    S.sn_nrows[s] = nc + below + 1;
    S.sn_rowptr[s + 1] = S.sn_rowptr[s] + S.sn_nrows[s];
    S.sn_valptr[s + 1] = S.sn_valptr[s] + (int64_t)S.sn_nrows[s] * nc;
    for (int k = 0; k < nc; ++k) S.col2sn[c0 + k] = s;
    for (int v = sn_first[s]; v <= sn_last[s]; ++v) var2sn[v] = s;
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
  // ---- update lists: for descendant d, group its below-rows by target supernode
  std::vector<std::vector<int>> ul(S.n_sn);   // triples (d, a, b)
  for (int d = 0; d < S.n_sn; ++d) {
    const int* r = &S.rowidx[S.sn_rowptr[d]];
    int nr = S.sn_nrows[d] - 1;   // exclude rhs row
    int a = S.sn_ncols[d];
    while (a < nr) {
      int t = S.col2sn[r[a]];
      int b = a + 1;
      while (b < nr && S.col2sn[r[b]] == t) ++b;
      ul[t].push_back(d); ul[t].push_back(a); ul[t].push_back(b);
      a = b;
    }
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
  return FG_OK;
}

}  // namespace fg
