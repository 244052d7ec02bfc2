// fg_internal.h -- host-side graph store, symbolic structure and device views (not part of the ABI).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/fg_abi.h"
#include "fg_factors.cuh"

// Every kernel launch of the library passes its stream through FGS(): the process-wide launch count that bench.py reports as
// `gpu_launches` (fg_debug_counts(ctx, 100)) is counted, not estimated.
extern long long g_fg_launches;
#define FGS(stream) (++g_fg_launches, (stream))

namespace fg {

enum FactorKind { K_PP = 0, K_PV, K_PB, K_BT, K_GE, K_IMU, K_PL, K_COUNT };     // pose-side factor kinds (colour tables)
enum VarType { T_POSE = 0, T_VEC3 = 1, T_BIAS = 2, T_POINT = 3, T_PLANE = 4, T_COUNT = 5 };
static const int kStore[T_COUNT] = {12, 3, 6, 3, 4};   // doubles stored per value
static const int kDim[T_COUNT] = {6, 3, 6, 3, 3};      // tangent dimension

struct VarRef { int type; int idx; };

// ------------------------------------------------------------------ host graph store
struct HostGraph {
  std::unordered_map<fg_key, VarRef> index;
  std::vector<double> val[T_COUNT];          // kStore[t] doubles per value (during an incremental session: the ESTIMATE)
  std::vector<double> lin[T_COUNT];          // incremental session only: the linearisation point theta of every value
  std::vector<fg_key> keys[T_COUNT];

  // factor SoA (host)
  std::vector<int> pp_var;  std::vector<double> pp_mean, pp_info;                 // prior pose: 12, 36
  std::vector<int> pv_var;  std::vector<double> pv_mean, pv_info;                 // prior vec3: 3, 9
  std::vector<int> pb_var;  std::vector<double> pb_mean, pb_info;                 // prior bias: 6, 36
  std::vector<int> pq_var;  std::vector<double> pq_mean, pq_w;                    // prior point: 3, 1 (1/sigma^2)
  std::vector<int> bt_i, bt_j; std::vector<double> bt_meas, bt_info;             // between: 12, 36
  std::vector<int> ge_i, ge_j; std::vector<double> ge_meas, ge_info;             // g2o EdgeSE3: 12, 36 (tangent order [trans, rot])
  std::vector<char> fixed_pose;                                                   // VertexSE3::setFixed (sized on demand; g2o graphs only)
  std::vector<int> imu_var;                                                       // 6 per factor (pose_i, vel_i, pose_j, vel_j, bias_i, bias_j)
  std::vector<ImuRec> imu_rec;
  std::vector<int> pj_pose, pj_point; std::vector<double> pj_uv; std::vector<double> pj_w;  // projection: uv 2, w = 1/sigma^2
  std::vector<int> pj_calib, pj_sensor;
  std::vector<int> pl_pose, pl_plane; std::vector<double> pl_meas, pl_info;      // plane factor: 4, 9
  std::vector<double> calib;    // 9 per id
  std::vector<double> sensor;   // 12 per id
  std::vector<int> se_a, se_b;  // structure-only pose couplings (other ranks' landmark co-visibility)

  int64_t count(int t) const { return (int64_t)keys[t].size(); }
};

// one (tile, leaf member) work item of k_front_syrk
struct __align__(16) FrontRec {
  long long val_off;   // offset of the member's panel in L
  long long pm_off;    // offset of the member's position map
  int nrd, K;          // panel leading dimension and width
  int pad[2];
};

// one descendant update, everything the factorisation kernel needs in a single 32-byte load
struct __align__(16) UpdRec {
  long long val_off;   // offset in L of descendant row a, column 0
  int row_off;         // offset in rowidx of descendant row a
  int nrd;             // descendant leading dimension (rows)
  int nrows_u;         // descendant rows taking part (nrd - a)
  short K, nb;         // descendant width, rows of it inside the target's columns
  int pad[2];
};

// ------------------------------------------------------------------ symbolic structure of the reduced system
struct Symbolic {
  int n_r = 0;                       // reduced scalar dimension (poses, vels, biases, planes)
  int n_sn = 0;                      // supernodes
  int64_t nnz = 0;                   // doubles in panel storage (incl. rhs rows)
  int64_t nnz_S = 0;                 // structural non-zeros of the reduced system before the factorisation (lower + rhs row)
  std::vector<int> off[T_COUNT];     // scalar offset of each reduced variable (T_POINT unused)
  std::vector<int> sn_col0, sn_ncols, sn_nrows, sn_rowptr;  // nrows includes diag rows and the rhs row
  std::vector<int64_t> sn_valptr;
  std::vector<int> rowidx;           // concatenated row lists (global scalar rows; last = n_r = rhs row)
  std::vector<int> col2sn;           // scalar column -> supernode
  std::vector<int> upd_ptr, upd_d, upd_a, upd_b;   // per target supernode: (descendant, row range [a,b) in d)
  std::vector<int> anc_ptr, anc_t, anc_a, anc_b;   // per supernode: (ancestor t, row range [a,b) of this supernode inside t's columns)
  std::vector<int> level, sched;                   // dependency level per supernode; supernodes sorted by level
  int n_levels = 0;
  std::vector<int64_t> cov_ptr; std::vector<int> cov_pose;   // pose co-visibility through landmarks (incl. self)
  int max_nrows = 0, max_ncols = 0;
  double flops_factor = 0;
  // ---- leaf fronts (multifrontal-style path for the separators, fg_front.cu)
  int n_leaves = 0;
  bool use_fronts = false;
  std::vector<int> sn_leaf;                        // nested-dissection leaf of each supernode (-1: separator / plane)
  std::vector<int> leaf_sn_lo, leaf_sn_hi;         // member supernodes of a leaf: [lo, hi)
  std::vector<int> fr_rowptr, fr_rows;             // per leaf: sorted global rows below the leaf (its front), incl. the rhs row
  std::vector<int64_t> fr_uptr;                    // per leaf: offset of its dense nR x nR update matrix in the U buffer
  std::vector<int64_t> pm_ptr; std::vector<int> posmap;   // per member supernode: position of every front row in its row list (-1)
  std::vector<int64_t> pmne_ptr; std::vector<unsigned char> pm_nonempty;   // per member: one flag per 64-row block of the front
  std::vector<int> updr_ptr, updr_d, updr_a, updr_b;   // update lists without leaf -> outside entries
  std::vector<int> tf_ptr, tf_leaf;                // per supernode: leaves whose front must be subtracted from it
  std::vector<int> sched_a, sched_c;               // schedules: leaf members, then everything else (both level sorted)
  std::vector<int> tile_leaf, tile_i, tile_j;      // 64 x 64 tiles of the lower triangles of all fronts
  std::vector<int> tile_mptr; std::vector<FrontRec> tile_mrec;   // per tile: the members that have rows in both of its blocks
  int n_levels_fronts = 0;
  // ---- row-split work units (fg_chol_rs.cu): (supernode, first own row, end own row, blocks of the supernode), level sorted;
  //      the first rs_units_a units are phase A (leaf members, or everything when fronts are off)
  bool rs_ok = false;
  int rs_units_a = 0;
  std::vector<int4> rs_units;
  std::vector<int2> rs_sn_units;                   // per supernode: (global index of its first unit, number of units)
  std::vector<int64_t> rs_moff;                    // per unit: offset of its row maps
  std::vector<short> rs_map;                       // per (unit, update, local row): descendant row (from row a) landing on that row, or -1
  std::vector<signed char> rs_colinv;              // per rsu entry, 32 entries: descendant row (from a) holding target column c, or -1
  // the update list k_chol_rs walks: the list in use (upd_* or updr_*) with descendants wider than 16 columns cut into
  // column slices (UpdRec.val_off points at the slice, pad[0] = 8-column groups of the target it reaches, pad[1] = k0)
  std::vector<int> rsu_ptr, rsu_d, rsu_src; std::vector<UpdRec> rsu_rec;
  std::vector<int64_t> pk_idx;                     // multi-GPU: sorted panel offsets of the entries exchanged by the allreduce
  int n_levels_rs = 0;                             // dependency levels of the factorisation as run (fronts included)
};

// ------------------------------------------------------------------ device view passed to kernels
struct SysView {
  double* L;                 // panel storage (values)
  const int* col2sn;
  const int* sn_col0;
  const int* sn_ncols;
  const int* sn_nrows;
  const int* sn_rowptr;
  const int64_t* sn_valptr;
  const int* rowidx;
  int n_r;
};

// offset of entry (row R, col C), R >= C, in panel storage; ld returned through *ld
__device__ __forceinline__ int64_t sys_find(const SysView& s, int R, int C, int* ld) {
  int sn = s.col2sn[C];
  int c0 = s.sn_col0[sn], nc = s.sn_ncols[sn], nr = s.sn_nrows[sn];
  *ld = nr;
  int r;
  if (R < c0 + nc) {
    r = R - c0;
  } else {
    const int* rows = s.rowidx + s.sn_rowptr[sn];
    int lo = nc, hi = nr - 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (rows[mid] < R) lo = mid + 1; else hi = mid;
    }
    r = lo;
  }
  return s.sn_valptr[sn] + r + (int64_t)(C - c0) * nr;
}

// add a dense block H (da x db, row-major, ld ldh) coupling reduced offsets oa (rows of H) and ob (cols of H)
// into the lower-triangular panel storage.  For oa == ob only the lower triangle is written.
__device__ __forceinline__ void sys_add_block(const SysView& s, int oa, int da, int ob, int db,
                                              const double* H, int ldh) {
  int ld;
  if (oa == ob) {
    int64_t base = sys_find(s, oa, oa, &ld);
    for (int i = 0; i < da; ++i)
      for (int j = 0; j <= i; ++j) atomicAdd(&s.L[base + i + (int64_t)j * ld], H[i * ldh + j]);
  } else if (oa > ob) {
    int64_t base = sys_find(s, oa, ob, &ld);
    for (int i = 0; i < da; ++i)
      for (int j = 0; j < db; ++j) atomicAdd(&s.L[base + i + (int64_t)j * ld], H[i * ldh + j]);
  } else {
    int64_t base = sys_find(s, ob, oa, &ld);   // transposed block: rows = ob.., cols = oa..
    for (int i = 0; i < da; ++i)
      for (int j = 0; j < db; ++j) atomicAdd(&s.L[base + j + (int64_t)i * ld], H[i * ldh + j]);
  }
}

// leaf fronts as seen by phase C of k_chol_reg (all null when the phase has nothing to subtract)
struct FrontView {
  const int* tf_ptr; const int* tf_leaf; const int* fr_rowptr; const int* fr_rows; const int64_t* fr_uptr; const double* U;
};

// ------------------------------------------------------------------ device graph (raw pointers owned by the ctx)
struct DevGraph {
  // values, current and trial
  double* val[T_COUNT] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  double* val_new[T_COUNT] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  int64_t n[T_COUNT] = {0, 0, 0, 0, 0};
  int* off[T_COUNT] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // reduced offsets (not for points)
  // factors
  int n_pp = 0; int* pp_var = nullptr; double* pp_mean = nullptr; double* pp_info = nullptr;
  int n_pv = 0; int* pv_var = nullptr; double* pv_mean = nullptr; double* pv_info = nullptr;
  int n_pb = 0; int* pb_var = nullptr; double* pb_mean = nullptr; double* pb_info = nullptr;
  int n_bt = 0; int* bt_i = nullptr; int* bt_j = nullptr; double* bt_meas = nullptr; double* bt_info = nullptr;
  int n_bt_eblk = 0; int* bt_end = nullptr; int* bt_eblk = nullptr;   // between factor ends (2 f + end) sorted by (pose, factor); block ranges on pose boundaries (0 blocks: coloured path)
  int n_ge = 0; int* ge_i = nullptr; int* ge_j = nullptr; double* ge_meas = nullptr; double* ge_info = nullptr;
  int n_fixed = 0; int* fixed_list = nullptr; char* fixed_pose = nullptr; char* fixed_col = nullptr;   // fixed poses: list, per-pose flag, per reduced column flag
  int pose_chart = 0;               // Pose3 chart (fg_math.cuh): 0 EXPMAP, 1 g2o VertexSE3::oplus, 2 FIRST_ORDER / Rot3 EXPMAP, 3 FIRST_ORDER / Rot3 CAYLEY
  int n_imu = 0; int* imu_var = nullptr; ImuRec* imu_rec = nullptr;
  int n_pl = 0; int* pl_pose = nullptr; int* pl_plane = nullptr; double* pl_meas = nullptr; double* pl_info = nullptr;
  // landmarks: observations sorted by landmark (CSR), plus CSR by pose
  int64_t n_obs = 0;
  int64_t* lm_ptr = nullptr;        // L+1
  int* obs_pose = nullptr;          // M
  double* obs_uv = nullptr;         // 2M
  double* obs_w = nullptr;          // M  (1/sigma^2)
  int64_t* pose_obs_ptr = nullptr;  // P+1
  int64_t* pose_obs = nullptr;      // M  observation ids grouped by pose
  int* obs_point = nullptr;         // M
  int n_oblk = 0; int64_t* oblk_ptr = nullptr;   // observation ranges of the blocks of k_proj_obs / k_lm_backsub_obs (on landmark boundaries)
  double* part = nullptr; double* part2 = nullptr; int part_cap = 0;   // per-block partial sums of the scalar reductions
  int n_dup = 0; int* dup_prim = nullptr; int* dup_sec = nullptr;   // extra factors on an already seen (pose, landmark) pair: (primary, secondary) observation ids
  double* lm_prior_mean = nullptr;  // 3L (NaN weight = no prior)
  double* lm_prior_w = nullptr;     // L
  double* W = nullptr;              // M x 18 (AoS): w Jp^T Jl per observation
  double* ul = nullptr;             // 3 L : G^-1 g_l with V + lambda I = G G^T (so that W Vinv g = Z u)
  double* Cf = nullptr;             // 6 L : upper-triangular C = G^-T (00 01 02 11 12 22); Vinv = C C^T
  double* Zp = nullptr;             // 3 planes x M x 6 (plane k = column k of Z_o = W_o C_l, POSE-major order): W Vinv W^T = Z Z^T
  int* obs_ppos = nullptr;          // M : position of observation o in pose-major order (index into Zp / pose_obs)
  int* pz_point = nullptr;          // M : landmark of the k-th pose-major observation
  // Schur tiles (fg_schur.cu): 16 x 16 pose tiles of the reduced Hessian, landmarks cut into chunks of schur_ch
  int64_t n_pairs = 0;              // observation pairs (a, b) of a common landmark with pose(b) <= pose(a): Schur work units
  int schur_ch = 0;                 // landmarks per table word (32)
  int n_tiles = 0;
  int4* tile_desc = nullptr;        // (row pose group, column pose group, first chunk, end chunk), heaviest first
  int* pc_lo = nullptr;             // P : first chunk in which the pose has an observation
  int* pc_n = nullptr;              // P : number of chunks from pc_lo to its last one
  int64_t* pc_ptr = nullptr;        // P : offset of the pose's chunk entries
  uint2* pc_ent = nullptr;          // per (pose, chunk): x = pose-major index of its first observation there, y = landmark bit mask
  double* V = nullptr;              // 6 L  : upper of sum Jl^T Jl w + prior
  double* gl = nullptr;             // 3 L
  double* Vinv = nullptr;           // 6 L
  double* tl = nullptr;             // 3 L  : sum_o W_o^T delta_p (back-substitution scratch)
  double* calib = nullptr;          // 9
  struct ProjCal { double K[9], S[12]; } cal;   // host copy of calib / sensor: passed to the projection kernels by value (constant bank)
  int n_cal = 1;                    // distinct (Cal3DS2, body_P_sensor) pairs among the projection factors
  ProjCal* cals = nullptr;          // n_cal entries; only used (with obs_cal) when n_cal > 1
  unsigned char* obs_cal = nullptr; // M : index into cals per observation (landmark-sorted order)
  double* sensor = nullptr;         // 12
  // reduced system
  double* L = nullptr;              // panels being factored
  double* U0 = nullptr;             // undamped pose-side Hessian in panel layout (lambda independent)
  double* g_r = nullptr;            // reduced gradient J^T Omega r  (n_r)
  double* delta = nullptr;          // n_r
  double* scal = nullptr;           // scalars: [0] chi2, [1] g^T delta, [2] |delta|^2, [3] new chi2
  int* flags = nullptr;             // per supernode epoch flags
  int* status = nullptr;            // factorisation status (0 ok)
  // symbolic on device
  int *col2sn = nullptr, *sn_col0 = nullptr, *sn_ncols = nullptr, *sn_nrows = nullptr, *sn_rowptr = nullptr, *rowidx = nullptr;
  int64_t* sn_valptr = nullptr;
  int *upd_ptr = nullptr, *upd_d = nullptr, *upd_a = nullptr, *upd_b = nullptr;
  // leaf fronts
  int *updr_ptr = nullptr, *updr_d = nullptr;
  int *rsu_ptr = nullptr, *rsu_d = nullptr; UpdRec* rsu_rec = nullptr;
  int *sched_a = nullptr, *sched_c = nullptr;
  int *fr_rowptr = nullptr, *fr_rows = nullptr; int64_t* fr_uptr = nullptr;
  int64_t* pm_ptr = nullptr; int* posmap = nullptr; int64_t* pmne_ptr = nullptr; unsigned char* pm_nonempty = nullptr;
  int *leaf_sn_lo = nullptr, *leaf_sn_hi = nullptr;
  int *tf_ptr = nullptr, *tf_leaf = nullptr;
  int *tile_leaf = nullptr, *tile_i = nullptr, *tile_j = nullptr, *tile_mptr = nullptr; FrontRec* tile_mrec = nullptr;
  double* U = nullptr;              // dense update matrices of all leaves
  int *anc_ptr = nullptr, *anc_t = nullptr, *anc_a = nullptr, *anc_b = nullptr;
  int *sched = nullptr;
  int *bs_order = nullptr;          // backward-solve processing order (reverse level order)
  int4* rs_units = nullptr; int64_t* rs_moff = nullptr; short* rs_map = nullptr; signed char* rs_colinv = nullptr; int* rs_done = nullptr;   // rs_done: one done flag per unit
  int2* rs_sn_units = nullptr;   // per supernode: (first unit, number of units)
  int64_t n_pk = 0; int64_t* pk_idx = nullptr; double* pk_buf = nullptr;   // packed exchange buffer: n_pk entries + 1 scalar (chi2)
  int *flags2 = nullptr;            // per supernode epoch flags of the backward solve
  int *counters = nullptr;          // [0] next schedule slot (factor), [1] next schedule slot (backsolve), [2] phase C slot, [5] relinearised variables
};

}  // namespace fg

// ------------------------------------------------------------------ the context
struct fg_ctx {
  int device = 0, rank = 0, nranks = 1;
  std::string err;
  fg::HostGraph h;
  fg::Symbolic sym;
  fg::DevGraph d;
  bool finalized = false;
  bool values_dirty = false;     // host values newer than device
  bool device_newer = false;     // device values newer than host
  // incremental session (fg_update_incremental): d.val holds theta, d.val_new the estimate; h.val / h.lin mirror them
  int pose_chart = 0;            // fg_set_pose_chart (GTSAM graphs; the g2o back-end uses its own)
  std::vector<int> color_ptr[fg::K_COUNT];   // per pose-side factor kind: offsets of its colour classes in the (colour-sorted) device arrays
  bool inc_active = false;
  int inc_updates = 0;
  int64_t inc_known[fg::T_COUNT] = {0, 0, 0, 0, 0};   // variables per type at the previous update
  std::vector<std::pair<void*, size_t>> pool;          // device buffers of the previous build, reused by the next (a graph that grows by one frame per update)
  cudaStream_t stream = nullptr;
  int epoch = 0;
  int num_sms = 148;
  void* nccl_comm = nullptr;
  cudaEvent_t kev[4] = {nullptr, nullptr, nullptr, nullptr};   // around k_proj_obs<JAC> and k_schur_tiles (roofline timing)
  std::vector<void*> allocs;
  std::vector<size_t> alloc_bytes;
};

namespace fg {
// host side of the Schur tile kernel (fg_api.cu: build_schur_tables)
struct SchurTables {
  int ch = 32; int64_t npairs = 0;
  std::vector<int> ppos, pzp, pc_lo, pc_n; std::vector<int64_t> pc_ptr; std::vector<uint2> pc_ent; std::vector<int4> tiles;
};
int build_schur_tables(fg_ctx* c, int64_t L, int64_t M, int64_t P, const std::vector<int64_t>& lm_ptr, const std::vector<int64_t>& pose_ptr,
                       const std::vector<int>& pose_nprim, const std::vector<int64_t>& pose_obs, const std::vector<int>& s_point, SchurTables& T);
// host.cpp
int build_symbolic(fg_ctx* c);
// kernels (launch wrappers), all asynchronous on c->stream
void launch_linearize(fg_ctx* c);                         // U0, g_r, V, gl, W, chi2 -> scal[0]
void launch_build_and_schur(fg_ctx* c, double lambda);    // L = U0 + lambda I - W V'^-1 W^T ; rhs row = -(g_red)
void launch_schur(fg_ctx* c, double lambda);              // fg_schur.cu: the landmark part of the line above
// fg_chol_rs.cu: cholesky (row-split units, width <= 32); the rhs row makes it the forward solve too
void launch_factor_rs(fg_ctx* c);
void launch_front_syrk(fg_ctx* c);                        // fg_front.cu: dense update matrix of every leaf onto its front
void launch_backsolve(fg_ctx* c);
void launch_marginal(fg_ctx* c, int col0, int dim, double* work, double* out36);   // fg_chol.cu: [S^-1] block of one variable from the factor in d.L                         // backward solve -> delta
void launch_retract_error(fg_ctx* c, double lambda);      // val_new = val (+) delta (incl. landmarks), scal[1..3]
void launch_pack(fg_ctx* c, bool with_chi2);               // multi-GPU: gather the exchanged entries of d.L (+ scal[0]) into d.pk_buf
void launch_unpack(fg_ctx* c, bool with_chi2);             // ... and scatter the reduced values back
void launch_max_diag(fg_ctx* c, double* d_out);          // g2o computeLambdaInit: max diagonal entry of the assembled Hessian (free variables) -> *d_out
void launch_inc_gate(fg_ctx* c, double threshold, int* d_count);   // move theta to the estimate where |delta| >= threshold
void launch_error_only(fg_ctx* c, bool trial);            // chi2 of val (or val_new) -> scal[0] (or scal[3])
void launch_preintegrate(int n, const int* d_off, const double* d_imu, double dt, const ImuParamsDev* d_par,
                         const double* d_bias, fg_pim* d_out, cudaStream_t st);
}  // namespace fg
