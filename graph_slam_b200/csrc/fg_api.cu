// fg_api.cu -- C ABI (include/fg_abi.h): graph store, device upload, Levenberg-Marquardt control.
//
// The LM control flow restates gtsam::LevenbergMarquardtOptimizer with default parameters as invoked by
// CGraphGT::optimizeGraphBatch (gtsam/gtsam_graph.cpp:1784-1788); see SURVEY.md A.7.  All arithmetic over
// factors and variables runs in the CUDA kernels of fg_kernels.cu / fg_chol.cu; the host only takes the
// accept/reject decision from four scalars per trial.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <dlfcn.h>
#include <limits>
#include "fg_internal.h"

using namespace fg;

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      char buf[512];                                                                     \
      snprintf(buf, sizeof buf, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      if (c) c->err = buf;                                                               \
      return FG_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

static int fail(fg_ctx* c, int code, const char* msg) {
  if (c) c->err = msg;
  return code;
}

// ------------------------------------------------------------------ NCCL (loaded at run time; single-GPU use needs no NCCL)
namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef int (*p_ncclGetUniqueId)(nccl_uid*);
typedef int (*p_ncclCommInitRank)(void**, int, nccl_uid, int);
typedef int (*p_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*p_ncclCommDestroy)(void*);
struct NcclApi {
  void* h = nullptr;
  p_ncclGetUniqueId GetUniqueId = nullptr;
  p_ncclCommInitRank CommInitRank = nullptr;
  p_ncclAllReduce AllReduce = nullptr;
  p_ncclCommDestroy CommDestroy = nullptr;
} g_nccl;
bool load_nccl() {
  if (g_nccl.h) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return false;
  g_nccl.GetUniqueId = (p_ncclGetUniqueId)dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (p_ncclCommInitRank)dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.AllReduce = (p_ncclAllReduce)dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.CommDestroy = (p_ncclCommDestroy)dlsym(g_nccl.h, "ncclCommDestroy");
  return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy;
}
const int kNcclDouble = 8, kNcclSum = 0;   // ncclFloat64, ncclSum
}  // namespace

// ------------------------------------------------------------------ device memory helpers
// Device buffers come from the context's pool when a buffer of the previous build fits (a graph that grows by one frame
// per fg_update_incremental re-uploads ~100 arrays per frame: no cudaMalloc / cudaFree churn); fresh buffers get 25 %
// headroom for the same reason.
static int dev_alloc(fg_ctx* c, void** dst, size_t bytes) {
  size_t best = (size_t)-1, bi = 0;
  for (size_t i = 0; i < c->pool.size(); ++i)
    if (c->pool[i].second >= bytes && c->pool[i].second <= 2 * bytes + 4096 && c->pool[i].second < best) { best = c->pool[i].second; bi = i; }
  if (best != (size_t)-1) {
    *dst = c->pool[bi].first;
    c->allocs.push_back(*dst); c->alloc_bytes.push_back(best);
    c->pool[bi] = c->pool.back(); c->pool.pop_back();
    return FG_OK;
  }
  const size_t cap = bytes + bytes / 4 + 256;
  CK(cudaMalloc(dst, cap));
  c->allocs.push_back(*dst); c->alloc_bytes.push_back(cap);
  return FG_OK;
}
template <typename T>
static int dev_upload(fg_ctx* c, T** dst, const T* src, size_t n) {
  *dst = nullptr;
  if (n == 0) n = 1;
  int rc = dev_alloc(c, (void**)dst, n * sizeof(T));
  if (rc != FG_OK) return rc;
  if (src) CK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  else CK(cudaMemsetAsync(*dst, 0, n * sizeof(T), c->stream));
  return FG_OK;
}
template <typename T>
static int dev_upload(fg_ctx* c, T** dst, const std::vector<T>& v) {
  return dev_upload(c, dst, v.empty() ? (const T*)nullptr : v.data(), v.size());
}
// retire the buffers of the current build into the pool (keep = true) or release everything
static void dev_free_all(fg_ctx* c, bool keep = false) {
  if (keep) {
    for (size_t i = 0; i < c->allocs.size(); ++i) c->pool.push_back({c->allocs[i], c->alloc_bytes[i]});
  } else {
    for (void* p : c->allocs) cudaFree(p);
    for (auto& p : c->pool) cudaFree(p.first);
    c->pool.clear();
  }
  c->allocs.clear(); c->alloc_bytes.clear();
  c->d = DevGraph();
}
// buffers of the previous build that the new one did not take
static void pool_trim(fg_ctx* c) {
  for (auto& p : c->pool) cudaFree(p.first);
  c->pool.clear();
}

// During an incremental session the device holds two states per variable: d.val = linearisation point theta (host mirror
// h.lin), d.val_new = estimate (host mirror h.val -- what fg_get_value returns).  Otherwise d.val <-> h.val.
static int pull_values(fg_ctx* c) {
  if (!c->device_newer) return FG_OK;
  for (int t = 0; t < T_COUNT; ++t)
    if (c->d.n[t]) {
      const size_t nb = sizeof(double) * c->h.val[t].size();
      if (c->inc_active) {
        CK(cudaMemcpyAsync(c->h.val[t].data(), c->d.val_new[t], nb, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(c->h.lin[t].data(), c->d.val[t], nb, cudaMemcpyDeviceToHost, c->stream));
      } else {
        CK(cudaMemcpyAsync(c->h.val[t].data(), c->d.val[t], nb, cudaMemcpyDeviceToHost, c->stream));
      }
    }
  CK(cudaStreamSynchronize(c->stream));
  c->device_newer = false;
  return FG_OK;
}
static int push_values(fg_ctx* c) {
  if (!c->finalized || !c->values_dirty) return FG_OK;
  for (int t = 0; t < T_COUNT; ++t)
    if (c->d.n[t]) {
      const size_t nb = sizeof(double) * c->h.val[t].size();
      if (c->inc_active) {
        CK(cudaMemcpyAsync(c->d.val[t], c->h.lin[t].data(), nb, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->d.val_new[t], c->h.val[t].data(), nb, cudaMemcpyHostToDevice, c->stream));
      } else {
        CK(cudaMemcpyAsync(c->d.val[t], c->h.val[t].data(), nb, cudaMemcpyHostToDevice, c->stream));
      }
    }
  c->values_dirty = false;
  return FG_OK;
}
// leave the incremental session: the estimate becomes the (single) state -- the reference copies calculateEstimate() into
// mp_node_values (gtsam_graph.cpp:1773) and every batch entry point starts from there
static int end_incremental(fg_ctx* c) {
  if (!c->inc_active) return FG_OK;
  if (c->finalized && !c->values_dirty) {
    for (int t = 0; t < T_COUNT; ++t)
      if (c->d.n[t]) CK(cudaMemcpyAsync(c->d.val[t], c->d.val_new[t], sizeof(double) * c->h.val[t].size(), cudaMemcpyDeviceToDevice, c->stream));
  }
  for (int t = 0; t < T_COUNT; ++t) { c->h.lin[t].clear(); c->h.lin[t].shrink_to_fit(); }
  c->inc_active = false;
  c->inc_updates = 0;
  return FG_OK;
}

// ------------------------------------------------------------------ lifetime
extern "C" fg_ctx* fg_create(int device, int rank, int nranks) {
  if (device == -1) {
    // detached context: graph construction and symbolic analysis only (host logic tests);
    // every numeric entry point fails with FG_ERR_CUDA -- there is no CPU solver in this library.
    fg_ctx* c = new fg_ctx();
    c->device = -1; c->rank = rank; c->nranks = nranks < 1 ? 1 : nranks;
    c->h.calib.assign(9, 0.0);
    c->h.sensor = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    return c;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  fg_ctx* c = new fg_ctx();
  c->device = device; c->rank = rank; c->nranks = nranks < 1 ? 1 : nranks;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
  for (auto& e : c->kev) cudaEventCreate(&e);
  c->h.calib.assign(9, 0.0);
  c->h.sensor = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
  return c;
}
extern "C" void fg_destroy(fg_ctx* c) {
  if (!c) return;
  if (c->device < 0) { delete c; return; }
  cudaSetDevice(c->device);
  if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm);
  dev_free_all(c);
  for (auto& e : c->kev) if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}
extern "C" const char* fg_last_error(fg_ctx* c) { return c ? c->err.c_str() : "null context"; }
extern "C" int fg_abi_version(void) { return 1; }

// ------------------------------------------------------------------ values
static int add_value(fg_ctx* c, fg_key key, int type, const double* v) {
  if (!c || !v) return fail(c, FG_ERR_INVALID, "null argument");
  if (c->h.index.count(key)) return fail(c, FG_ERR_DUPLICATE_KEY, "key already exists in Values");
  if (pull_values(c) != FG_OK) return FG_ERR_CUDA;
  int idx = (int)c->h.keys[type].size();
  c->h.index[key] = VarRef{type, idx};
  c->h.keys[type].push_back(key);
  c->h.val[type].insert(c->h.val[type].end(), v, v + kStore[type]);
  if (c->inc_active) c->h.lin[type].insert(c->h.lin[type].end(), v, v + kStore[type]);     // a new variable: theta = estimate = initial value
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_add_pose(fg_ctx* c, fg_key k, const double T[12]) { return add_value(c, k, T_POSE, T); }
extern "C" int fg_add_vec3(fg_ctx* c, fg_key k, const double v[3]) { return add_value(c, k, T_VEC3, v); }
extern "C" int fg_add_bias(fg_ctx* c, fg_key k, const double b[6]) { return add_value(c, k, T_BIAS, b); }
extern "C" int fg_add_point(fg_ctx* c, fg_key k, const double p[3]) { return add_value(c, k, T_POINT, p); }
extern "C" int fg_add_plane(fg_ctx* c, fg_key k, const double pl[4]) {
  if (!pl) return fail(c, FG_ERR_INVALID, "null argument");
  double n = std::sqrt(pl[0] * pl[0] + pl[1] * pl[1] + pl[2] * pl[2]);
  if (!(n > 0)) return fail(c, FG_ERR_INVALID, "zero plane normal");
  double q[4] = {pl[0] / n, pl[1] / n, pl[2] / n, pl[3]};   // OrientedPlane3(a,b,c,d): Unit3(a,b,c), d kept
  return add_value(c, k, T_PLANE, q);
}
extern "C" int fg_add_points(fg_ctx* c, int64_t n, const fg_key* keys, const double* p3) {
  if (!c || !keys || !p3 || n < 0) return fail(c, FG_ERR_INVALID, "bad argument");
  if (pull_values(c) != FG_OK) return FG_ERR_CUDA;
  // all or nothing: a duplicate (against Values or inside the batch) leaves the graph untouched
  const size_t n0 = c->h.keys[T_POINT].size();
  c->h.index.reserve(c->h.index.size() + (size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    if (!c->h.index.emplace(keys[i], VarRef{T_POINT, (int)(n0 + i)}).second) {
      for (int64_t j = 0; j < i; ++j) c->h.index.erase(keys[j]);
      return fail(c, FG_ERR_DUPLICATE_KEY, "key already exists in Values");
    }
  }
  c->h.keys[T_POINT].insert(c->h.keys[T_POINT].end(), keys, keys + n);
  c->h.val[T_POINT].insert(c->h.val[T_POINT].end(), p3, p3 + 3 * n);
  if (c->inc_active) c->h.lin[T_POINT].insert(c->h.lin[T_POINT].end(), p3, p3 + 3 * n);
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_update_value(fg_ctx* c, fg_key key, const double* v) {
  if (!c || !v) return fail(c, FG_ERR_INVALID, "null argument");
  auto it = c->h.index.find(key);
  if (it == c->h.index.end()) return fail(c, FG_ERR_UNKNOWN_KEY, "key does not exist in Values");
  if (pull_values(c) != FG_OK) return FG_ERR_CUDA;
  int t = it->second.type;
  std::copy(v, v + kStore[t], c->h.val[t].begin() + (size_t)it->second.idx * kStore[t]);
  if (c->inc_active) std::copy(v, v + kStore[t], c->h.lin[t].begin() + (size_t)it->second.idx * kStore[t]);   // Values::update: both states
  c->values_dirty = true;
  return FG_OK;
}
extern "C" int fg_exists(fg_ctx* c, fg_key key) { return (c && c->h.index.count(key)) ? 1 : 0; }
extern "C" int fg_get_value(fg_ctx* c, fg_key key, double* out, int* n_out) {
  if (!c || !out) return fail(c, FG_ERR_INVALID, "null argument");
  auto it = c->h.index.find(key);
  if (it == c->h.index.end()) return fail(c, FG_ERR_UNKNOWN_KEY, "key does not exist in Values");
  if (pull_values(c) != FG_OK) return FG_ERR_CUDA;
  int t = it->second.type;
  const double* src = c->h.val[t].data() + (size_t)it->second.idx * kStore[t];
  std::copy(src, src + kStore[t], out);
  if (n_out) *n_out = kStore[t];
  return FG_OK;
}
extern "C" int64_t fg_num_values(fg_ctx* c, int type) {
  if (!c || type < 0 || type >= T_COUNT) return -1;
  return c->h.count(type);
}
extern "C" int fg_get_values(fg_ctx* c, int type, double* out) {
  if (!c || !out || type < 0 || type >= T_COUNT) return fail(c, FG_ERR_INVALID, "bad argument");
  size_t nb = sizeof(double) * c->h.val[type].size();
  if (c->finalized && c->device_newer && c->d.n[type]) {
    CK(cudaMemcpyAsync(out, c->inc_active ? c->d.val_new[type] : c->d.val[type], nb, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  } else {
    std::memcpy(out, c->h.val[type].data(), nb);
  }
  return FG_OK;
}
extern "C" int fg_set_values(fg_ctx* c, int type, const double* in) {
  if (!c || !in || type < 0 || type >= T_COUNT) return fail(c, FG_ERR_INVALID, "bad argument");
  const size_t nb = sizeof(double) * c->h.val[type].size();
  if (c->inc_active) { if (pull_values(c) != FG_OK) return FG_ERR_CUDA; end_incremental(c); }
  if (c->finalized && c->d.n[type] && !c->values_dirty) {
    // fast path: one host -> device copy of this array straight from the caller's buffer (which may be pinned); the
    // device copy becomes the authoritative one, the host mirror is refreshed lazily by pull_values
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->d.val[type], in, nb, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));        // the caller owns `in` again on return
    c->device_newer = true;
    return FG_OK;
  }
  if (pull_values(c) != FG_OK) return FG_ERR_CUDA;
  std::memcpy(c->h.val[type].data(), in, nb);
  c->values_dirty = true;
  return FG_OK;
}

// ------------------------------------------------------------------ factors
static int find_var(fg_ctx* c, fg_key key, int type, int* idx) {
  auto it = c->h.index.find(key);
  if (it == c->h.index.end()) return fail(c, FG_ERR_UNKNOWN_KEY, "factor refers to a key that is not in Values");
  if (it->second.type != type) return fail(c, FG_ERR_INVALID, "factor key has the wrong value type");
  *idx = it->second.idx;
  return FG_OK;
}
#define FIND(key, type, out) do { int rc_ = find_var(c, key, type, out); if (rc_ != FG_OK) return rc_; } while (0)

extern "C" int fg_add_prior_pose(fg_ctx* c, fg_key key, const double T[12], const double info[36]) {
  if (!c || !T || !info) return fail(c, FG_ERR_INVALID, "null argument");
  int v; FIND(key, T_POSE, &v);
  c->h.pp_var.push_back(v);
  c->h.pp_mean.insert(c->h.pp_mean.end(), T, T + 12);
  c->h.pp_info.insert(c->h.pp_info.end(), info, info + 36);
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_add_prior_vec3(fg_ctx* c, fg_key key, const double m[3], const double info[9]) {
  if (!c || !m || !info) return fail(c, FG_ERR_INVALID, "null argument");
  int v; FIND(key, T_VEC3, &v);
  c->h.pv_var.push_back(v);
  c->h.pv_mean.insert(c->h.pv_mean.end(), m, m + 3);
  c->h.pv_info.insert(c->h.pv_info.end(), info, info + 9);
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_add_prior_bias(fg_ctx* c, fg_key key, const double m[6], const double info[36]) {
  if (!c || !m || !info) return fail(c, FG_ERR_INVALID, "null argument");
  int v; FIND(key, T_BIAS, &v);
  c->h.pb_var.push_back(v);
  c->h.pb_mean.insert(c->h.pb_mean.end(), m, m + 6);
  c->h.pb_info.insert(c->h.pb_info.end(), info, info + 36);
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_add_prior_point(fg_ctx* c, fg_key key, const double m[3], double sigma) {
  if (!c || !m || !(sigma > 0)) return fail(c, FG_ERR_INVALID, "bad argument");
  int v; FIND(key, T_POINT, &v);
  c->h.pq_var.push_back(v);
  c->h.pq_mean.insert(c->h.pq_mean.end(), m, m + 3);
  c->h.pq_w.push_back(1.0 / (sigma * sigma));
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_add_prior_points(fg_ctx* c, int64_t n, const fg_key* keys, const double* m3, double sigma) {
  if (!c || !keys || !m3 || n < 0 || !(sigma > 0)) return fail(c, FG_ERR_INVALID, "bad argument");
  const size_t n0 = c->h.pq_var.size();
  for (int64_t i = 0; i < n; ++i) {
    int v;
    const int rc = find_var(c, keys[i], T_POINT, &v);
    if (rc != FG_OK) { c->h.pq_var.resize(n0); return rc; }      // all or nothing
    c->h.pq_var.push_back(v);
  }
  c->h.pq_mean.insert(c->h.pq_mean.end(), m3, m3 + 3 * n);
  c->h.pq_w.insert(c->h.pq_w.end(), (size_t)n, 1.0 / (sigma * sigma));
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_add_between(fg_ctx* c, fg_key k1, fg_key k2, const double T[12], const double info[36]) {
  if (!c || !T || !info) return fail(c, FG_ERR_INVALID, "null argument");
  int a, b; FIND(k1, T_POSE, &a); FIND(k2, T_POSE, &b);
  c->h.bt_i.push_back(a); c->h.bt_j.push_back(b);
  c->h.bt_meas.insert(c->h.bt_meas.end(), T, T + 12);
  c->h.bt_info.insert(c->h.bt_info.end(), info, info + 36);
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_set_pose_chart(fg_ctx* c, int chart) {
  if (!c) return FG_ERR_INVALID;
  if (chart != FG_CHART_EXPMAP && chart != FG_CHART_FIRST_ORDER_EXPMAP && chart != FG_CHART_FIRST_ORDER_CAYLEY) return fail(c, FG_ERR_INVALID, "unknown Pose3 chart");
  c->pose_chart = chart;
  c->d.pose_chart = c->d.n_ge ? 1 : chart;      // takes effect at once on a finalized graph too
  return FG_OK;
}
extern "C" int fg_add_g2o_edge(fg_ctx* c, fg_key k1, fg_key k2, const double T[12], const double info[36]) {
  if (!c || !T || !info) return fail(c, FG_ERR_INVALID, "null argument");
  int a, b; FIND(k1, T_POSE, &a); FIND(k2, T_POSE, &b);
  c->h.ge_i.push_back(a); c->h.ge_j.push_back(b);
  c->h.ge_meas.insert(c->h.ge_meas.end(), T, T + 12);
  c->h.ge_info.insert(c->h.ge_info.end(), info, info + 36);
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_set_fixed(fg_ctx* c, fg_key key, int fixed) {
  if (!c) return FG_ERR_INVALID;
  int v; FIND(key, T_POSE, &v);
  if (c->h.fixed_pose.size() <= (size_t)v) c->h.fixed_pose.resize((size_t)v + 1, 0);
  c->h.fixed_pose[v] = fixed ? 1 : 0;
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_add_structure_edges(fg_ctx* c, int64_t n, const fg_key* ka, const fg_key* kb) {
  if (!c || !ka || !kb || n < 0) return fail(c, FG_ERR_INVALID, "bad argument");
  for (int64_t i = 0; i < n; ++i) {
    int a, b; FIND(ka[i], T_POSE, &a); FIND(kb[i], T_POSE, &b);
    c->h.se_a.push_back(a); c->h.se_b.push_back(b);
  }
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_set_calibration(fg_ctx* c, int id, const double K[9]) {
  if (!c || !K || id < 0) return fail(c, FG_ERR_INVALID, "bad argument");
  if ((size_t)(id + 1) * 9 > c->h.calib.size()) c->h.calib.resize((size_t)(id + 1) * 9, 0.0);
  std::copy(K, K + 9, c->h.calib.begin() + (size_t)id * 9);
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_set_sensor(fg_ctx* c, int id, const double T[12]) {
  if (!c || !T || id < 0) return fail(c, FG_ERR_INVALID, "bad argument");
  if ((size_t)(id + 1) * 12 > c->h.sensor.size()) c->h.sensor.resize((size_t)(id + 1) * 12, 0.0);
  std::copy(T, T + 12, c->h.sensor.begin() + (size_t)id * 12);
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_add_projections(fg_ctx* c, int64_t n, const fg_key* kp, const fg_key* kq, const double* uv2,
                                  double sigma, int calib_id, int sensor_id) {
  if (!c || !kp || !kq || !uv2 || n < 0 || !(sigma > 0) || calib_id < 0 || sensor_id < 0)
    return fail(c, FG_ERR_INVALID, "bad argument");
  if ((size_t)(calib_id + 1) * 9 > c->h.calib.size() || (size_t)(sensor_id + 1) * 12 > c->h.sensor.size())
    return fail(c, FG_ERR_INVALID, "unknown calibration or sensor id");
  size_t base = c->h.pj_pose.size();
  c->h.pj_pose.resize(base + n); c->h.pj_point.resize(base + n);
  for (int64_t i = 0; i < n; ++i) {
    int a, b;
    int rc = find_var(c, kp[i], T_POSE, &a);
    if (rc == FG_OK) rc = find_var(c, kq[i], T_POINT, &b);
    if (rc != FG_OK) { c->h.pj_pose.resize(base); c->h.pj_point.resize(base); return rc; }
    c->h.pj_pose[base + i] = a; c->h.pj_point[base + i] = b;
  }
  c->h.pj_uv.insert(c->h.pj_uv.end(), uv2, uv2 + 2 * n);
  c->h.pj_w.insert(c->h.pj_w.end(), (size_t)n, 1.0 / (sigma * sigma));
  c->h.pj_calib.insert(c->h.pj_calib.end(), (size_t)n, calib_id);
  c->h.pj_sensor.insert(c->h.pj_sensor.end(), (size_t)n, sensor_id);
  c->finalized = false;
  return FG_OK;
}
extern "C" int fg_add_projection(fg_ctx* c, fg_key kp, fg_key kq, const double uv[2], double sigma, int calib_id, int sensor_id) {
  return fg_add_projections(c, 1, &kp, &kq, uv, sigma, calib_id, sensor_id);
}

// symmetric positive definite inverse by Cholesky (host, n <= 15)
static bool spd_inverse(const double* A, int n, double* Ai) {
  double L[225], Li[225];
  for (int i = 0; i < n * n; ++i) { L[i] = 0; Li[i] = 0; }
  for (int j = 0; j < n; ++j) {
    double s = A[j * n + j];
    for (int k = 0; k < j; ++k) s -= L[j * n + k] * L[j * n + k];
    if (!(s > 0)) return false;
    L[j * n + j] = std::sqrt(s);
    for (int i = j + 1; i < n; ++i) {
      double t = 0.5 * (A[i * n + j] + A[j * n + i]);
      for (int k = 0; k < j; ++k) t -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = t / L[j * n + j];
    }
  }
  for (int j = 0; j < n; ++j) {       // Li = L^-1 (lower), column by column
    Li[j * n + j] = 1.0 / L[j * n + j];
    for (int i = j + 1; i < n; ++i) {
      double t = 0;
      for (int k = j; k < i; ++k) t -= L[i * n + k] * Li[k * n + j];
      Li[i * n + j] = t / L[i * n + i];
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double t = 0;
      for (int k = i; k < n; ++k) t += Li[k * n + i] * Li[k * n + j];
      Ai[i * n + j] = t; Ai[j * n + i] = t;
    }
  return true;
}

extern "C" int fg_add_plane_factor(fg_ctx* c, fg_key kpose, fg_key kplane, const double z[4], const double cov[9]) {
  if (!c || !z || !cov) return fail(c, FG_ERR_INVALID, "null argument");
  int a, b; FIND(kpose, T_POSE, &a); FIND(kplane, T_PLANE, &b);
  double info[9];
  if (!spd_inverse(cov, 3, info)) return fail(c, FG_ERR_INVALID, "plane covariance is not positive definite");
  double n = std::sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
  if (!(n > 0)) return fail(c, FG_ERR_INVALID, "zero plane normal");
  double q[4] = {z[0] / n, z[1] / n, z[2] / n, z[3]};
  c->h.pl_pose.push_back(a); c->h.pl_plane.push_back(b);
  c->h.pl_meas.insert(c->h.pl_meas.end(), q, q + 4);
  c->h.pl_info.insert(c->h.pl_info.end(), info, info + 9);
  c->finalized = false;
  return FG_OK;
}

extern "C" int fg_add_imu(fg_ctx* c, const fg_key keys[6], const fg_pim* pim) {
  if (!c || !keys || !pim) return fail(c, FG_ERR_INVALID, "null argument");
  int v[6];
  FIND(keys[0], T_POSE, &v[0]); FIND(keys[1], T_VEC3, &v[1]); FIND(keys[2], T_POSE, &v[2]);
  FIND(keys[3], T_VEC3, &v[3]); FIND(keys[4], T_BIAS, &v[4]); FIND(keys[5], T_BIAS, &v[5]);
  ImuRec r;
  r.dt = pim->dt;
  std::copy(pim->preint, pim->preint + 9, r.preint);
  std::copy(pim->H_ba, pim->H_ba + 27, r.Hba);
  std::copy(pim->H_bg, pim->H_bg + 27, r.Hbg);
  std::copy(pim->bias_hat, pim->bias_hat + 6, r.bias_hat);
  std::copy(pim->gravity, pim->gravity + 3, r.gravity);
  if (!spd_inverse(pim->cov, 15, r.info)) return fail(c, FG_ERR_INVALID, "preintMeasCov is not positive definite");
  c->h.imu_var.insert(c->h.imu_var.end(), v, v + 6);
  c->h.imu_rec.push_back(r);
  c->finalized = false;
  return FG_OK;
}

// ------------------------------------------------------------------ IMU preintegration
extern "C" int fg_preintegrate(fg_ctx* c, int n, const int* offsets, const double* imu6, double dt,
                               const fg_imu_params* params, const double* bias_hat6, fg_pim* out) {
  if (n < 0 || !offsets || !imu6 || !params || !bias_hat6 || !out || !(dt > 0)) return fail(c, FG_ERR_INVALID, "bad argument");
  if (n == 0) return FG_OK;
  if (c && c->device < 0) return fail(c, FG_ERR_CUDA, "detached context (device -1): no CUDA device");
  cudaStream_t st = c ? c->stream : (cudaStream_t)0;
  if (c) cudaSetDevice(c->device);
  int ns = offsets[n];
  ImuParamsDev hp;
  std::memcpy(hp.acc_cov, params->acc_cov, sizeof hp.acc_cov);
  std::memcpy(hp.gyro_cov, params->gyro_cov, sizeof hp.gyro_cov);
  std::memcpy(hp.int_cov, params->int_cov, sizeof hp.int_cov);
  std::memcpy(hp.bias_acc_cov, params->bias_acc_cov, sizeof hp.bias_acc_cov);
  std::memcpy(hp.bias_gyro_cov, params->bias_gyro_cov, sizeof hp.bias_gyro_cov);
  std::memcpy(hp.bint, params->bias_acc_omega_int, sizeof hp.bint);
  std::memcpy(hp.gravity, params->gravity, sizeof hp.gravity);
  int* d_off = nullptr; double* d_imu = nullptr; ImuParamsDev* d_par = nullptr; double* d_bias = nullptr; fg_pim* d_out = nullptr;
  int rc = FG_OK;
  cudaError_t e;
  e = cudaMalloc((void**)&d_off, sizeof(int) * (n + 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_imu, sizeof(double) * 6 * (size_t)(ns > 0 ? ns : 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_par, sizeof(ImuParamsDev));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_bias, sizeof(double) * 6 * n);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, sizeof(fg_pim) * n);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, offsets, sizeof(int) * (n + 1), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && ns > 0) e = cudaMemcpyAsync(d_imu, imu6, sizeof(double) * 6 * (size_t)ns, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_par, &hp, sizeof hp, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_bias, bias_hat6, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    launch_preintegrate(n, d_off, d_imu, dt, d_par, d_bias, d_out, st);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(fg_pim) * n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { if (c) c->err = cudaGetErrorString(e); rc = FG_ERR_CUDA; }
  cudaFree(d_off); cudaFree(d_imu); cudaFree(d_par); cudaFree(d_bias); cudaFree(d_out);
  return rc;
}

extern "C" int fg_pim_predict(const fg_pim* pim, const double Xi[12], const double vi[3], const double bi[6],
                              double Xj[12], double vj[3]) {
  if (!pim || !Xi || !vi || !bi || !Xj || !vj) return FG_ERR_INVALID;
  double inc[6], bc[9];
  for (int i = 0; i < 6; ++i) inc[i] = bi[i] - pim->bias_hat[i];
  for (int i = 0; i < 9; ++i)
    bc[i] = pim->preint[i] + pim->H_ba[3 * i] * inc[0] + pim->H_ba[3 * i + 1] * inc[1] + pim->H_ba[3 * i + 2] * inc[2]
          + pim->H_bg[3 * i] * inc[3] + pim->H_bg[3 * i + 1] * inc[4] + pim->H_bg[3 * i + 2] * inc[5];
  double dt = pim->dt, dt22 = 0.5 * dt * dt, Rtv[3], Rtg[3], xp[3], xv[3], dR[9], t[3];
  m3_tvec(Xi, vi, Rtv);
  m3_tvec(Xi, pim->gravity, Rtg);
  for (int i = 0; i < 3; ++i) { xp[i] = bc[3 + i] + dt * Rtv[i] + dt22 * Rtg[i]; xv[i] = bc[6 + i] + dt * Rtg[i]; }
  so3_exp(bc, dR);
  m3_mul(Xi, dR, Xj);
  m3_vec(Xi, xp, t);
  for (int i = 0; i < 3; ++i) Xj[9 + i] = Xi[9 + i] + t[i];
  m3_vec(Xi, xv, t);
  for (int i = 0; i < 3; ++i) vj[i] = vi[i] + t[i];
  return FG_OK;
}

// ------------------------------------------------------------------ Schur tile tables (host; fg_schur.cu consumes them)
// pose_obs IS the pose-major order: position k holds observation pose_obs[k], and the observations of one pose are sorted
// by landmark.  Landmarks are cut into chunks of CH consecutive ids; per (pose, chunk) the kernel needs the first
// pose-major position and the bit mask of the landmarks seen; tiles are pairs of 16-pose groups that share a landmark.
namespace fg {
int build_schur_tables(fg_ctx* c, int64_t L, int64_t M, int64_t P, const std::vector<int64_t>& lm_ptr, const std::vector<int64_t>& pose_ptr,
                   const std::vector<int>& pose_nprim, const std::vector<int64_t>& pose_obs, const std::vector<int>& s_point, SchurTables& T) {
  const Symbolic& S = c->sym;
  std::vector<int>& ppos = T.ppos; std::vector<int>& pzp = T.pzp; std::vector<int>& pc_lo = T.pc_lo; std::vector<int>& pc_n = T.pc_n;
  std::vector<int64_t>& pc_ptr = T.pc_ptr; std::vector<uint2>& pc_ent = T.pc_ent; std::vector<int4>& tiles = T.tiles;
  if (6 * M >= (int64_t)1 << 31) return fail(c, FG_ERR_INVALID, "more than 2^31 / 6 projection factors on one rank");
  const int CH = 32;                                     // table words of 32 landmarks; k_schur_tiles walks three of them per staging round (fg_schur.cu)
  ppos.assign(M, 0); pzp.assign(M, 0);
  for (int64_t k = 0; k < M; ++k) { ppos[pose_obs[k]] = (int)k; pzp[k] = s_point[pose_obs[k]]; }
  pc_lo.assign(P, 0); pc_n.assign(P, 0);
  pc_ptr.assign(P + 1, 0);
  // the first pose_nprim[p] pose-major records of a pose are its PRIMARY observations, one per landmark, sorted by
  // landmark; further factors on an already seen (pose, landmark) pair sit behind them (their W is folded into the
  // primary's by k_merge_dup, so they take no part in the tile products)
  for (int64_t p = 0; p < P; ++p) {
    if (pose_nprim[p] > 0) {
      pc_lo[p] = pzp[pose_ptr[p]] / CH;
      pc_n[p] = pzp[pose_ptr[p] + pose_nprim[p] - 1] / CH - pc_lo[p] + 1;
    }
    pc_ptr[p + 1] = pc_ptr[p] + pc_n[p];
  }
  pc_ent.assign(pc_ptr[P] ? pc_ptr[P] : 1, make_uint2(0u, 0u));
  for (int64_t p = 0; p < P; ++p)
    for (int64_t k = pose_ptr[p]; k < pose_ptr[p] + pose_nprim[p]; ++k) {
      const int l = pzp[k];
      uint2& e = pc_ent[pc_ptr[p] + (l / CH - pc_lo[p])];
      if (e.y == 0u) e.x = (unsigned)k;
      if (e.y & (1u << (l % CH))) return fail(c, FG_ERR_STATE, "internal: duplicate primary observation in the Schur tables");
      e.y |= 1u << (l % CH);
    }
  // tiles: pairs of 16-pose groups that share a landmark, with the chunk range both sides cover
  const int G = (int)((P + 15) / 16);
  std::vector<int> g_lo(G, INT32_MAX), g_hi(G, 0);
  for (int64_t p = 0; p < P; ++p)
    if (pc_n[p]) { g_lo[p / 16] = std::min(g_lo[p / 16], pc_lo[p]); g_hi[p / 16] = std::max(g_hi[p / 16], pc_lo[p] + pc_n[p]); }
  std::vector<int64_t> tkeys;
  if (!S.cov_ptr.empty())
    for (int64_t p = 0; p < P; ++p) {
      int last = -1;
      for (int64_t k = S.cov_ptr[p]; k < S.cov_ptr[p + 1]; ++k) {
        const int q = S.cov_pose[k];
        if (q > p) continue;
        const int gq = q / 16;
        if (gq != last) { tkeys.push_back((int64_t)(p / 16) * G + gq); last = gq; }
      }
    }
  std::sort(tkeys.begin(), tkeys.end());
  tkeys.erase(std::unique(tkeys.begin(), tkeys.end()), tkeys.end());
  tiles.clear();
  for (int64_t key : tkeys) {
    const int gi = (int)(key / G), gj = (int)(key % G);
    const int cb = std::max(g_lo[gi], g_lo[gj]), ce = std::min(g_hi[gi], g_hi[gj]);
    if (cb < ce) tiles.push_back(make_int4(gi, gj, cb, ce));
  }
  // heaviest first.  (Measured: issuing the tiles in bands of neighbouring row groups, so that the CTAs in flight share
  // Z records in L2, is 15 % slower -- the long diagonal tiles must all start early.)
  std::stable_sort(tiles.begin(), tiles.end(), [](const int4& a, const int4& b) { return (a.w - a.z) > (b.w - b.z); });
  T.ch = CH;
  T.npairs = 0;
  std::vector<int64_t> nprim_l(L, 0);
  for (int64_t p = 0; p < P; ++p) for (int64_t k = pose_ptr[p]; k < pose_ptr[p] + pose_nprim[p]; ++k) nprim_l[pzp[k]]++;
  for (int64_t l = 0; l < L; ++l) T.npairs += nprim_l[l] * (nprim_l[l] + 1) / 2;
  return FG_OK;
}
}  // namespace fg

// ------------------------------------------------------------------ distribution of the leaf phase (multi-GPU; host only)
// ------------------------------------------------------------------ finalize: symbolic + upload
extern "C" int fg_finalize(fg_ctx* c) {
  if (!c) return FG_ERR_INVALID;
  if (c->device < 0) return fail(c, FG_ERR_CUDA, "detached context (device -1): no CUDA device, and there is no CPU solver");
  if (c->finalized) return push_values(c);
  CK(cudaSetDevice(c->device));
  if (pull_values(c) != FG_OK) return FG_ERR_CUDA;
  HostGraph& h = c->h;
  if (h.count(T_POSE) + h.count(T_VEC3) + h.count(T_BIAS) + h.count(T_PLANE) == 0)
    return fail(c, FG_ERR_STATE, "graph has no pose-side variables");
  dev_free_all(c, /*keep=*/true);
  int rc = build_symbolic(c);
  if (rc != FG_OK) return fail(c, rc, "symbolic analysis failed");
  Symbolic& S = c->sym;
  DevGraph& d = c->d;
  // values
  for (int t = 0; t < T_COUNT; ++t) {
    d.n[t] = h.count(t);
    if ((rc = dev_upload(c, &d.val[t], c->inc_active ? h.lin[t] : h.val[t])) != FG_OK) return rc;
    if (c->inc_active) { if ((rc = dev_upload(c, &d.val_new[t], h.val[t])) != FG_OK) return rc; }
    else if ((rc = dev_upload<double>(c, &d.val_new[t], nullptr, h.val[t].size())) != FG_OK) return rc;
    if (t != T_POINT) if ((rc = dev_upload(c, &d.off[t], S.off[t])) != FG_OK) return rc;
  }
  // pose-side factors, uploaded COLOUR-SORTED: within a colour no two factors share a variable, so the assembly kernels
  // (launched colour by colour, fg_kernels.cu: run_factors) never add to one address from two threads of a launch --
  // the assembled system is bitwise repeatable without giving up the thread-per-factor kernels (greedy colouring in
  // insertion order, PER KIND: the kinds are launched one after the other on one stream, so only factors of one kind can meet;
  // a VIO graph with 5 look-back edges takes ~12 colours for its edges and 2 for its IMU chain)
  d.n_pp = (int)h.pp_var.size(); d.n_pv = (int)h.pv_var.size(); d.n_pb = (int)h.pb_var.size();
  d.n_bt = (int)h.bt_i.size(); d.n_imu = (int)h.imu_rec.size(); d.n_pl = (int)h.pl_pose.size();
  {
    const int64_t nP = h.count(T_POSE), nV = h.count(T_VEC3), nB = h.count(T_BIAS);
    std::vector<std::vector<uint64_t>> used((size_t)(nP + nV + nB + h.count(T_PLANE)));
    auto pick = [&](const int* vars, int nv) {
      for (int col = 0;; ++col) {
        const size_t w = (size_t)col >> 6; const uint64_t bit = 1ull << (col & 63);
        bool free_ = true;
        for (int k = 0; k < nv && free_; ++k) { const auto& u = used[vars[k]]; if (w < u.size() && (u[w] & bit)) free_ = false; }
        if (!free_) continue;
        for (int k = 0; k < nv; ++k) { auto& u = used[vars[k]]; if (u.size() <= w) u.resize(w + 1, 0); u[w] |= bit; }
        return col;
      }
    };
    auto sort_kind = [&](int kind, const std::vector<int>& col) {
      std::vector<int> ord(col.size());
      for (size_t i = 0; i < ord.size(); ++i) ord[i] = (int)i;
      std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return col[a] < col[b]; });
      std::vector<int>& cp = c->color_ptr[kind];
      cp.assign(1, 0);
      for (size_t i = 0; i < ord.size(); ++i) { while ((int)cp.size() <= col[ord[i]]) cp.push_back((int)i); }
      cp.push_back((int)ord.size());
      return ord;
    };
    auto permute_i = [](const std::vector<int>& a, const std::vector<int>& ord, int w) { std::vector<int> o(a.size()); for (size_t i = 0; i < ord.size(); ++i) for (int k = 0; k < w; ++k) o[i * w + k] = a[(size_t)ord[i] * w + k]; return o; };
    auto permute_d = [](const std::vector<double>& a, const std::vector<int>& ord, int w) { std::vector<double> o(a.size()); for (size_t i = 0; i < ord.size(); ++i) for (int k = 0; k < w; ++k) o[i * w + k] = a[(size_t)ord[i] * w + k]; return o; };
    std::vector<int> col_pp(d.n_pp), col_pv(d.n_pv), col_pb(d.n_pb), col_bt(d.n_bt), col_ge(h.ge_i.size()), col_imu(d.n_imu), col_pl(d.n_pl);
    for (int f = 0; f < d.n_pp; ++f) { int v[1] = {h.pp_var[f]}; col_pp[f] = pick(v, 1); }
    for (auto& u : used) u.clear();
    for (int f = 0; f < d.n_pv; ++f) { int v[1] = {(int)(nP + h.pv_var[f])}; col_pv[f] = pick(v, 1); }
    for (auto& u : used) u.clear();
    for (int f = 0; f < d.n_pb; ++f) { int v[1] = {(int)(nP + nV + h.pb_var[f])}; col_pb[f] = pick(v, 1); }
    for (auto& u : used) u.clear();
    for (int f = 0; f < d.n_bt; ++f) { int v[2] = {h.bt_i[f], h.bt_j[f]}; col_bt[f] = pick(v, 2); }
    for (auto& u : used) u.clear();
    for (size_t f = 0; f < h.ge_i.size(); ++f) { int v[2] = {h.ge_i[f], h.ge_j[f]}; col_ge[f] = pick(v, 2); }
    for (auto& u : used) u.clear();
    for (int f = 0; f < d.n_imu; ++f) {
      const int* q = &h.imu_var[6 * (size_t)f];
      int v[6] = {q[0], (int)(nP + q[1]), q[2], (int)(nP + q[3]), (int)(nP + nV + q[4]), (int)(nP + nV + q[5])};
      col_imu[f] = pick(v, 6);
    }
    for (auto& u : used) u.clear();
    for (int f = 0; f < d.n_pl; ++f) { int v[2] = {h.pl_pose[f], (int)(nP + nV + nB + h.pl_plane[f])}; col_pl[f] = pick(v, 2); }
    std::vector<int> o;
    o = sort_kind(K_PP, col_pp);
    if ((rc = dev_upload(c, &d.pp_var, permute_i(h.pp_var, o, 1))) || (rc = dev_upload(c, &d.pp_mean, permute_d(h.pp_mean, o, 12))) || (rc = dev_upload(c, &d.pp_info, permute_d(h.pp_info, o, 36)))) return rc;
    o = sort_kind(K_PV, col_pv);
    if ((rc = dev_upload(c, &d.pv_var, permute_i(h.pv_var, o, 1))) || (rc = dev_upload(c, &d.pv_mean, permute_d(h.pv_mean, o, 3))) || (rc = dev_upload(c, &d.pv_info, permute_d(h.pv_info, o, 9)))) return rc;
    o = sort_kind(K_PB, col_pb);
    if ((rc = dev_upload(c, &d.pb_var, permute_i(h.pb_var, o, 1))) || (rc = dev_upload(c, &d.pb_mean, permute_d(h.pb_mean, o, 6))) || (rc = dev_upload(c, &d.pb_info, permute_d(h.pb_info, o, 36)))) return rc;
    o = sort_kind(K_BT, col_bt);
    if ((rc = dev_upload(c, &d.bt_i, permute_i(h.bt_i, o, 1))) || (rc = dev_upload(c, &d.bt_j, permute_i(h.bt_j, o, 1))) || (rc = dev_upload(c, &d.bt_meas, permute_d(h.bt_meas, o, 12))) || (rc = dev_upload(c, &d.bt_info, permute_d(h.bt_info, o, 36)))) return rc;
    {
      // colour-free Jacobian pass of the between factors (k_between_ends): factor ends sorted by (pose, factor), cut into blocks
      // of <= 128 ends on pose boundaries.  Needs one factor per pose pair at most (the j-end owns the off-diagonal block);
      // a graph with parallel or reversed edges on one pair keeps the coloured kernel.
      const std::vector<int> bi = permute_i(h.bt_i, o, 1), bj = permute_i(h.bt_j, o, 1);
      std::vector<std::pair<int, int>> pairs(bi.size());
      bool simple = !bi.empty();
      for (size_t f = 0; f < bi.size(); ++f) { pairs[f] = {std::min(bi[f], bj[f]), std::max(bi[f], bj[f])}; if (bi[f] == bj[f]) simple = false; }
      std::sort(pairs.begin(), pairs.end());
      if (std::adjacent_find(pairs.begin(), pairs.end()) != pairs.end()) simple = false;
      d.n_bt_eblk = 0;
      if (simple) {
        std::vector<std::pair<int, int>> ends;      // (pose, 2 f + end)
        ends.reserve(2 * bi.size());
        for (size_t f = 0; f < bi.size(); ++f) { ends.push_back({bi[f], (int)(2 * f)}); ends.push_back({bj[f], (int)(2 * f + 1)}); }
        std::sort(ends.begin(), ends.end());
        std::vector<int> rec(ends.size()), blk(1, 0);
        for (size_t q = 0; q < ends.size(); ++q) rec[q] = ends[q].second;
        size_t q = 0;
        while (q < ends.size()) {
          size_t r = q;
          while (r < ends.size() && ends[r].first == ends[q].first) ++r;            // the run of one pose
          if ((int)(r - blk.back()) > 128 && (int)q > blk.back()) blk.push_back((int)q);      // close the block before this pose
          q = r;
        }
        blk.push_back((int)ends.size());
        if ((rc = dev_upload(c, &d.bt_end, rec)) || (rc = dev_upload(c, &d.bt_eblk, blk))) return rc;
        d.n_bt_eblk = (int)blk.size() - 1;
      }
    }
    o = sort_kind(K_GE, col_ge);
    if ((rc = dev_upload(c, &d.ge_i, permute_i(h.ge_i, o, 1))) || (rc = dev_upload(c, &d.ge_j, permute_i(h.ge_j, o, 1))) || (rc = dev_upload(c, &d.ge_meas, permute_d(h.ge_meas, o, 12))) || (rc = dev_upload(c, &d.ge_info, permute_d(h.ge_info, o, 36)))) return rc;
    o = sort_kind(K_IMU, col_imu);
    {
      std::vector<ImuRec> recs(h.imu_rec.size());
      for (size_t i = 0; i < o.size(); ++i) recs[i] = h.imu_rec[o[i]];
      if ((rc = dev_upload(c, &d.imu_var, permute_i(h.imu_var, o, 6))) || (rc = dev_upload(c, &d.imu_rec, recs))) return rc;
    }
    o = sort_kind(K_PL, col_pl);
    if ((rc = dev_upload(c, &d.pl_pose, permute_i(h.pl_pose, o, 1))) || (rc = dev_upload(c, &d.pl_plane, permute_i(h.pl_plane, o, 1))) || (rc = dev_upload(c, &d.pl_meas, permute_d(h.pl_meas, o, 4))) || (rc = dev_upload(c, &d.pl_info, permute_d(h.pl_info, o, 9)))) return rc;
    CK(cudaStreamSynchronize(c->stream));       // the permuted host copies go out of scope
  }
  // g2o back-end: EdgeSE3 factors, fixed vertices, VertexSE3::oplus as the pose retraction
  d.n_ge = (int)h.ge_i.size();
  d.pose_chart = d.n_ge ? 1 : c->pose_chart;
  if (d.n_ge && (h.pp_var.size() + h.bt_i.size() + h.imu_rec.size() + h.pl_pose.size() + h.pj_pose.size()))
    return fail(c, FG_ERR_INVALID, "g2o edges and GTSAM pose factors cannot share a graph (different pose charts)");
  {
    std::vector<int> fixed_list;
    h.fixed_pose.resize(h.count(T_POSE), 0);
    for (size_t i = 0; i < h.fixed_pose.size(); ++i) if (h.fixed_pose[i]) fixed_list.push_back((int)i);
    d.n_fixed = (int)fixed_list.size();
    if (d.n_fixed) {
      if (!d.n_ge) return fail(c, FG_ERR_INVALID, "fg_set_fixed: only g2o edges may touch a fixed vertex");
      std::vector<char> fixed_col(S.n_r, 0);
      for (int i : fixed_list) for (int k = 0; k < 6; ++k) fixed_col[S.off[T_POSE][i] + k] = 1;
      if ((rc = dev_upload(c, &d.fixed_list, fixed_list)) || (rc = dev_upload(c, &d.fixed_pose, h.fixed_pose)) || (rc = dev_upload(c, &d.fixed_col, fixed_col))) return rc;
    }
  }

  if ((rc = dev_upload(c, &d.pl_pose, h.pl_pose)) || (rc = dev_upload(c, &d.pl_plane, h.pl_plane)) || (rc = dev_upload(c, &d.pl_meas, h.pl_meas)) || (rc = dev_upload(c, &d.pl_info, h.pl_info))) return rc;
  // landmarks: sort observations by landmark (stable), CSR by pose
  const int64_t L = h.count(T_POINT), M = (int64_t)h.pj_pose.size(), P = h.count(T_POSE);
  d.n_obs = M;
  if (L) {
    // distinct (calibration, body_P_sensor) pairs in order of first appearance: a graph with one pair (what the reference
    // builds) passes it to the kernels by value; a mixed graph (GTSAM accepts them) carries an index per observation
    std::vector<DevGraph::ProjCal> cals;
    std::vector<unsigned char> cal_of(M, 0);
    {
      std::vector<std::pair<int, int>> pairs;
      for (int64_t o = 0; o < M; ++o) {
        const std::pair<int, int> key(h.pj_calib[o], h.pj_sensor[o]);
        size_t k = 0;
        while (k < pairs.size() && pairs[k] != key) ++k;
        if (k == pairs.size()) {
          if (pairs.size() == 255) return fail(c, FG_ERR_INVALID, "more than 255 distinct (calibration, body_P_sensor) pairs");
          pairs.push_back(key);
          DevGraph::ProjCal pc;
          std::copy(h.calib.begin() + 9 * key.first, h.calib.begin() + 9 * key.first + 9, pc.K);
          std::copy(h.sensor.begin() + 12 * key.second, h.sensor.begin() + 12 * key.second + 12, pc.S);
          cals.push_back(pc);
        }
        cal_of[o] = (unsigned char)k;
      }
      if (cals.empty()) {
        DevGraph::ProjCal pc;
        std::copy(h.calib.begin(), h.calib.begin() + 9, pc.K);
        std::copy(h.sensor.begin(), h.sensor.begin() + 12, pc.S);
        cals.push_back(pc);
      }
    }
    int cid = M ? h.pj_calib[0] : 0, sid = M ? h.pj_sensor[0] : 0;
    std::vector<int64_t> lm_ptr(L + 1, 0), pose_ptr(P + 1, 0);
    for (int64_t o = 0; o < M; ++o) { lm_ptr[h.pj_point[o] + 1]++; pose_ptr[h.pj_pose[o] + 1]++; }
    for (int64_t l = 0; l < L; ++l) lm_ptr[l + 1] += lm_ptr[l];
    for (int64_t p = 0; p < P; ++p) pose_ptr[p + 1] += pose_ptr[p];
    std::vector<int> s_pose(M), s_point(M);
    std::vector<double> s_uv(2 * M), s_w(M);
    std::vector<unsigned char> s_cal(M);
    {
      std::vector<int64_t> cur(lm_ptr.begin(), lm_ptr.end() - 1);
      for (int64_t o = 0; o < M; ++o) {
        int64_t k = cur[h.pj_point[o]]++;
        s_cal[k] = cal_of[o];
        s_pose[k] = h.pj_pose[o]; s_point[k] = h.pj_point[o];
        s_uv[2 * k] = h.pj_uv[2 * o]; s_uv[2 * k + 1] = h.pj_uv[2 * o + 1]; s_w[k] = h.pj_w[o];
      }
    }
    // several projection factors on one (pose, landmark) pair (GTSAM accepts them; CGraphGT's BA builder can produce them
    // when two features of a frame match the same landmark, gtsam_graph.cpp:398-434): the first is the PRIMARY
    // observation, the others are folded into it where the pair acts as one block (W), and are kept as observations
    // everywhere else (residuals, V_l, g_l, U_pp, g_p)
    std::vector<int> dup_prim, dup_sec, pose_nprim(P, 0);
    std::vector<char> is_sec(M, 0);
    {
      std::vector<int64_t> seen_l(P, -1); std::vector<int> seen_k(P, 0);
      for (int64_t l = 0; l < L; ++l)
        for (int64_t k = lm_ptr[l]; k < lm_ptr[l + 1]; ++k) {
          const int p = s_pose[k];
          if (seen_l[p] == l) { is_sec[k] = 1; dup_prim.push_back(seen_k[p]); dup_sec.push_back((int)k); }
          else { seen_l[p] = l; seen_k[p] = (int)k; pose_nprim[p]++; }
        }
    }
    std::vector<int64_t> pose_obs(M);
    {
      std::vector<int64_t> cur(pose_ptr.begin(), pose_ptr.end() - 1);
      for (int64_t k = 0; k < M; ++k) if (!is_sec[k]) pose_obs[cur[s_pose[k]]++] = k;
      for (int64_t k = 0; k < M; ++k) if (is_sec[k]) pose_obs[cur[s_pose[k]]++] = k;
    }
    if (!dup_sec.empty()) {                                  // pairs of one primary consecutive, in observation order (k_merge_dup)
      std::vector<int> ord(dup_sec.size());
      for (size_t i = 0; i < ord.size(); ++i) ord[i] = (int)i;
      std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return dup_prim[a] != dup_prim[b] ? dup_prim[a] < dup_prim[b] : dup_sec[a] < dup_sec[b]; });
      std::vector<int> a(ord.size()), b(ord.size());
      for (size_t i = 0; i < ord.size(); ++i) { a[i] = dup_prim[ord[i]]; b[i] = dup_sec[ord[i]]; }
      dup_prim.swap(a); dup_sec.swap(b);
    }
    // observation ranges of the blocks of k_proj_obs / k_lm_backsub_obs: whole landmarks, at most 256 observations (a landmark
    // with more gets a block of its own), so that every per-landmark sum is formed inside one block in a fixed order
    {
      std::vector<int64_t> ob(1, 0);
      int64_t cur = 0;
      for (int64_t l = 0; l < L; ++l) {
        const int64_t k = lm_ptr[l + 1] - lm_ptr[l];
        if (k == 0) continue;
        if (cur > 0 && cur + k > 256) { ob.push_back(lm_ptr[l]); cur = 0; }
        cur += k;
      }
      if (M > 0) ob.push_back(M);
      d.n_oblk = (int)ob.size() - 1;
      if ((rc = dev_upload(c, &d.oblk_ptr, ob)) != FG_OK) return rc;
    }
    d.n_dup = (int)dup_sec.size();
    if ((rc = dev_upload(c, &d.dup_prim, dup_prim)) || (rc = dev_upload(c, &d.dup_sec, dup_sec))) return rc;
    std::vector<double> pm(3 * L, 0.0), pw(L, 0.0);
    for (size_t i = 0; i < h.pq_var.size(); ++i) {
      int l = h.pq_var[i];
      if (pw[l] != 0.0) return fail(c, FG_ERR_INVALID, "more than one PriorFactor<Point3> on a landmark");
      pw[l] = h.pq_w[i];
      pm[3 * l] = h.pq_mean[3 * i]; pm[3 * l + 1] = h.pq_mean[3 * i + 1]; pm[3 * l + 2] = h.pq_mean[3 * i + 2];
    }
    if ((rc = dev_upload(c, &d.lm_ptr, lm_ptr)) || (rc = dev_upload(c, &d.obs_pose, s_pose)) || (rc = dev_upload(c, &d.obs_point, s_point)) ||
        (rc = dev_upload(c, &d.obs_uv, s_uv)) || (rc = dev_upload(c, &d.obs_w, s_w)) || (rc = dev_upload(c, &d.pose_obs_ptr, pose_ptr)) ||
        (rc = dev_upload(c, &d.pose_obs, pose_obs)) || (rc = dev_upload(c, &d.lm_prior_mean, pm)) || (rc = dev_upload(c, &d.lm_prior_w, pw))) return rc;
    if ((rc = dev_upload<double>(c, &d.W, nullptr, (size_t)18 * M)) || (rc = dev_upload<double>(c, &d.V, nullptr, (size_t)6 * L)) ||
        (rc = dev_upload<double>(c, &d.gl, nullptr, (size_t)3 * L)) || (rc = dev_upload<double>(c, &d.Vinv, nullptr, (size_t)6 * L)) ||
        (rc = dev_upload<double>(c, &d.tl, nullptr, (size_t)3 * L))) return rc;
    if ((rc = dev_upload(c, &d.calib, h.calib.data() + 9 * cid, 9)) || (rc = dev_upload(c, &d.sensor, h.sensor.data() + 12 * sid, 12))) return rc;
    d.cal = cals[0];
    d.n_cal = (int)cals.size();
    if (d.n_cal > 1 && ((rc = dev_upload(c, &d.cals, cals)) || (rc = dev_upload(c, &d.obs_cal, s_cal)))) return rc;
    if ((rc = dev_upload<double>(c, &d.ul, nullptr, (size_t)3 * L)) || (rc = dev_upload<double>(c, &d.Cf, nullptr, (size_t)6 * L)) ||
        (rc = dev_upload<double>(c, &d.Zp, nullptr, (size_t)18 * M))) return rc;
    {
      SchurTables T;
      if ((rc = build_schur_tables(c, L, M, P, lm_ptr, pose_ptr, pose_nprim, pose_obs, s_point, T)) != FG_OK) return rc;
      d.schur_ch = T.ch; d.n_tiles = (int)T.tiles.size(); d.n_pairs = T.npairs;
      if ((rc = dev_upload(c, &d.obs_ppos, T.ppos)) || (rc = dev_upload(c, &d.pz_point, T.pzp)) || (rc = dev_upload(c, &d.pc_lo, T.pc_lo)) ||
          (rc = dev_upload(c, &d.pc_n, T.pc_n)) || (rc = dev_upload(c, &d.pc_ptr, T.pc_ptr)) || (rc = dev_upload(c, &d.pc_ent, T.pc_ent)) ||
          (rc = dev_upload(c, &d.tile_desc, T.tiles))) return rc;
      CK(cudaStreamSynchronize(c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));   // host staging vectors go out of scope
  }
  // per-block partial sums of the scalar reductions (one slot per block of every kernel of a pass)
  {
    size_t cap = 64;
    for (int k = 0; k < K_COUNT; ++k) cap += c->color_ptr[k].size() + (size_t)c->color_ptr[k].back() / 4 + 4;      // >= sum over colours of cdiv(n, 64) or cdiv(n, 4)
    cap += (size_t)h.count(T_POINT) / 256 + 2 + (size_t)d.n_oblk + (size_t)d.n_bt_eblk;
    for (int t = 0; t < T_COUNT; ++t) cap += (size_t)h.count(t) / 128 + 2;
    d.part_cap = (int)cap;
    if ((rc = dev_upload<double>(c, &d.part, nullptr, cap)) || (rc = dev_upload<double>(c, &d.part2, nullptr, cap))) return rc;
  }
  // reduced system
  if ((rc = dev_upload<double>(c, &d.L, nullptr, (size_t)S.nnz + 8)) || (rc = dev_upload<double>(c, &d.U0, nullptr, (size_t)S.nnz + 8)) ||
      (rc = dev_upload<double>(c, &d.g_r, nullptr, S.n_r)) || (rc = dev_upload<double>(c, &d.delta, nullptr, S.n_r)) ||
      (rc = dev_upload<double>(c, &d.scal, nullptr, 8)) || (rc = dev_upload<int>(c, &d.flags, nullptr, S.n_sn)) ||
      (rc = dev_upload<int>(c, &d.status, nullptr, 1))) return rc;
  if ((rc = dev_upload(c, &d.col2sn, S.col2sn)) || (rc = dev_upload(c, &d.sn_col0, S.sn_col0)) || (rc = dev_upload(c, &d.sn_ncols, S.sn_ncols)) ||
      (rc = dev_upload(c, &d.sn_nrows, S.sn_nrows)) || (rc = dev_upload(c, &d.sn_rowptr, S.sn_rowptr)) || (rc = dev_upload(c, &d.sn_valptr, S.sn_valptr)) ||
      (rc = dev_upload(c, &d.rowidx, S.rowidx)) || (rc = dev_upload(c, &d.upd_ptr, S.upd_ptr)) || (rc = dev_upload(c, &d.upd_d, S.upd_d)) ||
      (rc = dev_upload(c, &d.upd_a, S.upd_a)) || (rc = dev_upload(c, &d.upd_b, S.upd_b)) ||
      (rc = dev_upload(c, &d.anc_ptr, S.anc_ptr)) || (rc = dev_upload(c, &d.anc_t, S.anc_t)) || (rc = dev_upload(c, &d.anc_a, S.anc_a)) ||
      (rc = dev_upload(c, &d.anc_b, S.anc_b)) || (rc = dev_upload(c, &d.sched, S.sched)) ||
      (rc = dev_upload<int>(c, &d.flags2, nullptr, S.n_sn)) || (rc = dev_upload<int>(c, &d.counters, nullptr, 8))) return rc;
  if (S.use_fronts) {
    if ((rc = dev_upload(c, &d.updr_ptr, S.updr_ptr)) || (rc = dev_upload(c, &d.updr_d, S.updr_d)) ||
        (rc = dev_upload(c, &d.sched_a, S.sched_a)) || (rc = dev_upload(c, &d.sched_c, S.sched_c)) ||
        (rc = dev_upload(c, &d.fr_rowptr, S.fr_rowptr)) || (rc = dev_upload(c, &d.fr_rows, S.fr_rows)) || (rc = dev_upload(c, &d.fr_uptr, S.fr_uptr)) ||
        (rc = dev_upload(c, &d.pm_ptr, S.pm_ptr)) || (rc = dev_upload(c, &d.posmap, S.posmap)) || (rc = dev_upload(c, &d.pmne_ptr, S.pmne_ptr)) ||
        (rc = dev_upload(c, &d.pm_nonempty, S.pm_nonempty)) || (rc = dev_upload(c, &d.leaf_sn_lo, S.leaf_sn_lo)) || (rc = dev_upload(c, &d.leaf_sn_hi, S.leaf_sn_hi)) ||
        (rc = dev_upload(c, &d.tf_ptr, S.tf_ptr)) || (rc = dev_upload(c, &d.tf_leaf, S.tf_leaf)) ||
        (rc = dev_upload(c, &d.tile_mptr, S.tile_mptr)) || (rc = dev_upload(c, &d.tile_mrec, S.tile_mrec)) ||
        (rc = dev_upload(c, &d.tile_leaf, S.tile_leaf)) || (rc = dev_upload(c, &d.tile_i, S.tile_i)) || (rc = dev_upload(c, &d.tile_j, S.tile_j)) ||
        (rc = dev_upload<double>(c, &d.U, nullptr, (size_t)S.fr_uptr[S.n_leaves]))) return rc;
  }
  if (!S.rs_ok)
    return fail(c, FG_ERR_INVALID, "reduced system outside the limits of the factorisation kernel (a supernode with more than 32767 rows)");
  {
    if ((rc = dev_upload(c, &d.rsu_ptr, S.rsu_ptr)) || (rc = dev_upload(c, &d.rsu_d, S.rsu_d)) || (rc = dev_upload(c, &d.rsu_rec, S.rsu_rec)) ||
        (rc = dev_upload(c, &d.rs_units, S.rs_units)) || (rc = dev_upload(c, &d.rs_moff, S.rs_moff)) || (rc = dev_upload(c, &d.rs_map, S.rs_map)) ||
        (rc = dev_upload(c, &d.rs_colinv, S.rs_colinv)) ||
        (rc = dev_upload<int>(c, &d.rs_done, nullptr, S.rs_units.size() + S.n_sn)) || (rc = dev_upload(c, &d.rs_sn_units, S.rs_sn_units))) return rc;
  }
  {
    const std::vector<int> bs_order(S.sched.rbegin(), S.sched.rend());      // backward solve: reverse level order
    if ((rc = dev_upload(c, &d.bs_order, bs_order)) != FG_OK) return rc;
  }
  if (c->nranks > 1) {
    d.n_pk = (int64_t)S.pk_idx.size();
    if ((rc = dev_upload(c, &d.pk_idx, S.pk_idx)) || (rc = dev_upload<double>(c, &d.pk_buf, nullptr, (size_t)d.n_pk + 1))) return rc;
  }
  CK(cudaStreamSynchronize(c->stream));
  pool_trim(c);
  c->epoch = 0;
  c->finalized = true;
  c->values_dirty = false;
  c->device_newer = false;
  return FG_OK;
}

// ------------------------------------------------------------------ optimise
extern "C" void fg_lm_params_default(fg_lm_params* p) {
  if (!p) return;
  p->lambda_initial = 1e-5; p->lambda_factor = 10.0; p->lambda_upper = 1e5; p->lambda_lower = 0.0;
  p->min_model_fidelity = 1e-3; p->max_iterations = 100; p->relative_error_tol = 1e-5;
  p->absolute_error_tol = 1e-5; p->error_tol = 0.0; p->force_iterations = 0; p->verbosity = 0;
}

static int allreduce(fg_ctx* c, double* buf, size_t n) {
  if (c->nranks <= 1) return FG_OK;
  if (!c->nccl_comm) return fail(c, FG_ERR_NCCL, "nranks > 1 but fg_comm_init was not called");
  int r = g_nccl.AllReduce(buf, buf, n, kNcclDouble, kNcclSum, c->nccl_comm, c->stream);
  return r == 0 ? FG_OK : fail(c, FG_ERR_NCCL, "ncclAllReduce failed");
}

// The one large collective of a trial: every rank holds its partial reduced system in d.L; only the entries
// that can be non-zero before the factorisation travel (packed), with the chi2 of the linearisation point in the same
// buffer when it is fresh.
static int reduce_system(fg_ctx* c, bool with_chi2) {
  if (c->nranks <= 1) return FG_OK;
  launch_pack(c, with_chi2);
  int rc = allreduce(c, c->d.pk_buf, (size_t)c->d.n_pk + 1);
  if (rc != FG_OK) return rc;
  launch_unpack(c, with_chi2);
  return FG_OK;
}

static int read_scalars(fg_ctx* c, double* hs, int* status) {
  CK(cudaMemcpyAsync(hs, c->d.scal, sizeof(double) * 4, cudaMemcpyDeviceToHost, c->stream));
  if (status) CK(cudaMemcpyAsync(status, c->d.status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FG_OK;
}

extern "C" int fg_error(fg_ctx* c, double* error) {
  if (!c || !error) return fail(c, FG_ERR_INVALID, "null argument");
  int rc = fg_finalize(c);
  if (rc != FG_OK) return rc;
  CK(cudaSetDevice(c->device));
  if ((rc = end_incremental(c)) != FG_OK) return rc;
  launch_error_only(c, false);
  if ((rc = allreduce(c, c->d.scal, 1)) != FG_OK) return rc;
  double hs[4];
  if ((rc = read_scalars(c, hs, nullptr)) != FG_OK) return rc;
  CK(cudaGetLastError());
  *error = 0.5 * hs[0];
  return FG_OK;
}

extern "C" int fg_optimize_lm(fg_ctx* c, const fg_lm_params* params, fg_lm_report* rep) {
  if (!c) return FG_ERR_INVALID;
  fg_lm_params p;
  if (params) p = *params; else fg_lm_params_default(&p);
  fg_lm_report local;
  if (!rep) rep = &local;
  std::memset(rep, 0, sizeof *rep);
  int rc = fg_finalize(c);
  if (rc != FG_OK) { rep->status = rc; return rc; }
  CK(cudaSetDevice(c->device));
  if ((rc = end_incremental(c)) != FG_OK) { rep->status = rc; return rc; }
  DevGraph& d = c->d;
  struct Events {          // destroyed on every return path (CK returns early)
    cudaEvent_t e[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ~Events() { for (auto& x : e) if (x) cudaEventDestroy(x); }
  } evs;
  for (auto& e : evs.e) CK(cudaEventCreate(&e));
  cudaEvent_t* ev = evs.e;
  const cudaEvent_t ev_begin = evs.e[6], ev_end = evs.e[7];

  // graph.error(values) of the starting point comes out of the first linearisation (scal[0]): no separate error pass
  double err = std::numeric_limits<double>::quiet_NaN();
  bool have_err = false;
  rep->n_reduced_dims = c->sym.n_r; rep->n_supernodes = c->sym.n_sn; rep->nnz_L = c->sym.nnz;
  rep->n_projections = d.n_obs; rep->n_landmarks = d.n[T_POINT];
  rep->n_schur_pairs = d.n_pairs; rep->n_levels = c->sym.n_levels_rs;
  rep->nnz_S = c->sym.nnz_S;
  rep->allreduce_bytes = c->nranks > 1 ? (int64_t)sizeof(double) * (d.n_pk + 1) : 0;
  double lam = p.lambda_initial;
  const double inf = std::numeric_limits<double>::infinity();
  int it = 0;
  CK(cudaEventRecord(ev_begin, c->stream));
  while (true) {
    double cur = err;
    // ---- LevenbergMarquardtOptimizer::iterate()
    CK(cudaEventRecord(ev[0], c->stream));
    launch_linearize(c);
    CK(cudaEventRecord(ev[1], c->stream));
    bool first = true;
    while (true) {
      if (!first) CK(cudaEventRecord(ev[1], c->stream));
      launch_build_and_schur(c, lam);
      if ((rc = reduce_system(c, first)) != FG_OK) break;
      CK(cudaEventRecord(ev[2], c->stream));
      launch_factor_rs(c);
      CK(cudaEventRecord(ev[3], c->stream));
      launch_backsolve(c);
      CK(cudaEventRecord(ev[4], c->stream));
      launch_retract_error(c, lam);
      // [0] chi2 at the linearisation point (fresh only on the first trial of an iteration), [1] g^T delta, [2] |delta|^2, [3] new chi2
      if ((rc = allreduce(c, d.scal + 1, 3)) != FG_OK) break;     // the small collective: the three scalars that depend on the solve
      CK(cudaEventRecord(ev[5], c->stream));
      double hs[4]; int st = 0;
      if ((rc = read_scalars(c, hs, &st)) != FG_OK) break;
      CK(cudaGetLastError());
      if (!have_err) { err = 0.5 * hs[0]; cur = err; rep->initial_error = err; have_err = true; }
      float ms;
      if (first) {
        cudaEventElapsedTime(&ms, ev[0], ev[1]); rep->ms_linearize += ms;
        if (d.n_obs && c->kev[0] && cudaEventElapsedTime(&ms, c->kev[0], c->kev[1]) == cudaSuccess) rep->ms_proj_obs += ms;
      }
      if (d.n_tiles && c->kev[2] && cudaEventElapsedTime(&ms, c->kev[2], c->kev[3]) == cudaSuccess) rep->ms_schur_blocks += ms;
      cudaEventElapsedTime(&ms, ev[1], ev[2]); rep->ms_schur += ms;
      cudaEventElapsedTime(&ms, ev[2], ev[3]); rep->ms_factor += ms;
      cudaEventElapsedTime(&ms, ev[3], ev[4]); rep->ms_solve += ms;
      cudaEventElapsedTime(&ms, ev[4], ev[5]); rep->ms_retract_error += ms;
      first = false;
      rep->trials++;

      const double gTd = hs[1], dd = hs[2];
      double new_err = 0.5 * hs[3];
      bool solved = (st == 0) && std::isfinite(gTd) && std::isfinite(dd);
      bool step_ok = false, stop = false;
      if (solved) {
        // linear.error(delta) = err + g^T d + 1/2 d^T H d with (H + lam I) d = -g
        const double lin_change = -0.5 * gTd + 0.5 * lam * dd;
        if (lin_change >= 0) {
          const double cost_change = err - new_err;
          if (lin_change > 1e-20) step_ok = (cost_change / lin_change) > p.min_model_fidelity;
          else stop = true;
          if (std::fabs(cost_change) < p.relative_error_tol * err) stop = true;
        } else {
          new_err = inf;
        }
      } else {
        new_err = inf;
      }
      if (rep->trace_len < FG_TRACE_MAX) {
        int k = rep->trace_len++;
        rep->trace_lambda[k] = lam; rep->trace_error[k] = err; rep->trace_new_error[k] = new_err; rep->trace_accepted[k] = step_ok;
      }
      if (p.verbosity > 0)
        fprintf(stderr, "[fg] iter %d lambda %.3e error %.9e -> %.9e %s\n", it, lam, err, new_err, step_ok ? "accepted" : "rejected");
      if (step_ok) {
        for (int t = 0; t < T_COUNT; ++t) std::swap(d.val[t], d.val_new[t]);
        c->device_newer = true;
        err = new_err;
        lam = std::max(p.lambda_lower, lam / p.lambda_factor);
        break;
      } else if (!stop) {
        lam *= p.lambda_factor;
        if (lam >= p.lambda_upper) break;
      } else {
        break;
      }
    }
    if (rc != FG_OK) break;
    ++it;
    if (it >= p.max_iterations || !std::isfinite(err)) break;
    if (!p.force_iterations) {
      if (p.error_tol >= err) break;
      const double absdec = cur - err, reldec = cur != 0.0 ? absdec / cur : 0.0;
      if ((p.relative_error_tol != 0.0 && reldec <= p.relative_error_tol) || absdec <= p.absolute_error_tol) break;
    }
  }
  cudaEventRecord(ev_end, c->stream);
  cudaEventSynchronize(ev_end);
  float total = 0;
  cudaEventElapsedTime(&total, ev_begin, ev_end);
  rep->ms_total = total;
  rep->iterations = it;
  rep->final_error = err;
  rep->lambda = lam;
  rep->status = rc;
  return rc;
}

// ------------------------------------------------------------------ g2o Levenberg (config 1)
extern "C" void fg_g2o_params_default(fg_g2o_params* p) {
  if (!p) return;
  p->iterations = 20; p->iterations_per_call = 2; p->tau = 1e-5; p->max_trials = 10;
}

extern "C" int fg_g2o_chi2(fg_ctx* c, double* chi2) {
  if (!c || !chi2) return fail(c, FG_ERR_INVALID, "null argument");
  double e;
  int rc = fg_error(c, &e);
  if (rc != FG_OK) return rc;
  *chi2 = 2.0 * e;            // fg_error reports GTSAM's 1/2 sum; g2o's chi2() has no 1/2
  return FG_OK;
}

extern "C" int fg_optimize_g2o(fg_ctx* c, const fg_g2o_params* params, fg_g2o_report* rep) {
  if (!c) return FG_ERR_INVALID;
  fg_g2o_params p;
  if (params) p = *params; else fg_g2o_params_default(&p);
  fg_g2o_report local;
  if (!rep) rep = &local;
  std::memset(rep, 0, sizeof *rep);
  if (p.iterations_per_call < 1 || p.max_trials < 1) return fail(c, FG_ERR_INVALID, "bad g2o parameters");
  if (c->nranks > 1) return fail(c, FG_ERR_INVALID, "fg_optimize_g2o runs on one GPU: a pose graph has no landmarks to shard");
  int rc = fg_finalize(c);
  if (rc != FG_OK) { rep->status = rc; return rc; }
  CK(cudaSetDevice(c->device));
  if ((rc = end_incremental(c)) != FG_OK) { rep->status = rc; return rc; }
  DevGraph& d = c->d;
  if (!d.n_ge) return fail(c, FG_ERR_STATE, "fg_optimize_g2o: the graph has no g2o edges");
  struct Events {
    cudaEvent_t e[2] = {nullptr, nullptr};
    ~Events() { for (auto& x : e) if (x) cudaEventDestroy(x); }
  } ev;
  for (auto& e : ev.e) CK(cudaEventCreate(&e));
  CK(cudaEventRecord(ev.e[0], c->stream));
  double lam = 0.0, ni = 2.0, cur = 0.0;
  bool have_initial = false;
  int done_total = 0;
  while (done_total < p.iterations) {                      // for (i = 0; i < iter; i += currIt) currIt = optimize(per_call)
    int done = 0;
    bool ok = true;
    for (int it = 0; it < p.iterations_per_call && ok; ++it) {
      // ---- OptimizationAlgorithmLevenberg::solve(it)
      launch_linearize(c);
      if (it == 0) launch_max_diag(c, d.scal + 4);
      double hs0[5];
      CK(cudaMemcpyAsync(hs0, d.scal, sizeof(double) * 5, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      cur = hs0[0];
      if (!have_initial) { rep->initial_chi2 = cur; have_initial = true; }
      if (it == 0) { lam = p.tau * hs0[4]; ni = 2.0; }     // computeLambdaInit
      double rho = 0.0;
      int qmax = 0;
      do {
        launch_build_and_schur(c, lam);
        launch_factor_rs(c);
        launch_backsolve(c);
        launch_retract_error(c, lam);
        double hs[4]; int st = 0;
        if ((rc = read_scalars(c, hs, &st)) != FG_OK) { rep->status = rc; return rc; }
        CK(cudaGetLastError());
        const bool solved = st == 0 && std::isfinite(hs[1]) && std::isfinite(hs[2]);
        const double temp = solved ? hs[3] : std::numeric_limits<double>::max();
        const double scale = (solved ? lam * hs[2] - hs[1] : 0.0) + 1e-3;        // computeScale: sum dx_j (lambda dx_j + b_j), b = -g
        rho = (cur - temp) / scale;
        if (rho > 0 && std::isfinite(temp)) {
          double alpha = 1.0 - std::pow(2.0 * rho - 1.0, 3);
          alpha = std::min(alpha, 2.0 / 3.0);
          lam *= std::max(1.0 / 3.0, alpha);
          ni = 2.0;
          cur = temp;
          for (int t = 0; t < T_COUNT; ++t) std::swap(d.val[t], d.val_new[t]);
          c->device_newer = true;
        } else {
          lam *= ni;
          ni *= 2.0;
          if (!std::isfinite(lam)) break;
        }
        ++qmax;
      } while (rho < 0 && qmax < p.max_trials);
      ok = !(qmax == p.max_trials || rho == 0 || !std::isfinite(lam));          // otherwise: Terminate
      ++done;
      if (rep->trace_len < FG_G2O_TRACE_MAX) {
        const int k = rep->trace_len++;
        rep->trace_chi2[k] = cur; rep->trace_lambda[k] = lam; rep->trace_trials[k] = qmax;
      }
    }
    rep->calls += 1;
    done_total += done;
    if (done == 0) break;
  }
  CK(cudaEventRecord(ev.e[1], c->stream));
  CK(cudaEventSynchronize(ev.e[1]));
  float ms = 0; cudaEventElapsedTime(&ms, ev.e[0], ev.e[1]);
  rep->ms_total = ms;
  rep->iterations = done_total;
  rep->final_chi2 = cur;
  rep->lambda = lam;
  rep->status = FG_OK;
  return FG_OK;
}

// ------------------------------------------------------------------ incremental update (ISAM2 semantics)
extern "C" void fg_isam2_params_default(fg_isam2_params* p) {
  if (!p) return;
  p->relinearize_threshold = 0.1; p->relinearize_skip = 1;       // initISAM2Params, gtsam_graph.cpp:96-97
}

extern "C" int fg_update_incremental(fg_ctx* c, const fg_isam2_params* params, fg_inc_report* rep) {
  if (!c) return FG_ERR_INVALID;
  fg_isam2_params p;
  if (params) p = *params; else fg_isam2_params_default(&p);
  fg_inc_report local;
  if (!rep) rep = &local;
  std::memset(rep, 0, sizeof *rep);
  if (c->device < 0) return fail(c, FG_ERR_CUDA, "detached context (device -1): no CUDA device, and there is no CPU solver");
  CK(cudaSetDevice(c->device));
  int rc;
  if (!c->inc_active) {
    // first update of a session: theta = estimate = the current values
    if ((rc = pull_values(c)) != FG_OK) return rc;
    for (int t = 0; t < T_COUNT; ++t) c->h.lin[t] = c->h.val[t];
    c->inc_active = true; c->inc_updates = 0;
    for (int t = 0; t < T_COUNT; ++t) c->inc_known[t] = 0;
    if (c->finalized) c->values_dirty = true;        // the device has no estimate array content yet
  }
  const auto t_host = std::chrono::steady_clock::now();
  if ((rc = fg_finalize(c)) != FG_OK) { rep->status = rc; return rc; }
  rep->ms_rebuild = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host).count();
  DevGraph& d = c->d;
  for (int t = 0; t < T_COUNT; ++t) { rep->n_variables += d.n[t]; rep->n_new_variables += d.n[t] - c->inc_known[t]; c->inc_known[t] = d.n[t]; }
  struct Events {
    cudaEvent_t e[2] = {nullptr, nullptr};
    ~Events() { for (auto& x : e) if (x) cudaEventDestroy(x); }
  } ev;
  for (auto& e : ev.e) CK(cudaEventCreate(&e));
  CK(cudaEventRecord(ev.e[0], c->stream));
  // ---- fluid relinearisation: theta_j <- estimate_j where |delta_j| reaches the threshold
  c->inc_updates += 1;
  CK(cudaMemsetAsync(d.counters + 5, 0, sizeof(int), c->stream));     // [5]: relinearised variables (the factorisation owns [0..3])
  if (p.relinearize_skip <= 1 || c->inc_updates % p.relinearize_skip == 0) launch_inc_gate(c, p.relinearize_threshold, d.counters + 5);
  // ---- one undamped Gauss-Newton system at theta; estimate = theta (+) delta
  launch_linearize(c);
  launch_build_and_schur(c, 0.0);
  if ((rc = reduce_system(c, true)) != FG_OK) { rep->status = rc; return rc; }
  launch_factor_rs(c);
  launch_backsolve(c);
  launch_retract_error(c, 0.0);
  if ((rc = allreduce(c, d.scal + 1, 3)) != FG_OK) { rep->status = rc; return rc; }
  CK(cudaEventRecord(ev.e[1], c->stream));
  double hs[4]; int st = 0, nrel = 0;
  CK(cudaMemcpyAsync(&nrel, d.counters + 5, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if ((rc = read_scalars(c, hs, &st)) != FG_OK) { rep->status = rc; return rc; }
  CK(cudaGetLastError());
  float ms = 0; cudaEventElapsedTime(&ms, ev.e[0], ev.e[1]);
  rep->ms_update = ms;
  rep->n_relinearized = nrel;
  rep->error_before = 0.5 * hs[0];
  c->device_newer = true;
  if (st != 0 || !std::isfinite(hs[1]) || !std::isfinite(hs[2]) || !std::isfinite(hs[3])) {
    // IndeterminantLinearSystemException in GTSAM: the estimate stays where the linearisation point is
    for (int t = 0; t < T_COUNT; ++t)
      if (d.n[t]) CK(cudaMemcpyAsync(d.val_new[t], d.val[t], sizeof(double) * c->h.val[t].size(), cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    rep->error_after = rep->error_before;
    rep->status = FG_ERR_INDETERMINATE;
    return fail(c, FG_ERR_INDETERMINATE, "the undamped system of the incremental update is not positive definite (IndeterminantLinearSystemException in GTSAM)");
  }
  rep->error_after = 0.5 * hs[3];
  rep->status = FG_OK;
  return FG_OK;
}

// ------------------------------------------------------------------ marginal covariance
extern "C" int fg_marginal_cov(fg_ctx* c, fg_key key, double* cov, int* dim) {
  if (!c || !cov) return fail(c, FG_ERR_INVALID, "null argument");
  auto it = c->h.index.find(key);
  if (it == c->h.index.end()) return fail(c, FG_ERR_UNKNOWN_KEY, "key does not exist in Values");
  const int type = it->second.type;
  if (type == T_POINT) return fail(c, FG_ERR_INVALID, "marginal covariance of a Point3 landmark is not supported (eliminated by the Schur complement)");
  int rc = fg_finalize(c);
  if (rc != FG_OK) return rc;
  CK(cudaSetDevice(c->device));
  if ((rc = end_incremental(c)) != FG_OK) return rc;
  DevGraph& d = c->d;
  const int col0 = c->sym.off[type][it->second.idx], dm = kDim[type];
  launch_linearize(c);
  launch_build_and_schur(c, 0.0);
  if ((rc = reduce_system(c, false)) != FG_OK) return rc;
  launch_factor_rs(c);
  double* work = nullptr;
  CK(cudaMalloc((void**)&work, sizeof(double) * (6 * ((size_t)c->sym.n_r + 1) + 36)));
  double* out36 = work + 6 * ((size_t)c->sym.n_r + 1);
  launch_marginal(c, col0, dm, work, out36);
  double h36[36];
  int st = 0;
  cudaError_t e = cudaMemcpyAsync(h36, out36, sizeof h36, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&st, d.status, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(work);
  if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return FG_ERR_CUDA; }
  CK(cudaGetLastError());
  if (st != 0) return fail(c, FG_ERR_INDETERMINATE, "the undamped system is not positive definite (IndeterminantLinearSystemException in GTSAM)");
  for (int a = 0; a < dm; ++a) for (int b = 0; b < dm; ++b) cov[a * dm + b] = h36[a * 6 + b];
  if (dim) *dim = dm;
  return FG_OK;
}

// ------------------------------------------------------------------ multi-GPU
extern "C" int fg_comm_unique_id(char id[128]) {
  if (!id) return FG_ERR_INVALID;
  if (!load_nccl()) return FG_ERR_NCCL;
  nccl_uid u;
  if (g_nccl.GetUniqueId(&u) != 0) return FG_ERR_NCCL;
  std::memcpy(id, u.internal, 128);
  return FG_OK;
}
extern "C" int fg_comm_init(fg_ctx* c, const char id[128]) {
  if (!c || !id) return FG_ERR_INVALID;
  if (!load_nccl()) return fail(c, FG_ERR_NCCL, "libnccl.so.2 could not be loaded");
  CK(cudaSetDevice(c->device));
  nccl_uid u;
  std::memcpy(u.internal, id, 128);
  if (g_nccl.CommInitRank(&c->nccl_comm, c->nranks, u, c->rank) != 0) return fail(c, FG_ERR_NCCL, "ncclCommInitRank failed");
  return FG_OK;
}

// Host-only symbolic analysis for tests: fills `out` (capacity cap int64 entries) with array `which` and
// returns the array length (or a negative status).  Works on a detached context (fg_create(-1, ...)).
extern "C" int64_t fg_debug_symbolic(fg_ctx* c, int which, int64_t* out, int64_t cap) {
  if (!c) return FG_ERR_INVALID;
  if (which == 0) {
    int rc = build_symbolic(c);
    if (rc != FG_OK) return rc;
  }
  const Symbolic& S = c->sym;
  std::vector<int64_t> v;
  auto put = [&](const std::vector<int>& a) { v.assign(a.begin(), a.end()); };
  switch (which) {
    case 0: v = {S.n_r, S.n_sn, S.nnz, S.max_nrows, S.max_ncols, (int64_t)S.flops_factor, S.n_levels}; break;
    case 1: put(S.sn_col0); break;
    case 2: put(S.sn_ncols); break;
    case 3: put(S.sn_nrows); break;
    case 4: put(S.sn_rowptr); break;
    case 5: v.assign(S.sn_valptr.begin(), S.sn_valptr.end()); break;
    case 6: put(S.rowidx); break;
    case 7: put(S.upd_ptr); break;
    case 8: put(S.upd_d); break;
    case 9: put(S.upd_a); break;
    case 10: put(S.upd_b); break;
    case 11: put(S.off[T_POSE]); break;
    case 12: put(S.off[T_VEC3]); break;
    case 13: put(S.off[T_BIAS]); break;
    case 14: put(S.off[T_PLANE]); break;
    case 15: put(S.sched); break;
    case 16: put(S.level); break;
    case 17: put(S.anc_ptr); break;
    case 18: put(S.anc_t); break;
    case 19: put(S.anc_a); break;
    case 20: put(S.anc_b); break;
    case 21: put(S.sn_leaf); break;
    case 22: put(S.updr_ptr); break;
    case 23: put(S.updr_d); break;
    case 24: put(S.updr_a); break;
    case 25: put(S.updr_b); break;
    case 26: put(S.fr_rowptr); break;
    case 27: put(S.fr_rows); break;
    case 28: v.clear(); for (size_t l = 0; l < S.leaf_sn_lo.size(); ++l) { v.push_back(S.leaf_sn_lo[l]); v.push_back(S.leaf_sn_hi[l]); } break;
    case 29: put(S.tf_ptr); break;
    case 30: put(S.tf_leaf); break;
    case 31: v.assign(S.pm_ptr.begin(), S.pm_ptr.end()); break;
    case 32: put(S.posmap); break;
    case 33: v = {S.use_fronts ? 1 : 0, S.n_leaves, S.n_levels_fronts, (int64_t)S.tile_leaf.size()}; break;
    case 34: put(S.sched_a); break;
    case 35: put(S.sched_c); break;
    case 36: v.clear(); for (const int4& u : S.rs_units) { v.push_back(u.x); v.push_back(u.y); v.push_back(u.z); v.push_back(u.w); } break;
    case 37: v.assign(S.rs_moff.begin(), S.rs_moff.end()); break;
    case 38: v.assign(S.rs_map.begin(), S.rs_map.end()); break;
    case 40: v.assign(S.rs_colinv.begin(), S.rs_colinv.end()); break;
    case 39: v = {S.rs_ok ? 1 : 0, S.rs_units_a, (int64_t)S.rs_units.size(), S.n_levels_rs}; break;
    case 47: put(S.rsu_ptr); break;
    case 49: v.assign(S.pk_idx.begin(), S.pk_idx.end()); break;
    case 48: v.clear(); for (size_t q = 0; q < S.rsu_d.size(); ++q) { const UpdRec& r = S.rsu_rec[q]; v.push_back(S.rsu_d[q]); v.push_back(S.rsu_src[q]); v.push_back(r.pad[1]); v.push_back(r.K); v.push_back(r.pad[0]); } break;
    case 41: case 42: case 43: case 44: case 45: case 46: {
      // Schur tile tables of the projection factors held by this context (host only): 41 header [CH, n_tiles, n_pairs],
      // 42 tiles (4 ints each), 43 pc_lo, 44 pc_n, 45 pc_ptr, 46 pc_ent (start, mask pairs)
      const HostGraph& h = c->h;
      const int64_t L = h.count(T_POINT), M = (int64_t)h.pj_pose.size(), P = h.count(T_POSE);
      std::vector<int64_t> lm_ptr(L + 1, 0), pose_ptr(P + 1, 0);
      for (int64_t o = 0; o < M; ++o) { lm_ptr[h.pj_point[o] + 1]++; pose_ptr[h.pj_pose[o] + 1]++; }
      for (int64_t l = 0; l < L; ++l) lm_ptr[l + 1] += lm_ptr[l];
      for (int64_t p = 0; p < P; ++p) pose_ptr[p + 1] += pose_ptr[p];
      std::vector<int> s_pose(M), s_point(M);
      { std::vector<int64_t> cur(lm_ptr.begin(), lm_ptr.end() - 1); for (int64_t o = 0; o < M; ++o) { int64_t k = cur[h.pj_point[o]]++; s_pose[k] = h.pj_pose[o]; s_point[k] = h.pj_point[o]; } }
      std::vector<int64_t> pose_obs(M);
      std::vector<int> pose_nprim(P, 0);
      {
        std::vector<char> is_sec(M, 0);
        std::vector<int64_t> seen_l(P, -1);
        for (int64_t l = 0; l < L; ++l)
          for (int64_t k = lm_ptr[l]; k < lm_ptr[l + 1]; ++k) { const int p = s_pose[k]; if (seen_l[p] == l) is_sec[k] = 1; else { seen_l[p] = l; pose_nprim[p]++; } }
        std::vector<int64_t> cur(pose_ptr.begin(), pose_ptr.end() - 1);
        for (int64_t k = 0; k < M; ++k) if (!is_sec[k]) pose_obs[cur[s_pose[k]]++] = k;
        for (int64_t k = 0; k < M; ++k) if (is_sec[k]) pose_obs[cur[s_pose[k]]++] = k;
      }
      SchurTables T;
      int rc = build_schur_tables(c, L, M, P, lm_ptr, pose_ptr, pose_nprim, pose_obs, s_point, T);
      if (rc != FG_OK) return rc;
      if (which == 41) v = {T.ch, (int64_t)T.tiles.size(), T.npairs};
      else if (which == 42) { for (const int4& t : T.tiles) { v.push_back(t.x); v.push_back(t.y); v.push_back(t.z); v.push_back(t.w); } }
      else if (which == 43) put(T.pc_lo);
      else if (which == 44) put(T.pc_n);
      else if (which == 45) v.assign(T.pc_ptr.begin(), T.pc_ptr.end());
      else { for (const uint2& e : T.pc_ent) { v.push_back(e.x); v.push_back(e.y); } }
      break;
    }
    default: return FG_ERR_INVALID;
  }
  if (out) for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[i];
  return (int64_t)v.size();
}

long long g_fg_launches = 0;

// Host-side factor counts for tests: 0 point priors, 1 projections, 2 betweens, 3 imu, 4 plane factors, 5 pose priors;
// 100: kernels this process has launched so far (FGS() in fg_internal.h).
extern "C" int64_t fg_debug_counts(fg_ctx* c, int which) {
  if (which == 100) return (int64_t)g_fg_launches;
  if (!c) return FG_ERR_INVALID;
  const HostGraph& h = c->h;
  switch (which) {
    case 0: return (h.pq_var.size() == h.pq_w.size() && 3 * h.pq_var.size() == h.pq_mean.size()) ? (int64_t)h.pq_var.size() : -1;
    case 1: return (int64_t)h.pj_pose.size();
    case 2: return (int64_t)h.bt_i.size();
    case 3: return (int64_t)h.imu_rec.size();
    case 4: return (int64_t)h.pl_pose.size();
    case 5: return (int64_t)h.pp_var.size();
    default: return FG_ERR_INVALID;
  }
}

extern "C" int fg_debug_sizes(fg_ctx* c, int64_t out[8]) {
  if (!c || !out) return FG_ERR_INVALID;
  int rc = fg_finalize(c);
  if (rc != FG_OK) return rc;
  out[0] = c->sym.n_r; out[1] = c->sym.n_sn; out[2] = c->sym.nnz; out[3] = c->sym.max_nrows;
  out[4] = c->sym.max_ncols; out[5] = (int64_t)c->sym.flops_factor; out[6] = c->d.n_obs; out[7] = c->d.n[T_POINT];
  return FG_OK;
}
