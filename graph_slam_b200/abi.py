"""ctypes binding of include/fg_abi.h (libfg_b200.so) -- the only way Python reaches the solver.

There is no CPU fallback: if the shared library is missing this module raises at import time with the
build command, and every numeric call raises FgError when no CUDA device is usable.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libfg_b200.so')

FG_TRACE_MAX = 512
T_POSE, T_VEC3, T_BIAS, T_POINT, T_PLANE = range(5)
STORE = {T_POSE: 12, T_VEC3: 3, T_BIAS: 6, T_POINT: 3, T_PLANE: 4}


class FgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('fg status %d: %s' % (code, msg))
        self.code = code


class ImuParams(C.Structure):
    _fields_ = [('acc_cov', C.c_double * 9), ('gyro_cov', C.c_double * 9), ('int_cov', C.c_double * 9),
                ('bias_acc_cov', C.c_double * 9), ('bias_gyro_cov', C.c_double * 9),
                ('bias_acc_omega_int', C.c_double * 36), ('gravity', C.c_double * 3)]


class Pim(C.Structure):
    _fields_ = [('dt', C.c_double), ('preint', C.c_double * 9), ('H_ba', C.c_double * 27), ('H_bg', C.c_double * 27),
                ('bias_hat', C.c_double * 6), ('cov', C.c_double * 225), ('gravity', C.c_double * 3)]


class LMParams(C.Structure):
    _fields_ = [('lambda_initial', C.c_double), ('lambda_factor', C.c_double), ('lambda_upper', C.c_double),
                ('lambda_lower', C.c_double), ('min_model_fidelity', C.c_double), ('max_iterations', C.c_int),
                ('relative_error_tol', C.c_double), ('absolute_error_tol', C.c_double), ('error_tol', C.c_double),
                ('force_iterations', C.c_int), ('verbosity', C.c_int)]


class Isam2Params(C.Structure):
    _fields_ = [('relinearize_threshold', C.c_double), ('relinearize_skip', C.c_int)]


class IncReport(C.Structure):
    _fields_ = [('error_before', C.c_double), ('error_after', C.c_double), ('n_variables', C.c_int64), ('n_relinearized', C.c_int64),
                ('n_new_variables', C.c_int64), ('ms_update', C.c_double), ('ms_rebuild', C.c_double), ('status', C.c_int)]


class G2OParams(C.Structure):
    _fields_ = [('iterations', C.c_int), ('iterations_per_call', C.c_int), ('tau', C.c_double), ('max_trials', C.c_int)]


class G2OReport(C.Structure):
    _fields_ = [('iterations', C.c_int), ('calls', C.c_int), ('initial_chi2', C.c_double), ('final_chi2', C.c_double), ('lambda_', C.c_double),
                ('status', C.c_int), ('trace_len', C.c_int), ('trace_chi2', C.c_double * 64), ('trace_lambda', C.c_double * 64),
                ('trace_trials', C.c_int * 64), ('ms_total', C.c_double)]


class LMReport(C.Structure):
    _fields_ = [('iterations', C.c_int), ('trials', C.c_int), ('initial_error', C.c_double),
                ('final_error', C.c_double), ('lambda_', C.c_double), ('status', C.c_int), ('trace_len', C.c_int),
                ('trace_lambda', C.c_double * FG_TRACE_MAX), ('trace_error', C.c_double * FG_TRACE_MAX),
                ('trace_new_error', C.c_double * FG_TRACE_MAX), ('trace_accepted', C.c_int * FG_TRACE_MAX),
                ('ms_linearize', C.c_double), ('ms_schur', C.c_double), ('ms_factor', C.c_double),
                ('ms_solve', C.c_double), ('ms_retract_error', C.c_double), ('ms_total', C.c_double),
                ('n_reduced_dims', C.c_int64), ('n_supernodes', C.c_int64), ('nnz_L', C.c_int64),
                ('n_projections', C.c_int64), ('n_landmarks', C.c_int64),
                ('ms_proj_obs', C.c_double), ('ms_schur_blocks', C.c_double), ('n_schur_pairs', C.c_int64), ('n_levels', C.c_int64), ('allreduce_bytes', C.c_int64), ('nnz_S', C.c_int64)]

    def trace(self):
        n = self.trace_len
        return [dict(lam=self.trace_lambda[i], err=self.trace_error[i], new_err=self.trace_new_error[i],
                     accepted=bool(self.trace_accepted[i])) for i in range(n)]


_dp = C.POINTER(C.c_double)
_kp = C.POINTER(C.c_uint64)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

# name -> (restype, argtypes); this table is also what tests check against include/fg_abi.h
SIGNATURES = {
    'fg_create': (_vp, [C.c_int, C.c_int, C.c_int]),
    'fg_destroy': (None, [_vp]),
    'fg_last_error': (C.c_char_p, [_vp]),
    'fg_abi_version': (C.c_int, []),
    'fg_add_pose': (C.c_int, [_vp, C.c_uint64, _dp]),
    'fg_add_vec3': (C.c_int, [_vp, C.c_uint64, _dp]),
    'fg_add_bias': (C.c_int, [_vp, C.c_uint64, _dp]),
    'fg_add_point': (C.c_int, [_vp, C.c_uint64, _dp]),
    'fg_add_points': (C.c_int, [_vp, C.c_int64, _kp, _dp]),
    'fg_add_plane': (C.c_int, [_vp, C.c_uint64, _dp]),
    'fg_update_value': (C.c_int, [_vp, C.c_uint64, _dp]),
    'fg_exists': (C.c_int, [_vp, C.c_uint64]),
    'fg_get_value': (C.c_int, [_vp, C.c_uint64, _dp, _ip]),
    'fg_num_values': (C.c_int64, [_vp, C.c_int]),
    'fg_get_values': (C.c_int, [_vp, C.c_int, _dp]),
    'fg_set_values': (C.c_int, [_vp, C.c_int, _dp]),
    'fg_add_prior_pose': (C.c_int, [_vp, C.c_uint64, _dp, _dp]),
    'fg_add_prior_vec3': (C.c_int, [_vp, C.c_uint64, _dp, _dp]),
    'fg_add_prior_bias': (C.c_int, [_vp, C.c_uint64, _dp, _dp]),
    'fg_add_prior_point': (C.c_int, [_vp, C.c_uint64, _dp, C.c_double]),
    'fg_add_prior_points': (C.c_int, [_vp, C.c_int64, _kp, _dp, C.c_double]),
    'fg_add_between': (C.c_int, [_vp, C.c_uint64, C.c_uint64, _dp, _dp]),
    'fg_set_calibration': (C.c_int, [_vp, C.c_int, _dp]),
    'fg_set_sensor': (C.c_int, [_vp, C.c_int, _dp]),
    'fg_add_projection': (C.c_int, [_vp, C.c_uint64, C.c_uint64, _dp, C.c_double, C.c_int, C.c_int]),
    'fg_add_projections': (C.c_int, [_vp, C.c_int64, _kp, _kp, _dp, C.c_double, C.c_int, C.c_int]),
    'fg_add_plane_factor': (C.c_int, [_vp, C.c_uint64, C.c_uint64, _dp, _dp]),
    'fg_preintegrate': (C.c_int, [_vp, C.c_int, _ip, _dp, C.c_double, C.POINTER(ImuParams), _dp, C.POINTER(Pim)]),
    'fg_pim_predict': (C.c_int, [C.POINTER(Pim), _dp, _dp, _dp, _dp, _dp]),
    'fg_add_imu': (C.c_int, [_vp, _kp, C.POINTER(Pim)]),
    'fg_lm_params_default': (None, [C.POINTER(LMParams)]),
    'fg_finalize': (C.c_int, [_vp]),
    'fg_optimize_lm': (C.c_int, [_vp, C.POINTER(LMParams), C.POINTER(LMReport)]),
    'fg_error': (C.c_int, [_vp, _dp]),
    'fg_marginal_cov': (C.c_int, [_vp, C.c_uint64, _dp, C.POINTER(C.c_int)]),
    'fg_set_pose_chart': (C.c_int, [_vp, C.c_int]),
    'fg_add_g2o_edge': (C.c_int, [_vp, C.c_uint64, C.c_uint64, _dp, _dp]),
    'fg_set_fixed': (C.c_int, [_vp, C.c_uint64, C.c_int]),
    'fg_g2o_params_default': (None, [C.POINTER(G2OParams)]),
    'fg_optimize_g2o': (C.c_int, [_vp, C.POINTER(G2OParams), C.POINTER(G2OReport)]),
    'fg_g2o_chi2': (C.c_int, [_vp, _dp]),
    'fg_isam2_params_default': (None, [C.POINTER(Isam2Params)]),
    'fg_update_incremental': (C.c_int, [_vp, C.POINTER(Isam2Params), C.POINTER(IncReport)]),
    'fg_debug_counts': (C.c_int64, [_vp, C.c_int]),
    'fg_debug_fp64_peak': (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    'fg_comm_unique_id': (C.c_int, [C.c_char_p]),
    'fg_comm_init': (C.c_int, [_vp, C.c_char_p]),
    'fg_add_structure_edges': (C.c_int, [_vp, C.c_int64, _kp, _kp]),
    'fg_debug_sizes': (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    'fg_debug_symbolic': (C.c_int64, [_vp, C.c_int, C.POINTER(C.c_int64), C.c_int64]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError('libfg_b200.so not built: run `python -c "import __graft_entry__ as g; g.build()"` '
                              'or `make -C graph_slam_b200/csrc` (there is no CPU fallback)')
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def symbol(ch, j):
    """gtsam::Symbol(c, j) key (gtsam_graph.cpp:50-54)."""
    return (ord(ch) << 56) | int(j)


def symbols(ch, idx):
    return (np.uint64(ord(ch)) << np.uint64(56)) | np.asarray(idx).astype(np.uint64)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _k(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(_kp)


def pose12(R, t):
    return np.concatenate([np.asarray(R, dtype=np.float64).reshape(-1, 9), np.asarray(t, dtype=np.float64).reshape(-1, 3)], 1)


class Context:
    """Thin RAII wrapper around fg_ctx*; method names follow the C ABI."""

    def __init__(self, device=0, rank=0, nranks=1):
        self.l = lib()
        self.h = self.l.fg_create(device, rank, nranks)
        if not self.h:
            raise FgError(-4, 'fg_create failed: no usable CUDA device %d (no CPU fallback exists)' % device)

    def close(self):
        if self.h:
            self.l.fg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise FgError(rc, self.l.fg_last_error(self.h).decode())
        return rc

    def call(self, name, *args):
        return self._ck(getattr(self.l, name)(self.h, *args))

    # values
    def add_pose(self, key, T12): a, p = _d(T12); self.call('fg_add_pose', key, p)
    def add_vec3(self, key, v): a, p = _d(v); self.call('fg_add_vec3', key, p)
    def add_bias(self, key, v): a, p = _d(v); self.call('fg_add_bias', key, p)
    def add_point(self, key, v): a, p = _d(v); self.call('fg_add_point', key, p)
    def add_plane(self, key, v): a, p = _d(v); self.call('fg_add_plane', key, p)

    def add_points(self, keys, pts):
        ka, kp = _k(keys); a, p = _d(pts)
        self.call('fg_add_points', len(ka), kp, p)

    def update_value(self, key, v): a, p = _d(v); self.call('fg_update_value', key, p)
    def exists(self, key): return bool(self.l.fg_exists(self.h, key))

    def get_value(self, key):
        out = np.zeros(12)
        n = C.c_int(0)
        self.call('fg_get_value', key, out.ctypes.data_as(_dp), C.byref(n))
        return out[:n.value].copy()

    def num_values(self, t): return int(self.l.fg_num_values(self.h, t))

    def get_values(self, t, out=None):
        n = self.num_values(t)
        if out is None:
            out = np.zeros((n, STORE[t]))
        self.call('fg_get_values', t, out.ctypes.data_as(_dp))
        return out

    def set_values(self, t, arr):
        a, p = _d(arr)
        self.call('fg_set_values', t, p)

    # factors
    def add_prior_pose(self, key, T12, info): a, p = _d(T12); b, q = _d(info); self.call('fg_add_prior_pose', key, p, q)
    def add_prior_vec3(self, key, m, info): a, p = _d(m); b, q = _d(info); self.call('fg_add_prior_vec3', key, p, q)
    def add_prior_bias(self, key, m, info): a, p = _d(m); b, q = _d(info); self.call('fg_add_prior_bias', key, p, q)
    def add_prior_point(self, key, m, sigma): a, p = _d(m); self.call('fg_add_prior_point', key, p, float(sigma))

    def add_prior_points(self, keys, means, sigma):
        ka, kp = _k(keys); a, p = _d(means)
        self.call('fg_add_prior_points', len(ka), kp, p, float(sigma))

    def add_between(self, k1, k2, T12, info): a, p = _d(T12); b, q = _d(info); self.call('fg_add_between', k1, k2, p, q)
    def set_calibration(self, cid, K9): a, p = _d(K9); self.call('fg_set_calibration', cid, p)
    def set_sensor(self, sid, T12): a, p = _d(T12); self.call('fg_set_sensor', sid, p)

    def add_projections(self, kpose, kpoint, uv, sigma, cid=0, sid=0):
        a, pa = _k(kpose); b, pb = _k(kpoint); u, pu = _d(uv)
        self.call('fg_add_projections', len(a), pa, pb, pu, float(sigma), cid, sid)

    def add_plane_factor(self, kpose, kplane, z, cov): a, p = _d(z); b, q = _d(cov); self.call('fg_add_plane_factor', kpose, kplane, p, q)

    def add_imu(self, keys6, pim):
        ka, kp = _k(keys6)
        self.call('fg_add_imu', kp, C.byref(pim))

    def preintegrate(self, offsets, imu6, dt, params, bias_hat):
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        n = len(offsets) - 1
        a, pa = _d(imu6); b, pb = _d(bias_hat)
        out = (Pim * n)()
        self.call('fg_preintegrate', n, offsets.ctypes.data_as(_ip), pa, float(dt), C.byref(params), pb, out)
        return out

    # optimise
    def finalize(self): self.call('fg_finalize')

    def error(self):
        e = C.c_double(0)
        self.call('fg_error', C.byref(e))
        return e.value

    def optimize(self, **kw):
        p = LMParams()
        self.l.fg_lm_params_default(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        rep = LMReport()
        self.call('fg_optimize_lm', C.byref(p), C.byref(rep))
        return rep

    def set_pose_chart(self, chart): self.call('fg_set_pose_chart', int(chart))
    def add_g2o_edge(self, k1, k2, T12, info): a, p = _d(T12); b, q = _d(info); self.call('fg_add_g2o_edge', k1, k2, p, q)
    def set_fixed(self, key, fixed=True): self.call('fg_set_fixed', key, int(bool(fixed)))

    def optimize_g2o(self, **kw):
        """CGraphG2O::optimizeGraph (g2o/g2o_graph.cpp:241-252)."""
        p = G2OParams()
        self.l.fg_g2o_params_default(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        rep = G2OReport()
        self.call('fg_optimize_g2o', C.byref(p), C.byref(rep))
        return rep

    def g2o_chi2(self):
        out = C.c_double(0)
        self.call('fg_g2o_chi2', C.byref(out))
        return out.value

    def update_incremental(self, **kw):
        """isam2->update(new factors, new values) + calculateEstimate() (CGraphGT::optimizeGraphIncremental, gtsam_graph.cpp:1768-1776)."""
        p = Isam2Params()
        self.l.fg_isam2_params_default(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        rep = IncReport()
        self.call('fg_update_incremental', C.byref(p), C.byref(rep))
        return rep

    def marginal_covariance(self, key):
        """Marginals(graph, values).marginalCovariance(key) (gtsam_graph.cpp:598-601) for a pose / velocity / bias / plane key."""
        out = np.zeros(36)
        n = C.c_int(0)
        self.call('fg_marginal_cov', int(key), out.ctypes.data_as(_dp), C.byref(n))
        return out[:n.value * n.value].reshape(n.value, n.value).copy()

    def debug_sizes(self):
        """[n_r, n_supernodes, nnz_L, max_nrows, max_ncols, flops of one factorisation, n_projections, n_landmarks]"""
        out = (C.c_int64 * 8)()
        self.call('fg_debug_sizes', out)
        return [int(v) for v in out]

    def comm_init(self, uid): self.call('fg_comm_init', uid)

    def add_structure_edges(self, ka, kb):
        a, pa = _k(ka); b, pb = _k(kb)
        self.call('fg_add_structure_edges', len(a), pa, pb)

    def symbolic(self, which):
        n = self.l.fg_debug_symbolic(self.h, which, None, 0)
        if n < 0:
            raise FgError(int(n), self.l.fg_last_error(self.h).decode())
        out = np.zeros(max(n, 1), dtype=np.int64)
        self.l.fg_debug_symbolic(self.h, which, out.ctypes.data_as(C.POINTER(C.c_int64)), n)
        return out[:n]


def fp64_peak(device=0):
    """(DFMA, DMMA) TFLOP/s measured on `device` now (fg_debug_fp64_peak)."""
    out = (C.c_double * 2)()
    rc = lib().fg_debug_fp64_peak(device, out)
    if rc != 0:
        raise FgError(rc, 'fg_debug_fp64_peak failed')
    return float(out[0]), float(out[1])


def launch_count():
    """Kernels this process has launched through the library so far (fg_debug_counts(NULL, 100))."""
    return int(lib().fg_debug_counts(None, 100))


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib().fg_comm_unique_id(buf)
    if rc != 0:
        raise FgError(rc, 'fg_comm_unique_id failed (libnccl.so.2 not loadable?)')
    return buf.raw


def make_imu_params(acc_cov, gyro_cov, int_cov, bias_acc_cov, bias_gyro_cov, bias_acc_omega_int, gravity):
    p = ImuParams()
    for name, val, n in (('acc_cov', acc_cov, 9), ('gyro_cov', gyro_cov, 9), ('int_cov', int_cov, 9),
                         ('bias_acc_cov', bias_acc_cov, 9), ('bias_gyro_cov', bias_gyro_cov, 9),
                         ('bias_acc_omega_int', bias_acc_omega_int, 36), ('gravity', gravity, 3)):
        arr = np.asarray(val, dtype=np.float64).ravel()
        assert arr.size == n
        getattr(p, name)[:] = arr.tolist()
    return p


def vn100_imu_params():
    """CImuVn100::getIMUParams (gtsam/imu_vn100.cpp:24-67) + MakeSharedD(9.71) (gtsam/imu_base.cpp:258-263)."""
    g = 9.81
    fps = 200.0
    I3 = np.eye(3)
    return make_imu_params(I3 * (0.14e-3 * g) ** 2, I3 * np.deg2rad(0.0035) ** 2, I3 * 1e-4,
                           I3 * ((0.04e-3 * g) * np.sqrt(fps)) ** 2,
                           I3 * ((np.deg2rad(10.0) / 3600.0) * np.sqrt(fps)) ** 2, np.eye(6) * 1e-3, [0, 0, 9.71])


def covisibility_band(spec):
    """Pose pairs (p, q), p < q, covering every landmark's pose span [lo, hi]: a superset of the true
    co-visibility that every rank can compute cheaply and declare with fg_add_structure_edges."""
    P = spec['n_poses']
    pose, pt = spec['proj_pose'].astype(np.int64), spec['proj_point'].astype(np.int64)
    L = int(pt.max()) + 1 if len(pt) else 0
    lo = np.full(L, P, dtype=np.int64); hi = np.full(L, -1, dtype=np.int64)
    np.minimum.at(lo, pt, pose); np.maximum.at(hi, pt, pose)
    reach = np.full(P, -1, dtype=np.int64)
    ok = hi >= 0
    np.maximum.at(reach, lo[ok], hi[ok])
    reach = np.maximum.accumulate(reach)
    a = np.repeat(np.arange(P), np.maximum(reach - np.arange(P), 0))
    b = np.concatenate([np.arange(p + 1, r + 1) for p, r in enumerate(reach) if r > p]) if len(a) else np.zeros(0, dtype=np.int64)
    return a, b


def shard_landmarks(L, rank, world, block=96):
    """Landmark ids of one rank's shard: blocks of `block` consecutive landmarks dealt round-robin.  Landmarks are ordered along
    the trajectory, so a contiguous slice would give a rank 1/world of the POSES -- 1/world of the Schur tiles, too few to fill
    the GPU, the longest tile setting the time (measured at 8 GPUs: k_schur_tiles 1.02 ms against 3.61 / 8 = 0.45 ms).  Dealt in
    blocks of one staging round of k_schur_tiles (96 landmarks), every rank holds every tile with 1/world of its rounds."""
    nb = (L + block - 1) // block
    ids = [np.arange(b * block, min(L, (b + 1) * block)) for b in range(rank, nb, world)]
    return np.concatenate(ids) if ids else np.zeros(0, dtype=np.int64)


def load_spec(ctx, spec, landmark_slice=None, preintegrated=None, landmark_ids=None):
    """Build the graph of a synth.GraphSpec in `ctx` through the C ABI, wiring factors the way the
    reference does (firstNode priors gtsam_graph.cpp:320-368, Between :630-695, CombinedImuFactor
    test_vro_imu_graph.cpp:191-196, point priors + projections :370-448, plane factors :1265).
    landmark_slice = (lo, hi) or landmark_ids = increasing id array restricts point landmarks to a shard (multi-GPU).
    Returns the Pim array (device-preintegrated) when the spec has IMU data."""
    P = spec['n_poses']
    X = symbols('x', np.arange(P)); V = symbols('v', np.arange(P)); B = symbols('b', np.arange(P))
    T = pose12(spec['pose_init_R'], spec['pose_init_t'])
    for i in range(P):
        ctx.add_pose(int(X[i]), T[i])
    ctx.add_prior_pose(int(X[0]), pose12(spec['prior_pose_R'], spec['prior_pose_t'])[0], np.eye(6) / 1e-7 ** 2)
    pims = None
    if 'imu_samples' in spec:
        for i in range(P):
            ctx.add_vec3(int(V[i]), spec['vel_init'][i])
            ctx.add_bias(int(B[i]), spec['bias_init'][i])
        ctx.add_prior_vec3(int(V[0]), spec['prior_vel_mean'], np.eye(3) / 1e-3 ** 2)
        ctx.add_prior_bias(int(B[0]), np.zeros(6), np.eye(6) / 1e-3 ** 2)
        if preintegrated is None:
            S = spec['imu_samples'].shape[1]
            offsets = np.arange(P) * S
            pims = ctx.preintegrate(offsets, spec['imu_samples'].reshape(-1, 6), spec['imu_dt'], vn100_imu_params(),
                                    np.zeros((P - 1, 6)))
        else:
            pims = preintegrated
        for i in range(P - 1):
            ctx.add_imu([X[i], V[i], X[i + 1], V[i + 1], B[i], B[i + 1]], pims[i])
    if 'between_i' in spec:
        Tm = pose12(spec['between_R'], spec['between_t'])
        for n in range(len(spec['between_i'])):
            ctx.add_between(int(X[spec['between_i'][n]]), int(X[spec['between_j'][n]]), Tm[n], spec['between_info'][n])
    if 'proj_pose' in spec:
        L = len(spec['point_init'])
        if landmark_ids is None:
            lo, hi = (0, L) if landmark_slice is None else landmark_slice
            ids = np.arange(lo, hi)
            m = (spec['proj_point'] >= lo) & (spec['proj_point'] < hi)
        else:
            ids = np.asarray(landmark_ids, dtype=np.int64)
            mine = np.zeros(L, dtype=bool); mine[ids] = True
            m = mine[spec['proj_point']]
        Q = symbols('q', ids)
        ctx.set_calibration(0, spec['K'])
        ctx.set_sensor(0, pose12(spec['Rs'], spec['ts'])[0])
        ctx.add_points(Q, spec['point_init'][ids])
        ctx.add_prior_points(Q, spec['point_init'][ids], spec['point_prior_sigma'])
        if landmark_slice is not None or landmark_ids is not None:
            a, b = covisibility_band(spec)
            ctx.add_structure_edges(X[a], X[b])
        if 'proj_cal' in spec:                                   # several (Cal3DS2, body_P_sensor) pairs: slot c holds pair c
            for cidx, (Kc, Rsc, tsc) in enumerate(spec['cals']):
                ctx.set_calibration(cidx, np.asarray(Kc, dtype=np.float64))
                ctx.set_sensor(cidx, pose12(Rsc, tsc)[0])
                mc = m & (np.asarray(spec['proj_cal']) == cidx)
                if mc.any():
                    ctx.add_projections(X[spec['proj_pose'][mc]], symbols('q', spec['proj_point'][mc]), spec['proj_uv'][mc],
                                        spec['proj_sigma'], cid=cidx, sid=cidx)
        else:
            ctx.add_projections(X[spec['proj_pose'][m]], symbols('q', spec['proj_point'][m]), spec['proj_uv'][m],
                                spec['proj_sigma'])
    if 'plane_init' in spec:
        for l, pl in enumerate(spec['plane_init']):
            ctx.add_plane(symbol('l', l), pl)
        for n in range(len(spec['plane_obs_pose'])):
            ctx.add_plane_factor(int(X[spec['plane_obs_pose'][n]]), symbol('l', int(spec['plane_obs_plane'][n])),
                                 spec['plane_meas'][n], spec['plane_cov'][n])
    return pims
