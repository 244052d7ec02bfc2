"""Timing baseline on the host cores (test infrastructure): ctypes front end of oracle/cpu_lm.cpp, an OpenMP C++ port of
one LM iteration of the BA + IMU graph (see that file's header).  Used by bench.py's `cpu_baseline` / `--impl reference`
legs and checked against the numpy oracle in tests/test_cpu_baseline.py.  Never imported by the product."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'cpu_lm.cpp')
LIB = os.path.join(HERE, '_build', 'libcpu_lm.so')
_lib = None


def build(force=False):
    deps = [SRC] + [os.path.join(HERE, '..', 'graph_slam_b200', 'csrc', f) for f in ('fg_math.cuh', 'fg_factors.cuh')]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(['g++', '-O3', '-mavx2', '-mfma', '-fopenmp', '-shared', '-fPIC', '-Wno-unknown-pragmas', SRC, '-o', LIB])
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.cpu_lm_iteration.restype = C.c_int
        _lib.cpu_lm_max_threads.restype = C.c_int
    return _lib


class ImuRec(C.Structure):      # graph_slam_b200/csrc/fg_factors.cuh: struct ImuRec
    _fields_ = [('dt', C.c_double), ('preint', C.c_double * 9), ('Hba', C.c_double * 27), ('Hbg', C.c_double * 27),
                ('bias_hat', C.c_double * 6), ('gravity', C.c_double * 3), ('info', C.c_double * 225)]


class State:
    """Mutable state + constant factor data of a BA + IMU graph spec (graph_slam_b200.synth), laid out for cpu_lm_iteration."""

    def __init__(self, spec):
        from . import imu as oimu
        P = spec['n_poses']
        self.P = P
        self.pose = np.ascontiguousarray(np.concatenate([spec['pose_init_R'].reshape(P, 9), spec['pose_init_t']], 1))
        self.vel = np.ascontiguousarray(spec['vel_init'], dtype=np.float64).copy()
        self.bias = np.ascontiguousarray(spec['bias_init'], dtype=np.float64).copy()
        self.pp_T = np.ascontiguousarray(np.concatenate([np.asarray(spec['prior_pose_R']).reshape(9), np.asarray(spec['prior_pose_t']).reshape(3)]))
        self.pp_info = np.ascontiguousarray(np.eye(6) / 1e-7 ** 2)
        self.pv_mean = np.ascontiguousarray(spec['prior_vel_mean'], dtype=np.float64)
        self.pv_info = np.ascontiguousarray(np.eye(3) / 1e-3 ** 2)
        self.pb_mean = np.zeros(6)
        self.pb_info = np.ascontiguousarray(np.eye(6) / 1e-3 ** 2)
        pim = oimu.preintegrate(spec['imu_samples'], spec['imu_dt'], oimu.vn100_params(), np.zeros((P - 1, 6)))
        info = np.linalg.inv(pim['cov'])
        self.imu = (ImuRec * (P - 1))()
        for i in range(P - 1):
            r = self.imu[i]
            r.dt = float(pim['dt'][i])
            r.preint[:] = pim['preint'][i].tolist(); r.Hba[:] = pim['Hba'][i].ravel().tolist(); r.Hbg[:] = pim['Hbg'][i].ravel().tolist()
            r.bias_hat[:] = pim['bias_hat'][i].tolist(); r.gravity[:] = np.asarray(pim['gravity']).tolist(); r.info[:] = info[i].ravel().tolist()
        order = np.argsort(spec['proj_point'], kind='stable')
        self.pts = np.ascontiguousarray(spec['point_init'], dtype=np.float64).copy()
        self.pt_mean = np.ascontiguousarray(spec['point_init'], dtype=np.float64).copy()
        self.pt_sigma = float(spec['point_prior_sigma'])
        self.obs_pose = np.ascontiguousarray(spec['proj_pose'][order], dtype=np.int32)
        self.obs_point = np.ascontiguousarray(spec['proj_point'][order], dtype=np.int32)
        self.obs_uv = np.ascontiguousarray(spec['proj_uv'][order], dtype=np.float64)
        self.obs_sigma = float(spec['proj_sigma'])
        self.K = np.ascontiguousarray(spec['K'], dtype=np.float64)
        self.sensor = np.ascontiguousarray(np.concatenate([np.asarray(spec['Rs']).reshape(9), np.asarray(spec['ts']).reshape(3)]))

    def iterate(self, lam, nthreads=0):
        """One LM trial at damping `lam`; the state moves to the trial point.  Returns (status, error before, error after, seconds[5])."""
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        out = np.zeros(8)
        rc = lib().cpu_lm_iteration(C.c_int(self.P), dp(self.pose), dp(self.vel), dp(self.bias), dp(self.pp_T), dp(self.pp_info), dp(self.pv_mean),
                                    dp(self.pv_info), dp(self.pb_mean), dp(self.pb_info), C.c_int(self.P - 1), C.byref(self.imu),
                                    C.c_int64(len(self.pts)), dp(self.pts), dp(self.pt_mean), C.c_double(self.pt_sigma), C.c_int64(len(self.obs_pose)),
                                    ip(self.obs_pose), ip(self.obs_point), dp(self.obs_uv), C.c_double(self.obs_sigma), dp(self.K), dp(self.sensor),
                                    C.c_double(lam), C.c_int(nthreads), dp(out))
        return rc, out[0], out[1], out[2:7].copy(), int(out[7])


def max_threads():
    return int(lib().cpu_lm_max_threads())
