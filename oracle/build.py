"""GraphSpec (graph_slam_b200.synth) -> oracle.Graph (oracle; test infrastructure).

Mirrors how the reference wires factors: firstNode priors (gtsam/gtsam_graph.cpp:320-368),
BetweenFactor per VRO edge (:630-695), CombinedImuFactor per frame pair
(gtsam/test_vro_imu_graph.cpp:191-196), PriorFactor<Point3> + GenericProjectionFactor
(gtsam/gtsam_graph.cpp:370-448), OrientedPlane3Factor (:1265).
"""
import numpy as np
from .graph import Graph
from . import imu as oimu


def from_spec(spec):
    g = Graph()
    P = spec['n_poses']
    g.R = spec['pose_init_R'].copy(); g.t = spec['pose_init_t'].copy()
    g.K = tuple(spec['K']); g.Rs = spec['Rs']; g.ts = spec['ts']
    f = {}
    s7 = 1e-7
    f['prior_pose'] = dict(i=np.array([0]), R=spec['prior_pose_R'][None], t=spec['prior_pose_t'][None],
                           info=np.eye(6)[None] / s7 ** 2)
    if 'imu_samples' in spec:
        g.vel = spec['vel_init'].copy(); g.bias = spec['bias_init'].copy()
        f['prior_vel'] = dict(i=np.array([0]), mean=spec['prior_vel_mean'][None].copy(), info=np.eye(3)[None] / 1e-3 ** 2)
        f['prior_bias'] = dict(i=np.array([0]), mean=np.zeros((1, 6)), info=np.eye(6)[None] / 1e-3 ** 2)
        par = oimu.vn100_params()
        pim = oimu.preintegrate(spec['imu_samples'], spec['imu_dt'], par, np.zeros((P - 1, 6)))
        a = np.arange(P - 1)
        f['imu'] = dict(pi=a, vi=a, pj=a + 1, vj=a + 1, bi=a, bj=a + 1, pim=pim, info=np.linalg.inv(pim['cov']))
    if 'between_i' in spec:
        f['between'] = dict(i=spec['between_i'].astype(np.int64), j=spec['between_j'].astype(np.int64),
                            R=spec['between_R'], t=spec['between_t'], info=spec['between_info'])
    if 'proj_pose' in spec:
        g.point = spec['point_init'].copy()
        L = len(g.point)
        f['prior_point'] = dict(i=np.arange(L), mean=spec['point_init'].copy(),
                                info=np.broadcast_to(np.eye(3) / spec['point_prior_sigma'] ** 2, (L, 3, 3)))
        f['proj'] = dict(i=spec['proj_pose'].astype(np.int64), l=spec['proj_point'].astype(np.int64),
                         uv=spec['proj_uv'], sigma=spec['proj_sigma'])
        if 'proj_cal' in spec:                                   # mixed calibrations: spec['cals'] = [(K, Rs, ts), ...]
            f['proj']['cal'] = np.asarray(spec['proj_cal'], dtype=np.int64)
            g.cals = [(tuple(K), Rs, ts) for K, Rs, ts in spec['cals']]
    if 'plane_init' in spec:
        g.plane = spec['plane_init'].copy()
        f['plane'] = dict(i=spec['plane_obs_pose'].astype(np.int64), l=spec['plane_obs_plane'].astype(np.int64),
                          meas=spec['plane_meas'], info=np.linalg.inv(spec['plane_cov']))
    g.f = f
    return g
