"""Writes tests/golden/export/: the BASELINE pose-graph / VIO configurations as files an OUTSIDE GTSAM / g2o installation can
load, with the numbers this repository's oracle gets on them, so that the "parity unpinned" rows (Between,
Prior, CombinedImu, LM control, g2o) can be closed by anyone who has the real libraries:

  c1.g2o                      C1 (100 poses, 500 edges): VERTEX_SE3:QUAT / EDGE_SE3:QUAT, information in g2o's [trans, rot] order
  c2_vro.log c2_imu.log c2_times.log          C2 at scale 0.05 in the reference's own log formats:
  c3_vro.log c3_imu.log c3_times.log c3_planes.txt   feed them to test_vro_imu_graph / tests/cpp/plane_driver.cpp
  expected.json               initial error, LM trace and optimum of the oracle (charts: full EXPMAP), g2o chi2 trace for c1

    python tests/golden/export_graphs.py
"""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'export')


def fmt(v):
    return '%.12g' % float(v)


def write_g2o(spec, path):
    from oracle import lie
    Pm = np.zeros((6, 6)); Pm[:3, 3:] = np.eye(3); Pm[3:, :3] = np.eye(3)
    with open(path, 'w') as f:
        q = lie.quat_from_rot(spec['pose_init_R'])
        for i in range(spec['n_poses']):
            t = spec['pose_init_t'][i]
            f.write('VERTEX_SE3:QUAT %d %s %s\n' % (i, ' '.join(fmt(v) for v in t), ' '.join(fmt(v) for v in (q[i, 1], q[i, 2], q[i, 3], q[i, 0]))))
        f.write('FIX 0\n')
        qm = lie.quat_from_rot(spec['between_R'])
        for n in range(len(spec['between_i'])):
            info = Pm @ spec['between_info'][n] @ Pm.T
            f.write('EDGE_SE3:QUAT %d %d %s %s %s\n' % (spec['between_i'][n], spec['between_j'][n], ' '.join(fmt(v) for v in spec['between_t'][n]),
                                                       ' '.join(fmt(v) for v in (qm[n, 1], qm[n, 2], qm[n, 3], qm[n, 0])),
                                                       ' '.join(fmt(info[i, j]) for i in range(6) for j in range(i, 6))))


def main():
    from graph_slam_b200 import synth
    from oracle import build, lm
    import driver_logs
    os.makedirs(OUT, exist_ok=True)
    exp = {}
    # ---- C1: the pose graph, as .g2o; GTSAM semantics (with the 1e-7 prior CGraphGT::firstNode adds) and g2o semantics (vertex 0 fixed)
    spec = synth.make_config('C1', seed=1)
    write_g2o(spec, os.path.join(OUT, 'c1.g2o'))
    g0 = build.from_spec(spec)
    g1, rep = lm.optimize_gtsam(g0)
    Pm = np.zeros((6, 6)); Pm[:3, 3:] = np.eye(3); Pm[3:, :3] = np.eye(3)
    pg = lm.PoseGraphG2O(spec['pose_init_R'], spec['pose_init_t'], spec['between_i'], spec['between_j'], spec['between_R'], spec['between_t'],
                         Pm @ spec['between_info'] @ Pm.T, fixed=(0,))
    chi0 = pg.chi2()
    pg, grep = lm.optimize_g2o_calls(pg)
    exp['c1'] = dict(file='c1.g2o', gtsam=dict(note='BetweenFactor<Pose3> per edge + PriorFactor<Pose3>(X0, the value of vertex 0, sigmas 1e-7); LevenbergMarquardtOptimizer defaults',
                                               initial_error=g0.error(), final_error=rep['error'], iterations=rep['iterations'],
                                               trace_lambda=[t['lam'] for t in rep['trace']], trace_accepted=[bool(t['accepted']) for t in rep['trace']],
                                               last_pose_t=g1.t[-1].tolist()),
                     g2o=dict(note='EdgeSE3 per edge, vertex 0 fixed, OptimizationAlgorithmLevenberg, 10 x optimize(2)', initial_chi2=chi0, final_chi2=grep['chi2'],
                              chi2_after_each_iteration=[t['chi2'] for t in grep['trace'][:6]], last_pose_t=pg.t[-1].tolist()))
    # ---- C2 / C3: reference-format logs
    for name in ('C2', 'C3'):
        spec = synth.make_config(name, seed=1, scale=0.05)
        k = name.lower()
        recs = driver_logs.write_logs(spec, *(os.path.join(OUT, '%s_%s.log' % (k, s)) for s in ('vro', 'imu', 'times')), fmt=fmt)
        g0 = driver_logs.oracle_graph_from_logs(spec, recs)
        entry = dict(files=['%s_vro.log' % k, '%s_imu.log' % k, '%s_times.log' % k], n_poses=int(spec['n_poses']),
                     note='CGraphGT::firstNode priors, BetweenFactor per VRO record (u2c = setCamera2IMU(0)), CombinedImuFactor per frame from the IMU log '
                          '(VN100 parameters, gravity 9.71, zero bias); values dead-reckoned through the (j-1, j) records and the IMU prediction')
        if name == 'C3':
            with open(os.path.join(OUT, 'c3_planes.txt'), 'w') as f:
                for n in range(len(spec['plane_obs_pose'])):
                    f.write('%d %d %s %s\n' % (spec['plane_obs_pose'][n], spec['plane_obs_plane'][n], ' '.join(fmt(v) for v in spec['plane_meas'][n]),
                                              ' '.join(fmt(v) for v in spec['plane_cov'][n].ravel())))
            g0.plane = spec['plane_init'].copy()
            g0.f['plane'] = dict(i=spec['plane_obs_pose'].astype(np.int64), l=spec['plane_obs_plane'].astype(np.int64), meas=spec['plane_meas'],
                                 info=np.linalg.inv(spec['plane_cov']))
            entry['files'].append('c3_planes.txt')
            entry['planes'] = dict(note='c3_planes.txt: pose_id landmark_id, measured plane (n, d) in the IMU frame, 3x3 covariance of the OrientedPlane3 '
                                        'tangent; OrientedPlane3Factor(z, Gaussian::Covariance(S), X(pose), L(landmark)); landmark initial values below',
                                   initial=spec['plane_init'].tolist())
        g1, rep = lm.optimize_gtsam(g0)
        entry.update(initial_error=g0.error(), final_error=rep['error'], iterations=rep['iterations'], trace_lambda=[t['lam'] for t in rep['trace']],
                     trace_accepted=[bool(t['accepted']) for t in rep['trace']], last_pose_t=g1.t[-1].tolist(), last_velocity=g1.vel[-1].tolist())
        exp[k] = entry
    json.dump(exp, open(os.path.join(OUT, 'expected.json'), 'w'), indent=1)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
