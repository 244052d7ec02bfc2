"""The reference's OWN wrapper sources (gtsam/gtsam_graph.cpp, imu_base.cpp, imu_vn100.cpp, g2o/g2o_graph.cpp) and its two
offline drivers, compiled UNCHANGED over compat/ + the gtsam / g2o facades and linked with libfg_b200.so
(compat/build_ref.py), run on reference-format text logs against the oracle on the graph those logs define."""
import os
import subprocess
import numpy as np
import pytest
from graph_slam_b200 import synth
from oracle import lm, lie
import driver_logs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_sources_compile_unchanged_and_link(refbin):
    """CPU-side check: every reference file on the path builds from /root/reference with -Icompat and
    links libfg_b200.so -- the wrapper libraries, both offline drivers and this repository's test programs."""
    for name in ('libgraphslam_gt.so', 'libgraphslam_g2o.so', 'test_vro_imu_graph', 'test_ba_imu_graph', 'vio_driver', 'ba_driver', 'format_io', 'plane_driver', 'plane_check', 'g2o_driver'):
        assert os.path.exists(refbin(name))


@pytest.mark.gpu
def test_vio_driver_matches_oracle(refbin, tmp_path):
    exe = refbin('vio_driver')
    spec = synth.make_config('C2', seed=1, scale=0.06)
    vro, imu, times, out = (str(tmp_path / n) for n in ('vro.log', 'imu.log', 'times.log', 'poses.txt'))
    recs = driver_logs.write_logs(spec, vro, imu, times)
    g0 = driver_logs.oracle_graph_from_logs(spec, recs)
    res = subprocess.run([exe, vro, imu, times, out], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith('RESULT')][0].split()
    e0, e1 = float(line[4]), float(line[6])
    assert int(line[2]) == spec['n_poses']
    assert abs(e0 - g0.error()) <= 1e-9 * g0.error()
    g1, rep = lm.optimize_gtsam(g0)
    assert abs(e1 - rep['error']) <= 1e-9 * rep['error']
    got = np.loadtxt(out)
    R = got[:, 1:10].reshape(-1, 3, 3); t = got[:, 10:13]; v = got[:, 13:16]
    assert np.linalg.norm(lie.so3_log(np.swapaxes(R, -1, -2) @ g1.R), axis=-1).max() <= 1e-8
    assert np.abs(t - g1.t).max() <= 1e-8 and np.abs(v - g1.vel).max() <= 1e-7


@pytest.mark.gpu
def test_vio_driver_incremental_per_frame(refbin, tmp_path):
    """The loop of gtsam/test_vro_imu_graph.cpp with its per-frame optimizeGraphIncremental() (:344-350): ISAM2 update per
    frame through the C++ facade (gtsam_lite.h: ISAM2 -> fg_update_incremental), the integrator re-seeded from the
    estimate, one batch LM at the end.  The ISAM2 estimate after the last frame is within the relinearisation threshold
    of the batch optimum; per-frame cost is printed."""
    exe = refbin('vio_driver')
    spec = synth.make_config('C2', seed=1, scale=float(os.environ.get('FG_INC_SCALE', '0.2')))
    vro, imu, times, out = (str(tmp_path / n) for n in ('vro.log', 'imu.log', 'times.log', 'poses.txt'))
    driver_logs.write_logs(spec, vro, imu, times)
    res = subprocess.run([exe, vro, imu, times, out, 'incremental'], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    inc = [l for l in res.stdout.splitlines() if l.startswith('INCREMENTAL')][0].split()
    line = [l for l in res.stdout.splitlines() if l.startswith('RESULT')][0].split()
    print(' '.join(inc))
    assert int(inc[2]) == spec['n_poses'] - 1 and int(line[2]) == spec['n_poses']
    e0, e1 = float(line[4]), float(line[6])
    assert e1 <= e0 * (1 + 1e-12)
    est = np.loadtxt(out + '.isam2'); fin = np.loadtxt(out)
    assert np.abs(est[:, 10:13] - fin[:, 10:13]).max() < 0.1 and np.abs(est[:, 1:10] - fin[:, 1:10]).max() < 0.1
    # the dead-reckoned chain drifts by metres over the sequence; the per-frame estimate must already be close to the optimum
    assert np.abs(est[:, 10:13] - fin[:, 10:13]).max() < 0.02


@pytest.mark.gpu
def test_ba_builder_and_bundle_adjust_match_oracle(refbin, tmp_path):
    """CGraphGT::addToGTSAM(CCameraNodeBA*, ...) (gtsam_graph.cpp:370-448) + optimizeGraphBatch, and
    CGraphGT::bundleAdjust (:500-610), through the C++ mirror, against the oracle on the graphs they define."""
    from oracle import build, factors as ofac
    from oracle.graph import Graph
    exe = refbin('ba_driver')
    rng = np.random.default_rng(31)
    P, n = 6, 60
    K = np.array([250.5773, 250.5773, 0, 90, 70, -0.8466, 0.5370, 0, 0])
    I3, z3 = np.eye(3), np.zeros(3)
    r_true = np.tile(np.array([0.01, -0.015, 0.02, 0.07, -0.01, 0.02]), (P - 1, 1)) * (1 + 0.1 * rng.normal(size=(P - 1, 1)))
    Rt, tt = [np.eye(3)], [np.zeros(3)]
    for k in range(P - 1):
        dR, dt = lie.se3_exp(r_true[k:k + 1])
        R, t = lie.pose_compose(Rt[-1], tt[-1], dR[0], dt[0]); Rt.append(R); tt.append(t)
    Rt, tt = np.array(Rt), np.array(tt)
    pts = np.column_stack([rng.uniform(-0.5, 0.5, n), rng.uniform(-0.4, 0.4, n), rng.uniform(1.5, 4.0, n)])      # world = camera 0
    pc = np.einsum('pji,pnj->pni', Rt, pts[None] - tt[:, None])                                                  # camera-frame points
    pc32 = (pc + rng.normal(size=pc.shape) * 0.01).astype(np.float32)
    uv = np.stack([ofac.projection(Rt[p], tt[p], pts, np.zeros((n, 2)), K, I3, z3, jac=False) for p in range(P)])
    uv32 = (uv + rng.normal(size=uv.shape)).astype(np.float32)
    with open(tmp_path / 'features.txt', 'w') as f:
        f.write('%d %d %r %r %r %r %r %r\n' % (P, n, *[float(v) for v in K[[0, 1, 3, 4, 5, 6]]]))
        for p in range(P):
            for k in range(n):
                f.write('%r %r %r %r %r\n' % (float(pc32[p, k, 0]), float(pc32[p, k, 1]), float(pc32[p, k, 2]), float(uv32[p, k, 0]), float(uv32[p, k, 1])))
    r_meas = r_true + rng.normal(size=r_true.shape) * np.array([0.005] * 3 + [0.01] * 3)
    info = np.diag([400.0, 400, 400, 100, 100, 100])
    with open(tmp_path / 'vro.log', 'w') as f:
        for k in range(P - 1):
            f.write('%d %d %s %s\n' % (k + 1, k, ' '.join(repr(float(x)) for x in r_meas[k]),
                                      ' '.join(repr(float(info[i, j])) for i in range(6) for j in range(i, 6))))
    res = subprocess.run([exe, str(tmp_path / 'features.txt'), str(tmp_path / 'vro.log'), str(tmp_path / 'out.txt')],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith('RESULT')][0].split()
    assert int(line[2]) == P and int(line[6]) == n and int(line[12]) == 1
    e0, e1 = float(line[8]), float(line[10])
    out = np.loadtxt(tmp_path / 'out.txt')
    # ---- oracle of the multi-frame graph: dead-reckoned poses, landmarks from frame 0, priors and projections
    Rm, tm = lie.se3_exp(r_meas)
    Ri, ti = [np.eye(3)], [np.zeros(3)]
    for k in range(P - 1):
        R, t = lie.pose_compose(Ri[-1], ti[-1], Rm[k], tm[k]); Ri.append(R); ti.append(t)
    spec = dict(name='ba_driver', seed=0, n_poses=P, K=K, Rs=I3, ts=z3, pose_init_R=np.array(Ri), pose_init_t=np.array(ti),
                prior_pose_R=I3, prior_pose_t=z3, between_i=np.arange(P - 1), between_j=np.arange(1, P), between_R=Rm, between_t=tm,
                between_info=np.broadcast_to(info, (P - 1, 6, 6)).copy(),
                point_init=pc32[0].astype(np.float64), point_prior_sigma=0.014,
                proj_pose=np.repeat(np.arange(P), n).astype(np.int32), proj_point=np.tile(np.arange(n), P).astype(np.int32),
                proj_uv=uv32.reshape(-1, 2).astype(np.float64), proj_sigma=1.0)
    g0 = build.from_spec(spec)
    assert abs(e0 - g0.error()) <= 1e-9 * g0.error()
    g1, rep = lm.optimize_gtsam(g0, solver='schur')
    assert abs(e1 - rep['error']) <= 1e-7 * rep['error']
    R = out[:P, :9].reshape(-1, 3, 3); t = out[:P, 9:12]
    assert np.linalg.norm(lie.so3_log(np.swapaxes(R, -1, -2) @ g1.R), axis=-1).max() <= 1e-6 and np.abs(t - g1.t).max() <= 1e-6
    # ---- oracle of bundleAdjust on the first edge: two poses, no body_P_sensor, SR4000 calibration
    spec2 = dict(name='two_view', seed=0, n_poses=2, K=K, Rs=I3, ts=z3, pose_init_R=np.stack([I3, I3]), pose_init_t=np.zeros((2, 3)),
                 prior_pose_R=I3, prior_pose_t=z3, point_init=pc32[0].astype(np.float64), point_prior_sigma=0.014,
                 proj_pose=np.repeat(np.arange(2), n).astype(np.int32), proj_point=np.tile(np.arange(n), 2).astype(np.int32),
                 proj_uv=uv32[:2].reshape(-1, 2).astype(np.float64), proj_sigma=1.0)
    h1, rep2 = lm.optimize_gtsam(build.from_spec(spec2), solver='schur')
    Tj = out[P]
    assert np.abs(Tj[9:12] - h1.t[1]).max() <= 1e-6 and np.abs(Tj[:9].reshape(3, 3) - h1.R[1]).max() <= 1e-6
    H, grad, err = h1.normal_equations()
    cov = np.linalg.inv(H.toarray())[6:12, 6:12]
    info_ref = np.linalg.inv(cov)
    info_got = out[P + 1:P + 7, :6]
    assert np.abs(info_got - info_ref).max() <= 1e-5 * np.abs(info_ref).max()


def _read_traj(path):
    a = np.loadtxt(path)                      # id x y z qx qy qz qw seq_id   (CGraphGT::writeTrajectory, gtsam_graph.cpp:1819-1840)
    q = np.concatenate([a[:, 7:8], a[:, 4:7]], 1)
    return a[:, 1:4], lie.rot_from_quat(q)


@pytest.mark.gpu
@pytest.mark.parametrize('driver', ['test_vro_imu_graph', 'test_ba_imu_graph'])
def test_reference_offline_drivers_run_unchanged(refbin, tmp_path, driver):
    """The reference's own drivers (their own main(), ROS parameters and all), compiled unchanged, on synthetic
    reference-format logs (no images: plane_aided false): gtsam/test_vro_imu_graph.cpp adds VRO + IMU factors and calls
    optimizeGraphIncremental() per frame; gtsam/test_ba_imu_graph.cpp does the same and ends with optimizeGraphBatch().
    Their trajectory logs are compared with this repository's own driver program running the same call sequence and with
    the batch optimum of the oracle on the graph the logs define."""
    exe = refbin(driver)
    spec = synth.make_config('C2', seed=1, scale=0.08)
    vro, imu, times, out = (str(tmp_path / n) for n in ('vro.log', 'imu.log', 'times.log', 'poses.txt'))
    recs = driver_logs.write_logs(spec, vro, imu, times)
    args = ['_sr_start_frame:=0', '_sr_end_frame:=%d' % (len(recs) + 10), '_sr_data_name:=synth', '_imu_file:=' + imu, '_imu_time_file:=' + times,
            '_vro_results_file:=' + vro, '_plane_aided:=false', '_use_imu:=true', '_gt_output_dir:=' + str(tmp_path)]
    res = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    name = 'synth_vio_trajectory.log' if driver == 'test_vro_imu_graph' else 'synth_ba_vio_after_trajectory.log'
    t, R = _read_traj(str(tmp_path / name))
    assert len(t) == spec['n_poses']
    # this repository's own test program on the same logs: identical API calls => the same ISAM2 estimate / batch optimum
    mine = subprocess.run([refbin('vio_driver'), vro, imu, times, out, 'incremental'], capture_output=True, text=True, timeout=600)
    assert mine.returncode == 0, mine.stderr[-2000:]
    ref = np.loadtxt(out + '.isam2') if driver == 'test_vro_imu_graph' else np.loadtxt(out)
    # (the reference's writeTrajectory prints with the default stream precision: 6 significant digits)
    assert np.abs(t - ref[:, 10:13]).max() <= 5e-5 and np.abs(R - ref[:, 1:10].reshape(-1, 3, 3)).max() <= 5e-5


@pytest.mark.gpu
def test_g2o_driver_matches_oracle(refbin, tmp_path):
    """BASELINE config 1 through the reference's own CGraphG2O (g2o/g2o_graph.cpp compiled unchanged over the g2o facade):
    100 poses, ~500 edges, first vertex fixed, optimizeGraph() = 10 x optimize(2) -- against oracle/lm.py."""
    from oracle import lm as olm
    exe = refbin('g2o_driver')
    spec = synth.make_config('C1', seed=1)
    Pm = np.zeros((6, 6)); Pm[:3, 3:] = np.eye(3); Pm[3:, :3] = np.eye(3)
    info = Pm @ spec['between_info'] @ Pm.T
    ei, ej = spec['between_i'], spec['between_j']
    order = sorted(range(len(ei)), key=lambda n: (ej[n], 0 if ej[n] - ei[n] == 1 else 1, ei[n]))     # the (j-1, j) edge reaches vertex j first
    with open(tmp_path / 'edges.txt', 'w') as f:
        for n in order:
            vals = list(spec['between_R'][n].ravel()) + list(spec['between_t'][n]) + list(info[n].ravel())
            f.write('%d %d %s\n' % (ei[n], ej[n], ' '.join(repr(float(v)) for v in vals)))
    res = subprocess.run([exe, str(tmp_path / 'edges.txt'), str(tmp_path / 'traj.log'), str(tmp_path / 'graph.g2o')], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith('RESULT')][0].split()
    assert int(line[2]) == spec['n_poses'] and int(line[4]) == len(ei)
    # the oracle starts from the poses CGraphG2O::addToGraph chained (X_j = X_i * T of the first edge that reaches j)
    R0, t0 = [np.eye(3)], [np.zeros(3)]
    first = {int(ej[n]): n for n in reversed(order)}
    for j in range(1, spec['n_poses']):
        n = first[j]
        R, t = lie.pose_compose(R0[ei[n]], t0[ei[n]], spec['between_R'][n], spec['between_t'][n]); R0.append(R); t0.append(t)
    pg = olm.PoseGraphG2O(np.array(R0), np.array(t0), ei, ej, spec['between_R'], spec['between_t'], info, fixed=(0,))
    chi0 = pg.chi2()
    assert abs(float(line[6]) - chi0) <= 1e-9 * chi0
    pg, orep = olm.optimize_g2o_calls(pg)
    assert abs(float(line[8]) - orep['chi2']) <= 1e-8 * orep['chi2']
    a = np.loadtxt(tmp_path / 'traj.log')
    assert np.abs(a[:, 1:4] - pg.t).max() <= 1e-6
    g2o = [l.split() for l in open(tmp_path / 'graph.g2o')]
    assert sum(l[0] == 'VERTEX_SE3:QUAT' for l in g2o) == spec['n_poses'] and sum(l[0] == 'EDGE_SE3:QUAT' for l in g2o) == len(ei)


def _plane_transform(pl, R, t):
    """OrientedPlane3::transform(pose) with the 3x3 Jacobian w.r.t. the plane (A.4): n' = R^T n, d' = n.t + d."""
    from oracle import factors as ofac
    n = pl[:3]
    q = R.T @ n
    out = np.concatenate([q, [n @ t + pl[3]]])
    Bn, Bq = ofac.unit3_basis(n), ofac.unit3_basis(q)
    Hp = np.zeros((3, 3))
    Hp[:2, :2] = Bq.T @ R.T @ Bn
    Hp[2, :2] = Bn.T @ t
    Hp[2, 2] = 1.0
    return out, Hp, Bn


@pytest.mark.gpu
def test_plane_branch_through_reference_add_plane_factor(refbin, tmp_path):
    """BASELINE config 3 (plane-aided VIO) through the reference's own CGraphGT::addPlaneFactor (gtsam_graph.cpp:1118-1298):
    the covariance J blkdiag(B^T S_n B, S_d) J^T, its conditioning (off-diagonal (0,1) dropped, diagonal truncated through a
    float), the world-frame landmark initialisation -- restated here in numpy for the oracle graph -- then batch LM on the
    device against the oracle."""
    from oracle import factors as ofac
    exe = refbin('plane_driver')
    spec = synth.make_config('C3', seed=2, scale=0.06)
    vro, imu, times = (str(tmp_path / n) for n in ('vro.log', 'imu.log', 'times.log'))
    recs = driver_logs.write_logs(spec, vro, imu, times)
    g0 = driver_logs.oracle_graph_from_logs(spec, recs)
    Ruc, tuc = spec['Rs'], spec['ts']                      # mp_u2c: camera pose in the IMU frame
    Rcu, tcu = lie.pose_inverse(Ruc, tuc)
    rng = np.random.default_rng(7)
    opose, opl = spec['plane_obs_pose'], spec['plane_obs_plane']
    order = np.lexsort((np.arange(len(opose)), opose))      # the driver adds the observations frame by frame, in file order within a frame
    # landmark ids in order of first appearance (the driver's file carries them)
    first_seen, lm_id = {}, []
    for k in order:
        first_seen.setdefault(int(opl[k]), len(first_seen))
    meas, infos, obs_i, obs_l, init = [], [], [], [], {}
    with open(tmp_path / 'planes.txt', 'w') as f:
        for k in order:
            zb = spec['plane_meas'][k]                       # plane in the IMU frame
            zc, _, _ = _plane_transform(zb, Ruc, tuc)        # ... as the camera sees it
            A = rng.normal(size=(4, 4)) * 0.003
            S = A @ A.T + np.diag([2e-5, 3e-5, 2.5e-5, 1e-4])
            lid = first_seen[int(opl[k])]
            f.write('%d %d %s %s\n' % (opose[k], lid, ' '.join(repr(float(v)) for v in zc), ' '.join(repr(float(v)) for v in S.ravel())))
            # --- what addPlaneFactor does with it
            onj, J, Bn = _plane_transform(zc, Rcu, tcu)      # ONJ = ONI.transform(Tcu, J)
            S_upi = np.zeros((3, 3)); S_upi[:2, :2] = Bn.T @ S[:3, :3] @ Bn; S_upi[2, 2] = S[3, 3]
            S_upj = J @ S_upi @ J.T
            dom = all(abs(S_upj[i, i]) >= sum(abs(S_upj[i, j]) for j in range(3) if j != i) for i in range(3))
            if not dom:
                S_upj = np.diag(np.diag(S_upj))
            S_upj[0, 1] = S_upj[1, 0] = 0.0
            for i in range(3):
                S_upj[i, i] = float(np.float32(int(S_upj[i, i] * 1e8))) * 1e-8 + 1e-8
            meas.append(onj); infos.append(np.linalg.inv(S_upj)); obs_i.append(int(opose[k])); obs_l.append(lid)
            if lid not in init:                              # new landmark: ONW = ONJ.transform(Twu^-1) at the pose's current value
                Rw, tw = lie.pose_inverse(g0.R[opose[k]], g0.t[opose[k]])
                init[lid] = _plane_transform(onj, Rw, tw)[0]
    g0.plane = np.array([init[l] for l in range(len(init))])
    g0.f['plane'] = dict(i=np.array(obs_i), l=np.array(obs_l), meas=np.array(meas), info=np.array(infos))
    res = subprocess.run([exe, vro, imu, times, str(tmp_path / 'planes.txt'), str(tmp_path / 'poses.txt'), str(tmp_path / 'planes_out.txt')],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    line = [l for l in res.stdout.splitlines() if l.startswith('RESULT')][0].split()
    assert int(line[2]) == spec['n_poses'] and int(line[4]) == len(init) and int(line[6]) == len(order) and int(line[8]) == 0
    ini = np.loadtxt(str(tmp_path / 'planes_out.txt') + '.init')
    assert np.abs(ini[:, 1:] - g0.plane).max() <= 1e-12
    e0, e1 = float(line[10]), float(line[12])
    assert abs(e0 - g0.error()) <= 1e-9 * g0.error(), (e0, g0.error())
    g1, rep = lm.optimize_gtsam(g0)
    assert abs(e1 - rep['error']) <= 1e-9 * rep['error']
    got = np.loadtxt(tmp_path / 'poses.txt')
    assert np.abs(got[:, 10:13] - g1.t).max() <= 1e-8 and np.abs(got[:, 1:10].reshape(-1, 3, 3) - g1.R).max() <= 1e-8
    pl = np.loadtxt(tmp_path / 'planes_out.txt')
    assert np.abs(pl[:, 1:] - g1.plane).max() <= 1e-7
