"""Text formats on the path through the reference's own CGraphGT (gtsam/gtsam_graph.cpp
compiled unchanged over compat/ + the gtsam facade, compat/build_ref.py) -- host only, no GPU: VRO edge log round trip with
the failed-match sentinel, trajectory log, trajectory PLY, g2o export."""
import os
import subprocess
import numpy as np
from oracle import lie

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vro_log_trajectory_ply_g2o(refbin, tmp_path):
    exe = refbin('format_io')
    res = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr[-2000:]
    out = dict(l.split(' ', 1) for l in res.stdout.splitlines() if l.split(' ')[0] in ('RECORDS', 'ADDED', 'X2', 'WRITE'))
    assert out['RECORDS'].strip() == '3'
    assert out['ADDED'].split()[0] == '2'                     # the record with information(0,0) == 10000 is a failed match
    # VRO log: id_to id_from, the 6-vector chart of the transform, 21 upper-triangular information entries
    log = np.loadtxt(tmp_path / 'vro.log')
    assert log.shape == (3, 29) and log[2, 8] == 10000 and list(log[:, 0]) == [1, 2, 3] and list(log[:, 1]) == [0, 1, 2]
    # dead-reckoned pose 2 = T(r1) * T(r2) (camera frame == IMU frame by default)
    R1, t1 = lie.se3_exp(log[0:1, 2:8]); R2, t2 = lie.se3_exp(log[1:2, 2:8])
    R, t = lie.pose_compose(R1, t1, R2, t2)
    assert np.allclose([float(v) for v in out['X2'].split()], t[0], atol=1e-12)
    traj = np.loadtxt(tmp_path / 'traj.log')
    assert traj.shape == (3, 9) and np.allclose(traj[2, 1:4], t[0], atol=1e-12) and np.allclose(np.linalg.norm(traj[:, 4:8], axis=1), 1.0)
    ply = open(tmp_path / 'traj.ply').read().splitlines()
    assert ply[:3] == ['ply', 'format ascii 1.0', 'element vertex 3'] and ply[9] == 'end_header' and ply[10].split()[3:] == ['0', '0', '255']
    g2o = [l.split() for l in open(tmp_path / 'graph.g2o')]
    vs = [l for l in g2o if l[0] == 'VERTEX_SE3:QUAT']; es = [l for l in g2o if l[0] == 'EDGE_SE3:QUAT']
    assert len(vs) == 3 and len(es) == 2 and len(es[0]) == 1 + 2 + 7 + 21
    # information re-ordered to [trans, rot]: first diagonal entry is the (3,3) entry of the [rot, trans] matrix
    assert float(es[0][10]) == 103.0 and float(es[0][10 + 6 + 5 + 4]) == 100.0
