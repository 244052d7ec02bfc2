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


def test_plane_propagation_and_inlier_check_of_the_reference(refbin):
    """SURVEY 8 f4 (front-end numerics next to the path): CGraphGT::computeSdj and CGraphGT::inThisPlane
    (gtsam/gtsam_graph.cpp:725-764), compiled unchanged over the facade (tests/cpp/plane_check.cpp), against a numpy
    restatement: N_j = transform(T_ij, N_i) (oracle/factors.py: plane_transform, pinned by the reference's own KATs),
    S_dj = S_di + n_i^T S_t n_i + D^T S_ni D with D = (I - n_i n_i^T) t, and the test d^2 <= S_dj or d^2 <= 0.014^2."""
    import subprocess
    import numpy as np
    from oracle import factors, lie
    rng = np.random.default_rng(7)
    cases, lines = [], []
    for k in range(40):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        ni = np.concatenate([n * rng.uniform(0.5, 2.0), [rng.uniform(0.5, 4.0)]])        # not normalised: OrientedPlane3(Vector4) normalises
        A = rng.normal(size=(3, 3)) * 1e-2; Sni = A @ A.T
        B = rng.normal(size=(3, 3)) * 2e-2; St = B @ B.T
        Sdi = rng.uniform(1e-6, 1e-3)
        R = lie.so3_exp(rng.normal(size=3) * 0.3); t = rng.normal(size=3) * 0.2
        pl = factors.plane_from_coeffs(ni)
        nj = factors.plane_transform(pl, R, t, jac=False)
        # a point near the propagated plane, sometimes inside the 1.4 cm / sqrt(S_dj) band, sometimes not
        foot = -nj[3] * nj[:3] + np.cross(nj[:3], rng.normal(size=3))
        pt = foot + nj[:3] * rng.choice([0.0, 0.005, 0.02, 0.05, 0.2]) * rng.choice([-1, 1])
        cases.append((pl, Sni, Sdi, R, t, St, nj, pt))
        lines.append(' '.join('%.17g' % x for x in np.concatenate([ni, Sni.ravel(), [Sdi], R.ravel(), t, St.ravel(), pt])))
    out = subprocess.run([refbin('plane_check')], input='\n'.join(lines) + '\n', capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [l.split() for l in out.stdout.splitlines() if len(l.split()) == 6]
    assert len(rows) == len(cases)
    n_in = 0
    for (pl, Sni, Sdi, R, t, St, nj, pt), row in zip(cases, rows):
        n = pl[:3]
        D = (np.eye(3) - np.outer(n, n)) @ t
        want = Sdi + n @ St @ n + D @ Sni @ D
        got = [float(x) for x in row[:5]]
        assert abs(got[0] - want) <= 1e-14 + 1e-12 * want
        assert np.allclose(got[1:], nj, rtol=0, atol=1e-14)
        d2 = (nj[:3] @ pt + nj[3]) ** 2
        assert int(row[5]) == int(d2 <= want or d2 <= 0.014 ** 2)
        n_in += int(row[5])
    assert 0 < n_in < len(cases)            # both outcomes exercised
