"""Reference-format text logs from a synth GraphSpec, plus the oracle graph that the
reference's offline VIO driver flow (gtsam/test_vro_imu_graph.cpp:158-357) builds from those logs."""
import numpy as np
from oracle import lie, imu as oimu
from oracle.graph import Graph


def write_logs(spec, vro_path, imu_path, times_path, fmt=None):
    """fmt: number formatter (default repr = round-trip exact); the returned records hold the values AS WRITTEN."""
    rp = fmt or (lambda x: repr(float(x)))
    P = spec['n_poses']
    Ruc, tuc = spec['Rs'], spec['ts']
    Ric, tic = lie.pose_inverse(Ruc, tuc)
    Ad = lie.adjoint(Ruc, tuc)
    Adi = np.linalg.inv(Ad)
    ei, ej = spec['between_i'], spec['between_j']
    # per new frame j: the (j-1, j) edge first, then the other edges whose id_to is j
    order = sorted(range(len(ei)), key=lambda n: (ej[n], 0 if ej[n] - ei[n] == 1 else 1, ei[n]))
    recs = []
    with open(vro_path, 'w') as f:
        for n in order:
            R, t = lie.pose_compose(*lie.pose_compose(Ric, tic, spec['between_R'][n], spec['between_t'][n]), Ruc, tuc)
            r = lie.se3_log(R, t)
            info = Adi @ spec['between_info'][n] @ Adi.T
            info = 0.5 * (info + info.T)
            vals = [rp(x) for x in r] + [rp(info[i, j]) for i in range(6) for j in range(i, 6)]
            f.write('%d %d %s\n' % (ej[n], ei[n], ' '.join(vals)))
            back = np.array([float(v) for v in vals])
            iw = np.zeros((6, 6)); iw[np.triu_indices(6)] = back[6:]; iw = iw + np.triu(iw, 1).T
            recs.append((int(ei[n]), int(ej[n]), back[:6].copy(), iw))
    S = spec['imu_samples'].shape[1]
    flat = spec['imu_samples'].reshape(-1, 6)
    # the reference reads the samples into float variables (gtsam/imu_vn100.cpp:86-90): 9 significant digits carry a float exactly
    ri = (lambda x: '%.9g' % float(np.float32(x))) if fmt else (lambda x: repr(float(x)))
    with open(imu_path, 'w') as f:
        # the recordings' IMU logs run past the last image; two trailing samples let findIndexAt bracket it
        for k in range(len(flat) + 2):
            m = flat[min(k, len(flat) - 1)]
            f.write('%s %s %s %s %s %s %s 0 0 0\n' % (rp(k * spec['imu_dt']), ri(m[3]), ri(m[4]), ri(m[5]), ri(m[0]), ri(m[1]), ri(m[2])))
    with open(times_path, 'w') as f:
        for j in range(P):
            f.write('%d %s\n' % (j, rp(j * S * spec['imu_dt'])))
    return recs


def oracle_graph_from_logs(spec, recs):
    """What CGraphGT + CImuVn100 hold after the driver loop: values chained through the k=1 edges and the IMU
    prediction, factors as wired by firstNode / addToGTSAM / the CombinedImuFactor lines."""
    P = spec['n_poses']
    S = spec['imu_samples'].shape[1]
    Ruc, tuc = spec['Rs'], spec['ts']
    Ric, tic = lie.pose_inverse(Ruc, tuc)
    Ad = lie.adjoint(Ruc, tuc)
    g = Graph()
    R = np.zeros((P, 3, 3)); t = np.zeros((P, 3)); v = np.zeros((P, 3))
    R[0] = np.eye(3)
    bi, bj, bR, bt, binfo = [], [], [], [], []
    edges = {}
    for (i, j, r, info) in recs:
        Rc, tc = lie.se3_exp(r)
        Rm, tm = lie.pose_compose(*lie.pose_compose(Ruc, tuc, Rc, tc), Ric, tic)
        bi.append(i); bj.append(j); bR.append(Rm); bt.append(tm); binfo.append(Ad @ info @ Ad.T)
        if j - i == 1:
            edges[j] = (Rm, tm)
    # the reference reads IMU samples into float variables (gtsam/imu_vn100.cpp:86-90)
    samples = spec['imu_samples'].astype(np.float32).astype(np.float64)
    par = oimu.vn100_params()
    # CImuBase keeps the sample period in a float member (`float m_dt`, gtsam/imu_base.h:73)
    pim = oimu.preintegrate(samples, float(np.float32(spec['imu_dt'])), par, np.zeros((P - 1, 6)))
    for j in range(1, P):
        R[j], t[j] = lie.pose_compose(R[j - 1], t[j - 1], *edges[j])
        one = {k: (val[j - 1] if isinstance(val, np.ndarray) and val.ndim >= 1 and len(val) == P - 1 else val) for k, val in pim.items()}
        _, _, v[j] = oimu.predict(one, R[j - 1], t[j - 1], v[j - 1], np.zeros(6))
    g.R, g.t, g.vel, g.bias = R, t, v, np.zeros((P, 6))
    a = np.arange(P - 1)
    g.f = dict(
        prior_pose=dict(i=np.array([0]), R=np.eye(3)[None], t=np.zeros((1, 3)), info=np.eye(6)[None] / 1e-7 ** 2),
        prior_vel=dict(i=np.array([0]), mean=np.zeros((1, 3)), info=np.eye(3)[None] / 1e-3 ** 2),
        prior_bias=dict(i=np.array([0]), mean=np.zeros((1, 6)), info=np.eye(6)[None] / 1e-3 ** 2),
        between=dict(i=np.array(bi), j=np.array(bj), R=np.array(bR), t=np.array(bt), info=np.array(binfo)),
        imu=dict(pi=a, vi=a, pj=a + 1, vj=a + 1, bi=a, bj=a + 1, pim=pim, info=np.linalg.inv(pim['cov'])))
    return g


def read_vro_log(path):
    """VRO edge log: `id_to id_from r(6) upper-triangular information(21)` per line -> [(i, j, r, info)]."""
    recs = []
    for line in open(path):
        v = line.split()
        if len(v) < 29:
            continue
        iw = np.zeros((6, 6)); iw[np.triu_indices(6)] = [float(x) for x in v[8:29]]; iw = iw + np.triu(iw, 1).T
        recs.append((int(v[1]), int(v[0]), np.array([float(x) for x in v[2:8]]), iw))
    return recs


def read_imu_log(imu_path, times_path):
    """IMU log `t ax ay az gx gy gz ...` + image time log `id t` -> samples (P-1, S, 6) as [gx gy gz ax ay az], dt."""
    a = np.loadtxt(imu_path)
    t_img = np.loadtxt(times_path)[:, 1]
    dt = float(a[1, 0] - a[0, 0])
    idx = np.rint(t_img / dt).astype(int)
    S = int(idx[1] - idx[0])
    flat = np.concatenate([a[:, 4:7], a[:, 1:4]], 1)
    return np.stack([flat[idx[k]:idx[k] + S] for k in range(len(t_img) - 1)]), dt
