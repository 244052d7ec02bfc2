"""Seeded synthetic graphs for the five BASELINE.json configs.

Pure numpy; produces a GraphSpec (plain dict of arrays) that is consumed both by the
C-ABI loader (graph_slam_b200.abi.load_spec) and by the test-side oracle builder.
No solver arithmetic happens here: IMU samples are emitted raw and preintegrated by
whoever consumes the spec.

Shapes (C = BASELINE.json configs[i]):
  C1  100 poses, ~500 VRO between edges (pose graph)
  C2  1k poses (X,V,B), ~5k between edges, 1k CombinedImu factors, 3 priors
  C3  C2 + 10 plane landmarks / 50 plane factors
  C4  2k poses (X,V,B), 100k points, 2M projections, 100k point priors, 2k IMU
  C5  5k poses (X,V,B), 500k points, 10M projections, 500k point priors, 5k IMU
"""
import numpy as np

CAL_SR4K = (250.5773, 250.5773, 0.0, 90.0, 70.0, -0.8466, 0.5370, 0.0, 0.0)   # gtsam_graph.cpp:544
GRAVITY = np.array([0.0, 0.0, 9.71])                                           # imu_base.cpp:261
IMU_DT = 0.005                                                                 # test_vro_imu_graph.cpp:111
SAMPLES_PER_FRAME = 20                                                         # 200 Hz IMU / 10 Hz frames

# imu_vn100.cpp:34-43 noise densities (as used there, per sqrt(Hz))
SIG_A = 0.14e-3 * 9.81
SIG_G = np.deg2rad(0.0035)
SIG_BA = 0.04e-3 * 9.81 * np.sqrt(200.0)
SIG_BG = np.deg2rad(10.0) / 3600.0 * np.sqrt(200.0)


def _skew(w):
    z = np.zeros(w.shape[:-1])
    return np.stack([np.stack([z, -w[..., 2], w[..., 1]], -1),
                     np.stack([w[..., 2], z, -w[..., 0]], -1),
                     np.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def _exp(w):
    th2 = np.sum(w * w, -1)
    small = th2 < 1e-16
    th = np.sqrt(np.where(small, 1.0, th2))
    a = np.where(small, 1.0, np.sin(th) / th)
    b = np.where(small, 0.5, (1 - np.cos(th)) / np.where(small, 1.0, th2))
    W = _skew(w)
    return np.eye(3) + a[..., None, None] * W + b[..., None, None] * (W @ W)


def _log(R):
    c = np.clip((np.trace(R, axis1=-2, axis2=-1) - 1) * 0.5, -1, 1)
    th = np.arccos(c)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], -1)
    s = np.sin(th)
    k = np.where(th < 1e-7, 0.5, th / (2 * np.where(np.abs(s) < 1e-300, 1.0, s)))
    return k[..., None] * v


def _se3_exp(xi):
    w, v = xi[..., :3], xi[..., 3:]
    th2 = np.sum(w * w, -1)
    small = th2 < 1e-16
    th = np.sqrt(np.where(small, 1.0, th2))
    b = np.where(small, 0.5, (1 - np.cos(th)) / np.where(small, 1.0, th2))
    c = np.where(small, 1.0 / 6, (th - np.sin(th)) / np.where(small, 1.0, th2 * th))
    W = _skew(w)
    V = np.eye(3) + b[..., None, None] * W + c[..., None, None] * (W @ W)
    return _exp(w), np.einsum('...ij,...j->...i', V, v)


def rzryrx(x, y, z):
    cx, sx, cy, sy, cz, sz = np.cos(x), np.sin(x), np.cos(y), np.sin(y), np.cos(z), np.sin(z)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def camera_to_imu(p=0.0):
    """CGraphGT::setCamera2IMU(p) rotation (gtsam_graph.cpp:218-254), zero translation."""
    return rzryrx(np.pi / 2, 0.0, np.pi / 2) @ rzryrx(p, 0.0, 0.0)


def _trajectory(tt, period):
    """Closed 3-D Lissajous; returns position, velocity, acceleration, body rotation (x along velocity
    with +-10 deg roll/pitch wobble).  Z is DOWN (MakeSharedD nav frame)."""
    w = 2 * np.pi / period
    A = np.array([6.0, 4.0, 0.6]) * (period / 100.0)
    k = np.array([1.0, 2.0, 3.0])
    ph = np.array([0.0, 0.5, 1.0])
    arg = w * k[None, :] * tt[:, None] + ph[None, :]
    p = A * np.sin(arg)
    v = A * w * k * np.cos(arg)
    a = -A * (w * k) ** 2 * np.sin(arg)
    yaw = np.unwrap(np.arctan2(v[:, 1], v[:, 0]))
    pitch = -np.arctan2(v[:, 2], np.hypot(v[:, 0], v[:, 1])) + np.deg2rad(10) * np.sin(0.7 * tt)
    roll = np.deg2rad(10) * np.sin(0.9 * tt + 0.3)
    cx, sx, cy, sy, cz, sz = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    R = np.zeros((len(tt), 3, 3))
    R[:, 0, 0] = cz * cy; R[:, 0, 1] = cz * sy * sx - sz * cx; R[:, 0, 2] = cz * sy * cx + sz * sx
    R[:, 1, 0] = sz * cy; R[:, 1, 1] = sz * sy * sx + cz * cx; R[:, 1, 2] = sz * sy * cx - cz * sx
    R[:, 2, 0] = -sy;     R[:, 2, 1] = cy * sx;                R[:, 2, 2] = cy * cx
    return p, v, a, R


def _project(R, t, p, K, Rs, ts):
    fx, fy, s, u0, v0, k1, k2, p1, p2 = K
    Rc = R @ Rs
    tc = np.einsum('...ij,j->...i', R, ts) + t
    q = np.einsum('...ji,...j->...i', Rc, p - tc)
    x = q[..., 0] / q[..., 2]; y = q[..., 1] / q[..., 2]
    r2 = x * x + y * y
    g = 1 + k1 * r2 + k2 * r2 * r2
    xd = g * x + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = g * y + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([fx * xd + s * yd + u0, fy * yd + v0], -1), q[..., 2]


def make_graph(n_poses, seed=1, vro=True, imu=True, n_landmarks=0, obs_per_landmark=20, window=50,
               n_planes=0, n_plane_obs=0, loop_closure_frac=0.02, lookback=5, name='custom'):
    rng = np.random.Generator(np.random.PCG64(seed))
    P = n_poses
    S = SAMPLES_PER_FRAME
    frame_dt = S * IMU_DT
    period = max(P * frame_dt, 10.0)
    # IMU-rate truth
    n_imu = (P - 1) * S + 1
    ti = np.arange(n_imu) * IMU_DT
    p_i, v_i, a_i, R_i = _trajectory(ti, period)
    h = 1e-4
    _, _, _, Rp = _trajectory(ti + h, period)
    _, _, _, Rm = _trajectory(ti - h, period)
    omega = _log(np.swapaxes(Rm, -1, -2) @ Rp) / (2 * h)                # body rate
    f_body = np.einsum('nji,nj->ni', R_i, a_i - GRAVITY[None, :])         # specific force R^T (a - g)
    fidx = np.arange(P) * S
    Rt, tt_, vt = R_i[fidx], p_i[fidx], v_i[fidx]

    spec = dict(name=name, seed=seed, n_poses=P, K=np.array(CAL_SR4K), Rs=camera_to_imu(0.0), ts=np.zeros(3),
                truth_R=Rt, truth_t=tt_, truth_v=vt, has_vel=bool(imu))

    # ---- IMU samples with white noise + bias random walk (discrete: sigma/sqrt(dt), sigma_b*sqrt(dt))
    if imu:
        nb = n_imu - 1
        ba = np.cumsum(rng.normal(size=(nb, 3)) * SIG_BA * np.sqrt(IMU_DT), 0) * 0.05
        bg = np.cumsum(rng.normal(size=(nb, 3)) * SIG_BG * np.sqrt(IMU_DT), 0) * 0.05
        gyro = omega[:-1] + bg + rng.normal(size=(nb, 3)) * SIG_G / np.sqrt(IMU_DT)
        acc = f_body[:-1] + ba + rng.normal(size=(nb, 3)) * SIG_A / np.sqrt(IMU_DT)
        spec['imu_samples'] = np.concatenate([gyro, acc], -1).reshape(P - 1, S, 6)   # [gx gy gz ax ay az]
        spec['imu_dt'] = IMU_DT
        spec['truth_bias'] = np.concatenate([np.zeros((1, 6)), np.concatenate([ba, bg], -1)[fidx[1:] - 1]], 0)
        spec['vel_init'] = vt + rng.normal(size=(P, 3)) * 0.02
        # firstNode puts a sigma=1e-3 prior on V0 (gtsam_graph.cpp:351-362; mean 0 there because the
        # recordings start at rest).  The Lissajous does not start at rest, so the prior mean is the true v0.
        spec['vel_init'][0] = vt[0]
        spec['prior_vel_mean'] = vt[0].copy()
        spec['bias_init'] = np.zeros((P, 6))

    # ---- VRO between edges
    if vro:
        ei, ej = [], []
        for k in range(1, lookback + 1):
            ej.append(np.arange(k, P)); ei.append(np.arange(k, P) - k)
        ei = np.concatenate(ei); ej = np.concatenate(ej)
        order = np.lexsort((ei, ej))          # per new frame j: edges (j-1..j-5, j)
        ei, ej = ei[order], ej[order]
        n_lc = int(round(loop_closure_frac * P)) if loop_closure_frac > 0 else 0
        if n_lc:
            a = rng.integers(0, P, size=n_lc); b = rng.integers(0, P, size=n_lc)
            keep = np.abs(a - b) > lookback
            a, b = a[keep], b[keep]
            ei = np.concatenate([ei, np.minimum(a, b)]); ej = np.concatenate([ej, np.maximum(a, b)])
        ne = len(ei)
        Rrel = np.swapaxes(Rt[ei], -1, -2) @ Rt[ej]
        trel = np.einsum('nji,nj->ni', Rt[ei], tt_[ej] - tt_[ei])
        sig = np.array([np.deg2rad(0.5)] * 3 + [0.01] * 3)
        noise = rng.normal(size=(ne, 6)) * sig
        dR, dt_ = _se3_exp(noise)
        Rmeas = Rrel @ dR
        tmeas = np.einsum('nij,nj->ni', Rrel, dt_) + trel
        info0 = np.diag(1.0 / sig ** 2)
        # mild conjugation (keeps blocks full but well-conditioned)
        mix = _exp6_small(rng, ne)
        info = np.swapaxes(mix, -1, -2) @ info0[None] @ mix
        info = 0.5 * (info + np.swapaxes(info, -1, -2))
        spec.update(between_i=ei.astype(np.int32), between_j=ej.astype(np.int32), between_R=Rmeas, between_t=tmeas,
                    between_info=info)
        # initial poses: dead-reckoning through k=1 edges (gtsam_graph.cpp:657-660)
        R0 = np.zeros((P, 3, 3)); t0 = np.zeros((P, 3))
        R0[0] = Rt[0]; t0[0] = tt_[0]
        first = {int(j): n for n, (i, j) in enumerate(zip(ei, ej)) if j - i == 1}
        for j in range(1, P):
            n = first[j]
            R0[j] = R0[j - 1] @ Rmeas[n]
            t0[j] = R0[j - 1] @ tmeas[n] + t0[j - 1]
        spec['pose_init_R'], spec['pose_init_t'] = R0, t0
    else:
        nz = rng.normal(size=(P, 6)) * np.array([np.deg2rad(0.3)] * 3 + [0.02] * 3)
        nz[0] = 0
        dR, dt_ = _se3_exp(nz)
        spec['pose_init_R'] = Rt @ dR
        spec['pose_init_t'] = np.einsum('nij,nj->ni', Rt, dt_) + tt_
    # gauge prior at pose 0 on its initial value (firstNode, gtsam_graph.cpp:338-341)
    spec['prior_pose_R'] = spec['pose_init_R'][0].copy()
    spec['prior_pose_t'] = spec['pose_init_t'][0].copy()

    # ---- point landmarks + projection factors
    L = n_landmarks
    if L:
        k = obs_per_landmark
        Rs, ts, K = spec['Rs'], spec['ts'], CAL_SR4K
        half = window // 2
        centre = np.sort(rng.integers(half, max(P - half, half + 1), size=L))     # sorted along the trajectory
        depth = rng.uniform(2.0, 6.0, size=L)
        lat = rng.uniform(-0.3, 0.3, size=(L, 2)) * depth[:, None]
        pc = np.stack([lat[:, 0], lat[:, 1], depth], -1)                           # camera frame of centre pose
        Rc = Rt[centre] @ Rs
        pw = np.einsum('nij,nj->ni', Rc, pc) + tt_[centre] + np.einsum('nij,j->ni', Rt[centre], ts)
        # candidate window [c-half, c-half+window), choose k with depth > 0.1
        cand = centre[:, None] - half + np.arange(window)[None, :]
        cand = np.clip(cand, 0, P - 1)
        uv0, z = _project(Rt[cand], tt_[cand], pw[:, None, :], K, Rs, ts)
        score = rng.random(size=cand.shape)
        # visible = in front (depth > 0.5 m) and inside a generous image window around the 176x144 sensor
        vis = (z > 0.5) & (np.abs(uv0[..., 0] - K[3]) < 160.0) & (np.abs(uv0[..., 1] - K[4]) < 140.0)
        score[~vis] = 2.0
        # drop duplicates created by clipping at the ends
        dup = np.zeros_like(score, dtype=bool)
        dup[:, 1:] = cand[:, 1:] == cand[:, :-1]
        score[dup] = 3.0
        sel = np.sort(np.argsort(score, axis=1)[:, :k], axis=1)
        obs_pose = np.take_along_axis(cand, sel, 1)
        valid = np.take_along_axis(score, sel, 1) < 1.5
        uv, _ = _project(Rt[obs_pose], tt_[obs_pose], pw[:, None, :], K, Rs, ts)
        uv = uv + rng.normal(size=uv.shape)
        lid = np.repeat(np.arange(L), k).reshape(L, k)
        m = valid.ravel()
        spec.update(proj_pose=obs_pose.ravel()[m].astype(np.int32), proj_point=lid.ravel()[m].astype(np.int32),
                    proj_uv=uv.reshape(-1, 2)[m], proj_sigma=1.0,
                    point_init=pw + rng.normal(size=(L, 3)) * 0.014, point_prior_sigma=0.014, truth_point=pw)

    # ---- planes
    if n_planes:
        nrm = rng.normal(size=(n_planes, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        dd = rng.uniform(2.0, 8.0, size=n_planes)
        planes = np.concatenate([nrm, dd[:, None]], -1)
        opose = np.sort(rng.integers(0, P, size=n_plane_obs))
        opl = rng.integers(0, n_planes, size=n_plane_obs)
        opl[:n_planes] = np.arange(n_planes)                       # every plane observed at least once
        n_b = np.einsum('nji,nj->ni', Rt[opose], planes[opl, :3])
        d_b = np.einsum('ni,ni->n', planes[opl, :3], tt_[opose]) + planes[opl, 3]
        n_b = n_b + rng.normal(size=n_b.shape) * 0.01
        n_b /= np.linalg.norm(n_b, axis=1, keepdims=True)
        d_b = d_b + rng.normal(size=d_b.shape) * 0.01
        meas = np.concatenate([n_b, d_b[:, None]], -1)
        # landmark initial value from its first observation and the initial pose (gtsam_graph.cpp:1195-1209)
        init = np.zeros((n_planes, 4))
        R0, t0 = spec['pose_init_R'], spec['pose_init_t']
        for l in range(n_planes):
            o = int(np.nonzero(opl == l)[0][0])
            i = opose[o]
            # plane in world: inverse of (n' = R^T n, d' = n.t + d)
            nw = R0[i] @ meas[o, :3]
            init[l] = np.concatenate([nw, [meas[o, 3] - nw @ t0[i]]])
        spec.update(plane_init=init, plane_obs_pose=opose.astype(np.int32), plane_obs_plane=opl.astype(np.int32),
                    plane_meas=meas, plane_cov=np.broadcast_to(np.eye(3) * 1e-4, (n_plane_obs, 3, 3)).copy(),
                    truth_plane=planes)
    return spec


def _exp6_small(rng, n):
    """Random near-orthogonal 6x6 mixers (orthogonal via QR of I + small noise) to make information blocks full."""
    A = np.eye(6)[None] + 0.2 * rng.normal(size=(n, 6, 6))
    Q, _ = np.linalg.qr(A)
    return Q


CONFIGS = {
    'C1': dict(n_poses=100, vro=True, imu=False, loop_closure_frac=0.02),
    'C2': dict(n_poses=1000, vro=True, imu=True),
    'C3': dict(n_poses=1000, vro=True, imu=True, n_planes=10, n_plane_obs=50),
    'C4': dict(n_poses=2000, vro=False, imu=True, n_landmarks=100000),
    'C5': dict(n_poses=5000, vro=False, imu=True, n_landmarks=500000),
}


def make_config(name, seed=1, scale=1.0):
    """BASELINE.json configs[i]; scale<1 shrinks poses and landmarks proportionally (parity-test sizes)."""
    kw = dict(CONFIGS[name])
    if scale != 1.0:
        kw['n_poses'] = max(int(round(kw['n_poses'] * scale)), 60)
        if 'n_landmarks' in kw:
            kw['n_landmarks'] = max(int(round(kw['n_landmarks'] * scale)), 10)
    return make_graph(seed=seed, name=name if scale == 1.0 else '%s@%g' % (name, scale), **kw)
