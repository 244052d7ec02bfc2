"""Factor residuals and Jacobians, batched numpy fp64 (oracle; test infrastructure).

Every function restates the [ext] GTSAM 4.0 semantics listed in SURVEY.md Appendix A
for the factor types the reference instantiates:
  between_pose     BetweenFactor<Pose3>         gtsam/gtsam_graph.cpp:691-692      (A.3)
  prior_pose/vec   PriorFactor<...>             gtsam/gtsam_graph.cpp:341,362-367  (A.3)
  projection       GenericProjectionFactor<Pose3,Point3,Cal3DS2>  :405-434        (A.2)
  plane            OrientedPlane3Factor         gtsam/gtsam_graph.cpp:1265         (A.4)
  imu_combined     CombinedImuFactor            gtsam/test_vro_imu_graph.cpp:191-196 (A.5/A.6)
Residuals are UNWHITENED; the caller applies the information matrix.
Jacobians are w.r.t. the right/body tangent of each key (A.1).
"""
import numpy as np
from . import lie


# --------------------------------------------------------------------------- Between / Prior
def between_pose(R1, t1, R2, t2, Rm, tm, jac=True, chart=0):
    """r = Local(Z, X1^-1 * X2) = ChartAtOrigin::Local(Z^-1 h);  H1 = -Ad(h^-1), H2 = I  (A.3, fast path: the derivative of
    Local is not chained in, whatever the chart)."""
    Rh, th = lie.pose_between(R1, t1, R2, t2)
    Re, te = lie.pose_between(Rm, tm, Rh, th)
    r = lie.chart_local0(Re, te, chart)
    if not jac:
        return r
    Rhi, thi = lie.pose_inverse(Rh, th)
    H1 = -lie.adjoint(Rhi, thi)
    H2 = np.broadcast_to(np.eye(6), H1.shape).copy()
    return r, H1, H2


def prior_pose(R, t, Rp, tp, jac=True, chart=0):
    """r = Local(prior, x) = ChartAtOrigin::Local(prior^-1 x), H = I."""
    r = lie.pose_local(Rp, tp, R, t, chart)
    if not jac:
        return r
    return r, np.broadcast_to(np.eye(6), r.shape[:-1] + (6, 6)).copy()


def prior_vec(x, p, jac=True):
    r = np.asarray(x, dtype=np.float64) - np.asarray(p, dtype=np.float64)
    if not jac:
        return r
    d = r.shape[-1]
    return r, np.broadcast_to(np.eye(d), r.shape[:-1] + (d, d)).copy()


# --------------------------------------------------------------------------- Projection (Cal3DS2)
def projection(R, t, p, uv, K, Rs, ts, jac=True):
    """GenericProjectionFactor<Pose3,Point3,Cal3DS2> with body_P_sensor (A.2).

    K = (fx, fy, s, u0, v0, k1, k2, p1, p2).  throwCheirality=false: q.z<=0 ->
    residual (2fx, 2fx), zero Jacobians.
    Returns r (...,2), Jpose (...,2,6), Jpoint (...,2,3).
    """
    fx, fy, s, u0, v0, k1, k2, p1, p2 = [float(v) for v in K]
    Rc, tc = lie.pose_compose(R, t, Rs, ts)
    q = np.einsum('...ji,...j->...i', Rc, p - tc)
    z = q[..., 2]
    bad = z <= 0
    zs = np.where(bad, 1.0, z)
    d = 1.0 / zs
    x = q[..., 0] * d
    y = q[..., 1] * d
    r2 = x * x + y * y
    g = 1.0 + k1 * r2 + k2 * r2 * r2
    dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    xd = g * x + dx
    yd = g * y + dy
    u = fx * xd + s * yd + u0
    v = fy * yd + v0
    r = np.stack([u, v], -1) - uv
    r = np.where(bad[..., None], 2.0 * fx, r)
    if not jac:
        return r
    # d(xd,yd)/d(x,y)
    gx = 2 * x * (k1 + 2 * k2 * r2)
    gy = 2 * y * (k1 + 2 * k2 * r2)
    D = np.zeros(r.shape[:-1] + (2, 2))
    D[..., 0, 0] = g + x * gx + 2 * p1 * y + 6 * p2 * x
    D[..., 0, 1] = x * gy + 2 * p1 * x + 2 * p2 * y
    D[..., 1, 0] = y * gx + 2 * p1 * x + 2 * p2 * y
    D[..., 1, 1] = g + y * gy + 6 * p1 * y + 2 * p2 * x
    DK = np.array([[fx, s], [0.0, fy]])
    Dpi = DK @ D
    # d(x,y)/d camera pose (right tangent of the camera pose)
    Dpose = np.zeros(r.shape[:-1] + (2, 6))
    Dpose[..., 0, 0] = x * y
    Dpose[..., 0, 1] = -(1 + x * x)
    Dpose[..., 0, 2] = y
    Dpose[..., 0, 3] = -d
    Dpose[..., 0, 5] = d * x
    Dpose[..., 1, 0] = 1 + y * y
    Dpose[..., 1, 1] = -x * y
    Dpose[..., 1, 2] = -x
    Dpose[..., 1, 4] = -d
    Dpose[..., 1, 5] = d * y
    Dpt = np.zeros(r.shape[:-1] + (2, 3))
    Dpt[..., 0, 0] = d
    Dpt[..., 0, 2] = -d * x
    Dpt[..., 1, 1] = d
    Dpt[..., 1, 2] = -d * y
    Dpt = Dpt @ np.swapaxes(Rc, -1, -2)
    # body pose -> camera pose: compose Jacobian Ad(sensor^-1)
    Rsi, tsi = lie.pose_inverse(Rs, ts)
    H0 = lie.adjoint(Rsi, tsi)
    Jpose = Dpi @ Dpose @ H0
    Jpoint = Dpi @ Dpt
    Jpose = np.where(bad[..., None, None], 0.0, Jpose)
    Jpoint = np.where(bad[..., None, None], 0.0, Jpoint)
    return r, Jpose, Jpoint


# --------------------------------------------------------------------------- OrientedPlane3
def unit3_basis(n):
    """Unit3::basis(): axis of smallest |component| (ties x, then y); b1=normalize(n x axis); b2 = n x b1."""
    n = np.asarray(n, dtype=np.float64)
    a = np.abs(n)
    mx, my, mz = a[..., 0], a[..., 1], a[..., 2]
    selx = (mx <= my) & (mx <= mz)
    sely = (~selx) & (my <= mx) & (my <= mz)
    axis = np.zeros_like(n)
    axis[..., 0] = selx
    axis[..., 1] = sely
    axis[..., 2] = ~(selx | sely)
    b1 = np.cross(n, axis)
    b1 = b1 / np.linalg.norm(b1, axis=-1, keepdims=True)
    b2 = np.cross(n, b1)
    return np.stack([b1, b2], -1)  # (...,3,2)


def plane_from_coeffs(c):
    """OrientedPlane3(a,b,c,d): normal normalised, d kept as given."""
    c = np.asarray(c, dtype=np.float64)
    n = c[..., :3] / np.linalg.norm(c[..., :3], axis=-1, keepdims=True)
    return np.concatenate([n, c[..., 3:4]], -1)


def unit3_local(n, q):
    """Unit3::localCoordinates(q) at n."""
    x = np.sum(n * q, -1)
    z = 1.0 - x * x
    tiny = z < np.finfo(np.float64).eps
    y = np.where(tiny, 1.0 - (x - 1.0) / 3.0, np.arccos(np.clip(x, -1, 1)) / np.sqrt(np.where(tiny, 1.0, z)))
    B = unit3_basis(n)
    out = np.einsum('...ji,...j->...i', B, y[..., None] * (q - x[..., None] * n))
    cop = tiny & (x <= 0)
    out = np.where(cop[..., None], np.array([np.pi, 0.0]), out)
    return out


def unit3_retract(n, v):
    B = unit3_basis(n)
    xi = np.einsum('...ij,...j->...i', B, v)
    th = np.linalg.norm(xi, axis=-1)
    tiny = th < np.finfo(np.float64).eps
    sc = np.where(tiny, 1.0, np.sin(th) / np.where(tiny, 1.0, th))
    p = np.cos(th)[..., None] * n + sc[..., None] * xi
    return p / np.linalg.norm(p, axis=-1, keepdims=True)


def plane_retract(pl, v):
    n = unit3_retract(pl[..., :3], v[..., :2])
    return np.concatenate([n, pl[..., 3:4] + v[..., 2:3]], -1)


def plane_local(pl, other):
    return np.concatenate([unit3_local(pl[..., :3], other[..., :3]), other[..., 3:4] - pl[..., 3:4]], -1)


def plane_transform(pl, R, t, jac=True):
    """OrientedPlane3::transform(pose): n' = R^T n, d' = n.t + d (A.4; pinned by the KAT).
    Returns plane', Hpose (...,3,6), Hplane (...,3,3)."""
    n, d = pl[..., :3], pl[..., 3]
    q = np.einsum('...ji,...j->...i', R, n)
    dp = np.sum(n * t, -1) + d
    out = np.concatenate([q, dp[..., None]], -1)
    if not jac:
        return out
    Bq = unit3_basis(q)
    Bn = unit3_basis(n)
    BqT = np.swapaxes(Bq, -1, -2)
    Hr = np.zeros(q.shape[:-1] + (3, 6))
    Hr[..., :2, :3] = BqT @ lie.skew(q)
    Hr[..., 2, 3:] = q
    Hp = np.zeros(q.shape[:-1] + (3, 3))
    Hp[..., :2, :2] = BqT @ np.swapaxes(R, -1, -2) @ Bn
    Hp[..., 2, :2] = np.einsum('...ji,...j->...i', Bn, t)
    Hp[..., 2, 2] = 1.0
    return out, Hr, Hp


def plane_error(pl, other):
    """OrientedPlane3::error(other) = (-n.localCoordinates(other.n), d - other.d)."""
    e2 = -unit3_local(pl[..., :3], other[..., :3])
    return np.concatenate([e2, (pl[..., 3] - other[..., 3])[..., None]], -1)


def plane_error_vector(pl, other):
    """OrientedPlane3::errorVector(other) = (B^T other.n, d - other.d) (regression KAT)."""
    B = unit3_basis(pl[..., :3])
    e2 = np.einsum('...ji,...j->...i', B, other[..., :3])
    return np.concatenate([e2, (pl[..., 3] - other[..., 3])[..., None]], -1)


def plane_factor(R, t, pl, meas, jac=True):
    """OrientedPlane3Factor::evaluateError = transform(plane, pose).error(measured);
    Jacobians are those of transform (A.4)."""
    if not jac:
        return plane_error(plane_transform(pl, R, t, jac=False), meas)
    pred, Hr, Hp = plane_transform(pl, R, t)
    return plane_error(pred, meas), Hr, Hp


# --------------------------------------------------------------------------- CombinedImuFactor
def imu_combined(Ri, ti, vi, Rj, tj, vj, bi, bj, pim, jac=True):
    """CombinedImuFactor::evaluateError (A.5/A.6), TangentPreintegration.

    pim: dict with dt (...), preint (...,9), Hba (...,9,3), Hbg (...,9,3), bias_hat (...,6), gravity (3,).
    Bias layout [acc, gyro].  Returns r (...,15) and Jacobians for
    (pose_i 15x6, vel_i 15x3, pose_j 15x6, vel_j 15x3, bias_i 15x6, bias_j 15x6).
    """
    dt = np.asarray(pim['dt'], dtype=np.float64)
    g = np.asarray(pim['gravity'], dtype=np.float64)
    inc = bi - pim['bias_hat']
    bc = pim['preint'] + np.einsum('...ij,...j->...i', pim['Hba'], inc[..., :3]) \
        + np.einsum('...ij,...j->...i', pim['Hbg'], inc[..., 3:])
    RiT = np.swapaxes(Ri, -1, -2)
    Rtv = np.einsum('...ij,...j->...i', RiT, vi)
    Rtg = np.einsum('...ij,j->...i', RiT, g)
    dt22 = 0.5 * dt * dt
    xi_th = bc[..., 0:3]
    xi_p = bc[..., 3:6] + dt[..., None] * Rtv + dt22[..., None] * Rtg
    xi_v = bc[..., 6:9] + dt[..., None] * Rtg
    bRc = lie.so3_exp(xi_th)
    Rp = Ri @ bRc
    tp = ti + np.einsum('...ij,...j->...i', Ri, xi_p)
    vp = vi + np.einsum('...ij,...j->...i', Ri, xi_v)
    RjT = np.swapaxes(Rj, -1, -2)
    dR = RjT @ Rp
    dtr = np.einsum('...ij,...j->...i', RjT, tp - tj)
    dvr = np.einsum('...ij,...j->...i', RjT, vp - vj)
    eth = lie.so3_log(dR)
    r = np.concatenate([eth, dtr, dvr, bi - bj], -1)
    if not jac:
        return r
    shp = r.shape[:-1]
    I3 = np.eye(3)
    bRcT = np.swapaxes(bRc, -1, -2)
    # correctPIM Jacobian wrt state_i (NavState tangent [theta, p, v])
    Dds = np.zeros(shp + (9, 9))
    Dds[..., 3:6, 0:3] = dt[..., None, None] * lie.skew(Rtv) + dt22[..., None, None] * lie.skew(Rtg)
    Dds[..., 3:6, 6:9] = dt[..., None, None] * I3
    Dds[..., 6:9, 0:3] = dt[..., None, None] * lie.skew(Rtg)
    # retract Jacobians
    Dps = np.zeros(shp + (9, 9))
    Dps[..., 0:3, 0:3] = bRcT
    Dps[..., 3:6, 0:3] = -bRcT @ lie.skew(xi_p)
    Dps[..., 3:6, 3:6] = bRcT
    Dps[..., 6:9, 0:3] = -bRcT @ lie.skew(xi_v)
    Dps[..., 6:9, 6:9] = bRcT
    Dpd = np.zeros(shp + (9, 9))
    Dpd[..., 0:3, 0:3] = lie.so3_jr(xi_th)
    Dpd[..., 3:6, 3:6] = bRcT
    Dpd[..., 6:9, 6:9] = bRcT
    H1p = Dps + Dpd @ Dds
    Dbc = np.concatenate([pim['Hba'], pim['Hbg']], -1)  # (...,9,6)
    H2p = Dpd @ Dbc
    # localCoordinates Jacobians
    Jri = lie.so3_jr_inv(eth)
    Dej = np.zeros(shp + (9, 9))
    Dej[..., 0:3, 0:3] = -Jri @ np.swapaxes(dR, -1, -2)
    Dej[..., 3:6, 0:3] = lie.skew(dtr)
    Dej[..., 3:6, 3:6] = -I3
    Dej[..., 6:9, 0:3] = lie.skew(dvr)
    Dej[..., 6:9, 6:9] = -I3
    Dep = np.zeros(shp + (9, 9))
    Dep[..., 0:3, 0:3] = Jri
    Dep[..., 3:6, 3:6] = dR
    Dep[..., 6:9, 6:9] = dR
    DH = Dep @ H1p
    J_pi = np.zeros(shp + (15, 6)); J_pi[..., :9, :] = DH[..., :, 0:6]
    J_vi = np.zeros(shp + (15, 3)); J_vi[..., :9, :] = DH[..., :, 6:9] @ RiT
    J_pj = np.zeros(shp + (15, 6)); J_pj[..., :9, :] = Dej[..., :, 0:6]
    J_vj = np.zeros(shp + (15, 3)); J_vj[..., :9, :] = Dej[..., :, 6:9] @ RjT
    J_bi = np.zeros(shp + (15, 6)); J_bi[..., :9, :] = Dep @ H2p; J_bi[..., 9:, :] = np.eye(6)
    J_bj = np.zeros(shp + (15, 6)); J_bj[..., 9:, :] = -np.eye(6)
    return r, (J_pi, J_vi, J_pj, J_vj, J_bi, J_bj)


# --------------------------------------------------------------------------- g2o EdgeSE3 (A.8)
def g2o_edge_se3(R1, t1, R2, t2, Rm, tm):
    """EdgeSE3 error = toVectorMQT(Z^-1 X1^-1 X2) = [t, q_xyz] (order [trans, rot])."""
    Rh, th = lie.pose_between(R1, t1, R2, t2)
    Re, te = lie.pose_between(Rm, tm, Rh, th)
    q = lie.quat_from_rot(Re)
    return np.concatenate([te, q[..., 1:]], -1)


def g2o_oplus(R, t, d):
    """VertexSE3::oplus: X <- X * fromVectorMQT(d), d = [t, q_xyz]."""
    d = np.asarray(d, dtype=np.float64)
    qv = d[..., 3:]
    w2 = 1.0 - np.sum(qv * qv, -1)
    w = np.sqrt(np.maximum(w2, 0.0))
    q = np.concatenate([w[..., None], qv], -1)
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    return lie.pose_compose(R, t, lie.rot_from_quat(q), d[..., :3])
