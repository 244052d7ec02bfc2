"""SO(3)/SE(3) primitives, batched numpy fp64 (oracle; test infrastructure).

Conventions follow SURVEY.md Appendix A.1 (GTSAM 4.0 semantics, full EXPMAP charts):
  * Pose3 tangent = [omega(3), v(3)], right/body retraction  X (+) xi = X * Expmap(xi).
  * Pose3 is carried as (R[...,3,3], t[...,3]).
All functions broadcast over leading dimensions.
"""
import numpy as np

EPS = np.finfo(np.float64).eps


def skew(w):
    w = np.asarray(w, dtype=np.float64)
    z = np.zeros(w.shape[:-1])
    return np.stack([
        np.stack([z, -w[..., 2], w[..., 1]], -1),
        np.stack([w[..., 2], z, -w[..., 0]], -1),
        np.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def _abc(theta2):
    """sin(t)/t, (1-cos t)/t^2, (t - sin t)/t^3 with series near zero."""
    theta2 = np.asarray(theta2, dtype=np.float64)
    small = theta2 < 1e-10
    t2 = np.where(small, 1.0, theta2)
    t = np.sqrt(t2)
    a = np.where(small, 1.0 - theta2 / 6.0, np.sin(t) / t)
    b = np.where(small, 0.5 - theta2 / 24.0, (2.0 * np.sin(0.5 * t) ** 2) / t2)
    c = np.where(small, 1.0 / 6.0 - theta2 / 120.0, (t - np.sin(t)) / (t2 * t))
    return a, b, c


def so3_exp(w):
    """Rodrigues: Rot3::Expmap (A.1)."""
    w = np.asarray(w, dtype=np.float64)
    W = skew(w)
    th2 = np.sum(w * w, -1)
    a, b, _ = _abc(th2)
    I = np.eye(3)
    return I + a[..., None, None] * W + b[..., None, None] * (W @ W)


def so3_log(R):
    """Rot3::Logmap, restating the trace-based branches of GTSAM's SO3::Logmap."""
    R = np.asarray(R, dtype=np.float64)
    tr = R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]
    v = np.stack([R[..., 2, 1] - R[..., 1, 2],
                  R[..., 0, 2] - R[..., 2, 0],
                  R[..., 1, 0] - R[..., 0, 1]], -1)
    tr3 = tr - 3.0
    c = np.clip((tr - 1.0) * 0.5, -1.0, 1.0)
    theta = np.arccos(c)
    s = np.sin(theta)
    near0 = tr3 >= -1e-7
    s_safe = np.where(np.abs(s) < 1e-300, 1.0, s)
    mag = np.where(near0, 0.5 - tr3 * tr3 / 12.0, theta / (2.0 * s_safe))
    out = mag[..., None] * v
    # theta ~ pi: use the column-based formula
    nearpi = np.abs(tr + 1.0) < 1e-10
    if np.any(nearpi):
        Rf = R.reshape(-1, 3, 3)
        of = out.reshape(-1, 3).copy()
        for i in np.nonzero(nearpi.reshape(-1))[0]:
            Ri = Rf[i]
            if abs(Ri[2, 2] + 1.0) > 1e-10:
                of[i] = (np.pi / np.sqrt(2.0 + 2.0 * Ri[2, 2])) * np.array([Ri[0, 2], Ri[1, 2], 1.0 + Ri[2, 2]])
            elif abs(Ri[1, 1] + 1.0) > 1e-10:
                of[i] = (np.pi / np.sqrt(2.0 + 2.0 * Ri[1, 1])) * np.array([Ri[0, 1], 1.0 + Ri[1, 1], Ri[2, 1]])
            else:
                of[i] = (np.pi / np.sqrt(2.0 + 2.0 * Ri[0, 0])) * np.array([1.0 + Ri[0, 0], Ri[1, 0], Ri[2, 0]])
        out = of.reshape(out.shape)
    return out


def so3_jr(w):
    """Right Jacobian of Exp (GTSAM ExpmapDerivative / DexpFunctor::dexp): I - b W + c W^2."""
    w = np.asarray(w, dtype=np.float64)
    W = skew(w)
    th2 = np.sum(w * w, -1)
    _, b, c = _abc(th2)
    return np.eye(3) - b[..., None, None] * W + c[..., None, None] * (W @ W)


def so3_jr_inv(w):
    """Inverse right Jacobian (GTSAM LogmapDerivative): I + W/2 + k W^2."""
    w = np.asarray(w, dtype=np.float64)
    W = skew(w)
    th2 = np.sum(w * w, -1)
    small = th2 < 1e-10
    t2 = np.where(small, 1.0, th2)
    t = np.sqrt(t2)
    k = np.where(small, 1.0 / 12.0 + th2 / 720.0,
                 1.0 / t2 - (1.0 + np.cos(t)) / (2.0 * t * np.where(small, 1.0, np.sin(t))))
    return np.eye(3) + 0.5 * W + k[..., None, None] * (W @ W)


def se3_exp(xi):
    """Pose3::Expmap([omega, v]) -> (R, t), t = V(omega) v  (A.1)."""
    xi = np.asarray(xi, dtype=np.float64)
    w, v = xi[..., :3], xi[..., 3:]
    R = so3_exp(w)
    W = skew(w)
    th2 = np.sum(w * w, -1)
    _, b, c = _abc(th2)
    V = np.eye(3) + b[..., None, None] * W + c[..., None, None] * (W @ W)
    t = np.einsum('...ij,...j->...i', V, v)
    return R, t


def se3_log(R, t):
    """Pose3::Logmap -> [omega, u] (Agrawal06 closed form as in GTSAM)."""
    w = so3_log(R)
    t = np.asarray(t, dtype=np.float64)
    th = np.sqrt(np.sum(w * w, -1))
    small = th < 1e-10
    ths = np.where(small, 1.0, th)
    W = skew(w / ths[..., None])
    WT = np.einsum('...ij,...j->...i', W, t)
    WWT = np.einsum('...ij,...j->...i', W, WT)
    coef = 1.0 - ths / (2.0 * np.tan(0.5 * ths))
    u = t - (0.5 * ths)[..., None] * WT + coef[..., None] * WWT
    u = np.where(small[..., None], t, u)
    return np.concatenate([w, u], -1)


def pose_compose(Ra, ta, Rb, tb):
    return Ra @ Rb, np.einsum('...ij,...j->...i', Ra, tb) + ta


def pose_inverse(R, t):
    Rt = np.swapaxes(R, -1, -2)
    return Rt, -np.einsum('...ij,...j->...i', Rt, t)


def pose_between(Ra, ta, Rb, tb):
    """a^-1 * b  (Pose3::between / transform_pose_to)."""
    Ri, ti = pose_inverse(Ra, ta)
    return pose_compose(Ri, ti, Rb, tb)


def adjoint(R, t):
    """Pose3::AdjointMap = [[R,0],[[t]x R, R]] for tangent order [rot, trans] (A.1)."""
    R = np.asarray(R, dtype=np.float64)
    Ad = np.zeros(R.shape[:-2] + (6, 6))
    Ad[..., :3, :3] = R
    Ad[..., 3:, 3:] = R
    Ad[..., 3:, :3] = skew(t) @ R
    return Ad


def pose_retract(R, t, xi, chart=0):
    """Values::retract for Pose3: X * ChartAtOrigin::Retract(xi)  (chart 0 = full EXPMAP, see the end of this file)."""
    dR, dt = chart_retract0(xi, chart)
    return pose_compose(R, t, dR, dt)


def pose_local(Ra, ta, Rb, tb, chart=0):
    """Pose3::localCoordinates: ChartAtOrigin::Local(a^-1 b)."""
    R, t = pose_between(Ra, ta, Rb, tb)
    return chart_local0(R, t, chart)


def rzryrx(x, y, z):
    """Rot3::RzRyRx(x,y,z) = Rz(z) Ry(y) Rx(x)  (A.1)."""
    cx, sx, cy, sy, cz, sz = np.cos(x), np.sin(x), np.cos(y), np.sin(y), np.cos(z), np.sin(z)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def quat_from_rot(R):
    """Unit quaternion (w,x,y,z), w >= 0, batched (Shepperd's method)."""
    R = np.asarray(R, dtype=np.float64)
    Rf = R.reshape(-1, 3, 3)
    q = np.zeros((Rf.shape[0], 4))
    for i, M in enumerate(Rf):
        tr = M[0, 0] + M[1, 1] + M[2, 2]
        if tr > 0:
            s = np.sqrt(tr + 1.0) * 2
            q[i] = [0.25 * s, (M[2, 1] - M[1, 2]) / s, (M[0, 2] - M[2, 0]) / s, (M[1, 0] - M[0, 1]) / s]
        elif M[0, 0] > M[1, 1] and M[0, 0] > M[2, 2]:
            s = np.sqrt(1.0 + M[0, 0] - M[1, 1] - M[2, 2]) * 2
            q[i] = [(M[2, 1] - M[1, 2]) / s, 0.25 * s, (M[0, 1] + M[1, 0]) / s, (M[0, 2] + M[2, 0]) / s]
        elif M[1, 1] > M[2, 2]:
            s = np.sqrt(1.0 + M[1, 1] - M[0, 0] - M[2, 2]) * 2
            q[i] = [(M[0, 2] - M[2, 0]) / s, (M[0, 1] + M[1, 0]) / s, 0.25 * s, (M[1, 2] + M[2, 1]) / s]
        else:
            s = np.sqrt(1.0 + M[2, 2] - M[0, 0] - M[1, 1]) * 2
            q[i] = [(M[1, 0] - M[0, 1]) / s, (M[0, 2] + M[2, 0]) / s, (M[1, 2] + M[2, 1]) / s, 0.25 * s]
        if q[i, 0] < 0:
            q[i] = -q[i]
    return q.reshape(R.shape[:-2] + (4,))


def rot_from_quat(q):
    q = np.asarray(q, dtype=np.float64)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)


# --------------------------------------------------------------------------- chart options
# GTSAM chooses the Pose3 / Rot3 retraction at compile time (GTSAM_POSE3_EXPMAP, GTSAM_ROT3_EXPMAP); the reference's build
# flags are unknown, so the charts are an explicit option everywhere a chart is used (Values::retract of a Pose3, the
# residual Local(measured, h) of BetweenFactor / PriorFactor<Pose3>, the 6-vector of the VRO log):
EXPMAP, FIRST_ORDER_EXPMAP, FIRST_ORDER_CAYLEY = 0, 2, 3        # ids shared with include/fg_abi.h (1 is g2o's chart)


def cayley_retract(w):
    """Rot3::CayleyChart::Retract: the Cayley transform of [w/2]x."""
    w = np.asarray(w, dtype=np.float64)
    x, y, z = w[..., 0], w[..., 1], w[..., 2]
    x2, y2, z2, xy, xz, yz = x * x, y * y, z * z, x * y, x * z, y * z
    f = 1.0 / (4.0 + x2 + y2 + z2); f2 = 2.0 * f
    return np.stack([np.stack([(4 + x2 - y2 - z2) * f, (xy - 2 * z) * f2, (xz + 2 * y) * f2], -1),
                     np.stack([(xy + 2 * z) * f2, (4 - x2 + y2 - z2) * f, (yz - 2 * x) * f2], -1),
                     np.stack([(xz - 2 * y) * f2, (yz + 2 * x) * f2, (4 - x2 - y2 + z2) * f], -1)], -2)


def cayley_local(R):
    """Rot3::CayleyChart::Local: the inverse of cayley_retract, [w/2]x = (R - I)(R + I)^-1."""
    R = np.asarray(R, dtype=np.float64)
    I = np.eye(3)
    A = (R - I) @ np.linalg.inv(R + I)
    return np.stack([A[..., 2, 1] - A[..., 1, 2], A[..., 0, 2] - A[..., 2, 0], A[..., 1, 0] - A[..., 0, 1]], -1)      # vee: [a]x -> 2a = w


def chart_retract0(xi, chart=EXPMAP):
    """Pose3::ChartAtOrigin::Retract under the chosen charts."""
    if chart == EXPMAP:
        return se3_exp(xi)
    xi = np.asarray(xi, dtype=np.float64)
    R = so3_exp(xi[..., :3]) if chart == FIRST_ORDER_EXPMAP else cayley_retract(xi[..., :3])
    return R, xi[..., 3:].copy()


def chart_local0(R, t, chart=EXPMAP):
    """Pose3::ChartAtOrigin::Local under the chosen charts."""
    if chart == EXPMAP:
        return se3_log(R, t)
    w = so3_log(R) if chart == FIRST_ORDER_EXPMAP else cayley_local(R)
    return np.concatenate([w, np.asarray(t, dtype=np.float64)], -1)
