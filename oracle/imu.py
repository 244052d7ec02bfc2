"""IMU preintegration (oracle; test infrastructure).

Restates gtsam::PreintegratedCombinedMeasurements with TangentPreintegration
 as driven by CImuBase::predictNext (gtsam/imu_base.cpp:72-87):
for every 200 Hz sample in the interval, integrateMeasurement(acc=imu.tail<3>,
gyro=imu.head<3>, dt); then predict(prev_state, prev_bias).
Samples are stored [gx gy gz ax ay az] as the reference stores them
(gtsam/imu_vn100.cpp:96).  Batched over intervals, sequential over samples.
"""
import numpy as np
from . import lie


def vn100_params():
    """Noise spec of CImuVn100::getIMUParams (gtsam/imu_vn100.cpp:24-67) and
    CImuBase::getParam MakeSharedD(9.71) (gtsam/imu_base.cpp:258-263)."""
    g = 9.81
    fps = 200.0
    acc_sigma = 0.14 * 1e-3 * g
    gyro_sigma = np.deg2rad(0.0035)
    acc_bias_rw = (0.04 * 1e-3 * g) * np.sqrt(fps)
    gyro_bias_rw = (np.deg2rad(10.0) / 3600.0) * np.sqrt(fps)
    I3 = np.eye(3)
    return dict(
        acc_cov=I3 * acc_sigma ** 2,
        gyro_cov=I3 * gyro_sigma ** 2,
        int_cov=I3 * 1e-4,
        bias_acc_cov=I3 * acc_bias_rw ** 2,
        bias_gyro_cov=I3 * gyro_bias_rw ** 2,
        bias_acc_omega_int=np.eye(6) * 1e-3,
        gravity=np.array([0.0, 0.0, 9.71]),   # MakeSharedD(g): n_gravity = (0,0,+g)
    )


def _d_jr_c(theta, c):
    """d/dtheta [ Jr(theta) c ] at fixed c (exact; replaces DexpFunctor::applyDexp's H1)."""
    th2 = np.sum(theta * theta, -1)
    small = th2 < 1e-10
    t2 = np.where(small, 1.0, th2)
    t = np.sqrt(t2)
    b = np.where(small, 0.5 - th2 / 24.0, (1 - np.cos(t)) / t2)
    c3 = np.where(small, 1.0 / 6.0 - th2 / 120.0, (t - np.sin(t)) / (t2 * t))
    db = np.where(small, -1.0 / 12.0, (t * np.sin(t) - 2 * (1 - np.cos(t))) / (t2 * t) / t)      # (db/dt)/t
    dc = np.where(small, -1.0 / 60.0, ((1 - np.cos(t)) / (t2 * t) - 3 * (t - np.sin(t)) / (t2 * t2)) / t)
    txc = np.cross(theta, c)
    ttc = np.cross(theta, txc)
    tc = np.sum(theta * c, -1)
    I3 = np.eye(3)
    D = b[..., None, None] * lie.skew(c) \
        - txc[..., :, None] * (db[..., None] * theta)[..., None, :] \
        + c3[..., None, None] * (tc[..., None, None] * I3 + theta[..., :, None] * c[..., None, :]
                                 - 2 * c[..., :, None] * theta[..., None, :]) \
        + ttc[..., :, None] * (dc[..., None] * theta)[..., None, :]
    return D


def update_preintegrated(a_body, w_body, dt, pre):
    """TangentPreintegration::UpdatePreintegrated -> (pre+, A 9x9, B 9x3, C 9x3)."""
    theta, pos, vel = pre[..., 0:3], pre[..., 3:6], pre[..., 6:9]
    Jr = lie.so3_jr(theta)
    invH = np.linalg.inv(Jr)
    w_tan = np.einsum('...ij,...j->...i', invH, w_body)
    w_tan_H_theta = -invH @ _d_jr_c(theta, w_tan)
    R = lie.so3_exp(theta)
    a_nav = np.einsum('...ij,...j->...i', R, a_body)
    dt22 = 0.5 * dt * dt
    out = np.concatenate([theta + w_tan * dt, pos + vel * dt + a_nav * dt22, vel + a_nav * dt], -1)
    shp = pre.shape[:-1]
    a_nav_H_theta = R @ lie.skew(-a_body) @ Jr
    A = np.broadcast_to(np.eye(9), shp + (9, 9)).copy()
    A[..., 0:3, 0:3] += w_tan_H_theta * dt
    A[..., 3:6, 0:3] = a_nav_H_theta * dt22
    A[..., 3:6, 6:9] = np.eye(3) * dt
    A[..., 6:9, 0:3] = a_nav_H_theta * dt
    B = np.zeros(shp + (9, 3))
    B[..., 3:6, :] = R * dt22
    B[..., 6:9, :] = R * dt
    C = np.zeros(shp + (9, 3))
    C[..., 0:3, :] = invH * dt
    return out, A, B, C


def preintegrate(samples, dt, params, bias_hat):
    """samples (N,S,6) [gyro, acc]; bias_hat (N,6) [acc, gyro] -> pim dict (batched over N)."""
    samples = np.asarray(samples, dtype=np.float64)
    N, S, _ = samples.shape
    bias_hat = np.broadcast_to(np.asarray(bias_hat, dtype=np.float64), (N, 6)).copy()
    pre = np.zeros((N, 9))
    Hba = np.zeros((N, 9, 3))
    Hbg = np.zeros((N, 9, 3))
    P = np.zeros((N, 15, 15))
    aCov, wCov, iCov = params['acc_cov'], params['gyro_cov'], params['int_cov']
    bint = params['bias_acc_omega_int']
    for s in range(S):
        acc = samples[:, s, 3:6] - bias_hat[:, 0:3]
        omega = samples[:, s, 0:3] - bias_hat[:, 3:6]
        pre, A, B, C = update_preintegrated(acc, omega, dt, pre)
        Hba = A @ Hba - B
        Hbg = A @ Hbg - C
        th_H_bg = -C[:, 0:3, :]
        v_H_ba = -B[:, 6:9, :]
        F = np.zeros((N, 15, 15))
        F[:, 0:9, 0:9] = A
        F[:, 0:3, 12:15] = th_H_bg
        F[:, 6:9, 9:12] = v_H_ba
        F[:, 9:15, 9:15] = np.eye(6)
        G = np.zeros((N, 15, 15))
        G[:, 3:6, 3:6] = dt * iCov
        G[:, 6:9, 6:9] = (1.0 / dt) * v_H_ba @ (aCov + bint[0:3, 0:3]) @ np.swapaxes(v_H_ba, -1, -2)
        G[:, 0:3, 0:3] = (1.0 / dt) * th_H_bg @ (wCov + bint[3:6, 3:6]) @ np.swapaxes(th_H_bg, -1, -2)
        G[:, 9:12, 9:12] = dt * params['bias_acc_cov']
        G[:, 12:15, 12:15] = dt * params['bias_gyro_cov']
        temp = v_H_ba @ bint[3:6, 0:3] @ np.swapaxes(th_H_bg, -1, -2)
        G[:, 6:9, 0:3] = temp
        G[:, 0:3, 6:9] = np.swapaxes(temp, -1, -2)
        P = F @ P @ np.swapaxes(F, -1, -2) + G
    return dict(dt=np.full(N, S * dt), preint=pre, Hba=Hba, Hbg=Hbg, cov=P,
                bias_hat=bias_hat, gravity=np.asarray(params['gravity'], dtype=np.float64))


def predict(pim, Ri, ti, vi, bias_i):
    """PreintegrationBase::predict(state_i, bias_i) -> (Rj, tj, vj)  (A.5)."""
    inc = bias_i - pim['bias_hat']
    bc = pim['preint'] + np.einsum('...ij,...j->...i', pim['Hba'], inc[..., :3]) \
        + np.einsum('...ij,...j->...i', pim['Hbg'], inc[..., 3:])
    dt = np.asarray(pim['dt'])
    g = pim['gravity']
    RiT = np.swapaxes(Ri, -1, -2)
    Rtv = np.einsum('...ij,...j->...i', RiT, vi)
    Rtg = np.einsum('...ij,j->...i', RiT, g)
    xi_p = bc[..., 3:6] + dt[..., None] * Rtv + (0.5 * dt * dt)[..., None] * Rtg
    xi_v = bc[..., 6:9] + dt[..., None] * Rtg
    Rj = Ri @ lie.so3_exp(bc[..., 0:3])
    tj = ti + np.einsum('...ij,...j->...i', Ri, xi_p)
    vj = vi + np.einsum('...ij,...j->...i', Ri, xi_v)
    return Rj, tj, vj
