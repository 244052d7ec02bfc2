"""The oracle's analytic Jacobians against manifold central differences, and closed-form minimisers on tiny graphs
(what validates the oracle where the reference holds no golden vectors -- BetweenFactor, PriorFactor,
GenericProjectionFactor/Cal3DS2, CombinedImuFactor, Levenberg-Marquardt)."""
import numpy as np
from graph_slam_b200 import synth
from oracle import lie, factors as F, imu as oimu, lm
from oracle.graph import Graph

EPS = 1e-6


def num_jac_pose(fun, R, t):
    cols = []
    for k in range(6):
        d = np.zeros(6); d[k] = EPS
        cols.append((fun(*lie.pose_retract(R, t, d)) - fun(*lie.pose_retract(R, t, -d))) / (2 * EPS))
    return np.stack(cols, -1)


def num_jac_vec(fun, x):
    cols = []
    for k in range(len(x)):
        d = np.zeros(len(x)); d[k] = EPS
        cols.append((fun(x + d) - fun(x - d)) / (2 * EPS))
    return np.stack(cols, -1)


def rand_pose(rng, s=0.6):
    return lie.se3_exp(rng.normal(size=6) * s)


def test_projection_jacobians():
    rng = np.random.default_rng(1)
    spec = synth.make_config('C4', seed=1, scale=0.01)
    K, Rs, ts = spec['K'], spec['Rs'], spec['ts']
    n = 0
    while n < 50:
        R, t = rand_pose(rng)
        Rc, tc = lie.pose_compose(R, t, Rs, ts)
        p = Rc @ np.array([rng.uniform(-0.8, 0.8), rng.uniform(-0.6, 0.6), rng.uniform(1.0, 5.0)]) + tc      # in front of the camera
        uv = rng.normal(size=2) * 50 + 80
        r, Jp, Jl = F.projection(R, t, p, uv, K, Rs, ts)
        Np = num_jac_pose(lambda Rq, tq: F.projection(Rq, tq, p, uv, K, Rs, ts, jac=False), R, t)
        Nl = num_jac_vec(lambda q: F.projection(R, t, q, uv, K, Rs, ts, jac=False), p)
        scale = max(1.0, np.abs(Np).max())
        assert np.abs(Jp - Np).max() <= 1e-6 * scale and np.abs(Jl - Nl).max() <= 1e-6 * scale
        n += 1
    # cheirality: a point behind the camera gives the constant residual (2 fx, 2 fx) and zero Jacobians (A.2)
    R, t = np.eye(3), np.zeros(3)
    Rc, tc = lie.pose_compose(R, t, Rs, ts)
    behind = Rc @ np.array([0.1, 0.1, -2.0]) + tc
    r, Jp, Jl = F.projection(R, t, behind, np.zeros(2), K, Rs, ts)
    assert np.allclose(r, [2 * K[0], 2 * K[0]]) and not Jp.any() and not Jl.any()


def test_between_and_prior_jacobians_at_zero_residual():
    # GTSAM's fast path H1 = -Ad(h^-1), H2 = I is the exact derivative only where the residual vanishes (A.3)
    rng = np.random.default_rng(2)
    for _ in range(30):
        R1, t1 = rand_pose(rng); R2, t2 = rand_pose(rng)
        Rz, tz = lie.pose_between(R1, t1, R2, t2)
        r, H1, H2 = F.between_pose(R1, t1, R2, t2, Rz, tz)
        assert np.abs(r).max() < 1e-12
        N1 = num_jac_pose(lambda Ra, ta: F.between_pose(Ra, ta, R2, t2, Rz, tz, jac=False), R1, t1)
        N2 = num_jac_pose(lambda Rb, tb: F.between_pose(R1, t1, Rb, tb, Rz, tz, jac=False), R2, t2)
        assert np.abs(H1 - N1).max() <= 1e-6 * max(1.0, np.abs(N1).max()) and np.abs(H2 - N2).max() <= 1e-6
        r, H = F.prior_pose(R1, t1, R1, t1)
        N = num_jac_pose(lambda Ra, ta: F.prior_pose(Ra, ta, R1, t1, jac=False), R1, t1)
        assert np.abs(H - N).max() <= 1e-6
    # away from it the fast path is a first-order approximation: the gap shrinks with the residual
    R1, t1 = rand_pose(rng); R2, t2 = rand_pose(rng)
    gaps = []
    for s in (1e-1, 1e-2):
        Rz, tz = lie.pose_retract(*lie.pose_between(R1, t1, R2, t2), np.full(6, s))
        r, H1, H2 = F.between_pose(R1, t1, R2, t2, Rz, tz)
        N2 = num_jac_pose(lambda Rb, tb: F.between_pose(R1, t1, Rb, tb, Rz, tz, jac=False), R2, t2)
        gaps.append(np.abs(H2 - N2).max())
    assert gaps[1] < 0.2 * gaps[0]


def test_combined_imu_factor_jacobians():
    rng = np.random.default_rng(3)
    spec = synth.make_config('C2', seed=2, scale=0.02)
    pim_all = oimu.preintegrate(spec['imu_samples'], spec['imu_dt'], oimu.vn100_params(), rng.normal(size=(spec['n_poses'] - 1, 6)) * 0.01)
    for k in range(min(8, spec['n_poses'] - 1)):
        pim = {key: (pim_all[key][k] if key != 'gravity' else pim_all[key]) for key in ('dt', 'preint', 'Hba', 'Hbg', 'bias_hat', 'gravity')}
        Ri, ti = rand_pose(rng); vi = rng.normal(size=3)
        bi = pim['bias_hat'] + rng.normal(size=6) * 0.01
        # state j near the prediction so that the residual is small but not zero
        r0 = F.imu_combined(Ri, ti, vi, Ri, ti, vi, bi, bi, pim, jac=False)
        Rj, tj = lie.pose_retract(Ri, ti, rng.normal(size=6) * 0.05); vj = vi + rng.normal(size=3) * 0.05
        bj = bi + rng.normal(size=6) * 0.001
        r, J = F.imu_combined(Ri, ti, vi, Rj, tj, vj, bi, bj, pim)
        nums = [num_jac_pose(lambda R, t: F.imu_combined(R, t, vi, Rj, tj, vj, bi, bj, pim, jac=False), Ri, ti),
                num_jac_vec(lambda v: F.imu_combined(Ri, ti, v, Rj, tj, vj, bi, bj, pim, jac=False), vi),
                num_jac_pose(lambda R, t: F.imu_combined(Ri, ti, vi, R, t, vj, bi, bj, pim, jac=False), Rj, tj),
                num_jac_vec(lambda v: F.imu_combined(Ri, ti, vi, Rj, tj, v, bi, bj, pim, jac=False), vj),
                num_jac_vec(lambda b: F.imu_combined(Ri, ti, vi, Rj, tj, vj, b, bj, pim, jac=False), bi),
                num_jac_vec(lambda b: F.imu_combined(Ri, ti, vi, Rj, tj, vj, bi, b, pim, jac=False), bj)]
        for a, n in zip(J, nums):
            assert np.abs(a - n).max() <= 2e-6 * max(1.0, np.abs(n).max()), np.abs(a - n).max()


def test_two_pose_chain_closed_form_minimiser():
    # prior on X0 and one BetweenFactor: the optimum is X1 = X0 * Z with zero error, whatever the initial guess
    rng = np.random.default_rng(4)
    Rz, tz = rand_pose(rng)
    g = Graph()
    R1, t1 = lie.pose_retract(Rz, tz, rng.normal(size=6) * 0.3)
    g.R = np.stack([np.eye(3), R1]); g.t = np.stack([np.zeros(3), t1])
    A = rng.normal(size=(6, 6)); info = A @ A.T + 6 * np.eye(6)
    g.f = dict(prior_pose=dict(i=np.array([0]), R=np.eye(3)[None], t=np.zeros((1, 3)), info=np.eye(6)[None] / 1e-7 ** 2),
               between=dict(i=np.array([0]), j=np.array([1]), R=Rz[None], t=tz[None], info=info[None]))
    g1, rep = lm.optimize_gtsam(g)
    assert rep['error'] < 1e-12
    assert np.abs(g1.R[1] - Rz).max() < 1e-7 and np.abs(g1.t[1] - tz).max() < 1e-7


def test_single_landmark_triangulation_minimiser():
    # two views tied by a stiff BetweenFactor, one landmark with a weak prior and noise-free projections: LM recovers the point
    rng = np.random.default_rng(5)
    spec = synth.make_config('C4', seed=1, scale=0.01)
    K, Rs, ts = spec['K'], spec['Rs'], spec['ts']
    R1, t1 = lie.se3_exp(np.array([0.02, -0.05, 0.03, 0.3, 0.05, -0.02]))
    Rc, tc = lie.pose_compose(np.eye(3), np.zeros(3), Rs, ts)
    truth = Rc @ np.array([0.2, -0.1, 3.0]) + tc
    uv = np.stack([F.projection(np.eye(3), np.zeros(3), truth, np.zeros(2), K, Rs, ts, jac=False),
                   F.projection(R1, t1, truth, np.zeros(2), K, Rs, ts, jac=False)])
    g = Graph()
    g.R = np.stack([np.eye(3), R1]); g.t = np.stack([np.zeros(3), t1]); g.K = tuple(K); g.Rs = Rs; g.ts = ts
    g.point = (truth + np.array([0.05, -0.04, 0.2]))[None]
    g.f = dict(prior_pose=dict(i=np.array([0]), R=np.eye(3)[None], t=np.zeros((1, 3)), info=np.eye(6)[None] / 1e-7 ** 2),
               between=dict(i=np.array([0]), j=np.array([1]), R=R1[None], t=t1[None], info=np.eye(6)[None] * 1e12),
               prior_point=dict(i=np.array([0]), mean=g.point.copy(), info=np.eye(3)[None] * 1e-6),
               proj=dict(i=np.array([0, 1]), l=np.array([0, 0]), uv=uv, sigma=1.0))
    g1, rep = lm.optimize_gtsam(g)
    assert np.abs(g1.point[0] - truth).max() < 1e-5 and rep['error'] < 1e-6


def test_preintegration_closed_forms():
    """Constant specific force without rotation, and constant rotation rate without force (SURVEY 8c ii): the preintegrated
    deltas, their bias Jacobians (checked by re-integrating with a shifted bias) and predict() have closed forms."""
    par = oimu.vn100_params()
    S, dt = 20, 0.005
    T = S * dt
    a = np.array([0.3, -0.2, 9.6])
    smp = np.zeros((1, S, 6)); smp[0, :, 3:6] = a
    pim = oimu.preintegrate(smp, dt, par, np.zeros((1, 6)))
    assert np.allclose(pim['preint'][0, 0:3], 0, atol=1e-15)
    assert np.allclose(pim['preint'][0, 3:6], 0.5 * a * T * T, atol=1e-13) and np.allclose(pim['preint'][0, 6:9], a * T, atol=1e-13)
    assert np.allclose(pim['Hba'][0, 3:6], -0.5 * T * T * np.eye(3), atol=1e-13) and np.allclose(pim['Hba'][0, 6:9], -T * np.eye(3), atol=1e-13)
    # predict from rest with gravity n_gravity = (0, 0, +g): p = 1/2 (a + g) T^2, v = (a + g) T
    Rj, tj, vj = oimu.predict({k: (v[0] if k != 'gravity' else v) for k, v in pim.items()}, np.eye(3), np.zeros(3), np.zeros(3), np.zeros(6))
    assert np.allclose(tj, 0.5 * (a + par['gravity']) * T * T, atol=1e-13) and np.allclose(vj, (a + par['gravity']) * T, atol=1e-13)
    # constant yaw rate, no force
    w = np.array([0.0, 0.0, 0.7])
    smp = np.zeros((1, S, 6)); smp[0, :, 0:3] = w
    pim = oimu.preintegrate(smp, dt, par, np.zeros((1, 6)))
    assert np.allclose(pim['preint'][0, 0:3], w * T, atol=1e-12) and np.allclose(pim['preint'][0, 3:9], 0, atol=1e-15)
    # first-order bias correction: preint(b) ~= preint(0) + Hba b_a + Hbg b_g  (error second order in the bias)
    rng = np.random.default_rng(6)
    smp = rng.normal(size=(1, S, 6)) * np.array([0.3, 0.3, 0.3, 2, 2, 2]) + np.array([0, 0, 0, 0, 0, 9.7])
    p0 = oimu.preintegrate(smp, dt, par, np.zeros((1, 6)))
    for scale in (1e-3, 1e-4):
        b = rng.normal(size=6) * scale
        p1 = oimu.preintegrate(smp, dt, par, b[None])
        lin = p0['preint'][0] + p0['Hba'][0] @ b[:3] + p0['Hbg'][0] @ b[3:]
        assert np.abs(p1['preint'][0] - lin).max() <= 5.0 * scale ** 2
    # the covariance is symmetric positive definite
    C = p0['cov'][0]
    assert np.allclose(C, C.T, atol=1e-18) and np.linalg.eigvalsh(C).min() > 0
