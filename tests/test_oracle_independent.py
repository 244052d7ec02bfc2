"""Independent evaluations of the oracle's "parity unpinned" rows: the reference holds no golden vectors for
Pose3 / BetweenFactor / GenericProjectionFactor / preintegration and its libraries cannot be built here, so the numpy oracle is
checked against implementations that share no code and no derivation with it:

  * SO(3) / SE(3) Expmap and Logmap  vs  scipy.linalg.expm / logm of the 4x4 twist matrix and scipy's Rotation (rotvec, quat);
  * the adjoint map                   vs  conjugation  T Exp(xi) T^-1  evaluated with expm / logm;
  * BetweenFactor / PriorFactor       vs  logm(Z^-1 X1^-1 X2) on homogeneous matrices;
  * Cal3DS2 projection                vs  a scalar re-derivation of the pinhole + Brown-Conrady model with a Newton undistortion round trip;
  * g2o EdgeSE3 error / oplus         vs  scipy quaternions;
  * TangentPreintegration             vs  brute-force integration of the continuous kinematics at 1/200 of the sample period.

CPU only (the CUDA kernels are compared with the same oracle in the -m gpu tests)."""
import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.spatial.transform import Rotation

from oracle import lie, factors


def hat6(xi):
    """GTSAM's Pose3 twist [omega, v] as a 4x4 matrix."""
    w, v = xi[:3], xi[3:]
    T = np.zeros((4, 4))
    T[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    T[:3, 3] = v
    return T


def vee6(T):
    return np.array([T[2, 1], T[0, 2], T[1, 0], T[0, 3], T[1, 3], T[2, 3]])


def homog(R, t):
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    return T


RNG = np.random.default_rng(123)
TWISTS = [RNG.normal(size=6) * s for s in (1e-9, 1e-4, 0.1, 1.0, 2.5) for _ in range(4)]


@pytest.mark.parametrize('k', range(len(TWISTS)))
def test_se3_expmap_is_the_matrix_exponential(k):
    xi = TWISTS[k]
    R, t = lie.se3_exp(xi)
    E = expm(hat6(xi))
    assert np.abs(homog(R, t) - E).max() < 1e-12
    # and the rotation part is scipy's rotation vector
    assert np.abs(R - Rotation.from_rotvec(xi[:3]).as_matrix()).max() < 1e-13


@pytest.mark.parametrize('k', range(len(TWISTS)))
def test_se3_logmap_is_the_matrix_logarithm(k):
    xi = TWISTS[k]
    if np.linalg.norm(xi[:3]) > 3.0:
        xi = xi * (3.0 / np.linalg.norm(xi[:3]))                 # stay inside the principal branch
    E = expm(hat6(xi))
    back = lie.se3_log(E[:3, :3], E[:3, 3])
    ref = vee6(np.real(logm(E))) if np.linalg.norm(xi) > 1e-6 else xi   # logm loses digits next to the identity
    assert np.abs(back - ref).max() < 1e-9
    assert np.abs(back - xi).max() < 1e-9
    assert np.abs(lie.so3_log(E[:3, :3]) - Rotation.from_matrix(E[:3, :3]).as_rotvec()).max() < 1e-10


def test_so3_logmap_near_pi_matches_scipy():
    """At pi - 1e-3 the trace formula is still in force (rounding-level agreement with scipy).  GTSAM's SO3::Logmap switches to
    its one-column formula when |trace + 1| < 1e-10, i.e. within ~1e-5 rad of pi, and that formula is exact only AT pi: the
    restatement keeps the branch (the device so3_log too), so at pi - 1e-6 the agreement is 1e-5, not rounding."""
    for axis in (np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), np.array([0, 0, 1.0]), np.array([1.0, 2.0, -0.5]) / np.linalg.norm([1.0, 2.0, -0.5])):
        for ang, tol in ((np.pi - 1e-3, 1e-9), (np.pi - 1e-6, 1e-5), (np.pi, 1e-9)):
            R = Rotation.from_rotvec(axis * ang).as_matrix()
            w = lie.so3_log(R)
            # the rotation vector is defined up to sign at pi: compare the rotations
            assert np.abs(lie.so3_exp(w) - R).max() < tol, (axis, ang)
            assert abs(np.linalg.norm(w) - ang) < tol


def test_adjoint_is_conjugation():
    for _ in range(5):
        R, t = lie.se3_exp(RNG.normal(size=6))
        T = homog(R, t)
        xi = RNG.normal(size=6) * 0.3
        lhs = vee6(np.real(logm(T @ expm(hat6(xi)) @ np.linalg.inv(T))))
        assert np.abs(lie.adjoint(R, t) @ xi - lhs).max() < 1e-10


def test_between_and_prior_errors_are_logs_of_homogeneous_products():
    for _ in range(5):
        R1, t1 = lie.se3_exp(RNG.normal(size=6)); R2, t2 = lie.se3_exp(RNG.normal(size=6))
        Rm, tm = lie.se3_exp(RNG.normal(size=6) * 0.5)
        r = factors.between_pose(R1, t1, R2, t2, Rm, tm, jac=False)
        r = r[0] if isinstance(r, tuple) else r
        D = np.linalg.inv(homog(Rm, tm)) @ np.linalg.inv(homog(R1, t1)) @ homog(R2, t2)
        assert np.abs(np.asarray(r).ravel() - vee6(np.real(logm(D)))).max() < 1e-9
        rp = factors.prior_pose(R2, t2, R1, t1, jac=False)
        rp = rp[0] if isinstance(rp, tuple) else rp
        Dp = np.linalg.inv(homog(R1, t1)) @ homog(R2, t2)
        assert np.abs(np.asarray(rp).ravel() - vee6(np.real(logm(Dp)))).max() < 1e-9


def _brown_conrady(x, y, K):
    fx, fy, s, u0, v0, k1, k2, p1, p2 = K
    r2 = x * x + y * y
    g = 1 + k1 * r2 + k2 * r2 * r2
    dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    xd, yd = g * x + dx, g * y + dy
    return fx * xd + s * yd + u0, fy * yd + v0


def test_cal3ds2_projection_scalar_rederivation_and_undistortion_round_trip():
    K = np.array([250.5773, 250.5773, 0.3, 90.0, 70.0, -0.8466, 0.5370, 1e-3, -2e-3])
    Rs, ts = lie.se3_exp(np.array([0.1, -0.2, 0.05, 0.02, -0.01, 0.03]))
    for _ in range(10):
        R, t = lie.se3_exp(RNG.normal(size=6) * 0.4)
        pc = np.array([RNG.uniform(-0.3, 0.3), RNG.uniform(-0.3, 0.3), RNG.uniform(2.0, 5.0)])   # in the camera frame
        Tc = homog(R, t) @ homog(Rs, ts)                          # camera pose = body pose * body_P_sensor
        pw = Tc[:3, :3] @ pc + Tc[:3, 3]
        r = factors.projection(R, t, pw[None, :], np.zeros((1, 2)), K, Rs, ts, jac=False)
        r = r[0] if isinstance(r, tuple) else r
        u, v = _brown_conrady(pc[0] / pc[2], pc[1] / pc[2], K)
        assert np.abs(np.asarray(r).ravel() - np.array([u, v])).max() < 1e-10
        # Newton undistortion of (u, v) gives back the normalised point (the model, not only the formula, is the standard one)
        x, y = pc[0] / pc[2], pc[1] / pc[2]
        xn, yn = (u - K[3]) / K[0], (v - K[4]) / K[1]
        for _it in range(50):
            h = 1e-7
            f0 = np.array(_brown_conrady(xn, yn, K)) - np.array([u, v])
            J = np.column_stack([(np.array(_brown_conrady(xn + h, yn, K)) - np.array(_brown_conrady(xn - h, yn, K))) / (2 * h),
                                 (np.array(_brown_conrady(xn, yn + h, K)) - np.array(_brown_conrady(xn, yn - h, K))) / (2 * h)])
            d = np.linalg.solve(J, f0)
            xn, yn = xn - d[0], yn - d[1]
        assert abs(xn - x) < 1e-8 and abs(yn - y) < 1e-8


def test_g2o_edge_error_and_oplus_against_scipy_quaternions():
    for _ in range(5):
        R1, t1 = lie.se3_exp(RNG.normal(size=6)); R2, t2 = lie.se3_exp(RNG.normal(size=6))
        Rm, tm = lie.se3_exp(RNG.normal(size=6) * 0.3)
        e = np.asarray(factors.g2o_edge_se3(R1, t1, R2, t2, Rm, tm)).ravel()
        D = np.linalg.inv(homog(Rm, tm)) @ np.linalg.inv(homog(R1, t1)) @ homog(R2, t2)
        q = Rotation.from_matrix(D[:3, :3]).as_quat()            # (x, y, z, w)
        if q[3] < 0:
            q = -q
        assert np.abs(e - np.concatenate([D[:3, 3], q[:3]])).max() < 1e-10
        d = np.concatenate([RNG.normal(size=3) * 0.1, RNG.normal(size=3) * 0.05])
        Rn, tn = factors.g2o_oplus(R1, t1, d)
        qw = np.sqrt(max(0.0, 1.0 - d[3:] @ d[3:]))
        Tn = homog(R1, t1) @ homog(Rotation.from_quat(np.concatenate([d[3:], [qw]])).as_matrix(), d[:3])
        assert np.abs(homog(Rn, tn) - Tn).max() < 1e-12


def test_preintegration_against_brute_force_integration():
    """Constant body-frame rate and specific force over one sample, many samples: the preintegrated (theta, p, v) must converge to
    the continuous kinematics  R' = R [w]x,  v' = R a,  p' = v  integrated with a 200 times finer midpoint rule, at the rate the
    sample-and-hold discretisation allows (first order in dt), and agree to rounding for a rate-only / force-only motion."""
    from oracle import imu as oimu
    dt, n = 0.005, 40
    w = np.array([0.3, -0.2, 0.5]); a = np.array([0.4, 9.0, -0.3])
    samples = np.tile(np.concatenate([w, a]), (1, n, 1))          # stored [gyro, acc] (gtsam/imu_vn100.cpp:96)
    pim = oimu.preintegrate(samples, dt, oimu.vn100_params(), np.zeros(6))
    theta, p, v = pim['preint'][0, :3], pim['preint'][0, 3:6], pim['preint'][0, 6:9]
    # brute force
    R = np.eye(3); pp = np.zeros(3); vv = np.zeros(3)
    h = dt / 200.0
    for _ in range(n * 200):
        Rm = R @ Rotation.from_rotvec(w * h * 0.5).as_matrix()
        pp = pp + vv * h + 0.5 * (Rm @ a) * h * h
        vv = vv + (Rm @ a) * h
        R = R @ Rotation.from_rotvec(w * h).as_matrix()
    assert np.abs(lie.so3_exp(theta) - R).max() < 1e-10                 # constant rate: exact
    T = n * dt
    assert np.abs(v - vv).max() < 2.0 * np.linalg.norm(a) * np.linalg.norm(w) * dt * T    # sample-and-hold error bound (first order in dt)
    assert np.abs(p - pp).max() < 2.0 * np.linalg.norm(a) * np.linalg.norm(w) * dt * T * T
    # halving dt halves the discretisation error
    samples2 = np.tile(np.concatenate([w, a]), (1, 2 * n, 1))
    pim2 = oimu.preintegrate(samples2, dt / 2, oimu.vn100_params(), np.zeros(6))
    e1, e2 = np.abs(v - vv).max(), np.abs(pim2['preint'][0, 6:9] - vv).max()
    assert e2 < 0.6 * e1
    # no rotation: exact
    samples0 = np.tile(np.concatenate([np.zeros(3), a]), (1, n, 1))
    pim0 = oimu.preintegrate(samples0, dt, oimu.vn100_params(), np.zeros(6))
    assert np.abs(pim0['preint'][0, 6:9] - a * T).max() < 1e-12 and np.abs(pim0['preint'][0, 3:6] - 0.5 * a * T * T).max() < 1e-12


def test_lm_minimum_of_a_small_pose_graph_against_scipy_least_squares():
    """The optimum, not the trace: a 6-pose graph (gauge prior + 9 Between factors, full 6x6 information matrices) minimised
    (i) by the oracle's restatement of GTSAM's Levenberg-Marquardt and (ii) by scipy.optimize.least_squares over global
    exponential coordinates with residuals written from scratch as logm(Z^-1 X_i^-1 X_j) -- different parametrisation,
    different solver, finite-difference Jacobians.

    The two objectives agree to rounding at the starting point.  The optima agree to 1e-5 relative and GTSAM's is the higher one:
    BetweenFactor's default Jacobians (H1 = -Ad(h^-1), H2 = I; SLOW_BUT_CORRECT_BETWEENFACTOR off) leave out the
    derivative of Logmap, so with non-zero residuals GTSAM's iteration stops where the APPROXIMATE gradient vanishes and rejects
    every further step -- the restatement reproduces that (and the CUDA path follows the restatement)."""
    from scipy.optimize import least_squares
    from graph_slam_b200 import synth
    from oracle import build, lm
    spec = synth.make_graph(6, seed=4, vro=True, imu=False, loop_closure_frac=0.0, lookback=2)
    g1, rep = lm.optimize_gtsam(build.from_spec(spec), lm.LMParams(relative_error_tol=1e-14, absolute_error_tol=1e-14, max_iterations=50))
    P = spec['n_poses']
    Zs = [homog(spec['between_R'][k], spec['between_t'][k]) for k in range(len(spec['between_i']))]
    Ls = [np.linalg.cholesky(spec['between_info'][k]).T for k in range(len(Zs))]          # r^T Omega r = |L r|^2
    prior = homog(spec['prior_pose_R'], spec['prior_pose_t'])

    def poses(x):
        return [expm(hat6(x[6 * i:6 * i + 6])) for i in range(P)]

    def resid(x):
        T = poses(x)
        out = [vee6(np.real(logm(np.linalg.inv(prior) @ T[0]))) / 1e-7]
        for k, (i, j) in enumerate(zip(spec['between_i'], spec['between_j'])):
            out.append(Ls[k] @ vee6(np.real(logm(np.linalg.inv(Zs[k]) @ np.linalg.inv(T[i]) @ T[j]))))
        return np.concatenate(out)

    x0 = np.concatenate([vee6(np.real(logm(homog(spec['pose_init_R'][i], spec['pose_init_t'][i])))) for i in range(P)])
    assert abs(0.5 * np.sum(resid(x0) ** 2) - build.from_spec(spec).error()) < 1e-6 * build.from_spec(spec).error()
    sol = least_squares(resid, x0, method='lm', xtol=1e-15, ftol=1e-15, gtol=1e-15, x_scale=1.0)
    best = 0.5 * np.sum(sol.fun ** 2)
    assert best <= rep['error'] * (1 + 1e-12) and rep['error'] - best < 2e-5 * best, (best, rep['error'])
    T = poses(sol.x)
    for i in range(P):
        assert np.abs(T[i] - homog(g1.R[i], g1.t[i])).max() < 1e-3


def test_mixed_calibration_graph_equals_per_camera_evaluation():
    """Oracle plumbing of several (Cal3DS2, body_P_sensor) pairs in one graph: the graph's chi2 is the sum of the two cameras'
    reprojection terms evaluated separately, and the Jacobian blocks land on the factors of the right camera."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_parity import mixed_calibration_spec
    from oracle import build
    spec = mixed_calibration_spec()
    g = build.from_spec(spec)
    q = g.f['proj']
    e_proj = 0.0
    for cidx, (K, Rs, ts) in enumerate(g.cals):
        m = q['cal'] == cidx
        r = factors.projection(g.R[q['i'][m]], g.t[q['i'][m]], g.point[q['l'][m]], q['uv'][m], K, Rs, ts, jac=False)
        e_proj += 0.5 * np.sum(r * r) / q['sigma'] ** 2
    g_no = build.from_spec({k: v for k, v in spec.items() if k not in ('proj_pose', 'proj_point', 'proj_uv', 'proj_cal', 'cals')} | {'proj_pose': spec['proj_pose'][:0], 'proj_point': spec['proj_point'][:0], 'proj_uv': spec['proj_uv'][:0]})
    assert abs(g.error() - (g_no.error() + e_proj)) < 1e-9 * g.error()
    blocks = [b for b in g.linearize(jac=True) if b[2] is not None and b[0].shape[-1] == 2][0]
    r, _, ((_, Jp), (_, Jl)) = blocks
    k = int(np.nonzero(q['cal'] == 1)[0][0])
    K, Rs, ts = g.cals[1]
    rr, Jpr, Jlr = factors.projection(g.R[q['i'][k]], g.t[q['i'][k]], g.point[q['l'][k]], q['uv'][k], K, Rs, ts, jac=True)
    assert np.allclose(r[k], rr) and np.allclose(Jp[k], Jpr) and np.allclose(Jl[k], Jlr)
