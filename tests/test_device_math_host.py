"""The formulas the CUDA kernels execute (graph_slam_b200/csrc/fg_math.cuh, fg_factors.cuh), compiled for the
host by the test-only shim tests/hostmath, compared with the numpy oracle on seeded inputs.  Runs without a GPU."""
import ctypes as C
import numpy as np
import pytest
from oracle import lie, factors as F, imu as oimu

DP = C.POINTER(C.c_double)


def p(a):
    return a.ctypes.data_as(DP)


def arr(*shape):
    return np.zeros(shape, dtype=np.float64)


def T12(R, t):
    return np.concatenate([R.ravel(), t.ravel()])


def rand_pose(rng, scale=0.7):
    return lie.se3_exp(rng.normal(size=6) * scale)


@pytest.fixture(scope='module')
def rng():
    return np.random.default_rng(123)


def test_lie(hostmath, rng):
    for _ in range(200):
        w = rng.normal(size=3) * rng.choice([1e-9, 1e-3, 1.0, 2.5])
        R = arr(9); hostmath.hm_so3_exp(p(w), p(R))
        assert np.allclose(R.reshape(3, 3), lie.so3_exp(w), atol=1e-14)
        w2 = arr(3); hostmath.hm_so3_log(p(R), p(w2))
        assert np.allclose(w2, lie.so3_log(lie.so3_exp(w)), atol=1e-12)
        J = arr(9); hostmath.hm_jr(p(w), p(J))
        assert np.allclose(J.reshape(3, 3), lie.so3_jr(w), atol=1e-13)
        hostmath.hm_jr_inv(p(w), p(J))
        assert np.allclose(J.reshape(3, 3), lie.so3_jr_inv(w), atol=1e-11)
        xi = np.concatenate([w, rng.normal(size=3)])
        T = arr(12); hostmath.hm_se3_exp(p(xi), p(T))
        Ro, to = lie.se3_exp(xi)
        assert np.allclose(T, T12(Ro, to), atol=1e-13)
        x2 = arr(6); hostmath.hm_se3_log(p(T), p(x2))
        assert np.allclose(x2, lie.se3_log(Ro, to), atol=1e-11)


def test_between_and_prior(hostmath, rng):
    for _ in range(100):
        R1, t1 = rand_pose(rng); R2, t2 = rand_pose(rng)
        Rz, tz = lie.pose_between(R1, t1, R2, t2)
        Rz, tz = lie.pose_retract(Rz, tz, rng.normal(size=6) * 0.05)
        r = arr(6); J = arr(36)
        hostmath.hm_between(p(T12(R1, t1)), p(T12(R2, t2)), p(T12(Rz, tz)), p(r), p(J))
        ro, H1, _ = F.between_pose(R1, t1, R2, t2, Rz, tz)
        assert np.allclose(r, ro, atol=1e-12)
        assert np.allclose(J.reshape(6, 6), H1, atol=1e-12)
        hostmath.hm_prior_pose(p(T12(R1, t1)), p(T12(R2, t2)), p(r))
        assert np.allclose(r, F.prior_pose(R1, t1, R2, t2, jac=False), atol=1e-12)


def test_projection(hostmath, rng):
    K = np.array([250.5773, 250.5773, 0.3, 90.0, 70.0, -0.8466, 0.5370, 1e-3, -2e-3])
    Rs = lie.rzryrx(np.pi / 2, 0, np.pi / 2); ts = np.array([0.02, -0.01, 0.03])
    for k in range(200):
        R, t = rand_pose(rng)
        Rc, tc = lie.pose_compose(R, t, Rs, ts)
        pc = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.5, 6)])
        if k % 10 == 0:
            pc[2] = -abs(pc[2])          # cheirality branch
        pw = Rc @ pc + tc
        uv = rng.normal(size=2) * 50 + 80
        r = arr(2); Jp = arr(12); Jl = arr(6)
        hostmath.hm_projection(p(T12(R, t)), p(pw), p(uv), p(K), p(T12(Rs, ts)), p(r), p(Jp), p(Jl))
        ro, Jpo, Jlo = F.projection(R, t, pw, uv, K, Rs, ts)
        assert np.allclose(r, ro, rtol=1e-12, atol=1e-10)
        assert np.allclose(Jp.reshape(2, 6), Jpo, rtol=1e-11, atol=1e-9)
        assert np.allclose(Jl.reshape(2, 3), Jlo, rtol=1e-11, atol=1e-9)


def test_plane(hostmath, rng):
    for _ in range(200):
        R, t = rand_pose(rng)
        pl = F.plane_from_coeffs(np.concatenate([rng.normal(size=3), rng.uniform(0.5, 8, size=1)]))
        z = F.plane_from_coeffs(np.concatenate([rng.normal(size=3), rng.uniform(0.5, 8, size=1)]))
        r = arr(3); Hr = arr(18); Hp = arr(9)
        hostmath.hm_plane(p(T12(R, t)), p(pl), p(z), p(r), p(Hr), p(Hp))
        ro, Hro, Hpo = F.plane_factor(R, t, pl, z)
        assert np.allclose(r, ro, atol=1e-12)
        assert np.allclose(Hr.reshape(3, 6), Hro, atol=1e-12)
        assert np.allclose(Hp.reshape(3, 3), Hpo, atol=1e-12)
        v = rng.normal(size=3) * 0.3
        out = arr(4); hostmath.hm_plane_retract(p(pl), p(v), p(out))
        assert np.allclose(out, F.plane_retract(pl, v), atol=1e-13)


def test_imu_factor(hostmath, rng):
    par = oimu.vn100_params()
    S = 20
    for _ in range(20):
        samples = np.concatenate([rng.normal(size=(1, S, 3)) * 0.3, rng.normal(size=(1, S, 3)) * 0.5 + np.array([0, 0, -9.7])], -1)
        bh = rng.normal(size=(1, 6)) * 0.01
        pim = oimu.preintegrate(samples, 0.005, par, bh)
        Ri, ti = rand_pose(rng); vi = rng.normal(size=3); bi = bh[0] + rng.normal(size=6) * 0.01
        Rj, tj, vj = oimu.predict({k: (v[0] if isinstance(v, np.ndarray) and v.ndim > 1 else v) for k, v in pim.items()} | {'dt': pim['dt'][0]}, Ri, ti, vi, bi)
        Rj, tj = lie.pose_retract(Rj, tj, rng.normal(size=6) * 0.01); vj = vj + rng.normal(size=3) * 0.01
        bj = bi + rng.normal(size=6) * 1e-3
        pim1 = dict(dt=pim['dt'][0], preint=pim['preint'][0], Hba=pim['Hba'][0], Hbg=pim['Hbg'][0], bias_hat=pim['bias_hat'][0], gravity=pim['gravity'])
        ro, Js = F.imu_combined(Ri, ti, vi, Rj, tj, vj, bi, bj, pim1)
        Jo = np.concatenate(Js, -1)
        r = arr(15); J = arr(450)
        hostmath.hm_imu(p(T12(Ri, ti)), p(vi), p(T12(Rj, tj)), p(vj), p(bi), p(bj), C.c_double(float(pim1['dt'])),
                        p(np.ascontiguousarray(pim1['preint'])), p(np.ascontiguousarray(pim1['Hba'])), p(np.ascontiguousarray(pim1['Hbg'])),
                        p(np.ascontiguousarray(pim1['bias_hat'])), p(np.ascontiguousarray(pim1['gravity'])), p(r), p(J))
        assert np.allclose(r, ro, atol=1e-12)
        assert np.allclose(J.reshape(15, 30), Jo, atol=1e-11)


def test_d_jr_c(hostmath, rng):
    for _ in range(100):
        th = rng.normal(size=3) * rng.choice([1e-7, 0.1, 1.5]); c = rng.normal(size=3)
        D = arr(9); hostmath.hm_d_jr_c(p(th), p(c), p(D))
        assert np.allclose(D.reshape(3, 3), oimu._d_jr_c(th, c), atol=1e-12)


def test_g2o_edge_and_oplus(hostmath, rng):
    """g2o::EdgeSE3 / VertexSE3::oplus as the device evaluates them: the error against the oracle, the analytic
    Jacobians against central differences of the oracle's error map through the oracle's oplus."""
    for _ in range(50):
        R1, t1 = rand_pose(rng); R2, t2 = rand_pose(rng)
        Rm, tm = lie.pose_compose(*lie.pose_between(R1, t1, R2, t2), *lie.se3_exp(rng.normal(size=6) * rng.choice([0.02, 0.4, 2.0])))
        e, J1, J2 = arr(6), arr(36), arr(36)
        hostmath.hm_g2o_edge(p(T12(R1, t1)), p(T12(R2, t2)), p(T12(Rm, tm)), p(e), p(J1), p(J2))
        ref = F.g2o_edge_se3(R1, t1, R2, t2, Rm, tm)
        assert np.allclose(e, ref, atol=1e-13)
        eps = 1e-6
        N1, N2 = np.zeros((6, 6)), np.zeros((6, 6))
        for k in range(6):
            d = np.zeros(6); d[k] = eps
            N1[:, k] = (F.g2o_edge_se3(*F.g2o_oplus(R1, t1, d), R2, t2, Rm, tm) - F.g2o_edge_se3(*F.g2o_oplus(R1, t1, -d), R2, t2, Rm, tm)) / (2 * eps)
            N2[:, k] = (F.g2o_edge_se3(R1, t1, *F.g2o_oplus(R2, t2, d), Rm, tm) - F.g2o_edge_se3(R1, t1, *F.g2o_oplus(R2, t2, -d), Rm, tm)) / (2 * eps)
        assert np.allclose(J1.reshape(6, 6), N1, atol=1e-8) and np.allclose(J2.reshape(6, 6), N2, atol=1e-8)
        d = rng.normal(size=6) * 0.2
        To = arr(12); hostmath.hm_g2o_oplus(p(T12(R1, t1)), p(d), p(To))
        Ro, to = F.g2o_oplus(R1, t1, d)
        assert np.allclose(To[:9].reshape(3, 3), Ro, atol=1e-14) and np.allclose(To[9:], to, atol=1e-14)


@pytest.mark.parametrize('chart', [lie.EXPMAP, lie.FIRST_ORDER_EXPMAP, lie.FIRST_ORDER_CAYLEY])
def test_pose_charts(hostmath, rng, chart):
    """Pose3::ChartAtOrigin::{Retract, Local} under the three chart options as the device evaluates them, against the oracle;
    Local inverts Retract; every chart agrees with EXPMAP to first order (so the Jacobians are chart independent)."""
    for _ in range(100):
        R, t = rand_pose(rng)
        xi = rng.normal(size=6) * rng.choice([1e-6, 0.05, 0.8])
        To = arr(12); hostmath.hm_chart_retract(p(T12(R, t)), p(xi), C.c_int(chart), p(To))
        Ro, to = lie.pose_retract(R, t, xi, chart)
        assert np.allclose(To[:9].reshape(3, 3), Ro, atol=1e-14) and np.allclose(To[9:], to, atol=1e-14)
        R0, t0 = lie.chart_retract0(xi, chart)
        back = arr(6); hostmath.hm_chart_local0(p(T12(R0, t0)), C.c_int(chart), p(back))
        assert np.allclose(back, xi, atol=1e-12) and np.allclose(back, lie.chart_local0(R0, t0, chart), atol=1e-12)
        Re, te = lie.se3_exp(xi)
        assert np.abs(R0 - Re).max() <= 2 * np.linalg.norm(xi) ** 2 + 1e-15 and np.abs(t0 - te).max() <= 2 * np.linalg.norm(xi) ** 2 + 1e-15
        R2, t2 = rand_pose(rng); Rm, tm = rand_pose(rng, 0.3)
        r, J1 = arr(6), arr(36)
        hostmath.hm_between_chart(p(T12(R, t)), p(T12(R2, t2)), p(T12(Rm, tm)), C.c_int(chart), p(r), p(J1))
        ro, H1, H2 = F.between_pose(R, t, R2, t2, Rm, tm, chart=chart)
        assert np.allclose(r, ro, atol=1e-12) and np.allclose(J1.reshape(6, 6), H1, atol=1e-13)


def test_device_lie_math_against_scipy_directly(hostmath, rng):
    """The device formulas (fg_math.cuh compiled for the host) against scipy, with no oracle in between: se3_exp is the matrix
    exponential of the 4x4 twist [omega, v], se3_log its logarithm, so3_exp / so3_log are scipy's rotation vectors, and the
    Between residual is the logarithm of Z^-1 X1^-1 X2 on homogeneous matrices."""
    from scipy.linalg import expm, logm
    from scipy.spatial.transform import Rotation

    def hat6(xi):
        w, v = xi[:3], xi[3:]
        T = np.zeros((4, 4))
        T[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
        T[:3, 3] = v
        return T

    def vee6(T):
        return np.array([T[2, 1], T[0, 2], T[1, 0], T[0, 3], T[1, 3], T[2, 3]])

    def homog12(T12v):
        H = np.eye(4); H[:3, :3] = T12v[:9].reshape(3, 3); H[:3, 3] = T12v[9:]
        return H

    for _ in range(50):
        xi = rng.normal(size=6) * rng.choice([1e-3, 0.3, 1.0, 2.0])
        if np.linalg.norm(xi[:3]) > 3.0:
            xi *= 3.0 / np.linalg.norm(xi[:3])
        T = arr(12); hostmath.hm_se3_exp(p(xi), p(T))
        E = expm(hat6(xi))
        assert np.abs(homog12(T) - E).max() < 1e-12
        R = arr(9); hostmath.hm_so3_exp(p(xi[:3].copy()), p(R))
        assert np.abs(R.reshape(3, 3) - Rotation.from_rotvec(xi[:3]).as_matrix()).max() < 1e-13
        x2 = arr(6); hostmath.hm_se3_log(p(T), p(x2))
        assert np.abs(x2 - vee6(np.real(logm(E)))).max() < 1e-8
        w2 = arr(3); hostmath.hm_so3_log(p(R), p(w2))
        assert np.abs(w2 - Rotation.from_matrix(R.reshape(3, 3)).as_rotvec()).max() < 1e-10
    for _ in range(20):
        X1 = arr(12); X2 = arr(12); Z = arr(12)
        for X, s in ((X1, 1.0), (X2, 1.0), (Z, 0.4)):
            xi = rng.normal(size=6) * s
            hostmath.hm_se3_exp(p(xi), p(X))
        r = arr(6); J = arr(36)
        hostmath.hm_between(p(X1), p(X2), p(Z), p(r), p(J))
        D = np.linalg.inv(homog12(Z)) @ np.linalg.inv(homog12(X1)) @ homog12(X2)
        assert np.abs(r - vee6(np.real(logm(D)))).max() < 1e-9
