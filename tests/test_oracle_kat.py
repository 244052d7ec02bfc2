"""Oracle pinned against the reference's own known-answer tests:
gtsam/test/testOrientedPlane3.cpp:61-91,111-140,143-164 and gtsam/test/testOrientedPlane3Factor.cpp:37-126."""
import numpy as np
from oracle import lie, factors as F, lm
from oracle.graph import Graph


def ypr(y, p, r):
    return lie.rzryrx(r, p, y)


def test_plane_transform_kat():
    # testOrientedPlane3.cpp:61-70
    R = ypr(-np.pi / 4, 0, 0); t = np.array([2.0, 3.0, 4.0])
    pl = F.plane_from_coeffs(np.array([-1.0, 0, 0, 5]))
    out, Hr, Hp = F.plane_transform(pl, R, t)
    assert np.allclose(out, [-np.sqrt(2) / 2, -np.sqrt(2) / 2, 0.0, 3.0], atol=1e-9)
    # :72-90 Jacobians vs numerical derivative @1e-9 (central differences limit us to ~1e-8)
    eps = 1e-6
    num_r = np.zeros((3, 6)); num_p = np.zeros((3, 3))
    for k in range(6):
        d = np.zeros(6); d[k] = eps
        a = F.plane_transform(pl, *lie.pose_retract(R, t, d), jac=False)
        b = F.plane_transform(pl, *lie.pose_retract(R, t, -d), jac=False)
        num_r[:, k] = (F.plane_local(out, a) - F.plane_local(out, b)) / (2 * eps)
    for k in range(3):
        d = np.zeros(3); d[k] = eps
        a = F.plane_transform(F.plane_retract(pl, d), R, t, jac=False)
        b = F.plane_transform(F.plane_retract(pl, -d), R, t, jac=False)
        num_p[:, k] = (F.plane_local(out, a) - F.plane_local(out, b)) / (2 * eps)
    assert np.allclose(Hr, num_r, atol=1e-8)
    assert np.allclose(Hp, num_p, atol=1e-8)


def test_plane_get_methods():
    # testOrientedPlane3.cpp:34-49
    pl = F.plane_from_coeffs(np.array([-1.0, 0, 0, 5]))
    assert np.allclose(pl, [-1, 0, 0, 5], atol=1e-8)


def test_plane_retract_local_roundtrip():
    # testOrientedPlane3.cpp:111-140 (10 000 random round trips @1e-6)
    rng = np.random.default_rng(7)
    n = 10000
    c = np.concatenate([rng.uniform(-1, 1, size=(n, 3)), rng.uniform(0.01, 10, size=(n, 1))], -1)
    p1 = F.plane_from_coeffs(c)
    v = np.stack([rng.uniform(-np.pi, np.pi, n), rng.uniform(-np.pi, np.pi, n), rng.uniform(-10, 10, n)], -1)
    big = np.linalg.norm(v[:, :2], axis=1) > np.pi       # |rotation| at most pi (Unit3 tangent is 2-d)
    v[big, :2] /= np.pi
    p2 = F.plane_retract(p1, v)
    v12 = F.plane_local(p1, p2)
    assert np.allclose(v12, v, atol=1e-6)
    assert np.allclose(F.plane_retract(p1, v12), p2, atol=1e-6)


def test_plane_error_vector_regression():
    # testOrientedPlane3.cpp:143-164
    p1 = F.plane_from_coeffs(np.array([-1, 0.1, 0.2, 5.0])); p2 = F.plane_from_coeffs(np.array([-1.1, 0.2, 0.3, 5.4]))
    assert np.allclose(F.plane_error_vector(p1, p1), 0, atol=1e-8)
    assert np.allclose(F.plane_error_vector(p1, p2), [-0.0677674148, -0.0760543588, -0.4], atol=1e-5)


def _plane_graph(meas):
    g = Graph()
    g.R = np.eye(3)[None]; g.t = np.zeros((1, 3))
    g.plane = F.plane_from_coeffs(np.array([[-1.0, 0, 0, 3.0]]))
    g.f = dict(prior_pose=dict(i=np.array([0]), R=np.eye(3)[None], t=np.zeros((1, 3)), info=np.eye(6)[None] / 1e-3 ** 2),
               plane=dict(i=np.array([0, 0]), l=np.array([0, 0]), meas=F.plane_from_coeffs(np.array(meas, dtype=float)),
                          info=np.broadcast_to(np.eye(3) / 0.1 ** 2, (2, 3, 3)).copy()))
    return g


def test_plane_factor_translation_kat():
    # testOrientedPlane3Factor.cpp:37-81: one ISAM2 update == one Gauss-Newton step (lambda = 0)
    g = _plane_graph([[-1, 0, 0, 3.0], [-1, 0, 0, 1.0]])
    H, grad, _ = g.normal_equations()
    from oracle.graph import solve_direct
    g1 = g.retract(solve_direct(H, grad, 0.0))
    assert np.allclose(g1.plane[0], [-1, 0, 0, 2.0], atol=1e-9)
    g2, rep = lm.optimize_gtsam(g)
    assert np.allclose(g2.plane[0], [-1, 0, 0, 2.0], atol=1e-6)


def test_plane_factor_rotation_kat():
    # testOrientedPlane3Factor.cpp:84-126
    g = _plane_graph([[-1, 0, 0, 3.0], [0, -1, 0, 3.0]])
    H, grad, _ = g.normal_equations()
    from oracle.graph import solve_direct
    g1 = g.retract(solve_direct(H, grad, 0.0))
    assert np.allclose(g1.plane[0], [-np.sqrt(2) / 2, -np.sqrt(2) / 2, 0, 3.0], atol=1e-9)
