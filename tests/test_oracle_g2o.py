"""Config 1 (BASELINE.json configs[0]): 100-pose SE3 pose graph, ~500 VRO edges, g2o semantics on the CPU
(CGraphG2O::optimizeGraph, g2o/g2o_graph.cpp:241-252: Levenberg, 20 iterations, first vertex fixed, chi2 without
the 1/2).  The reference marks this configuration "plumbing, no GPU": it is covered by the oracle only."""
import numpy as np
from graph_slam_b200 import synth
from oracle import lm, lie, factors as F


def make_pg(seed=1):
    spec = synth.make_config('C1', seed=seed)
    # g2o consumes the same 6x6 information with tangent order [trans, rot] (quirk D.4): reorder the blocks
    info = spec['between_info']
    P = np.zeros((6, 6)); P[:3, 3:] = np.eye(3); P[3:, :3] = np.eye(3)
    info_g2o = P @ info @ P.T
    return spec, lm.PoseGraphG2O(spec['pose_init_R'], spec['pose_init_t'], spec['between_i'], spec['between_j'],
                                 spec['between_R'], spec['between_t'], info_g2o, fixed=(0,))


def test_edge_error_and_oplus_are_consistent():
    rng = np.random.default_rng(0)
    R, t = lie.se3_exp(rng.normal(size=(5, 6)) * 0.5)
    d = rng.normal(size=(5, 6)) * 0.05
    R2, t2 = F.g2o_oplus(R, t, d)
    e = F.g2o_edge_se3(R, t, R2, t2, np.broadcast_to(np.eye(3), (5, 3, 3)), np.zeros((5, 3)))
    assert np.allclose(e, d, atol=1e-12)          # error of X^-1 (X (+) d) is d itself in the [t, q_xyz] chart


def test_g2o_lm_twenty_iterations():
    spec, pg = make_pg()
    chi0 = pg.chi2()
    R0 = pg.R[0].copy(); t0 = pg.t[0].copy()
    pg, rep = lm.optimize_g2o(pg, iterations=20)
    assert rep['chi2'] < 0.3 * chi0
    assert np.array_equal(pg.R[0], R0) and np.array_equal(pg.t[0], t0)      # first vertex fixed (g2o_graph.cpp:90)
    chis = [x['chi2'] for x in rep['trace']]
    assert all(b <= a * (1 + 1e-12) for a, b in zip(chis, chis[1:]))
    assert np.abs(pg.t - spec['truth_t']).max() < np.abs(spec['pose_init_t'] - spec['truth_t']).max()


def test_g2o_and_gtsam_semantics_agree_on_the_optimum():
    """Different error charts and LM schedules, same least-squares problem up to the chart: the optimised
    trajectories agree to well below the measurement noise."""
    from oracle import build
    spec, pg = make_pg(seed=2)
    pg, _ = lm.optimize_g2o(pg, iterations=20)
    g, _ = lm.optimize_gtsam(build.from_spec(spec))
    assert np.abs(pg.t - g.t).max() < 5e-3


def test_closed_form_edge_jacobians_match_central_differences():
    spec, pg = make_pg(seed=3)
    rng = np.random.default_rng(0)
    dR, dt = lie.se3_exp(rng.normal(size=(len(pg.R), 6)) * 0.5)
    pg.R, pg.t = lie.pose_compose(pg.R, pg.t, dR, dt)          # large residuals: the Jacobians are not near the identity
    Ji, Jj = pg.jacobians()
    Ni, Nj = pg.jacobians(eps=1e-6)
    assert np.abs(Ji - Ni).max() < 1e-8 and np.abs(Jj - Nj).max() < 1e-8
