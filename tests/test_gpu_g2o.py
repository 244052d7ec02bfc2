"""BASELINE config 1 on the device: the 100-pose / 500-edge SE3 pose graph with g2o semantics (CGraphG2O, g2o/g2o_graph.cpp:
EdgeSE3 error [t, q_xyz], VertexSE3::oplus, first vertex fixed, Levenberg with g2o's gain-ratio rule, 10 x optimize(2))
through the C ABI against the oracle's restatement (oracle/lm.py: PoseGraphG2O, optimize_g2o_calls)."""
import numpy as np
import pytest
from graph_slam_b200 import abi, synth
from oracle import lm, lie

pytestmark = pytest.mark.gpu


def load(spec, ctx):
    P = spec['n_poses']
    X = abi.symbols('x', np.arange(P))
    T = abi.pose12(spec['pose_init_R'], spec['pose_init_t'])
    for i in range(P):
        ctx.add_pose(int(X[i]), T[i])
    ctx.set_fixed(int(X[0]))                                   # CGraphG2O::firstNode: reference_pose->setFixed(true)
    Pm = np.zeros((6, 6)); Pm[:3, 3:] = np.eye(3); Pm[3:, :3] = np.eye(3)
    info = Pm @ spec['between_info'] @ Pm.T                    # the synthetic information is [rot, trans]; g2o's order is [trans, rot]
    Tm = abi.pose12(spec['between_R'], spec['between_t'])
    for n in range(len(spec['between_i'])):
        ctx.add_g2o_edge(int(X[spec['between_i'][n]]), int(X[spec['between_j'][n]]), Tm[n], info[n])
    return lm.PoseGraphG2O(spec['pose_init_R'], spec['pose_init_t'], spec['between_i'], spec['between_j'], spec['between_R'], spec['between_t'],
                           info, fixed=(0,))


@pytest.mark.parametrize('seed,noise', [(1, 0.0), (2, 0.0), (3, 0.6)])
def test_c1_g2o_matches_oracle(seed, noise):
    spec = synth.make_config('C1', seed=seed)
    if noise:                                                  # a bad start makes g2o's LM reject trials and raise lambda
        rng = np.random.default_rng(seed)
        dR, dt = lie.se3_exp(rng.normal(size=(spec['n_poses'], 6)) * noise); dR[0] = np.eye(3); dt[0] = 0
        spec['pose_init_R'], spec['pose_init_t'] = lie.pose_compose(spec['pose_init_R'], spec['pose_init_t'], dR, dt)
    ctx = abi.Context(device=0)
    pg = load(spec, ctx)
    chi0 = pg.chi2()
    assert abs(ctx.g2o_chi2() - chi0) <= 1e-11 * chi0
    rep = ctx.optimize_g2o()
    pg, orep = lm.optimize_g2o_calls(pg)
    assert 20 <= rep.iterations <= 21 and 20 <= orep['iterations'] <= 21      # a call that terminates after one iteration shifts the count (i += currIt)
    assert abs(rep.initial_chi2 - chi0) <= 1e-11 * chi0
    tr = orep['trace']
    # iteration by iteration while the optimisation is still making progress; once chi2 stalls at the optimum the accept /
    # reject decisions hang on the last bits of chi2 and are not comparable
    prev, n_cmp = chi0, 0
    for k, t in enumerate(tr):
        if prev - t['chi2'] <= 1e-7 * prev:
            break
        assert rep.trace_trials[k] == t['trials'] and abs(rep.trace_lambda[k] - t['lam']) <= 1e-5 * t['lam']
        assert abs(rep.trace_chi2[k] - t['chi2']) <= 1e-9 * t['chi2']
        prev, n_cmp = t['chi2'], n_cmp + 1
    assert n_cmp >= 2
    assert abs(rep.final_chi2 - orep['chi2']) <= 1e-9 * orep['chi2']
    T = ctx.get_values(abi.T_POSE)
    assert np.abs(T[:, 9:] - pg.t).max() <= 1e-7 and np.abs(T[:, :9].reshape(-1, 3, 3) - pg.R).max() <= 1e-7
    assert np.array_equal(T[0], abi.pose12(spec['pose_init_R'], spec['pose_init_t'])[0])     # the fixed vertex did not move
    assert abs(ctx.g2o_chi2() - rep.final_chi2) <= 1e-10 * rep.final_chi2
    ctx.close()


def test_g2o_and_gtsam_factors_do_not_mix():
    ctx = abi.Context(device=0)
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0])
    ctx.add_pose(abi.symbol('x', 0), I); ctx.add_pose(abi.symbol('x', 1), I)
    ctx.add_g2o_edge(abi.symbol('x', 0), abi.symbol('x', 1), I, np.eye(6))
    ctx.add_between(abi.symbol('x', 0), abi.symbol('x', 1), I, np.eye(6))
    with pytest.raises(abi.FgError) as e:
        ctx.optimize_g2o()
    assert e.value.code == -1
    ctx.close()
