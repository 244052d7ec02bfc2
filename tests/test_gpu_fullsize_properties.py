"""Full-size checks (BASELINE.json configs[3], C4: 2k poses, 100k landmarks, ~2M projections) through size-independent
properties, because the numpy oracle cannot finish this size in seconds:
  * LM accepted steps decrease graph.error monotonically and the report agrees with fg_error;
  * idempotence: optimising the converged state again changes nothing beyond the LM tolerances;
  * relabelling invariance: inserting the landmarks in a shuffled order (different internal numbering, different
    Schur tile tables) gives the same optimum;
  * the optimum is closer to the generator's ground truth than the initial guess."""
import numpy as np
import pytest
from graph_slam_b200 import abi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def c4():
    return synth.make_config('C4', seed=1)


def solve(spec):
    ctx = abi.Context(device=0)
    abi.load_spec(ctx, spec)
    e0 = ctx.error()
    rep = ctx.optimize()
    return ctx, e0, rep


def test_c4_full_size_properties(c4):
    ctx, e0, rep = solve(c4)
    tr = rep.trace()
    assert rep.n_projections == len(c4['proj_pose']) and rep.n_landmarks == len(c4['point_init'])
    assert abs(rep.initial_error - e0) <= 1e-12 * e0
    acc = [t for t in tr if t['accepted']]
    assert len(acc) >= 2
    for t in acc:
        assert t['new_err'] < t['err']
    assert abs(ctx.error() - rep.final_error) <= 1e-10 * rep.final_error
    assert rep.final_error < 0.25 * e0
    # closer to the truth than the initial guess
    T = ctx.get_values(abi.T_POSE)
    err_opt = np.abs(T[:, 9:] - c4['truth_t']).max()
    err_init = np.abs(c4['pose_init_t'] - c4['truth_t']).max()
    assert err_opt < 0.5 * err_init
    # idempotence
    rep2 = ctx.optimize()
    assert rep2.iterations <= 2
    assert abs(rep2.final_error - rep.final_error) <= 1e-5 * rep.final_error
    T2 = ctx.get_values(abi.T_POSE)
    assert np.abs(T2 - T).max() < 1e-4
    ctx.close()

    # relabelling invariance
    rng = np.random.default_rng(5)
    L = len(c4['point_init'])
    perm = rng.permutation(L)                 # new index -> old index
    inv = np.empty(L, dtype=np.int64); inv[perm] = np.arange(L)
    spec2 = dict(c4)
    spec2['point_init'] = c4['point_init'][perm]
    spec2['proj_point'] = inv[c4['proj_point']].astype(np.int32)
    order = rng.permutation(len(c4['proj_pose']))
    for k in ('proj_pose', 'proj_point', 'proj_uv'):
        spec2[k] = spec2[k][order]
    ctx2, e02, rep_b = solve(spec2)
    assert abs(e02 - e0) <= 1e-11 * e0
    assert rep_b.iterations == rep.iterations
    assert abs(rep_b.final_error - rep.final_error) <= 1e-9 * rep.final_error
    Tb = ctx2.get_values(abi.T_POSE)
    assert np.abs(Tb - T).max() < 1e-7
    ctx2.close()
