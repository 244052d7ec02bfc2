"""Committed fixtures of tests/golden/ (written by tests/golden/make_golden.py):
  * reference_kat.json -- the reference's own known-answer vectors (vendored GTSAM unit tests): the oracle must hit them;
  * oracle_lm.json     -- oracle regression vectors on seeded graphs: the oracle must still reproduce them (CPU) and the
                          CUDA path, through the C ABI, must match them without the oracle in the loop (GPU)."""
import json
import os
import numpy as np
import pytest
from graph_slam_b200 import abi, synth
from oracle import build, lm, lie, factors as F

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
KAT = json.load(open(os.path.join(HERE, 'reference_kat.json')))
LM = json.load(open(os.path.join(HERE, 'oracle_lm.json')))


def test_reference_plane_kats():
    k = KAT['plane_transform']
    y, p, r = k['pose_ypr']
    out = F.plane_transform(F.plane_from_coeffs(np.array(k['plane'])), lie.rzryrx(r, p, y), np.array(k['pose_t']), jac=False)
    assert np.allclose(out, k['expected'], atol=k['tol'])
    k = KAT['plane_error_vector']
    e = F.plane_error_vector(F.plane_from_coeffs(np.array(k['plane1'])), F.plane_from_coeffs(np.array(k['plane2'])))
    assert np.allclose(e, k['expected'], atol=k['tol'])


@pytest.mark.parametrize('case', LM[:3], ids=lambda c: '%s-s%d' % (c['config'], c['seed']))
def test_oracle_reproduces_its_regression_vectors(case):
    spec = synth.make_config(case['config'], seed=case['seed'], scale=case['scale'])
    g0 = build.from_spec(spec)
    assert abs(g0.error() - case['initial_error']) <= 1e-10 * case['initial_error']
    g1, rep = lm.optimize_gtsam(g0, solver=case['solver'])
    assert rep['iterations'] == case['iterations'] and [bool(t['accepted']) for t in rep['trace']] == case['trace_accepted']
    assert abs(rep['error'] - case['final_error']) <= 1e-9 * case['final_error']
    assert np.allclose(g1.t[-1], case['pose_t_last'], atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize('case', LM, ids=lambda c: '%s-s%d' % (c['config'], c['seed']))
def test_cuda_path_matches_golden_without_the_oracle(case):
    spec = synth.make_config(case['config'], seed=case['seed'], scale=case['scale'])
    ctx = abi.Context(device=0)
    abi.load_spec(ctx, spec)
    tol_chi2, tol_pose = (1e-7, 1e-6) if case['solver'] == 'schur' else (1e-9, 1e-8)
    assert abs(ctx.error() - case['initial_error']) <= 1e-11 * case['initial_error']
    rep = ctx.optimize()
    tr = rep.trace()
    assert rep.iterations == case['iterations'] and [bool(t['accepted']) for t in tr] == case['trace_accepted']
    assert np.allclose([t['lam'] for t in tr], case['trace_lambda'], rtol=1e-12)
    assert abs(rep.final_error - case['final_error']) <= tol_chi2 * case['final_error']
    T = ctx.get_values(abi.T_POSE)
    assert np.abs(T[-1, 9:] - case['pose_t_last']).max() <= tol_pose
    assert np.abs(T[:, 9:].sum(0) - case['pose_t_sum']).max() <= tol_pose * len(T)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['plane_factor_lm_1', 'plane_factor_lm_2'])
def test_cuda_path_hits_the_reference_plane_factor_kats(name):
    """gtsam/test/testOrientedPlane3Factor.cpp:37-126 through the C ABI: a pose prior (sigma 1e-3), one plane landmark
    initialised at (-1, 0, 0, 3) and two OrientedPlane3Factor measurements (sigma 0.1) -> the optimised landmark."""
    k = KAT[name]
    ctx = abi.Context(device=0)
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0])
    x0, l0 = abi.symbol('x', 0), abi.symbol('l', 0)
    ctx.add_pose(x0, I)
    ctx.add_prior_pose(x0, I, np.eye(6) / k['pose_prior_sigma'] ** 2)
    ctx.add_plane(l0, np.array([-1.0, 0.0, 0.0, 3.0]))
    for z in k['measurements']:
        ctx.add_plane_factor(x0, l0, np.array(z), np.eye(3) * k['sigma'] ** 2)
    ctx.optimize()
    got = ctx.get_value(l0)
    assert np.allclose(got, k['expected'], atol=1e-6), got
    ctx.close()
