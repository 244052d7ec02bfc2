"""The CUDA path against the ORACLE at the full BASELINE sizes ("configs tested only shrunk").
C2 / C3 (1k poses): the whole LM run against the numpy oracle run here (seconds).  C4 / C5: one full-size LM iteration
against tests/golden/oracle_fullsize.json (written by tests/golden/make_fullsize_golden.py; the oracle needs minutes per
iteration at these sizes): initial error, the accept decision, the error after the iteration and the sampled state."""
import json
import os
import numpy as np
import pytest
from graph_slam_b200 import abi, synth
from test_gpu_parity import check

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'oracle_fullsize.json')


def test_c2_vio_full_size():
    check(synth.make_config('C2', seed=1, scale=1.0), 1e-9, 1e-8)


def test_c3_vio_planes_full_size():
    check(synth.make_config('C3', seed=1, scale=1.0), 1e-9, 1e-8)


@pytest.mark.parametrize('name', ['C4', 'C5'])
def test_one_full_size_iteration_matches_the_oracle(name):
    gold = json.load(open(GOLD))
    if name not in gold:
        pytest.skip('no committed oracle vector for %s' % name)
    k = gold[name]
    spec = synth.make_config(name, seed=1, scale=1.0)
    assert spec['n_poses'] == k['n_poses'] and len(spec['proj_pose']) == k['n_projections']
    ctx = abi.Context(device=0)
    abi.load_spec(ctx, spec)
    e0 = ctx.error()
    assert abs(e0 - k['initial_error']) <= 1e-11 * k['initial_error']
    rep = ctx.optimize(max_iterations=1)
    tr = rep.trace()
    assert len(tr) == 1 and tr[0]['accepted'] and tr[0]['lam'] == k['lambda0'] and abs(rep.lambda_ - k["lambda_after"]) <= 1e-18
    assert abs(rep.final_error - k['error_after']) <= 1e-7 * k['error_after'], (rep.final_error, k['error_after'])
    T = ctx.get_values(abi.T_POSE); ip = np.array(k['pose_idx']); il = np.array(k['point_idx'])
    assert np.abs(T[ip, 9:] - np.array(k['pose_t'])).max() <= 1e-6
    assert np.abs(T[ip, :9] - np.array(k['pose_R'])).max() <= 1e-6
    assert np.abs(ctx.get_values(abi.T_VEC3)[ip] - np.array(k['vel'])).max() <= 1e-5
    assert np.abs(ctx.get_values(abi.T_BIAS)[ip] - np.array(k['bias'])).max() <= 1e-5
    Q = ctx.get_values(abi.T_POINT)
    assert np.abs(Q[il] - np.array(k['point'])).max() <= 1e-5
    assert np.abs(T[:, 9:].sum(0) - np.array(k['pose_t_sum'])).max() <= 1e-6 * len(T)
    assert np.abs(Q.sum(0) - np.array(k['point_sum'])).max() <= 1e-6 * len(Q)
    ctx.close()
