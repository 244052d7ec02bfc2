"""fg_update_incremental -- the entry point the reference's two drivers call every frame (CGraphGT::optimizeGraphIncremental,
gtsam/gtsam_graph.cpp:1768-1776; test_vro_imu_graph.cpp:344, test_ba_imu_graph.cpp:427) -- against the oracle's restatement of
the same ISAM2 semantics (oracle/incremental.py), frame by frame, and against the batch optimum at the end."""
import numpy as np
import pytest
from graph_slam_b200 import abi, synth
from oracle import build, lm, lie, incremental as oinc

pytestmark = pytest.mark.gpu


def prefix(spec, k):
    """The graph after frame k-1 arrived: the first k poses and every factor whose variables are among them."""
    out = dict(spec)
    out['n_poses'] = k
    for name in ('pose_init_R', 'pose_init_t', 'vel_init', 'bias_init'):
        if name in spec:
            out[name] = spec[name][:k]
    if 'imu_samples' in spec:
        out['imu_samples'] = spec['imu_samples'][:k - 1]
    if 'between_i' in spec:
        m = np.maximum(spec['between_i'], spec['between_j']) < k
        for name in ('between_i', 'between_j', 'between_R', 'between_t', 'between_info'):
            out[name] = spec[name][m]
    if 'plane_init' in spec:
        m = spec['plane_obs_pose'] < k
        seen = np.zeros(len(spec['plane_init']), dtype=bool); seen[spec['plane_obs_plane'][m]] = True
        assert np.all(np.diff(seen.astype(int)) <= 0), 'plane landmarks must appear in index order'
        for name in ('plane_obs_pose', 'plane_obs_plane', 'plane_meas', 'plane_cov'):
            out[name] = spec[name][m]
        out['plane_init'] = spec['plane_init'][:int(seen.sum())]
    return out


def drive(spec, k0, noise=0.0):
    """Frames k0..P-1 arrive one at a time; returns the device context and the oracle smoother after the last one."""
    P = spec['n_poses']
    if noise:
        rng = np.random.default_rng(5)
        dR, dt = lie.se3_exp(rng.normal(size=(P, 6)) * noise); dR[0] = np.eye(3); dt[0] = 0
        spec = dict(spec)
        spec['pose_init_R'], spec['pose_init_t'] = lie.pose_compose(spec['pose_init_R'], spec['pose_init_t'], dR, dt)
    X = abi.symbols('x', np.arange(P)); V = abi.symbols('v', np.arange(P)); B = abi.symbols('b', np.arange(P))
    ctx = abi.Context(device=0)
    sm = oinc.IncrementalSmoother()
    first = prefix(spec, k0)
    pims = abi.load_spec(ctx, first)
    has_imu = 'imu_samples' in spec
    T = abi.pose12(spec['pose_init_R'], spec['pose_init_t'])
    Tm = abi.pose12(spec['between_R'], spec['between_t']) if 'between_i' in spec else None
    S = spec['imu_samples'].shape[1] if has_imu else 0
    reports = []
    for k in range(k0, P + 1):
        if k > k0:
            j = k - 1                                     # the new frame
            ctx.add_pose(int(X[j]), T[j])
            if has_imu:
                ctx.add_vec3(int(V[j]), spec['vel_init'][j]); ctx.add_bias(int(B[j]), spec['bias_init'][j])
                pim = ctx.preintegrate(np.array([0, S]), spec['imu_samples'][j - 1], spec['imu_dt'], abi.vn100_imu_params(), np.zeros((1, 6)))
                ctx.add_imu([X[j - 1], V[j - 1], X[j], V[j], B[j - 1], B[j]], pim[0])
            if Tm is not None:
                for n in np.nonzero(np.maximum(spec['between_i'], spec['between_j']) == j)[0]:
                    ctx.add_between(int(X[spec['between_i'][n]]), int(X[spec['between_j'][n]]), Tm[n], spec['between_info'][n])
        rep = ctx.update_incremental()
        orep = sm.update(build.from_spec(prefix(spec, k)))
        assert rep.n_variables == (3 if has_imu else 1) * k and rep.n_new_variables == ((3 if has_imu else 1) * (k if k == k0 else 1))
        assert rep.n_relinearized == orep['n_relinearized'], (k, rep.n_relinearized, orep['n_relinearized'])
        assert abs(rep.error_before - orep['error_before']) <= 1e-9 * max(orep['error_before'], 1e-12), (k, rep.error_before, orep['error_before'])
        assert abs(rep.error_after - orep['error_after']) <= 1e-8 * max(orep['error_after'], 1e-9), (k, rep.error_after, orep['error_after'])
        Tk = ctx.get_values(abi.T_POSE)
        assert np.abs(Tk[:, 9:] - sm.est.t).max() <= 1e-8 and np.abs(Tk[:, :9].reshape(-1, 3, 3) - sm.est.R).max() <= 1e-8
        reports.append(rep)
    return ctx, sm, spec, reports


@pytest.mark.parametrize('name,scale,k0,noise', [('C1', 0.4, 2, 0.0), ('C2', 0.05, 2, 0.0), ('C2', 0.04, 3, 0.08)])
def test_frame_by_frame_matches_oracle_and_converges_to_batch(name, scale, k0, noise):
    ctx, sm, spec, reports = drive(synth.make_config(name, seed=2, scale=scale), k0, noise)
    if noise:
        assert max(r.n_relinearized for r in reports[1:]) > 0, 'the gating was never exercised'
    est = ctx.get_values(abi.T_POSE)
    # a few more updates without new factors settle the estimate (ISAM2: repeated update() calls)
    for _ in range(4):
        ctx.update_incremental()
        sm.update(build.from_spec(spec))
    est = ctx.get_values(abi.T_POSE)
    assert np.abs(est[:, 9:] - sm.est.t).max() <= 1e-8
    # batch LM from the same initial values: the incremental estimate is within the relinearisation threshold of it
    ref = abi.Context(device=0)
    abi.load_spec(ref, spec)
    ref_rep = ref.optimize()
    Tb = ref.get_values(abi.T_POSE)
    assert np.abs(est[:, 9:] - Tb[:, 9:]).max() < 0.1 and np.abs(est[:, :9] - Tb[:, :9]).max() < 0.1
    # ... and the batch entry point continues from the estimate (the reference copies calculateEstimate() into mp_node_values)
    e_est = ctx.error()
    rep = ctx.optimize()
    assert abs(rep.initial_error - e_est) <= 1e-12 * max(e_est, 1e-12)
    assert rep.final_error <= e_est * (1 + 1e-12) and abs(rep.final_error - ref_rep.final_error) <= 1e-4 * max(ref_rep.final_error, 1e-9)
    ref.close(); ctx.close()


def test_indeterminate_update_reports_and_keeps_the_estimate():
    """A pose with no factor at all makes the undamped system singular: GTSAM throws IndeterminantLinearSystemException."""
    ctx = abi.Context(device=0)
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0.0])
    ctx.add_pose(abi.symbol('x', 0), I); ctx.add_pose(abi.symbol('x', 1), I)
    ctx.add_prior_pose(abi.symbol('x', 0), I, np.eye(6))
    with pytest.raises(abi.FgError) as e:
        ctx.update_incremental()
    assert e.value.code == -5
    assert np.allclose(ctx.get_value(abi.symbol('x', 1)), I)
    ctx.close()
