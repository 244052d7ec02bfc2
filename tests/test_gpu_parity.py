"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.
Tolerances (BASELINE.md section 3 / SURVEY 8d): identical accept/reject sequence; final chi2 rel <= 1e-9 (C1-C3),
<= 1e-7 (C4-C5 shapes); poses <= 1e-8 rad / 1e-8 m (C1-C3), <= 1e-6 (C4)."""
import numpy as np
import pytest
from graph_slam_b200 import abi, synth
from oracle import build, lm, lie, imu as oimu

pytestmark = pytest.mark.gpu


def rot_angle(Ra, Rb):
    return np.linalg.norm(lie.so3_log(np.swapaxes(Ra, -1, -2) @ Rb), axis=-1)


def run_both(spec, solver='direct', chart=0, **lm_kw):
    ctx = abi.Context(device=0)
    abi.load_spec(ctx, spec)
    g0 = build.from_spec(spec)
    if chart:
        ctx.set_pose_chart(chart); g0.chart = chart
    e_dev, e_orc = ctx.error(), g0.error()
    rep = ctx.optimize(**lm_kw)
    g1, orep = lm.optimize_gtsam(g0, lm.LMParams(**lm_kw), solver=solver)
    return ctx, rep, g0, g1, orep, e_dev, e_orc


def check(spec, tol_chi2, tol_pose, solver='direct', chart=0, **lm_kw):
    ctx, rep, g0, g1, orep, e_dev, e_orc = run_both(spec, solver, chart, **lm_kw)
    assert abs(e_dev - e_orc) <= 1e-11 * e_orc, (e_dev, e_orc)
    dt, ot = rep.trace(), orep['trace']
    assert [t['accepted'] for t in dt] == [t['accepted'] for t in ot]
    assert np.allclose([t['lam'] for t in dt], [t['lam'] for t in ot], rtol=1e-12)
    assert rep.iterations == orep['iterations']
    assert abs(rep.final_error - orep['error']) <= tol_chi2 * orep['error'], (rep.final_error, orep['error'])
    T = ctx.get_values(abi.T_POSE)
    R = T[:, :9].reshape(-1, 3, 3); t = T[:, 9:]
    assert rot_angle(R, g1.R).max() <= tol_pose
    assert np.abs(t - g1.t).max() <= tol_pose
    if len(g1.vel):
        assert np.abs(ctx.get_values(abi.T_VEC3) - g1.vel).max() <= 10 * tol_pose
        assert np.abs(ctx.get_values(abi.T_BIAS) - g1.bias).max() <= 10 * tol_pose
    if len(g1.point):
        assert np.abs(ctx.get_values(abi.T_POINT) - g1.point).max() <= 10 * tol_pose
    if len(g1.plane):
        assert np.abs(ctx.get_values(abi.T_PLANE) - g1.plane).max() <= 10 * tol_pose
    # the optimiser moved the estimate (not a no-op) and graph.error agrees with the report
    assert rep.final_error < rep.initial_error
    assert abs(ctx.error() - rep.final_error) <= 1e-10 * rep.final_error
    ctx.close()
    return rep


def test_preintegration_matches_oracle():
    spec = synth.make_config('C2', seed=3, scale=0.06)
    P = spec['n_poses']; S = spec['imu_samples'].shape[1]
    rng = np.random.default_rng(0)
    bh = rng.normal(size=(P - 1, 6)) * 0.01
    ctx = abi.Context(device=0)
    pims = ctx.preintegrate(np.arange(P) * S, spec['imu_samples'].reshape(-1, 6), spec['imu_dt'], abi.vn100_imu_params(), bh)
    ref = oimu.preintegrate(spec['imu_samples'], spec['imu_dt'], oimu.vn100_params(), bh)
    for i in range(P - 1):
        assert abs(pims[i].dt - ref['dt'][i]) < 1e-15
        assert np.allclose(np.array(pims[i].preint), ref['preint'][i], rtol=1e-12, atol=1e-14)
        assert np.allclose(np.array(pims[i].H_ba).reshape(9, 3), ref['Hba'][i], rtol=1e-11, atol=1e-14)
        assert np.allclose(np.array(pims[i].H_bg).reshape(9, 3), ref['Hbg'][i], rtol=1e-11, atol=1e-14)
        c = np.array(pims[i].cov).reshape(15, 15)
        assert np.allclose(c, ref['cov'][i], rtol=1e-10, atol=1e-22)
    ctx.close()


def test_ragged_intervals_and_empty():
    spec = synth.make_config('C2', seed=4, scale=0.06)
    flat = spec['imu_samples'].reshape(-1, 6)
    offsets = np.array([0, 0, 7, 7 + 33, 7 + 33 + 1])          # empty, 7, 33, 1 samples
    ctx = abi.Context(device=0)
    pims = ctx.preintegrate(offsets, flat, 0.005, abi.vn100_imu_params(), np.zeros((4, 6)))
    assert pims[0].dt == 0.0 and np.all(np.array(pims[0].cov) == 0)
    for k, (a, b) in enumerate(zip(offsets[:-1], offsets[1:])):
        if b > a:
            ref = oimu.preintegrate(flat[a:b][None], 0.005, oimu.vn100_params(), np.zeros((1, 6)))
            assert np.allclose(np.array(pims[k].preint), ref['preint'][0], rtol=1e-12, atol=1e-15)
            assert np.allclose(np.array(pims[k].cov).reshape(15, 15), ref['cov'][0], rtol=1e-10, atol=1e-22)
    ctx.close()


def test_c1_pose_graph():
    check(synth.make_config('C1', seed=1), 1e-9, 1e-8)


def test_c2_vio():
    check(synth.make_config("C2", seed=1, scale=0.1), 1e-9, 1e-8)        # full size: tests/test_gpu_fullsize_oracle.py


def test_c3_vio_planes():
    check(synth.make_config('C3', seed=1, scale=0.1), 1e-9, 1e-8)


@pytest.mark.parametrize('chart', [lie.FIRST_ORDER_EXPMAP, lie.FIRST_ORDER_CAYLEY])
@pytest.mark.parametrize('name,scale', [('C1', 1.0), ('C3', 0.1)])
def test_pose_chart_options(name, scale, chart):
    """SURVEY A.1 / hard part 1: GTSAM's compile-time Pose3 / Rot3 charts as a context option -- FIRST_ORDER over Rot3 EXPMAP
    and over Rot3 CAYLEY (GTSAM 4.0's default build) against the oracle run with the same chart; the initial error differs
    from the EXPMAP one (non-zero residuals are chart dependent), the optimum does not."""
    spec = synth.make_config(name, seed=4, scale=scale)
    rep = check(spec, 1e-9, 1e-8, chart=chart)
    ctx = abi.Context(device=0); abi.load_spec(ctx, spec); rep0 = ctx.optimize(); ctx.close()
    assert abs(rep.initial_error - rep0.initial_error) > 1e-9 * rep0.initial_error
    assert abs(rep.final_error - rep0.final_error) < 1e-3 * rep0.final_error


def test_c4_ba_imu_schur():
    check(synth.make_config('C4', seed=1, scale=0.03), 1e-7, 1e-6, solver='schur')


def test_c4_multi_block_supernodes():
    """120 poses: panels taller than one row block of k_chol_rs (fg_chol_rs.cu), several leaf fronts."""
    check(synth.make_config('C4', seed=1, scale=0.06), 1e-7, 1e-6, solver='schur')


@pytest.mark.parametrize('seed', [2, 3])
def test_c4_other_seeds(seed):
    check(synth.make_config('C4', seed=seed, scale=0.02), 1e-7, 1e-6, solver='schur')


def test_rejected_steps_follow_oracle():
    """A badly initialised pose graph makes LM reject trials; lambda must climb exactly as in the oracle."""
    spec = synth.make_config('C1', seed=5)
    rng = np.random.default_rng(1)
    nz = rng.normal(size=(spec['n_poses'], 6)) * np.array([2.0] * 3 + [8.0] * 3)
    nz[0] = 0
    dR, dt = lie.se3_exp(nz)
    spec['pose_init_R'], spec['pose_init_t'] = lie.pose_compose(spec['pose_init_R'], spec['pose_init_t'], dR, dt)
    ctx, rep, g0, g1, orep, e_dev, e_orc = run_both(spec, max_iterations=6)
    dt_, ot = rep.trace(), orep['trace']
    assert [t['accepted'] for t in dt_] == [t['accepted'] for t in ot]
    assert not all(t['accepted'] for t in dt_), 'test graph did not trigger a rejection'
    assert np.allclose([t['lam'] for t in dt_], [t['lam'] for t in ot], rtol=1e-12)
    assert np.allclose([t['new_err'] for t in dt_ if np.isfinite(t['new_err'])], [t['new_err'] for t in ot if np.isfinite(t['new_err'])], rtol=1e-6)
    assert abs(rep.final_error - orep['error']) <= 1e-6 * orep['error']
    ctx.close()


def test_values_api_and_errors():
    ctx = abi.Context(device=0)
    I = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0.5, 0, 0.0])
    k0, k1 = abi.symbol('x', 0), abi.symbol('x', 1)
    ctx.add_pose(k0, I)
    with pytest.raises(abi.FgError) as e:
        ctx.add_pose(k0, I)
    assert e.value.code == -2
    with pytest.raises(abi.FgError) as e:
        ctx.add_between(k0, k1, I, np.eye(6))
    assert e.value.code == -3
    assert ctx.exists(k0) and not ctx.exists(k1)
    ctx.add_pose(k1, I)
    ctx.add_prior_pose(k0, I, np.eye(6) * 1e4)
    Z = I.copy(); Z[9:] = [1.0, 0.2, -0.1]
    ctx.add_between(k0, k1, Z, np.eye(6) * 100)
    rep = ctx.optimize()
    got = ctx.get_value(k1)
    assert np.allclose(got[9:], [1.5, 0.2, -0.1], atol=1e-6) and rep.final_error < 1e-10
    ctx.update_value(k1, I)
    assert abs(ctx.error() - 0.5 * 100 * (1 + 0.04 + 0.01)) < 1e-9
    ctx.close()


def test_single_pose_and_empty_landmark():
    """Edge cases: a landmark with zero observations (prior only) and a pose with no projection factors."""
    spec = synth.make_config('C4', seed=6, scale=0.02)
    keep = spec['proj_point'] != 3
    for k in ('proj_pose', 'proj_point', 'proj_uv'):
        spec[k] = spec[k][keep]
    check(spec, 1e-7, 1e-6, solver='schur')


def test_duplicate_pose_landmark_factors_match_oracle():
    """Several GenericProjectionFactors on one (pose, landmark) pair -- GTSAM accepts them and CGraphGT's BA builder can
    produce them (two features of frame i matched to one landmark, gtsam_graph.cpp:398-434): every one of them counts."""
    spec = synth.make_config('C4', seed=5, scale=0.03)
    rng = np.random.default_rng(11)
    M = len(spec['proj_pose'])
    pick = rng.choice(M, size=40, replace=False)
    pick = np.concatenate([pick, pick[:7], pick[:3]])         # some pairs three and four times
    spec['proj_pose'] = np.concatenate([spec['proj_pose'], spec['proj_pose'][pick]])
    spec['proj_point'] = np.concatenate([spec['proj_point'], spec['proj_point'][pick]])
    spec['proj_uv'] = np.concatenate([spec['proj_uv'], spec['proj_uv'][pick] + rng.normal(size=(len(pick), 2))])
    check(spec, 1e-7, 1e-6, solver='schur')


def test_loop_closure_covisibility_matches_oracle():
    """Landmarks re-observed after a long gap (loop closure): pose pairs far apart in the sequence share landmarks, so the reduced
    Hessian has blocks far off the band -- separators of the nested dissection grow, Schur tiles appear far from the diagonal
    and walk landmark ranges in which most of their poses see nothing (untested for fill and table limits)."""
    spec = synth.make_config('C4', seed=3, scale=0.06)
    rng = np.random.default_rng(5)
    P, L = spec['n_poses'], len(spec['point_init'])
    first = np.full(L, P, dtype=np.int64); last = np.zeros(L, dtype=np.int64)
    np.minimum.at(first, spec['proj_point'], spec['proj_pose']); np.maximum.at(last, spec['proj_point'], spec['proj_pose'])
    K = synth.CAL_SR4K
    add_p, add_l, add_uv = [], [], []
    for l in rng.permutation(L)[:2500]:
        far = np.array([q for q in range(P) if q < first[l] - 40 or q > last[l] + 40])
        if len(far) == 0: continue
        uv, z = synth._project(spec['truth_R'][far], spec['truth_t'][far], spec['truth_point'][l][None, :], K, spec['Rs'], spec['ts'])
        ok = (z > 0.5) & (np.abs(uv[:, 0] - K[3]) < 160.0) & (np.abs(uv[:, 1] - K[4]) < 140.0)      # really visible from there
        if not ok.any(): continue
        pick = rng.choice(np.nonzero(ok)[0], size=min(3, int(ok.sum())), replace=False)
        add_p.append(far[pick]); add_l.append(np.full(len(pick), l)); add_uv.append(uv[pick] + rng.normal(size=(len(pick), 2)))
    add_p = np.concatenate(add_p); assert len(add_p) > 100
    spec['proj_pose'] = np.concatenate([spec['proj_pose'], add_p.astype(np.int32)])
    spec['proj_point'] = np.concatenate([spec['proj_point'], np.concatenate(add_l).astype(np.int32)])
    spec['proj_uv'] = np.concatenate([spec['proj_uv'], np.concatenate(add_uv)])
    check(spec, 1e-7, 1e-6, solver='schur')


def mixed_calibration_spec():
    """C4 at 3 % with the observations of every third pose taken by a second camera: other Cal3DS2 (focal lengths, skew,
    principal point, distortion) and other body_P_sensor -- GTSAM accepts projection factors with different K / body_P_sensor in
    one graph (the reference builds one pair only; round 1 rejected mixed graphs)."""
    spec = synth.make_config('C4', seed=9, scale=0.03)
    rng = np.random.default_rng(17)
    K1 = np.array(spec['K'], dtype=np.float64)
    K2 = K1 * np.array([1.06, 0.97, 1.0, 1.03, 0.95, 0.8, 0.7, 1.0, 1.0]) + np.array([0, 0, 0.4, 0, 0, 0, 0, 2e-3, -1e-3])
    dR, dt = lie.se3_exp(np.array([0.03, -0.02, 0.04, 0.01, 0.02, -0.015]))
    Rs2, ts2 = lie.pose_compose(spec['Rs'], spec['ts'], dR, dt)
    cam2 = (spec['proj_pose'] % 3) == 1
    uv2, z = synth._project(spec['truth_R'][spec['proj_pose'][cam2]], spec['truth_t'][spec['proj_pose'][cam2]],
                            spec['truth_point'][spec['proj_point'][cam2]], K2, Rs2, ts2)
    assert np.all(z > 0.1)
    spec['proj_uv'] = spec['proj_uv'].copy()
    spec['proj_uv'][cam2] = uv2 + rng.normal(size=uv2.shape)
    spec['proj_cal'] = cam2.astype(np.int64)
    spec['cals'] = [(K1, spec['Rs'], spec['ts']), (K2, Rs2, ts2)]
    return spec


def test_mixed_calibrations_match_oracle():
    check(mixed_calibration_spec(), 1e-7, 1e-6, solver='schur')


def oracle_marginal(g, kind, idx):
    """Dense oracle: block of the inverse of the full (undamped) normal equations, landmarks included."""
    H, grad, err = g.normal_equations()
    d = g.dims
    o, n = {'pose': (6 * idx, 6), 'vel': (d['o_v'] + 3 * idx, 3), 'bias': (d['o_b'] + 6 * idx, 6), 'plane': (d['o_pl'] + 3 * idx, 3)}[kind]
    E = np.zeros((H.shape[0], n)); E[o + np.arange(n), np.arange(n)] = 1.0
    import scipy.sparse.linalg as spla
    X = spla.splu(H.tocsc()).solve(E)
    return X[o:o + n]


def test_marginal_covariance_matches_oracle():
    """Marginals(...).marginalCovariance(key): SURVEY 8f-2 (gtsam_graph.cpp:598-601, :1357)."""
    spec = synth.make_config('C4', seed=4, scale=0.03)
    ctx = abi.Context(device=0)
    abi.load_spec(ctx, spec)
    ctx.optimize()
    g = build.from_spec(spec)
    g.R = ctx.get_values(abi.T_POSE)[:, :9].reshape(-1, 3, 3); g.t = ctx.get_values(abi.T_POSE)[:, 9:]
    g.vel = ctx.get_values(abi.T_VEC3); g.bias = ctx.get_values(abi.T_BIAS); g.point = ctx.get_values(abi.T_POINT)
    P = spec['n_poses']
    for kind, ch, idx in (('pose', 'x', 0), ('pose', 'x', P // 2), ('pose', 'x', P - 1), ('vel', 'v', 3), ('bias', 'b', P - 2)):
        got = ctx.marginal_covariance(abi.symbol(ch, idx))
        ref = oracle_marginal(g, kind, idx)
        assert got.shape == ref.shape
        assert np.allclose(got, got.T, rtol=1e-9, atol=1e-30)
        assert np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max(), (kind, idx, np.abs(got - ref).max(), np.abs(ref).max())
    with pytest.raises(abi.FgError) as e:
        ctx.marginal_covariance(abi.symbol('q', 0))
    assert e.value.code == -1
    # the state is untouched and the optimiser still works afterwards
    assert ctx.optimize().iterations <= 2
    ctx.close()


def test_two_view_bundle_adjust_edge_information():
    """CGraphGT::bundleAdjust (gtsam_graph.cpp:500-610): prior on pose 0 (sigma 1e-7), PriorFactor<Point3> (0.014) and two
    GenericProjectionFactor per match (1 px), LM, then Marginals -> the VRO edge information = inverse of pose 1's
    marginal covariance."""
    rng = np.random.default_rng(11)
    full = synth.make_config('C4', seed=7, scale=0.02)
    K, Rs, ts = full['K'], full['Rs'], full['ts']
    n = 150
    T1 = lie.se3_exp(np.array([[0.02, -0.03, 0.05, 0.08, -0.02, 0.03]]))
    R = np.stack([np.eye(3), T1[0][0]]); t = np.stack([np.zeros(3), T1[1][0]])
    Rc, tc = lie.pose_compose(R, t, np.broadcast_to(Rs, (2, 3, 3)), np.broadcast_to(ts, (2, 3)))
    pc = np.column_stack([rng.uniform(-0.6, 0.6, n), rng.uniform(-0.5, 0.5, n), rng.uniform(1.5, 4.0, n)])
    pts = pc @ Rc[0].T + tc[0]
    from oracle import factors as ofac
    uv = np.concatenate([ofac.projection(R[k], t[k], pts, np.zeros((n, 2)), K, Rs, ts, jac=False) for k in range(2)])
    uv = uv + rng.normal(size=(2 * n, 2))
    spec = dict(name='two_view', seed=0, n_poses=2, K=K, Rs=Rs, ts=ts,
                pose_init_R=np.stack([np.eye(3), np.eye(3)]), pose_init_t=np.zeros((2, 3)),
                prior_pose_R=np.eye(3), prior_pose_t=np.zeros(3),
                point_init=pts + rng.normal(size=pts.shape) * 0.014, point_prior_sigma=0.014,
                proj_pose=np.repeat(np.arange(2), n).astype(np.int32), proj_point=np.tile(np.arange(n), 2).astype(np.int32),
                proj_uv=uv, proj_sigma=1.0)
    ctx, rep, g0, g1, orep, e_dev, e_orc = run_both(spec, solver='schur')
    assert rep.iterations == orep['iterations'] and abs(rep.final_error - orep['error']) <= 1e-7 * orep['error']
    cov = ctx.marginal_covariance(abi.symbol('x', 1))
    ref = oracle_marginal(g1, 'pose', 1)
    assert np.abs(cov - ref).max() <= 1e-6 * np.abs(ref).max()
    info = np.linalg.inv(cov)                      # what bundleAdjust hands back as the edge information (:601)
    assert np.all(np.linalg.eigvalsh(0.5 * (info + info.T)) > 0)
    ctx.close()


def _dense_spec(n, P, keep=None, seed=21):
    """P poses in a row that all see the same n landmarks (keep(pose, landmark) -> bool thins the observations)."""
    rng = np.random.default_rng(seed)
    full = synth.make_config('C4', seed=7, scale=0.02)
    K, Rs, ts = full['K'], full['Rs'], full['ts']
    xi = np.zeros((P, 6)); xi[:, 3] = 0.06 * np.arange(P) * 5.0 / P; xi[:, 1] = 0.01 * np.arange(P) * 5.0 / P
    R, t = lie.se3_exp(xi)
    Rc, tc = lie.pose_compose(R, t, np.broadcast_to(Rs, (P, 3, 3)), np.broadcast_to(ts, (P, 3)))
    pc = np.column_stack([rng.uniform(-0.5, 0.5, n), rng.uniform(-0.4, 0.4, n), rng.uniform(1.5, 4.0, n)])
    pts = pc @ Rc[0].T + tc[0]
    from oracle import factors as ofac
    uv = np.concatenate([ofac.projection(R[k], t[k], pts, np.zeros((n, 2)), K, Rs, ts, jac=False) for k in range(P)])
    uv = uv + rng.normal(size=uv.shape)
    nz = rng.normal(size=(P, 6)) * 0.01; nz[0] = 0
    dR, dt = lie.se3_exp(nz)
    Ri, ti = lie.pose_compose(R, t, dR, dt)
    pp, pl = np.repeat(np.arange(P), n), np.tile(np.arange(n), P)
    m = np.ones(len(pp), dtype=bool) if keep is None else np.array([keep(a, b) for a, b in zip(pp, pl)])
    return dict(name='dense', seed=0, n_poses=P, K=K, Rs=Rs, ts=ts, pose_init_R=Ri, pose_init_t=ti,
                prior_pose_R=np.eye(3), prior_pose_t=np.zeros(3),
                point_init=pts + rng.normal(size=pts.shape) * 0.014, point_prior_sigma=0.014,
                proj_pose=pp[m].astype(np.int32), proj_point=pl[m].astype(np.int32), proj_uv=uv[m], proj_sigma=1.0)


def test_dense_visibility_splits_schur_chunks():
    """Every pose sees every landmark: a pose then holds 96 records in a 96-landmark round of k_schur_tiles, more than its 56
    record slots, so every round is redone one 32-landmark word at a time (fg_schur.cu: the `single` path)."""
    check(_dense_spec(100, 5), 1e-7, 1e-6, solver='schur')


def test_shared_landmark_runs_overflow_the_hit_list():
    """40 of every 96 consecutive landmarks are seen by ALL 20 poses and the rest by two poses each: every pose stays within
    the 56 record slots of a round, but every pose pair has more than the 32 hits its hit list holds -- the CTA-wide vote
    sends the round to the word-by-word path."""
    def keep(pose, l):
        r = l % 96
        return r < 40 or pose == 2 + (l % 18) or pose == 2 + ((l + 7) % 18)
    spec = _dense_spec(480, 20, keep)
    per_pose = np.bincount(spec['proj_pose'] * 5 + spec['proj_point'] // 96, minlength=100)
    assert per_pose.max() <= 56 and per_pose.min() >= 40
    check(spec, 1e-7, 1e-6, solver='schur')


@pytest.mark.parametrize('name,scale', [('C2', 0.2), ('C3', 0.1), ('C4', 0.06)])
def test_runs_are_bitwise_repeatable(name, scale):
    """No fp64 atomics on the accumulation paths (fixed-order reductions): pose-side factors are assembled colour
    by colour, per-landmark sums are merged in order inside one block, scalars go through per-block partials -- two runs of
    the same graph give bit-identical errors, traces and states."""
    spec = synth.make_config(name, seed=3, scale=scale)
    out = []
    for _ in range(2):
        ctx = abi.Context(device=0)
        abi.load_spec(ctx, spec)
        e0 = ctx.error()
        rep = ctx.optimize()
        out.append((e0, rep.initial_error, rep.final_error, [t['new_err'] for t in rep.trace()], ctx.get_values(abi.T_POSE).tobytes(),
                    ctx.get_values(abi.T_POINT).tobytes() if ctx.num_values(abi.T_POINT) else b''))
        ctx.close()
    assert out[0] == out[1]
