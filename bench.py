#!/usr/bin/env python
"""bench.py -- LM iterations/s on the synthetic VIO-BA graph (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C5] [--impl ours|reference]

A "step" is one outer Levenberg-Marquardt iteration (linearise, damped Schur solve(s), retract, chi2, accept/
reject) of CGraphGT::optimizeGraphBatch's loop (gtsam/gtsam_graph.cpp:1784-1788) over the whole graph.
  value  : iterations/s with the graph and the state vector resident in HBM (timed region = K iterations)
  e2e    : iterations/s through the C ABI with HOST buffers: per step the state vector is copied in from
           pinned host memory (fg_set_values), one LM iteration runs, and the state is copied back out.
  N > 1  : landmarks are sharded across ranks, one ncclAllReduce of the reduced Hessian per trial;
           total work is fixed => scaling "strong".
The reference arm times the CPU restatement of the reference's GTSAM path on the host cores: for the BA graphs (C4, C5)
oracle/cpu_lm.cpp, an OpenMP C++ port of one LM iteration, on the full workload with every host thread.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'LM iterations/sec on the 5k-pose/500k-point VIO-BA graph'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--config', default=os.environ.get('FG_BENCH_CONFIG', 'C5'))
    ap.add_argument('--scale', type=float, default=float(os.environ.get('FG_BENCH_SCALE', '1.0')))
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--shard', default='blocks', choices=['blocks', 'slice'], help='multi-GPU landmark shards: 96-landmark blocks dealt round-robin, or one contiguous range per rank')
    return ap.parse_args()


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith('active'):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return dict(sm_mhz=float(np.median(self.samples)) if self.samples else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons))


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)).get('hbm_gbs', 6650.0), 'measured'
    return 6650.0, 'fallback'


def algorithmic_bytes(spec_sizes):
    """SURVEY 8d byte model for one accepted LM trial, from the generated graph's actual sizes."""
    M, L, P, nnz = spec_sizes['M'], spec_sizes['L'], spec_sizes['P'], spec_sizes['nnz']
    b_obs = 2 * (24 * M + 24 * L)                   # two passes over (idx, uv) + landmark xyz
    b_lm = L * (72 + 72 + 24)                       # Vinv, V^-1 g write+read, new landmark
    b_S = 8 * nnz * 4                               # reduced Hessian: write, read by factor, write factor, read in solve
    b_state = 3 * P * 176
    return dict(total=b_obs + b_lm + b_S + b_state, obs=b_obs, lm=b_lm, S=b_S, state=b_state)


def workload_name(config, scale):
    from graph_slam_b200 import synth
    full = synth.CONFIGS[config]
    P = max(2, int(round(full['n_poses'] * scale))); L = int(round(full.get('n_landmarks', 0) * scale))
    return '%s%s: %d poses (X,V,B), %d landmarks, ~%d projections, %d IMU factors' % (config, '' if scale == 1.0 else '@%g' % scale, P, L, 20 * L, P - 1)


def cpu_baseline_sample(config, steps=1, spec=None, scale=1.0, warmup=0, one_thread=False):
    """The reference's CPU path restated (the reference itself cannot be built here), timed on the host cores.
    BA + IMU graphs (C4, C5): oracle/cpu_lm.cpp, an OpenMP C++ port of one LM iteration, on the FULL workload with every
    host thread.  Other graphs: the numpy oracle (one core) on a bounded sample."""
    from graph_slam_b200 import synth
    full = synth.CONFIGS[config]
    if full.get('n_landmarks', 0) > 0:
        from oracle import cpu_baseline
        if spec is None:
            spec = synth.make_config(config, seed=1, scale=scale)
        threads = os.cpu_count() or 1
        st = cpu_baseline.State(spec)
        lam, n, dt = 1e-5, 0, 0.0
        ph = np.zeros(5)
        for k in range(warmup + max(steps, 1)):
            keep = [a.copy() for a in (st.pose, st.vel, st.bias, st.pts)]
            rc, e0, e1, secs, band = st.iterate(lam, threads)
            if k >= warmup:                                         # the first `warmup` iterations are untimed
                # the five phase timers of cpu_lm_iteration: linearise, Schur, factor, solve + retract, error.  Its index
                # build and buffer allocation (a one-time cost in a real solver) are outside them and not counted.
                dt += float(secs.sum()); ph += secs
                n += 1
            if rc == 0 and e1 < e0:
                lam = lam / 10.0
            else:                                                   # rejected trial: restore the state, raise lambda (GTSAM's rule)
                for a, b in zip((st.pose, st.vel, st.bias, st.pts), keep):
                    a[...] = b
                lam = lam * 10.0
        sample = ('full %s%s (%d poses, %d landmarks, %d projections): %d LM iteration(s) in %.1f s with %d OpenMP threads '
                  '(oracle/cpu_lm.cpp: linearise, Schur, block-banded Cholesky, back-substitution, retract, error)' %
                  (config, '' if scale == 1.0 else '@%g' % scale, spec['n_poses'], len(spec['point_init']), len(spec['proj_pose']), n, dt, threads))
        out = dict(value=n / dt, unit='iterations/s', cores=threads, kind='port', sample=sample, steps_run=n, warmup_run=warmup,
                   projections=len(spec['proj_pose']),
                   phases_s=dict(zip(('linearize', 'schur', 'factor', 'solve_retract', 'error'), (ph / n).tolist())))
        if one_thread:                                              # also the single-thread figure (one iteration)
            secs1 = st.iterate(lam, 1)[3]
            out['one_thread'] = dict(value=1.0 / float(secs1.sum()), unit='iterations/s', cores=1)
        return out
    from oracle import build, lm
    sc = 0.05
    spec = synth.make_config(config, seed=1, scale=sc)
    g = build.from_spec(spec)
    err = g.error()
    lam = 1e-5
    p = lm.LMParams()
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        g, lam, err = lm.lm_iterate(g, lam, p, err, solver='schur')
        n += 1
    dt = time.perf_counter() - t0
    m_s = len(spec.get('between_i', []))
    sample = ('%s scaled x%g (numpy oracle, one core): %d poses, %d relative-pose edges; %d LM iteration(s) in %.1f s; value = sample rate x %g' %
              (config, sc, spec['n_poses'], m_s, n, dt, sc))
    return dict(value=(n / dt) * sc, unit='iterations/s', cores=1, kind='port', sample=sample, steps_run=n, warmup_run=0)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t0 = time.perf_counter()
    # every step is one full LM iteration of the same workload (~1.7 s at C5 on 16 cores); the run is bounded at 40
    # iterations in all so that it ends within a few minutes, and the line reports the counts actually run
    steps = max(1, min(args.steps, 30)); warmup = max(0, min(args.warmup, 10))
    base = cpu_baseline_sample(args.config, steps=steps, scale=args.scale, warmup=warmup)
    line = dict(metric=METRIC, value=base['value'], unit='iterations/s', n_gpus=args.gpus, steps=base['steps_run'],
                warmup=base['warmup_run'], steps_requested=args.steps, warmup_requested=args.warmup,
                ms_per_step=1000.0 / base['value'] if base['value'] else None,
                higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f64', data='synthetic',
                impl='reference', config=dict(workload=workload_name(args.config, args.scale), projections=base.get('projections'),
                                              l2='inputs_exceed_L2', parallelism='%d host threads' % base['cores'],
                                              charts='Pose3 EXPMAP / Rot3 EXPMAP', lm='GTSAM defaults, forced iterations',
                                              note='CPU restatement of the reference path, see cpu_baseline.sample; '
                                              'timed region = the five phases of each iteration (setup outside)'),
                cpu_baseline=base,
                e2e=dict(value=base['value'], unit='iterations/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                wall_s=time.perf_counter() - t0)
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return
    import torch
    from graph_slam_b200 import abi, synth
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product has no CPU path')
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    spec = synth.make_config(args.config, seed=1, scale=args.scale)
    ctx = abi.Context(device=local, rank=rank, nranks=world)
    L = len(spec.get('point_init', []))
    sl = None
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device='cuda')
        if rank == 0:
            uid = torch.frombuffer(bytearray(abi.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
        if args.shard == 'slice':
            sl = (L * rank // world, L * (rank + 1) // world)
    t_build = time.perf_counter()
    # landmark shards: blocks of 96 consecutive landmarks dealt round-robin (abi.shard_landmarks) -- every rank then holds every
    # Schur tile with 1/world of its rounds; --shard slice gives each rank one contiguous range (1/world of the tiles)
    ids = abi.shard_landmarks(L, rank, world) if (world > 1 and args.shard == 'blocks' and L) else None
    abi.load_spec(ctx, spec, landmark_slice=sl, landmark_ids=ids)
    ctx.finalize()
    t_build = time.perf_counter() - t_build

    types = [t for t in range(5) if ctx.num_values(t) > 0]
    init = {t: ctx.get_values(t) for t in types}
    pinned = {t: torch.empty(init[t].shape, dtype=torch.float64).pin_memory() for t in types}
    for t in types:
        pinned[t].numpy()[:] = init[t]
    out_pinned = {t: torch.empty(init[t].shape, dtype=torch.float64).pin_memory() for t in types}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reset():
        for t in types:
            ctx.set_values(t, pinned[t].numpy())

    fp64_meas = abi.fp64_peak(local)
    # ---- device-resident arm
    ctx.optimize(max_iterations=max(args.warmup, 3), force_iterations=1)
    reset()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = abi.launch_count()
    t0 = time.perf_counter()
    rep = ctx.optimize(max_iterations=args.steps, force_iterations=1)
    barrier()
    dt = time.perf_counter() - t0
    launches = abi.launch_count() - launches0            # counted at every <<<>>> of the library (FGS() in fg_internal.h), rank 0's
    # ---- end-to-end arm: host buffers in and out every step
    reset()
    for _ in range(2):
        ctx.optimize(max_iterations=1, force_iterations=1)
    reset()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        for t in types:
            ctx.set_values(t, pinned[t].numpy())
        r1 = ctx.optimize(max_iterations=1, force_iterations=1)
        for t in types:
            ctx.get_values(t, out_pinned[t].numpy())
        pinned, out_pinned = out_pinned, pinned   # next step continues from this step's result
    barrier()
    dt_e2e = time.perf_counter() - t1
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if world > 1:
        tt = torch.tensor([dt, dt_e2e], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_e2e = tt.tolist()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    state_bytes = int(sum(init[t].nbytes for t in types))
    value = args.steps / dt
    peak, peak_src = measured_peaks()
    # the byte model of SURVEY 8d twice: with nnz(S) (the reduced Hessian as assembled: algorithmic) and with nnz(L) (what this
    # solver stores, nested-dissection fill included)
    M_all = len(spec.get('proj_pose', [])); L_all = len(spec.get('point_init', []))
    ab = algorithmic_bytes(dict(M=M_all, L=L_all, P=spec['n_poses'], nnz=int(rep.nnz_S)))
    ab_fill = algorithmic_bytes(dict(M=M_all, L=L_all, P=spec['n_poses'], nnz=int(rep.nnz_L)))
    trials = max(rep.trials, 1)
    phases = dict(linearize=rep.ms_linearize / max(rep.iterations, 1), schur=rep.ms_schur / trials,
                  factor=rep.ms_factor / trials, backsolve=rep.ms_solve / trials,
                  retract_error=rep.ms_retract_error / trials)
    # Rooflines.  Times are CUDA-event durations measured live around the single kernels on the
    # context's stream (fg_lm_report.ms_proj_obs / ms_schur_blocks) or around the phase for the factorisation / k_backsolve_w.
    m_rank, l_rank = int(rep.n_projections), int(rep.n_landmarks)
    iters = max(rep.iterations, 1)
    t_obs = rep.ms_proj_obs / iters                       # k_proj_obs<JAC>: 176 B per observation (idx 8, uv 16, w 8, W 144)
    b_obs = m_rank * 176 + l_rank * (24 + 72)
    t_sch = rep.ms_schur_blocks / trials                  # k_schur_tiles: every Z record read once (144 B/obs), reduced blocks RMW
    b_sch = m_rank * 144 + 2 * 8 * int(rep.nnz_L)
    f_sch = 216.0 * int(rep.n_schur_pairs)                # 108 DFMA per (observation a, observation b) pair of a landmark
    t_cho = phases['factor']                              # k_chol_rs (two launches) + k_front_syrk: one read of the assembled panels, one write of L
    b_cho = 2 * 8 * int(rep.nnz_L)

    # dram__bytes_read+write per launch from the committed ncu --set full capture (single GPU, C5 only)
    traffic = {}
    tp = os.path.join(ROOT, 'profiles', 'r2_traffic.json')   # refreshed with every committed ncu --set full capture
    if os.path.exists(tp) and args.config == 'C5' and args.scale == 1.0 and world == 1:
        traffic = json.load(open(tp)).get('dram_bytes_per_launch', {})

    # fp64 arithmetic peak of this GPU, measured in this run (fg_debug_fp64_peak: register-only DFMA / DMMA loops on every SM;
    # MEASURED_PEAKS.json has HBM and bf16 only)
    fp64_peak = dict(zip(('dfma', 'dmma'), fp64_meas), source='measured in this run (fg_debug_fp64_peak)')
    f_cho = float(ctx.debug_sizes()[5])                   # flops of one factorisation (fg_symbolic.cpp)

    def roof(name, nbytes, ms, note, flops=None, pipe='dfma'):
        a = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        r = dict(kernel=name, bound='hbm', achieved=a, peak=peak, unit='GB/s', frac=a / peak, traffic=traffic.get(name),
                 algorithmic_bytes=nbytes, ms=ms, note=note)
        if flops:
            tf = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            r['fp64'] = dict(achieved=tf, peak=fp64_peak[pipe], unit='TFLOP/s', frac=tf / fp64_peak[pipe], pipe=pipe.upper(),
                             flops=flops, peak_source=fp64_peak['source'])
        return r
    roofs = [roof('k_chol_rs', b_cho, t_cho, 'factorisation phase (k_chol_rs x2 + k_front_syrk): dependency chain of %d levels; ~5 GB of descendant panels through L2 per factorisation (profiles/r2_ncu_full_summary.md), fp64 on DMMA' % int(rep.n_levels), flops=f_cho, pipe='dmma'),
             roof('k_schur_tiles', b_sch, t_sch, 'fp64-FMA work: %d pairs x 216 flop = %.1f TFLOP/s achieved (fp64 DFMA peak measured in this run: %.1f TFLOP/s); the profile says shared-memory bandwidth bounds it (DESIGN 2.1, profiles/r2_ncu_full_summary.md)' % (
                 int(rep.n_schur_pairs), f_sch / (t_sch * 1e-3) / 1e12 if t_sch > 0 else 0.0, fp64_meas[0]), flops=f_sch, pipe='dfma'),
             roof('k_proj_obs<1>', b_obs, t_obs, 'streaming pass over the observations')]
    dominant = max(roofs, key=lambda r: r['ms'])
    line = dict(metric=METRIC, value=value, unit='iterations/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=1000.0 * dt / args.steps, ms_per_step_cuda_events=float(rep.ms_total) / max(rep.iterations, 1),
                higher_is_better=True, scaling='strong', vs_baseline=None,
                dtype='f64', data='synthetic',
                config=dict(workload=workload_name(args.config, args.scale), projections=len(spec.get('proj_pose', [])),
                    l2='inputs_exceed_L2' if ab['total'] > 126e6 else 'smaller_than_L2', parallelism='landmark-shard x%d (%s), reduced solve replicated' % (world, args.shard if world > 1 else 'all'),
                    charts='Pose3 EXPMAP / Rot3 EXPMAP', lm='GTSAM defaults, forced iterations',
                    graph_build_s=t_build),
                e2e=dict(value=args.steps / dt_e2e, unit='iterations/s', h2d_bytes_per_step=state_bytes,
                         d2h_bytes_per_step=state_bytes),
                gpu_launches=int(launches),
                clocks=sampler.summary(),
                roofline=dict(dominant, peak_source=peak_src, iteration_algorithmic_bytes=ab['total'],
                              iteration_frac=ab['total'] / (dt / args.steps) / 1e9 / peak,
                              iteration_bytes_with_fill=ab_fill['total'], iteration_frac_with_fill=ab_fill['total'] / (dt / args.steps) / 1e9 / peak),
                roofline_all=roofs,
                phases_ms=phases, lm=dict(iterations=rep.iterations, trials=rep.trials, initial_error=rep.initial_error,
                                          final_error=rep.final_error, e2e_last_error=r1.final_error),
                sizes=dict(reduced_dims=int(rep.n_reduced_dims), supernodes=int(rep.n_supernodes), nnz_L=int(rep.nnz_L), nnz_S=int(rep.nnz_S)),
                collectives=dict(per_trial=['ncclAllReduce f64 x %d (packed reduced Hessian + rhs + chi2)' % (int(rep.allreduce_bytes) // 8),
                                            'ncclAllReduce f64 x 3 (g.delta, |delta|^2, new chi2)'] if world > 1 else [],
                                 allreduce_bytes=int(rep.allreduce_bytes)))
    if not args.no_cpu_baseline and world == 1:                      # the CPU baseline is reported at N = 1 only
        try:
            line['cpu_baseline'] = cpu_baseline_sample(args.config, steps=2, spec=spec, scale=args.scale, one_thread=True)
        except Exception as e:                                   # the baseline must never take the bench down
            line['cpu_baseline'] = dict(value=None, unit='iterations/s', cores=1, kind='port', sample='failed: %r' % (e,))
    print(json.dumps(line))


if __name__ == '__main__':
    main()
