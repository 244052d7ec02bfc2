"""Writes tests/golden/oracle_fullsize.json (run from the repo root: python tests/golden/make_fullsize_golden.py [C4] [C5]).

One FULL-SIZE Levenberg-Marquardt iteration of the numpy oracle (oracle/lm.py: lm_iterate, Schur solver, lambda = 1e-5) on
the BASELINE configs C4 (2k poses, 100k landmarks, 2M projections) and C5 (5k poses, 500k landmarks, 10M projections):
initial error, gradient norm of the linearisation, error after the iteration, and the state after the iteration sampled at
fixed indices.  The oracle needs ~1 min (C4) / ~6 min (C5) per iteration on CPU, so the GPU tests compare the CUDA path
with these committed numbers instead of running it (tests/test_gpu_fullsize_oracle.py).  Like oracle_lm.json these are
ORACLE regression vectors -- the reference holds no goldens for these factor kinds."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, 'oracle_fullsize.json')


def one(name):
    from graph_slam_b200 import synth
    from oracle import build, lm
    spec = synth.make_config(name, seed=1, scale=1.0)
    g = build.from_spec(spec)
    e0 = g.error()
    H, grad, _ = g.normal_equations()
    n_r = g.dims['o_pt']
    g1, lam, e1 = lm.lm_iterate(g, 1e-5, lm.LMParams(), e0, solver='schur')
    P, L = spec['n_poses'], len(spec['point_init'])
    ip = np.unique(np.linspace(0, P - 1, 41).astype(int)); il = np.unique(np.linspace(0, L - 1, 41).astype(int))
    return dict(config=name, seed=1, scale=1.0, n_poses=int(P), n_landmarks=int(L), n_projections=int(len(spec['proj_pose'])),
                lambda0=1e-5, initial_error=float(e0), grad_norm_reduced=float(np.linalg.norm(grad[:n_r])),
                grad_norm_points=float(np.linalg.norm(grad[n_r:])), error_after=float(e1), lambda_after=float(lam),
                pose_idx=ip.tolist(), pose_t=g1.t[ip].tolist(), pose_R=g1.R[ip].reshape(len(ip), 9).tolist(),
                vel=g1.vel[ip].tolist(), bias=g1.bias[ip].tolist(), point_idx=il.tolist(), point=g1.point[il].tolist(),
                pose_t_sum=g1.t.sum(0).tolist(), point_sum=g1.point.sum(0).tolist())


if __name__ == '__main__':
    names = sys.argv[1:] or ['C4', 'C5']
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for n in names:
        out[n] = one(n)
        print(n, out[n]['initial_error'], out[n]['error_after'], flush=True)
        json.dump(out, open(OUT, 'w'), indent=1)
