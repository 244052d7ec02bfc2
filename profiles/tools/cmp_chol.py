"""Dev tool: one forced LM iteration with the row-split Cholesky (default) and with FG_CHOL_RS=0, compare the states."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graph_slam_b200 import abi, synth

def run(spec, rs):
    os.environ['FG_CHOL_RS'] = rs
    ctx = abi.Context(device=0)
    abi.load_spec(ctx, spec)
    rep = ctx.optimize(max_iterations=1, force_iterations=1)
    T = ctx.get_values(abi.T_POSE).copy()
    ctx.close()
    return rep, T

for name, scale in [(a.split('@')[0], float(a.split('@')[1])) for a in sys.argv[1:]]:
    spec = synth.make_config(name, seed=1, scale=scale)
    r1, T1 = run(spec, '1')
    r0, T0 = run(spec, '0')
    r2, T2 = run(spec, '1')
    print(name, scale, 'poses', spec['n_poses'], 'err rs %.10g reg %.10g' % (r1.final_error, r0.final_error),
          'max|dT| rs-reg %.3g rs-rs %.3g' % (np.abs(T1 - T0).max(), np.abs(T1 - T2).max()), 'trials', r1.trials, r0.trials, flush=True)
