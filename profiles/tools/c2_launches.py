import sys
sys.path.insert(0,'.')
from graph_slam_b200 import abi, synth
spec = synth.make_config('C2', seed=1, scale=1.0)
ctx = abi.Context(device=0)
abi.load_spec(ctx, spec)
n0 = abi.launch_count()
rep = ctx.optimize()
print('iterations', rep.iterations, 'trials', rep.trials, 'launches', abi.launch_count()-n0, 'ms_total', rep.ms_total, 'lin', rep.ms_linearize, 'fac', rep.ms_factor)
