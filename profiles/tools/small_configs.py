import sys, time, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graph_slam_b200 import abi, synth
for name in ['C1','C2','C3','C4']:
    spec=synth.make_config(name, seed=1)
    ctx=abi.Context(device=0); abi.load_spec(ctx, spec); ctx.finalize()
    rep=ctx.optimize()
    ctx2=abi.Context(device=0); abi.load_spec(ctx2, spec); ctx2.finalize(); ctx2.optimize(max_iterations=3, force_iterations=1)
    init={t: None for t in range(5)}
    t0=time.perf_counter(); rep2=ctx2.optimize(max_iterations=10, force_iterations=1); dt=time.perf_counter()-t0
    print(json.dumps(dict(config=name, iterations_to_converge=rep.iterations, trials=rep.trials, initial_error=rep.initial_error, final_error=rep.final_error,
          ms_per_iteration=1000*dt/10, levels=int(rep.n_levels), reduced_dims=int(rep.n_reduced_dims), supernodes=int(rep.n_supernodes),
          phases_ms=dict(linearize=rep2.ms_linearize/10, schur=rep2.ms_schur/max(rep2.trials,1), factor=rep2.ms_factor/max(rep2.trials,1), backsolve=rep2.ms_solve/max(rep2.trials,1), retract_error=rep2.ms_retract_error/max(rep2.trials,1)))))
    ctx.close(); ctx2.close()
