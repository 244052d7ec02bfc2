"""Host side of the multi-GPU path on CPU: two gloo ranks build their landmark shards in detached contexts and
must arrive at the SAME reduced-system structure and packed exchange index (what makes the single allreduce legal),
and that structure must equal the unsharded one declared with the same co-visibility band."""
import hashlib
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import hashlib, os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from graph_slam_b200 import abi, synth
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    spec = synth.make_config('C4', seed=3, scale=0.04)
    L = len(spec['point_init']); P = spec['n_poses']
    pims = (abi.Pim * (P - 1))()
    for i in range(P - 1):
        pims[i].dt = 0.1; pims[i].cov[:] = np.eye(15).ravel().tolist()
    def digest(sl, check_pack=False, ids=None):
        ctx = abi.Context(device=-1, rank=rank, nranks=world)
        abi.load_spec(ctx, spec, landmark_slice=sl, preintegrated=pims, landmark_ids=ids)
        h = hashlib.sha256()
        for w in list(range(0, 17)) + [49]:
            h.update(ctx.symbolic(w).tobytes())
        if check_pack:
            # the packed exchange index must cover every entry of the assembled reduced system (and its rhs row):
            # place the oracle's full reduced system in the panel layout and look its non-zeros up
            sys.path.insert(0, os.path.join(%r, 'tests'))
            from test_symbolic_host import reduced_system, perm_from_offsets
            g, S, b = reduced_system(spec, 1e-3)
            perm = perm_from_offsets(ctx, g)
            n_r = len(perm)
            Sp = np.zeros((n_r + 1, n_r)); Sp[np.ix_(perm, perm)] = S; Sp[n_r, perm] = b
            col0, ncols, nrows, rowptr, valptr, rowidx = (ctx.symbolic(w) for w in range(1, 7))
            pk = set(ctx.symbolic(49).tolist())
            assert len(pk) < ctx.symbolic(0)[2], 'the packed index is not smaller than the panel storage'
            for s_ in range(len(col0)):
                rows = rowidx[rowptr[s_]:rowptr[s_] + nrows[s_]]
                for c_ in range(ncols[s_]):
                    col = Sp[rows, col0[s_] + c_]
                    for r_ in np.nonzero(col)[0]:
                        if r_ >= c_:
                            assert int(valptr[s_] + r_ + c_ * nrows[s_]) in pk, 'assembled entry missing from the packed exchange index'
        ctx.close()
        return h.hexdigest()
    mine = digest((L * rank // world, L * (rank + 1) // world), check_pack=(rank == 0))
    full = digest((0, L))
    dealt = digest(None, ids=abi.shard_landmarks(L, rank, world))      # blocks of 96 landmarks dealt round-robin (bench.py's shards)
    assert dealt == full, 'round-robin shard: structure differs from the unsharded one'

    out = [None] * world
    dist.all_gather_object(out, (mine, full))
    if rank == 0:
        assert len(set(o[0] for o in out)) == 1, 'ranks disagree on the reduced-system structure'
        assert out[0][0] == out[0][1], 'sharded structure differs from the unsharded one'
        print('SHARDING_OK')
    dist.destroy_process_group()
''') % (ROOT, ROOT)


def test_two_rank_structures_agree(fglib, tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    res = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', '29533', str(script)],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    assert 'SHARDING_OK' in res.stdout
