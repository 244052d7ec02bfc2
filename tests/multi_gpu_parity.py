"""Run under torchrun on N GPUs of one box:
     python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_parity.py
Landmark-sharded LM (one ncclAllReduce of the reduced Hessian per trial) must agree with the
single-GPU solve of the same graph within the C4 tolerances (chi2 rel 1e-7, poses 1e-6)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_slam_b200 import abi, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    spec = synth.make_config('C4', seed=1, scale=float(os.environ.get('FG_MG_SCALE', '0.1')))
    L = len(spec['point_init'])
    ctx = abi.Context(device=local, rank=rank, nranks=world)
    uid = torch.zeros(128, dtype=torch.uint8, device='cuda')
    if rank == 0:
        uid = torch.frombuffer(bytearray(abi.comm_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(uid, 0)
    ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
    if os.environ.get('FG_MG_SHARD', 'blocks') == 'slice':
        ids = np.arange(L * rank // world, L * (rank + 1) // world)
    else:
        ids = abi.shard_landmarks(L, rank, world)                                         # what bench.py uses
    abi.load_spec(ctx, spec, landmark_ids=ids)
    e0 = ctx.error()
    rep = ctx.optimize()
    poses = ctx.get_values(abi.T_POSE)
    pts = ctx.get_values(abi.T_POINT)
    ok = True
    if rank == 0:
        ref = abi.Context(device=local)
        abi.load_spec(ref, spec)
        r0 = ref.error()
        rrep = ref.optimize()
        rposes = ref.get_values(abi.T_POSE)
        rpts = ref.get_values(abi.T_POINT)[ids]
        d_err0 = abs(e0 - r0) / r0
        d_err = abs(rep.final_error - rrep.final_error) / rrep.final_error
        d_pose = np.abs(poses - rposes).max()
        d_pts = np.abs(pts - rpts).max()
        same_trace = [t['accepted'] for t in rep.trace()] == [t['accepted'] for t in rrep.trace()]
        ok = d_err0 < 1e-10 and d_err < 1e-7 and d_pose < 1e-6 and d_pts < 1e-5 and same_trace and rep.iterations == rrep.iterations
        print('MULTI_GPU_PARITY world=%d iterations %d/%d err0 rel %.2e final rel %.2e pose %.2e points %.2e trace_equal %s -> %s'
              % (world, rep.iterations, rrep.iterations, d_err0, d_err, d_pose, d_pts, same_trace, 'OK' if ok else 'FAIL'))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
