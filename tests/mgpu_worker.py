"""Worker of tests/test_multi_gpu.py (one process per rank, launched with torch.distributed.run):
`SynchRad(ctx='mpi')` -- the replacement of the reference's mpi4py split + Reduce (calc.py:212,236,560-571) --
on real GPUs.  With >= WORLD_SIZE GPUs: one rank per GPU over NCCL (the product's own auto-initialisation).
With fewer GPUs (the single-GPU test box): the ranks share cuda:0 and the process group is gloo, so the split,
the rank-local weight normalisation, the GPU integration and the reduce-to-root semantics still run on hardware."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import cases  # noqa: E402


def problem(balanced=False):
    tracks, dt = cases.c5_tracks_numpy(7, 600)
    for i, t in enumerate(tracks):
        t[6] = 1.0 + 0.25 * i
    tracks = [t[:7] + [s] for t, s in zip(tracks, (0, 3, 0, 7, 1, 0, 2))]
    args = cases.c5_args(grid=(256, 6, 4))
    if balanced:        # extension Args['partition'] = 'balanced': tracks of very different lengths, contiguous slices
        tracks = [[c[:n] for c in t[:6]] + t[6:] for t, n in zip(tracks, (600, 40, 90, 500, 30, 600, 75))]
        args['partition'] = 'balanced'
        return args, tracks, dt, dict(comp='cartesian', nSnaps=2, it_range=(0, 610))
    kw = dict(comp='cartesian', Np_max=6, weights_normalize='mean', nSnaps=2, it_range=(0, 610))
    return args, tracks, dt, kw


def main():
    out_dir = sys.argv[1]
    world = int(os.environ['WORLD_SIZE'])
    rank = int(os.environ['RANK'])
    nccl = torch.cuda.device_count() >= world
    if not nccl:
        os.environ['LOCAL_RANK'] = '0'
        torch.cuda.set_device(0)
        dist.init_process_group('gloo')
    from synchrad.calc import SynchRad
    args, tracks, dt, kw = problem(balanced=len(sys.argv) > 2 and sys.argv[2] == 'balanced')
    args['ctx'] = 'mpi'
    calc = SynchRad(args)                       # NCCL group created here when none exists (one rank per GPU)
    assert calc.size == world and calc.rank == rank
    calc.calculate_spectrum(tracks, timeStep=dt, verbose=False, **kw)
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'),
             tw=np.array([np.nan if calc.total_weight is None else calc.total_weight]),
             passed=np.array([calc.last_run['passed_updates'], calc.last_run['updates']]),
             backend=np.array([dist.get_backend()]), device=np.array([str(calc.device)]),
             **{k: v for k, v in calc.Data['radiation'].items()})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
