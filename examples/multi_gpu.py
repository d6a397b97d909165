"""Particle-sharded run on all GPUs of one box (replaces `mpirun -n N python script.py` + 'ctx': 'mpi'):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/multi_gpu.py

Each rank integrates tracks[rank::size]; one NCCL reduce leaves the summed spectrum on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synchrad.calc import SynchRad             # noqa: E402
from synchrad_b200 import synthetic           # noqa: E402

# no explicit setup: with 'ctx': 'mpi' SynchRad joins the torchrun job itself (one NCCL group, one rank per
# GPU), the way the reference picks up MPI.COMM_WORLD
tracks = synthetic.batch_to_track_list(synthetic.c5_batch(64, 2000, seed=7))   # same list on every rank
args = synthetic.c5_args((256, 32, 32))
args['ctx'] = 'mpi'
calc = SynchRad(args)
calc.calculate_spectrum(tracks, timeStep=synthetic.C5_DT, comp='total')
if calc.rank == 0:
    S = calc.Data['radiation']['total']
    print(f'ranks={calc.size} total_weight={calc.total_weight} sum(S)={S.sum():.10e} '
          f'updates={calc.last_run["updates"]:.3e}')
if dist.is_initialized():
    dist.destroy_process_group()
