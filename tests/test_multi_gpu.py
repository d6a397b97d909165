"""On-hardware parity of the N > 1 path: `SynchRad(ctx='mpi')` under a 2-rank torch.distributed launch
(tests/mgpu_worker.py) against the oracle's emulation of a 2-rank mpirun of the reference
(calc.py:212,236,241-250,560-571: tracks[:Np][rank::size], rank-local weight normalisation, Reduce(SUM) to root,
zeros and total_weight=None on the other ranks)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import rel_errors

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(tmp_path, *extra):
    env = dict(os.environ)
    env.pop('OMP_NUM_THREADS', None)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(HERE, 'mgpu_worker.py'), str(tmp_path), *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-3000:]
    return np.load(tmp_path / 'rank0.npz'), np.load(tmp_path / 'rank1.npz')


def test_two_ranks_balanced_partition(cuda_lib, oracle, tmp_path):
    """Args['partition'] = 'balanced' (contiguous slices of equal work) over two ranks: the same sum over
    particles as a single-process run, up to the summation order."""
    import mgpu_worker
    r0, r1 = _launch(tmp_path, 'balanced')
    args, tracks, dt, kw = mgpu_worker.problem(balanced=True)
    a = dict(args)
    a.pop('partition')
    ref = oracle.calculate_spectrum(a, tracks, dt, **kw)
    for k in ('x', 'y', 'z'):
        assert max(rel_errors(r0[k], ref['radiation'][k])) <= 1e-9
        assert not r1[k].any()
    assert r0['tw'][0] == pytest.approx(ref['total_weight'], rel=1e-14)
    assert r0['passed'][0] == ref['passed']


def test_two_ranks_on_gpus_match_the_oracle(cuda_lib, oracle, tmp_path):
    import torch
    import mgpu_worker
    r0, r1 = _launch(tmp_path)
    want_backend = 'nccl' if torch.cuda.device_count() >= 2 else 'gloo'
    assert str(r0['backend'][0]) == want_backend
    if want_backend == 'nccl':
        assert str(r0['device'][0]) != str(r1['device'][0])          # one rank per GPU
    args, tracks, dt, kw = mgpu_worker.problem()
    ref = oracle.calculate_spectrum(args, tracks, dt, ranks=2, **kw)
    for k in ('x', 'y', 'z'):
        assert max(rel_errors(r0[k], ref['radiation'][k])) <= 1e-9, (k, rel_errors(r0[k], ref['radiation'][k]))
        assert r1[k].shape == r0[k].shape and not r1[k].any()        # non-root ranks end with zeros (calc.py:563-568)
    assert r0['tw'][0] == pytest.approx(ref['total_weight'], rel=1e-14) and np.isnan(r1['tw'][0])
    assert r0['passed'][0] == ref['passed']                          # guard decisions of both ranks, summed on root
    print(f'2 ranks, backend {want_backend}: parity ok')
