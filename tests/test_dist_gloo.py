"""N > 1 host logic on CPU: world_size-2 gloo run of the particle split ([rank::size], Np_max),
rank-local weight normalisation (Q7) and the single reduce-to-root of the spectra, with the
oracle standing in for the per-rank integration (the CUDA path needs a GPU; its multi-GPU run is
bench.py --gpus N / the driver's scaling step).  Rendezvous on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import reference_path as rp
    from synchrad_b200 import host
    from synchrad_b200.dist import reduce_to_root
    tracks, dt, info = cases.undulator_tracks(5, seed=7)
    for i, t in enumerate(tracks):
        t[6] = 1.0 + i
    args = cases.undulator_args(info, grid=(24, 3, 2))
    idx = host.select_tracks(len(tracks), 4, rank, world)              # Np_max = 4
    mine = [tracks[i] for i in idx]
    w = host.normalized_weights([t[6] for t in mine], 'mean')          # rank-local list
    mine = [t[:6] + [float(wi)] + t[7:] for t, wi in zip(mine, w)]
    res = rp.calculate_spectrum(args, mine, dt, comp='cartesian')      # stand-in integrator
    tens = [torch.from_numpy(res['radiation'][k]) for k in ('x', 'y', 'z')]
    cnt = torch.tensor([res['passed'], res['updates']], dtype=torch.int64)
    out, tw, c = reduce_to_root(dist, tens, float(np.sum(w)), cnt)
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), x=out[0].numpy(), y=out[1].numpy(), z=out[2].numpy(),
             tw=np.array([np.nan if tw is None else tw]), cnt=c.numpy())
    dist.destroy_process_group()


def test_two_rank_split_and_reduce(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / 'rank0.npz')
    r1 = np.load(tmp_path / 'rank1.npz')
    tracks, dt, info = cases.undulator_tracks(5, seed=7)
    for i, t in enumerate(tracks):
        t[6] = 1.0 + i
    args = cases.undulator_args(info, grid=(24, 3, 2))
    ref = oracle.calculate_spectrum(args, tracks, dt, comp='cartesian', Np_max=4, weights_normalize='mean', ranks=2)
    for k in ('x', 'y', 'z'):
        np.testing.assert_allclose(r0[k], ref['radiation'][k], rtol=1e-13, atol=0)
        assert not r1[k].any()                       # non-root ranks end with zeros (calc.py:563-568)
    assert r0['tw'][0] == ref['total_weight'] and np.isnan(r1['tw'][0])
    assert r0['cnt'][0] == ref['passed'] and r0['cnt'][1] == ref['updates']


class _Patch:
    """monkeypatch.setattr for a spawned worker (nothing to undo: the process ends)."""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)


def _worker_class(rank, world, port, out_dir):
    """The product class itself, `SynchRad(ctx='mpi')`, on two gloo ranks; the emulated device (tests/emu/fake_engine)
    stands in for the GPU."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from emu import fake_engine
    fake_engine.install(_Patch)
    from synchrad.calc import SynchRad
    tracks, dt, info = cases.undulator_tracks(5, seed=7, Periods=6)
    for i, t in enumerate(tracks):
        t[6] = 1.0 + i
    tracks[1] = [np.asarray(a)[:100].copy() for a in tracks[1][:6]] + [tracks[1][6]]
    args = cases.undulator_args(info, grid=(24, 3, 2))
    out = {}
    for name, extra in (('round_robin', {}), ('balanced', {'partition': 'balanced'})):
        calc = SynchRad({**args, 'ctx': 'mpi', **extra})          # creates the (gloo) process group on first use
        assert (calc.rank, calc.size) == (rank, world)
        calc.calculate_spectrum([list(t) for t in tracks], timeStep=dt, comp='cartesian', Np_max=4,
                                weights_normalize='mean', it_range=(0, 200), verbose=False)
        for k in 'xyz':
            out[f'{name}_{k}'] = calc.Data['radiation'][k]
        out[f'{name}_tw'] = np.array([np.nan if calc.total_weight is None else calc.total_weight])
        out[f'{name}_upd'] = np.array([calc.last_run['updates'], calc.last_run['passed_updates']])
    np.savez(os.path.join(out_dir, f'class_rank{rank}.npz'), **out)
    dist.destroy_process_group()


def test_two_rank_synchrad_class_on_the_emulated_device(tmp_path, oracle):
    world = 2
    mp.spawn(_worker_class, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / 'class_rank0.npz'), np.load(tmp_path / 'class_rank1.npz')
    tracks, dt, info = cases.undulator_tracks(5, seed=7, Periods=6)
    for i, t in enumerate(tracks):
        t[6] = 1.0 + i
    tracks[1] = [np.asarray(a)[:100].copy() for a in tracks[1][:6]] + [tracks[1][6]]
    args = cases.undulator_args(info, grid=(24, 3, 2))
    ref = oracle.calculate_spectrum(args, tracks, dt, comp='cartesian', Np_max=4, weights_normalize='mean', ranks=2,
                                    it_range=(0, 200))
    for k in 'xyz':
        np.testing.assert_allclose(r0[f'round_robin_{k}'], ref['radiation'][k], rtol=1e-9, atol=1e-9 * ref['radiation'][k].max())
        assert not r1[f'round_robin_{k}'].any() and not r1[f'balanced_{k}'].any()      # calc.py:563-568
        assert r0[f'balanced_{k}'].any()
    assert r0['round_robin_tw'][0] == pytest.approx(ref['total_weight'], rel=1e-15) and np.isnan(r1['round_robin_tw'][0])
    # root reports the job's totals (updates, guard-passed updates), whatever the partition; the balanced partition hands
    # out contiguous slices of about equal sum(n - 1): another split, so the rank-local 'mean' normalisation (Q7) gives
    # other weights and a genuinely different spectrum
    assert r0['balanced_upd'][0] == r0['round_robin_upd'][0] == ref['updates'] == (3 * 199 + 99) * 24 * 3 * 2
    assert r0['round_robin_upd'][1] == ref['passed']
    assert r1['round_robin_upd'][0] == 0
    assert not np.allclose(r0['balanced_x'], r0['round_robin_x'])
