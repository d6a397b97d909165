"""TEST INFRASTRUCTURE ONLY — a CPU stand-in for synchrad_b200.engine, so that the host flow of
`SynchRad.calculate_spectrum` (kwargs, precedence rules, weights, batches, tracks files, snapshots, result containers)
runs in the `not gpu` suite: srb_integrate is replaced by the single-warp emulation of the device code (emu.run_packed),
device tensors by CPU torch tensors, streams and events by no-ops.  Installed by the `emulated_device` fixture of
tests/test_calc_flow_emulated.py; never imported by the product."""
import types

import numpy as np
import torch

from synchrad_b200 import engine as real_engine
from synchrad_b200 import host

from . import emu

KIND_ID = {'direct': 0, 'recur': 1, 'pair': 3, 'pair_fma': 4, 'drec': 5}


class _Event:
    def elapsed_time(self, other):
        return 0.0

    def synchronize(self):
        pass


def _choose(Args, comp, phasor):
    if phasor not in ('auto', None):
        return {'recur': 'recur', 'direct': 'direct', 'pair': 'pair', 'pair_fma': 'pair_fma', 'drec': 'drec'}[phasor]
    if not host.omega_is_uniform(Args) or host.float_mode(Args) == 'literal':
        return 'direct'
    if Args['mode'] == 'near':
        return 'drec'
    return 'recur' if comp.startswith('spheric') else 'pair'


def integrate(Args, dtype, grid, packed, comp, nSnaps, native=False, phasor='auto', counters=True, device_tracks=None,
              timing=False, timeStep=None, max_scratch_bytes=None, spectra=None, counters_into=None, upload_stream=None):
    if device_tracks is not None:
        raise NotImplementedError('fake engine: host tracks only')
    if Args['mode'] == 'near' and comp.startswith('spheric'):
        raise AttributeError(f"no {comp!r} kernel in {Args['mode']}-field mode")
    kind = _choose(Args, comp, phasor)
    n_w, n_2, n_p = (int(v) for v in Args['gridNodeNums'])
    keys = host.COMP_KEYS[comp]
    if spectra is None:
        spectra = [torch.zeros((int(nSnaps), n_p, n_2, n_w), dtype=torch.float64) for _ in keys]
    views = [s.numpy() for s in spectra]
    dt = float(Args['timeStep']) if (timeStep is None or dtype is np.double) else float(timeStep)
    cnt = (0, 0)
    if packed.n:
        _, cnt = emu.run_packed(Args, dtype, packed, dt, comp, int(nSnaps), kind=kind, nPC=1 + (packed.n > 3),
                                spectra=views)
    res = types.SimpleNamespace()
    res.spectra = spectra
    c = torch.tensor([int(cnt[0]), int(cnt[1])], dtype=torch.int64)
    if counters_into is not None:
        counters_into += c
        c = counters_into
    res.counters = c
    res.kind = KIND_ID[kind] if host.float_mode(Args) != 'literal' else 2
    res.info = types.SimpleNamespace(tile_width=0, n_particle_chunks=1 + (packed.n > 3), n_time_segments=1,
                                     grid_blocks=0, kernels_launched=1, kind=res.kind)
    res.events = (_Event(), _Event())
    res.elapsed_ms, res.updates = None, None
    return res


def to_host_layout(spectra, nSnaps, n_w, n_2, n_p):
    return [s.swapaxes(-1, -3).contiguous() for s in spectra]


def install(monkeypatch):
    """Route synchrad_b200.calc onto this module and neutralise the torch.cuda calls of the host flow."""
    monkeypatch.setattr(real_engine, 'require_cuda', lambda index: torch.device('cpu'))
    monkeypatch.setattr(real_engine, 'integrate', integrate)
    monkeypatch.setattr(real_engine, 'to_host_layout', to_host_layout)
    monkeypatch.setattr(torch.cuda, 'get_device_name', lambda *a, **k: 'emulated device (tests/emu)')
    monkeypatch.setattr(torch.cuda, 'get_device_capability', lambda *a, **k: (10, 0))
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, 'Stream', lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, 'mem_get_info', lambda *a, **k: (8 << 30, 8 << 30))
