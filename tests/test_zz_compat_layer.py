"""The reference-side binding (synchrad_b200/compat): a `pyopencl` look-alike whose kernels marshal the
reference's per-particle positional argument list (calc.py:306-353) into the C ABI.

  * no GPU, reference present (build container): the UNMODIFIED reference (`/root/reference/synchrad/calc.py`)
    runs on top of the binding with the CPU emulation of the kernels standing in for `srb_integrate_host`, and must
    reproduce the stored outputs of the reference itself;
  * GPU box: the oracle's launch loop hands the same argument lists to the binding over the real library.
(Named test_zz_* so that it runs after every other test file.)
"""
import json
import os

import numpy as np
import pytest

from golden.make_golden import small_cases
from golden.make_reference_golden import extra_cases
from oracle import run_reference

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
ALL = dict(small_cases())
ALL.update(extra_cases())


def field_error(got, ref):
    scale = max(np.abs(r).max() for r in ref.values())
    return max(np.abs(got[k] - r).max() for k, r in ref.items()) / scale


@pytest.mark.skipif(not run_reference.available(), reason='/root/reference is not present (GPU box)')
def test_unmodified_reference_runs_on_the_binding():
    from emu import emu
    emu.build()
    stored = np.load(os.path.join(GOLD, 'reference_cases.npz'))
    meta = json.load(open(os.path.join(GOLD, 'reference_cases_meta.json')))
    names = sorted(ALL)
    out = run_reference.run_many([dict(args=ALL[n][0], tracks=ALL[n][1], kw=dict(ALL[n][3], timeStep=ALL[n][2]))
                                  for n in names], backend='compat_emu')
    for n, res in zip(names, out):
        assert 'error' not in res, (n, res.get('error'))
        assert 'CUDA' in res['device']                      # the reference saw the binding's device, not clshim's
        ref = {k: stored[f'{n}/{k}'] for k in meta[n]['keys']}
        assert list(res['radiation']) == meta[n]['keys']
        single = ALL[n][0].get('dtype') == 'float'
        # single precision: the binding selects the all-fp32 reproduction and adds in float32 like the reference
        assert field_error(res['radiation'], ref) <= (1e-6 if single else 1e-12), n
        assert res['total_weight'] == meta[n]['total_weight']


def test_binding_rejects_wrong_argument_lists():
    from synchrad_b200.compat import pyopencl as cl
    from synchrad_b200.compat.pyopencl import array as arr
    prog = cl.Program(None, '__kernel void total(__global double *spectrum)').build()
    assert hasattr(prog, 'total') and not hasattr(prog, 'cartesian_comps')
    with pytest.raises(TypeError):
        prog.total(None, (32,), (32,), arr.zeros(None, (4,), np.double).data)
    with pytest.raises(NotImplementedError):
        cl.Program(None, '__kernel void something_else(__global double *x)')
    near = cl.Program(None, '__kernel void total(__global float *spectrum, float distanceToScreen)')
    assert near.mode == 'near' and near.dtype == 'float'
    assert cl.device_type.to_string(cl.create_some_context(answers=[0, 3]).devices[0].type) == 'GPU'
    assert cl.create_some_context(answers=[0, 3]).devices[0].index == 3


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['far_cartesian_snaps', 'far_cartesian_complex', 'far_spheric_complex', 'near_total',
                                  'near_cartesian_complex', 'opt_it_start_late', 'wiggler_loggrid', 'float_far_total'])
def test_binding_on_the_gpu(cuda_lib, oracle, name):
    stored = np.load(os.path.join(GOLD, 'reference_cases.npz'))
    meta = json.load(open(os.path.join(GOLD, 'reference_cases_meta.json')))[name]
    args, tracks, dt, kw = ALL[name]
    res = oracle.calculate_spectrum(args, tracks, dt, lib='compat', **kw)
    ref = {k: stored[f'{name}/{k}'] for k in meta['keys']}
    assert field_error(res['radiation'], ref) <= (1e-4 if args.get('dtype') == 'float' else 1e-9), name


def test_binding_glue_with_emulated_kernels(oracle, monkeypatch):
    """The GPU test's own path -- oracle launch loop -> binding -> `srb_integrate_host` -- with the CPU emulation of
    the kernels substituted for the library call (no reference needed)."""
    import ctypes
    from emu import emu
    from synchrad_b200.compat import pyopencl as cl
    emu.build()
    lib = ctypes.CDLL(emu._SO)
    lib.srb_emu_integrate.restype = ctypes.c_int

    def emu_integrate_host(grid, tracks, spectra_ptrs, n_spectra, device):
        tw = next((o for o in (2, 4, 8) if 32 * o >= grid.nOmega), 8)
        cnt = (ctypes.c_ulonglong * 2)(0, 0)
        assert lib.srb_emu_integrate(ctypes.byref(grid), ctypes.byref(tracks), spectra_ptrs, n_spectra, 0, tw,
                                     ctypes.c_uint32(1), cnt, ctypes.c_int(1), ctypes.c_uint32(1)) == 0
    monkeypatch.setattr(cl, '_integrate_host', emu_integrate_host)
    stored = np.load(os.path.join(GOLD, 'reference_cases.npz'))
    meta = json.load(open(os.path.join(GOLD, 'reference_cases_meta.json')))
    for name in ('far_cartesian_snaps', 'far_spheric_complex', 'near_cartesian_complex', 'opt_it_start_late',
                 'wiggler_loggrid', 'float_far_total', 'float_near_total'):
        args, tracks, dt, kw = ALL[name]
        res = oracle.calculate_spectrum(args, tracks, dt, lib='compat', **kw)
        ref = {k: stored[f'{name}/{k}'] for k in meta[name]['keys']}
        assert field_error(res['radiation'], ref) <= (1e-6 if args.get('dtype') == 'float' else 1e-12), name
