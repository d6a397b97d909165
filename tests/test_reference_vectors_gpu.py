"""GPU parity against outputs of the UNMODIFIED reference (tests/golden/reference_cases.npz, reference_kat.json;
generated in the build container by tests/golden/make_reference_golden.py, see tests/test_reference_pin.py).

Tolerances (BASELINE.json north_star):
  double : 1e-9, max-norm and 2-norm, on the scale of the stored vector field (all components of a case
           together: a component that is pure rounding residue in the reference -- e.g. z at theta = 0,
           4e-34 against x = 6e7 -- has no digits of its own to match; the reference's own kernels move it by
           100 % when the compiler contracts FMAs);
  single : `float_mode='literal'` (every operation of the reference kernels in fp32) against the reference's fp32
           outputs, 1e-4 in the same norms.
"""
import json
import os

import numpy as np
import pytest

import cases
from golden.make_golden import small_cases
from golden.make_reference_golden import extra_cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
ALL = dict(small_cases())
ALL.update(extra_cases())


def run_gpu(args, tracks, dt, phasor='auto', **kw):
    from synchrad.calc import SynchRad
    a = dict(args)
    a['phasor'] = phasor
    calc = SynchRad(a)
    calc.calculate_spectrum([list(t) for t in tracks], timeStep=dt, verbose=False, **kw)
    return calc


def field_errors(got, ref):
    """(max-norm, 2-norm) error of a dict of components on the scale of the whole stored field."""
    scale = max(np.abs(r).max() for r in ref.values())
    n2 = np.sqrt(sum(np.linalg.norm(r) ** 2 for r in ref.values()))
    emax = max(np.abs(got[k] - r).max() for k, r in ref.items())
    e2 = np.sqrt(sum(np.linalg.norm(got[k] - r) ** 2 for k, r in ref.items()))
    return float(emax / scale), float(e2 / n2)


@pytest.mark.parametrize('name', sorted(ALL))
def test_cuda_path_matches_reference_vectors(cuda_lib, name):
    stored = np.load(os.path.join(GOLD, 'reference_cases.npz'))
    meta = json.load(open(os.path.join(GOLD, 'reference_cases_meta.json')))[name]
    args, tracks, dt, kw = ALL[name]
    single = args.get('dtype') == 'float'
    a = dict(args)
    if single:
        a['float_mode'] = 'literal'
    uniform = not args.get('Features')
    phasors = ('auto',) if single or not uniform else ('auto', 'recur', 'direct')
    ref = {k: stored[f'{name}/{k}'] for k in meta['keys']}
    for phasor in phasors:
        calc = run_gpu(a, tracks, dt, phasor=phasor, **kw)
        assert list(calc.Data['radiation']) == meta['keys']
        e = field_errors(calc.Data['radiation'], ref)
        assert max(e) <= (1e-4 if single else 1e-9), (name, phasor, e, calc.last_run)
        assert calc.total_weight == pytest.approx(meta['total_weight'], rel=1e-15)
        if 'it_range' in kw:
            np.testing.assert_array_equal(np.asarray(calc.snap_iterations), stored[f'{name}/snap_iterations'])
    # utils.py post-processing (host and on-device integrals) against the reference's own utils.py results
    if not single:
        from golden.make_reference_golden import POST
        for i, (meth, pkw) in enumerate(POST):
            want = stored[f'{name}/post{i}']
            got = getattr(calc, meth)(**pkw)
            tol = 1e-9 * np.abs(want).max()
            np.testing.assert_allclose(got, want, rtol=0, atol=tol, err_msg=f'{name} {meth}')
            if meth in ('get_energy', 'get_energy_spectrum'):
                got_dev = getattr(calc, meth)(on_device=True, **pkw)
                np.testing.assert_allclose(got_dev, want, rtol=0, atol=tol, err_msg=f'{name} {meth} on_device')


@pytest.mark.parametrize('tag,near,Np', [('C1_far_double_24', False, 24), ('C2_near_double_2', True, 2)])
def test_reference_test_scripts_full_size(cuda_lib, tag, near, Np):
    """BASELINE configs[0] and configs[1]: the reference's own test scripts (tests/test_undulator_analytic.py,
    ..._near.py; RNG seeded) at their full grids, against summary values of the reference's run."""
    kat = json.load(open(os.path.join(GOLD, 'reference_kat.json')))[tag]
    tracks, dt, info = cases.undulator_tracks(Np, near=near, seed=0)
    args = cases.undulator_args(info, near=near)
    kw = dict(L_screen=1e5) if near else {}
    calc = run_gpu(args, tracks, dt, comp='total', Np_max=Np, **kw)
    S = calc.Data['radiation']['total'][0]
    # at theta = 0 all phi nodes hold the same value up to rounding: compare the value, not the phi index
    assert list(np.unravel_index(S.argmax(), S.shape))[:2] == kat['argmax'][:2]
    assert abs(S[tuple(kat['argmax'])] - S.max()) <= 1e-9 * S.max()
    np.testing.assert_allclose(S.max(), kat['max'], rtol=1e-9)
    np.testing.assert_allclose(S.sum(), kat['sum'], rtol=1e-9)
    np.testing.assert_allclose(np.linalg.norm(S), kat['l2'], rtol=1e-9)
    for idx, val in kat['spots']:
        assert abs(S[tuple(idx)] - val) <= 1e-9 * kat['max'], (idx, S[tuple(idx)], val)
    E = calc.get_energy(lambda0_um=1)
    np.testing.assert_allclose(E, kat['energy_J'], rtol=1e-9)
    from synchrad.utils import J_in_um
    Et = cases.undulator_energy_theory(info, J_in_um)
    assert abs(abs(E - Et) / Et * 100 - kat['deviation_percent']) < 1e-6
    assert abs(E - Et) / Et < (0.12 if near else 0.02)      # the scripts' own (printed) analytic criterion
