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


# ------------------------------------------------------------------------------- BASELINE configs[3] (C4, spiral beam)
from golden.make_reference_golden import C4_POST, c4_cases  # noqa: E402

C4 = c4_cases()


@pytest.mark.parametrize('name', sorted(C4))
def test_c4_spiral_matches_reference_vectors(cuda_lib, name):
    """tutorials/Spiral_Beam_Part1.ipynb recipe (SI units) against the unmodified reference's outputs
    (tests/golden/reference_c4.npz): double 1e-9 on every uniform-grid kernel; single precision in the north star's
    literal sense (`float_mode='literal'` vs the reference's fp32 output, 1e-4); the default mixed-precision float
    mode is judged against the reference's DOUBLE output -- it must be closer to it than the reference's own fp32
    run is, and within 1e-4 of it."""
    stored = np.load(os.path.join(GOLD, 'reference_c4.npz'))
    meta = json.load(open(os.path.join(GOLD, 'reference_c4_meta.json')))[name]
    args, tracks, dt, kw = C4[name]
    ref = {k: stored[f'{name}/{k}'] for k in meta['keys']}
    if args['dtype'] == 'float':
        a = dict(args)
        a['float_mode'] = 'literal'
        calc = run_gpu(a, tracks, dt, **kw)
        assert calc.last_run['kernel'] == 'literal'
        e = field_errors(calc.Data['radiation'], ref)
        assert max(e) <= 1e-4, (name, 'literal', e)
        if name == 'c4_float_total':
            ref64 = {'total': stored['c4_double_total/total']}
            e_ref32 = field_errors(ref, ref64)                       # how far the reference's own fp32 run is
            for phasor in ('auto', 'recur', 'direct'):
                mixed = run_gpu(args, tracks, dt, phasor=phasor, **kw)
                e = field_errors(mixed.Data['radiation'], ref64)
                assert e[0] <= e_ref32[0] and e[1] <= e_ref32[1], (phasor, e, e_ref32)
                assert max(e) <= 1e-4, (phasor, e)
        return
    for phasor in ('auto', 'pair', 'recur', 'direct'):
        calc = run_gpu(args, tracks, dt, phasor=phasor, **kw)
        e = field_errors(calc.Data['radiation'], ref)
        assert max(e) <= 1e-9, (name, phasor, e, calc.last_run)
    for i, (meth, pkw) in enumerate(C4_POST):
        want = stored[f'{name}/post{i}']
        np.testing.assert_allclose(getattr(calc, meth)(**pkw), want, rtol=0, atol=1e-9 * np.abs(want).max())


def test_c4_coherent_gain_matches_the_reference(cuda_lib):
    """'Enhancement due to coherency' (Spiral_Beam_Part1.ipynb:283-287) of the stored reference run, from the GPU."""
    meta = json.load(open(os.path.join(GOLD, 'reference_c4_meta.json')))
    e = {}
    for name in ('c4_double_coherent', 'c4_double_total'):
        args, tracks, dt, kw = C4[name]
        e[name] = run_gpu(args, tracks, dt, **kw).get_energy(**C4_POST[0][1])
    assert e['c4_double_coherent'] / e['c4_double_total'] == pytest.approx(meta['_coherent_gain_double'], rel=1e-9)


def test_c4_full_size_single_precision(cuda_lib):
    """BASELINE configs[3] at full size: 10^4 spiral-beam particles x 192 samples, 512x64x64 grid, dtype='float'
    (4.0e12 updates).  The oracle cannot do the full grid in test time, so the check is on the (theta, phi) nodes
    {0, 63} x {0, 32}, which coincide bit for bit with the nodes of a 512x2x2 grid (linspace end points; phi = 0, pi):
    literal mode against the strict fp32 oracle (1e-4), mixed mode against the strict fp64 oracle (1e-4)."""
    from oracle import reference_path as rp
    rp.build()
    tracks, dt, info = cases.spiral_tracks(10000, seed=0)
    full32 = cases.spiral_args(info)
    assert tuple(full32['grid'][-1]) == (512, 64, 64) and full32['dtype'] == 'float'
    sub = lambda rad: {k: v[:, :, ::63, ::32] for k, v in rad.items()}
    mixed = run_gpu(full32, tracks, dt)
    assert mixed.last_run['updates'] == 10000 * 191 * 512 * 64 * 64
    r64 = rp.calculate_spectrum(cases.spiral_args(info, grid=(512, 2, 2), dtype='double'), tracks, dt)['radiation']
    e = field_errors(sub(mixed.Data['radiation']), r64)
    assert max(e) <= 1e-4, ('mixed vs fp64 oracle', e)
    lit_args = dict(full32)
    lit_args['float_mode'] = 'literal'
    lit = run_gpu(lit_args, tracks, dt)
    r32 = rp.calculate_spectrum(cases.spiral_args(info, grid=(512, 2, 2), dtype='float'), tracks, dt)['radiation']
    e = field_errors(sub(lit.Data['radiation']), r32)
    assert max(e) <= 1e-4, ('literal vs fp32 oracle', e)
    print(f"C4 full size: mixed {mixed.last_run['integrate_ms']:.0f} ms ({mixed.last_run['kernel']}), "
          f"literal {lit.last_run['integrate_ms']:.0f} ms")


def test_spiral_notebook_coherent_enhancement(cuda_lib):
    """The notebook's own experiment at its own size (Spiral_Beam_Part1.ipynb:255-287): 16 000 particles, grid
    1024x32x32, incoherent `total` over the first 1000 tracks and coherent `cartesian_complex` over all of them;
    it prints 'Enhancement due to coherency 12.5'.  Tracks here come from a seeded leap-frog instead of the notebook's
    unseeded Radau runs, so the figure is reproduced approximately, not digit for digit."""
    tracks, dt, info = cases.spiral_tracks(16000, seed=0)
    args = cases.spiral_args(info, grid=(1024, 32, 32), dtype='double')
    incoh = run_gpu(args, tracks, dt, Np_max=1000)
    coh = run_gpu(args, tracks, dt, comp='cartesian_complex')
    gain = coh.get_energy(normalize_to_weights=True, lambda0_um=1e6) / incoh.get_energy(normalize_to_weights=True, lambda0_um=1e6)
    print(f'Enhancement due to coherency {gain:.1f} (notebook: 12.5)')
    assert 9.0 < gain < 16.0, gain
