"""Host flow of `SynchRad.calculate_spectrum` (synchrad_b200/calc.py) in the `not gpu` suite: the class runs unchanged,
with srb_integrate replaced by the CPU emulation of the device code (tests/emu/fake_engine.py), and is compared with
the oracle's restatement of the reference's flow (calc.py:101-290 there).  The `-m gpu` tests check the same calls on the
real library; these keep the kwargs / precedence / batching / file logic covered where no GPU exists."""
import os

import numpy as np
import pytest

import cases
from conftest import rel_errors
from emu import fake_engine


@pytest.fixture
def emulated_device(monkeypatch):
    fake_engine.install(monkeypatch)
    from synchrad.calc import SynchRad
    return SynchRad


def close(calc, ref_rad, tol=1e-9):
    assert list(calc.Data['radiation']) == list(ref_rad)
    for key, ref in ref_rad.items():
        got = calc.Data['radiation'][key]
        assert got.dtype == np.float64 and got.shape == ref.shape and got.flags.c_contiguous
        assert max(rel_errors(got, ref)) <= tol, (key, rel_errors(got, ref))


def small(n=3, seed=4, grid=(40, 4, 3), **kw):
    tracks, dt, info = cases.undulator_tracks(n, seed=seed, Periods=6, **kw)
    return tracks, dt, cases.undulator_args(info, grid=grid, near=kw.get('near', False))


@pytest.mark.parametrize('comp', ['total', 'cartesian', 'cartesian_complex', 'spheric', 'spheric_complex'])
def test_list_of_tracks_all_components(emulated_device, oracle, comp):
    tracks, dt, args = small()
    calc = emulated_device(dict(args))
    calc.calculate_spectrum([list(t) for t in tracks], timeStep=dt, comp=comp, nSnaps=2, sigma_particle=1e-3, verbose=False)
    ref = oracle.calculate_spectrum(args, tracks, dt, comp=comp, nSnaps=2, sigma_particle=1e-3)
    close(calc, ref['radiation'])
    assert calc.total_weight == ref['total_weight']
    np.testing.assert_array_equal(calc.snap_iterations, ref['snap_iterations'])
    assert calc.last_run['passed_updates'] == ref['passed'] and calc.last_run['updates'] == ref['updates']
    assert calc.Args['comp'] == comp and calc.Data['FormFactor'].shape == (40,)


def test_near_field_needs_and_keeps_l_screen(emulated_device, oracle):
    tracks, dt, args = small(near=True, grid=(24, 5, 3))
    calc = emulated_device(dict(args))
    with pytest.raises(ValueError, match='L_screen'):
        calc.calculate_spectrum(tracks, timeStep=dt, verbose=False)
    calc.calculate_spectrum(tracks, timeStep=dt, L_screen=50.0, comp='cartesian', verbose=False)
    ref = oracle.calculate_spectrum(args, tracks, dt, L_screen=50.0, comp='cartesian')
    close(calc, ref['radiation'])
    np.testing.assert_array_equal(calc.Args['theta'], np.arctan2(calc.Args['radius'], 50.0))
    calc.calculate_spectrum(tracks, timeStep=dt, comp='cartesian', verbose=False)        # L_screen is remembered
    close(calc, ref['radiation'])
    with pytest.raises(AttributeError):                                                  # calc.py:342: no such kernel
        calc.calculate_spectrum(tracks, timeStep=dt, comp='spheric', verbose=False)


def test_selection_weights_and_ranges(emulated_device, oracle):
    tracks, dt, args = small(n=5, seed=2)
    tracks = [t[:6] + [w] + [s] for t, w, s in zip(tracks, (1.0, 2.5, 0.5, 4.0, 1.5), (0, 3, 0, 9, 1))]
    n = len(tracks[0][0])
    for kw in (dict(Np_max=3), dict(weights_normalize='mean'), dict(weights_normalize='max', it_range=(2, n - 7), nSnaps=3),
               dict(weights_normalize='ones', Np_max=4, it_range=(0, n + 9))):
        calc = emulated_device(dict(args))
        mine = [list(t) for t in tracks]
        calc.calculate_spectrum(mine, timeStep=dt, verbose=False, **kw)
        ref = oracle.calculate_spectrum(args, [list(t) for t in tracks], dt, **kw)
        close(calc, ref['radiation'])
        assert calc.total_weight == pytest.approx(ref['total_weight'], rel=1e-15), kw
        np.testing.assert_array_equal(calc.snap_iterations, ref['snap_iterations'])
    # per-track ranges (no it_range anywhere): snapshots of the LAST processed track stay on the object (calc.py:297-301)
    ragged = [[np.asarray(a)[:m].copy() for a in t[:6]] + [t[6]] for t, m in zip(tracks, (n, n - 40, n - 11, 60, n - 3))]
    calc = emulated_device(dict(args))
    calc.calculate_spectrum([list(t) for t in ragged], timeStep=dt, nSnaps=4, verbose=False)
    ref = oracle.calculate_spectrum(args, ragged, dt, nSnaps=4)
    close(calc, ref['radiation'])
    np.testing.assert_array_equal(calc.snap_iterations, ref['snap_iterations'])
    np.testing.assert_array_equal(calc.snap_iterations, np.linspace(0, n - 3, 5, dtype=np.uint32)[1:])


def test_argument_errors(emulated_device):
    tracks, dt, args = small(n=1)
    calc = emulated_device(dict(args))
    with pytest.raises(ValueError, match='timeStep'):
        calc.calculate_spectrum(tracks, verbose=False)
    with pytest.raises(ValueError, match='comp'):
        calc.calculate_spectrum(tracks, timeStep=dt, comp='everything', verbose=False)
    with pytest.raises(ValueError, match='nSnaps'):
        calc.calculate_spectrum(tracks, timeStep=dt, nSnaps=0, verbose=False)
    bad = [list(tracks[0])]
    bad[0][2] = bad[0][2][:-1]
    with pytest.raises(ValueError, match='differ in length'):
        calc.calculate_spectrum(bad, timeStep=dt, verbose=False)
    calc.calculate_spectrum([], timeStep=dt, verbose=False)          # no tracks: zero spectrum, zero weight
    assert calc.total_weight == 0.0 and not calc.Data['radiation']['total'].any()
    no_device = emulated_device({**args, 'ctx': False})
    with pytest.raises(RuntimeError, match='without a device'):
        no_device.calculate_spectrum(tracks, timeStep=dt)


def test_batches_forced_and_pipelined(emulated_device, oracle, monkeypatch, capsys):
    from synchrad_b200 import host
    tracks, dt, args = small(n=9, seed=6)
    n = len(tracks[0][0])
    ref = oracle.calculate_spectrum(args, tracks, dt, comp='cartesian_complex', nSnaps=2)
    one = emulated_device(dict(args))
    one.calculate_spectrum([list(t) for t in tracks], timeStep=dt, comp='cartesian_complex', nSnaps=2, verbose=False)
    close(one, ref['radiation'])
    forced = emulated_device({**args, 'max_batch_bytes': 96 * 2 * n + 96})                # two tracks per batch
    forced.calculate_spectrum([list(t) for t in tracks], timeStep=dt, comp='cartesian_complex', nSnaps=2, verbose=False)
    monkeypatch.setattr(host, 'PIPELINE_BATCH_BYTES', 48 * int(2.5 * n))                  # 4 balanced shares: 3, 2, 2, 2
    piped = emulated_device(dict(args))
    piped.calculate_spectrum([list(t) for t in tracks], timeStep=dt, comp='cartesian_complex', nSnaps=2, verbose=False)
    assert (one.last_run['batches'], forced.last_run['batches'], piped.last_run['batches']) == (1, 5, 4)
    # verbose: the reference's per-particle tqdm bar, here over the tracks of the finished batches
    loud = emulated_device(dict(args))
    loud.calculate_spectrum([list(t) for t in tracks], timeStep=dt, comp='cartesian_complex', nSnaps=2)
    err = capsys.readouterr().err
    assert '9/9' in err and loud.last_run['batches'] == 4
    for calc in (forced, piped, loud):
        close(calc, ref['radiation'])
        assert calc.total_weight == one.total_weight
        assert calc.last_run['passed_updates'] == one.last_run['passed_updates'] == ref['passed']
        assert calc.last_run['h2d_bytes'] >= one.last_run['h2d_bytes']


def test_tracks_file_to_spectrum_file(emulated_device, oracle, tmp_path):
    """tutorials/PIC/compute_spectrum.py:16-18 and the analysis-only re-load (calc.py:98-99)."""
    from synchrad_b200 import trackio
    tracks, dt, args = small(n=5, seed=9, grid=(32, 4, 4))
    tracks = [t[:7] + [s] for t, s in zip(tracks, [0, 3, 0, 11, 2])]
    n = len(tracks[0][0])
    ftr, fsp = str(tmp_path / 'tracks.h5'), str(tmp_path / 'spectrum.h5')
    trackio.write_tracks(ftr, tracks, cdt=dt, it_range=(0, n + 20))
    calc = emulated_device(dict(args))
    calc.calculate_spectrum(file_tracks=ftr, file_spectrum=fsp, timeStep=123.0, comp='cartesian', nSnaps=2, Np_max=4,
                            verbose=False)
    ref = oracle.calculate_spectrum(args, tracks, dt, comp='cartesian', nSnaps=2, Np_max=4, it_range=(0, n + 20))
    close(calc, ref['radiation'])
    assert float(calc.Args['timeStep']) == dt                       # misc/cdt overrides the kwarg (calc.py:189)
    assert calc.last_run['file_read_s'] > 0 and calc.last_run['host_pack_s'] >= calc.last_run['file_read_s']
    loaded = emulated_device(file_spectrum=fsp)
    for k in 'xyz':
        np.testing.assert_array_equal(loaded.Data['radiation'][k], calc.Data['radiation'][k])
    np.testing.assert_array_equal(loaded.snap_iterations, calc.snap_iterations)
    assert loaded.total_weight == calc.total_weight and loaded.Args['comp'] == 'cartesian'
    assert loaded.get_energy(lambda0_um=1) == pytest.approx(calc.get_energy(lambda0_um=1), rel=1e-14)
    # kwarg it_range beats the file's (calc.py:192-197); without either: per-track ranges (calc.py:199-201)
    calc.calculate_spectrum(file_tracks=ftr, it_range=(5, n - 5), verbose=False)
    close(calc, oracle.calculate_spectrum(args, tracks, dt, it_range=(5, n - 5))['radiation'])
    trackio.write_tracks(ftr, tracks, cdt=dt)
    calc.calculate_spectrum(file_tracks=ftr, verbose=False)
    close(calc, oracle.calculate_spectrum(args, tracks, dt)['radiation'])


def test_converter_to_spectrum(emulated_device, oracle, tmp_path):
    """Time series with particles entering / leaving -> tracksFromOPMD -> tracks file -> calculate_spectrum."""
    from golden import converter_cases as cc
    from synchrad.utils import tracksFromOPMD
    from synchrad_b200 import trackio
    tracks, dt, args = small(n=4, seed=4, grid=(32, 4, 3))
    n_it = len(tracks[0][0])
    series = {v: np.array([t[k] for t in tracks]).T.copy() for k, v in enumerate(('x', 'y', 'z', 'ux', 'uy', 'uz'))}
    series['w'] = np.tile([t[6] for t in tracks], (n_it, 1)).astype(np.double)
    for v in series:
        series[v][:40, 1] = np.nan
        series[v][150:, 2] = np.nan
        series[v][100:104, 3] = np.nan
    data = [{v: series[v][k] for v in series} for k in range(n_it)]
    ts = cc.FakeTimeSeries(data, np.arange(n_it), np.arange(n_it) * (dt / cc.C))
    ts.n_all = 4
    ftr = str(tmp_path / 'tracks.h5')
    tracksFromOPMD(ts, cc.FakeTracker(ts, species='e'), 0, fname=ftr)
    cdt, rng, n = trackio.read_header(ftr)
    assert n == 5 and rng == (0, n_it)
    lists = trackio.read_tracks(ftr, range(n))
    calc = emulated_device(dict(args))
    calc.calculate_spectrum(file_tracks=ftr, nSnaps=3, verbose=False)
    close(calc, oracle.calculate_spectrum(args, lists, cdt, nSnaps=3, it_range=rng)['radiation'])
