"""Kernel LOGIC against the oracle on the CPU: the device code of srb_core.cuh compiled by g++ and
run one warp at a time (tests/emu, test infrastructure only).  Covers what can go wrong
independently of the hardware: flush chains, Nyquist ranges, phasor seeds and recurrences, tile /
chunk indexing, particle-chunk partial spectra.  The GPU parity tests (-m gpu) are the real gate."""
import numpy as np
import pytest

import cases
from conftest import rel_errors
from emu import emu


def check(oracle, args, tracks, dt, kinds=('direct', 'recur'), tol=1e-9, nPC=1, tw=None, **kw):
    ref = oracle.calculate_spectrum(args, tracks, dt, **kw)
    if 'recur' in kinds and args.get('mode', 'far') == 'far':
        kinds = tuple(kinds) + ('pair',)      # the symmetric-pair kernel covers the same cases
    for kind in kinds:
        rad, cnt = emu.run(args, tracks, dt, kind=kind, nPC=nPC, tw=tw if kind == 'recur' else None, **kw)
        for key, r in ref['radiation'].items():
            e = rel_errors(rad[key], r)
            limit = 1e-13 if kind == 'direct' else tol
            assert max(e) < limit, (kind, key, e)
    return ref, cnt


@pytest.mark.parametrize('comp', ['total', 'cartesian', 'cartesian_complex', 'spheric', 'spheric_complex'])
def test_far_components(oracle, comp):
    tr, dt, info = cases.undulator_tracks(2, seed=1)
    check(oracle, cases.undulator_args(info, grid=(100, 4, 3)), tr, dt, comp=comp, sigma_particle=2e-5)


@pytest.mark.parametrize('comp', ['total', 'cartesian', 'cartesian_complex'])
def test_near_components(oracle, comp):
    tr, dt, info = cases.undulator_tracks(2, near=True, seed=2)
    check(oracle, cases.undulator_args(info, near=True, grid=(40, 5, 3)), tr, dt, comp=comp,
          L_screen=1e5, nPC=2)


def test_near_small_screen_distance_uses_recurrence(oracle):
    tr, dt, info = cases.undulator_tracks(1, near=True)
    args = cases.undulator_args(info, near=True, grid=(64, 4, 2), L_scr=2.0)
    args['grid'][0] = (1.0, 40.0)
    check(oracle, args, tr, dt, L_screen=2.0, tol=1e-10)


def test_snapshots_and_particle_chunks(oracle):
    tr, dt, info = cases.undulator_tracks(3, seed=1)
    _, cnt = check(oracle, cases.undulator_args(info, grid=(70, 5, 3)), tr, dt, nSnaps=4, nPC=2)
    assert cnt[1] == 3 * 1664 * 70 * 5 * 3


def test_global_it_range_with_it_start(oracle):
    tr, dt, info = cases.undulator_tracks(3, seed=1)
    tr2 = [t[:7] + [s] for t, s in zip(tr, [0, 5, 17])]
    check(oracle, cases.undulator_args(info, grid=(70, 5, 3)), tr2, dt, nSnaps=3, it_range=(0, 1500), nPC=3)
    check(oracle, cases.undulator_args(info, grid=(70, 5, 3)), tr2, dt, nSnaps=5, it_range=(3, 2000))


def test_flush_chain_quirks(oracle):
    """Q3: a track whose it_start sits exactly one below a snapshot never flushes; duplicate
    snapshot iterations stall the chain; it_range shorter than the track truncates it."""
    tr, dt, info = cases.undulator_tracks(3, seed=4)
    args = cases.undulator_args(info, grid=(33, 3, 2))
    short = [[c[:40] for c in t[:6]] + [t[6]] for t in tr]
    snaps = np.linspace(0, 30, 4, dtype=np.uint32)[1:]             # [10, 20, 30]
    tr3 = [short[0] + [int(snaps[0]) - 1], short[1] + [2], short[2] + [int(snaps[1]) - 2]]
    check(oracle, args, tr3, dt, nSnaps=3, it_range=(0, 30))
    check(oracle, args, short, dt, nSnaps=50)                       # more snapshots than steps -> duplicates
    check(oracle, args, short, dt, nSnaps=2, it_range=(0, 25))


@pytest.mark.parametrize('grid,tw', [((256, 3, 2), 16), ((300, 3, 2), 16), ((128, 3, 2), 8), ((36, 3, 2), 4), ((600, 2, 2), 16),
                                     ((1, 3, 2), None), ((2, 2, 1), None)])
def test_omega_chunking_and_tile_widths(oracle, grid, tw):
    tr, dt = cases.c5_tracks_numpy(2, 300)
    kinds = ('direct',) if grid[0] < 2 else ('direct', 'recur')
    check(oracle, cases.c5_args(grid=grid), tr, dt, kinds=kinds, tw=tw)


def test_pair_kernel_scalar_and_tensor_core_forms_agree(oracle):
    """KIND_PAIR (DMMA layout for fp64, TW*NC % 8 == 0) and KIND_PAIR_FMA (one tile per lane) are the same
    sums in a different association: both within tolerance of the oracle, and of each other far below it."""
    tr, dt, info = cases.undulator_tracks(2, seed=3)
    for grid, comp, tw in (((256, 3, 2), 'total', 8), ((100, 3, 2), 'cartesian', 4), ((200, 3, 2), 'spheric', 8)):
        args = cases.undulator_args(info, grid=grid)
        ref = oracle.calculate_spectrum(args, tr, dt, comp=comp, nSnaps=2)['radiation']
        a, _ = emu.run(args, tr, dt, kind='pair', tw=tw, comp=comp, nSnaps=2)
        b, _ = emu.run(args, tr, dt, kind='pair_fma', tw=tw, comp=comp, nSnaps=2)
        for key in ref:
            assert max(rel_errors(a[key], ref[key])) < 1e-9 and max(rel_errors(b[key], ref[key])) < 1e-9
            assert max(rel_errors(a[key], b[key])) < 1e-12, (grid, comp, key)


def test_guard_dominated_wiggler(oracle):
    tr, dt, info = cases.wiggler_tracks(4, 200)
    ref, cnt = check(oracle, cases.wiggler_args(info, grid=(256, 4, 3)), tr, dt, comp='cartesian', nPC=2)
    assert cnt[0] == ref['passed'] and cnt[0] < 0.05 * cnt[1]       # identical per-node guard decisions


def test_si_units_large_phase_falls_back_per_step(oracle):
    tr, dt, info = cases.wiggler_tracks(3, 200, si_scale=1e-3)
    check(oracle, cases.wiggler_args(info, grid=(200, 4, 3), si_scale=1e-3), tr, dt, tol=1e-12)


@pytest.mark.parametrize('feature', ['wavelengthGrid', 'logGrid'])
def test_nonuniform_grids_direct(oracle, feature):
    tr, dt, info = cases.wiggler_tracks(3, 200)
    ref, cnt = check(oracle, cases.wiggler_args(info, grid=(150, 4, 3), features=[feature]), tr, dt,
                     kinds=('direct',))
    assert cnt[0] == ref['passed']


def test_float_mixed_precision_is_no_worse_than_literal_fp32(oracle):
    """fp32 protocol: the kernel keeps tau / seeds in fp64 and only the per-omega work
    in fp32, so it must sit closer to the fp64 answer than the literal fp32 restatement does."""
    tr, dt, info = cases.undulator_tracks(1)
    a64 = cases.undulator_args(info, grid=(128, 6, 2))
    a32 = cases.undulator_args(info, grid=(128, 6, 2), dtype='float')
    r64 = oracle.calculate_spectrum(a64, tr, dt)['radiation']['total']
    lit = oracle.calculate_spectrum(a32, tr, dt)['radiation']['total']
    e_lit = rel_errors(lit, r64)
    for kind in ('direct', 'recur'):
        rad, _ = emu.run(a32, tr, dt, kind=kind)
        e = rel_errors(rad['total'], r64)
        assert e[0] <= e_lit[0] and e[1] <= e_lit[1], (kind, e, e_lit)
        assert max(e) < 2e-4, (kind, e)


def test_device_sincos_accuracy():
    """srb_core.cuh:sincos_big (host build of the same code) against long-double sin/cos."""
    import ctypes
    emu.build()
    lib = ctypes.CDLL(emu._SO)
    rs = np.random.RandomState(0)
    for scale in (1.0, 1e3, 1e6, 1e10, 1e13):
        x = np.ascontiguousarray(rs.uniform(-scale, scale, 200000))
        s, c = np.empty_like(x), np.empty_like(x)
        lib.srb_emu_sincos(x.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p),
                           c.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(x.size))
        xl = x.astype(np.longdouble)
        assert np.abs(s - np.sin(xl)).max() < 3e-16 and np.abs(c - np.cos(xl)).max() < 3e-16, scale
    x = np.array([0.0, np.pi / 2, -np.pi / 2, np.pi, 1e-300])
    s, c = np.empty_like(x), np.empty_like(x)
    lib.srb_emu_sincos(x.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p),
                       c.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(x.size))
    np.testing.assert_allclose(s, [0, 1, -1, 0, 1e-300], atol=2e-16)
    np.testing.assert_allclose(c, [1, 0, 0, -1, 1], atol=2e-16)


@pytest.mark.parametrize('near', [False, True])
def test_literal_fp32_mode_matches_the_fp32_oracle(oracle, near):
    """float_mode='literal' (srb_literal.cuh): tracks, tables, tau, amplitude and phase in fp32 in the reference's
    order, per-node guard on the fp32 phases; accumulation with fused multiply-adds.  Agreement with the strict fp32
    restatement to a few 1e-6 (north star: 1e-4), with IDENTICAL guard decisions (in the near field fp32 phases of
    1e10 rad make the guard fail at random)."""
    tr, dt, info = cases.undulator_tracks(2, seed=3)
    kw = dict(L_screen=1e5) if near else {}
    comps = ['total', 'cartesian_complex'] if near else ['total', 'cartesian', 'cartesian_complex', 'spheric_complex']
    for comp in comps:
        a32 = cases.undulator_args(info, near=near, grid=(100, 4, 3), dtype='float')
        lit = oracle.calculate_spectrum(a32, tr, dt, comp=comp, sigma_particle=2e-5, nSnaps=2, **kw)
        a = dict(a32)
        a['float_mode'] = 'literal'
        rad, cnt = emu.run(a, tr, dt, comp=comp, sigma_particle=2e-5, nSnaps=2, nPC=2, **kw)
        assert cnt[0] == lit['passed'], (comp, cnt, lit['passed'])
        for k in rad:
            assert max(rel_errors(rad[k], lit['radiation'][k])) < 5e-6, (comp, k, rel_errors(rad[k], lit['radiation'][k]))


@pytest.mark.parametrize('grid', [(256, 3, 2), (200, 2, 3), (33, 2, 2), (600, 2, 2)])
def test_warp_specialised_pair_kernel_logic(oracle, grid):
    """srb_ws.cuh (the headline kernel: DMMA consumer warps + producer warps over an mbarrier ring): the per-item functions
    shared with the GPU kernel -- item iterator (tracks / snapshot intervals / sub-batches / flush items), producer step
    (guard, amplitude, tile and pair phasors in the fragment layouts of the stage), consumer main phase (MMA spelled out),
    lane-by-lane path for partial steps, fragment <-> tile transposes around the flush -- run producer-then-consumer in
    sequence per item.  mbarriers, TMA and the role split are GPU-only (tests -m gpu)."""
    tr, dt = cases.c5_tracks_numpy(3, 500)
    tr = [t[:7] + [s] for t, s in zip(tr, (0, 4, 9))]
    args = cases.c5_args(grid=grid)
    for kw in (dict(), dict(comp='cartesian', nSnaps=3, it_range=(0, 480)), dict(comp='cartesian_complex', sigma_particle=1e-5)):
        ref = oracle.calculate_spectrum(args, tr, dt, **kw)
        for nPC in (1, 2):
            rad, cnt = emu.run(args, tr, dt, kind='pair_ws', nPC=nPC, **kw)
            for key, r in ref['radiation'].items():
                assert max(rel_errors(rad[key], r)) < 1e-10, (grid, kw, key, rel_errors(rad[key], r))
            if 'it_range' not in kw:
                assert cnt[0] == ref['passed']
    # guard-dominated input (partial steps dominate) and SI units (|phase| > 2^18: node-by-node fallback)
    trw, dtw, infow = cases.wiggler_tracks(4, 256)
    argw = cases.wiggler_args(infow, grid=(grid[0], 3, 2))
    ref = oracle.calculate_spectrum(argw, trw, dtw, comp='cartesian', nSnaps=3)
    rad, cnt = emu.run(argw, trw, dtw, kind='pair_ws', comp='cartesian', nSnaps=3)
    for key, r in ref['radiation'].items():
        assert max(rel_errors(rad[key], r)) < 1e-9, (key, rel_errors(rad[key], r))
    assert cnt[0] == ref['passed']
    trs, dts, infos = cases.wiggler_tracks(3, 256, si_scale=1e-3)
    args_si = cases.wiggler_args(infos, grid=(grid[0], 2, 2), si_scale=1e-3)
    ref = oracle.calculate_spectrum(args_si, trs, dts, comp='cartesian_complex')
    rad, _ = emu.run(args_si, trs, dts, kind='pair_ws', comp='cartesian_complex')
    for key, r in ref['radiation'].items():
        assert max(rel_errors(rad[key], r)) < 1e-9, (key, rel_errors(rad[key], r))


def test_corrected_recurrence_kernel_logic(oracle):
    """srb_drec.cuh (KIND_DREC): direct layout, per-lane recurrence along omega, per-update correction onto the
    reference's ROUNDED phase.  The near field of the reference's own near-field test has phases of 1e10 rad (one ulp =
    2e-6 rad): the rounding of w_j * tau has to be reproduced, not just the exact product.  Also SI-unit far fields
    (|phase| ~ 1e6), guard-dominated input (partial ranges), snapshots, ragged chunks."""
    tr, dt, info = cases.undulator_tracks(2, near=True, seed=4)
    for grid in ((128, 6, 3), (100, 4, 2), (300, 3, 2)):
        a = cases.undulator_args(info, near=True, grid=grid)
        for kw in (dict(), dict(comp='cartesian', nSnaps=3), dict(comp='cartesian_complex')):
            ref = oracle.calculate_spectrum(a, tr, dt, L_screen=1e5, **kw)
            for nPC in (1, 2):
                rad, cnt = emu.run(a, tr, dt, kind='drec', L_screen=1e5, nPC=nPC, **kw)
                for key, r in ref['radiation'].items():
                    assert max(rel_errors(rad[key], r)) < 3e-10, (grid, kw, key, rel_errors(rad[key], r))
                assert cnt[0] == ref['passed']
    trs, dts, infos = cases.wiggler_tracks(4, 256, si_scale=1e-3)
    args = cases.wiggler_args(infos, grid=(256, 4, 4), si_scale=1e-3)
    for comp in ('total', 'cartesian_complex', 'spheric'):
        ref = oracle.calculate_spectrum(args, trs, dts, comp=comp)
        rad, cnt = emu.run(args, trs, dts, kind='drec', comp=comp)
        for key, r in ref['radiation'].items():
            assert max(rel_errors(rad[key], r)) < 1e-12, (comp, key)
        assert cnt[0] == ref['passed']
    trb, dtb, infob = cases.betatron_tracks(5, seed=3, samples_per_osc=32)
    ab = cases.betatron_args(infob, grid=(128, 5, 4))
    ref = oracle.calculate_spectrum(ab, trb, dtb, comp='cartesian')
    rad, cnt = emu.run(ab, trb, dtb, kind='drec', comp='cartesian')
    assert max(max(rel_errors(rad[k], r)) for k, r in ref['radiation'].items()) < 1e-12 and cnt[0] == ref['passed']


@pytest.mark.parametrize('comp', ['total', 'cartesian_complex', 'spheric'])
def test_time_axis_split_far(oracle, comp):
    """Time-axis split (few-particle configurations): every track cut into nTS step segments, partial complex
    amplitudes summed before squaring.  Snapshots (cumulative), late it_start, tracks of different lengths, every
    far-field kernel form; the passed-update counts must not change."""
    tr, dt, info = cases.undulator_tracks(3, seed=4)
    tr[1] = [np.asarray(a)[:700].copy() if isinstance(a, np.ndarray) else a for a in tr[1]]     # a shorter track
    tr2 = [t[:7] + [s] for t, s in zip(tr, [0, 7, 40])]
    args = cases.undulator_args(info, grid=(70, 3, 2))
    kw = dict(comp=comp, nSnaps=3, it_range=(2, 1600), sigma_particle=2e-5)
    ref = oracle.calculate_spectrum(args, tr2, dt, **kw)
    kinds = ['direct', 'recur', 'pair', 'drec'] + (['pair_ws'] if not comp.startswith('spheric') else [])
    base = None
    for kind in kinds:
        for nTS in (3, 5):
            rad, cnt = emu.run(args, tr2, dt, kind=kind, nTS=nTS, **kw)
            for key, r in ref['radiation'].items():
                assert max(rel_errors(rad[key], r)) < 1e-10, (kind, nTS, key, rel_errors(rad[key], r))
            base = base or cnt
            assert cnt == base, (kind, nTS, cnt, base)
    assert base[1] > 0


def test_time_axis_split_near_and_per_track_ranges(oracle):
    tr, dt, info = cases.undulator_tracks(2, near=True, seed=2)
    args = cases.undulator_args(info, near=True, grid=(40, 3, 2))
    for kind in ('direct', 'drec'):
        ref = oracle.calculate_spectrum(args, tr, dt, comp='cartesian', L_screen=1e5, nSnaps=2)
        rad, _ = emu.run(args, tr, dt, kind=kind, nTS=4, comp='cartesian', L_screen=1e5, nSnaps=2)
        for key, r in ref['radiation'].items():
            assert max(rel_errors(rad[key], r)) < 1e-9, (kind, key)
    args = cases.undulator_args(info, near=True, grid=(64, 3, 2), L_scr=2.0)
    args['grid'][0] = (1.0, 40.0)
    ref = oracle.calculate_spectrum(args, tr, dt, L_screen=2.0)
    rad, _ = emu.run(args, tr, dt, kind='recur', nTS=7, L_screen=2.0)
    assert max(rel_errors(rad['total'], ref['radiation']['total'])) < 1e-10


def test_guard_dominated_sub_batches_lane_is_step(oracle):
    """Recurrence kernels, sub-batches where no step passes the guard at every node: evaluated with lane = step
    (main_sparse).  A variant build takes that path whenever it is eligible (the shipped threshold takes it when only
    a few nodes pass); wiggler regime, SI-unit phases (flag 3), near field, snapshots."""
    always = ('sparse_always', ('SRB_SPARSE_ROW=1',))
    tr, dt, info = cases.wiggler_tracks(3, 200)
    args = cases.wiggler_args(info, grid=(100, 3, 2))
    for kw in (dict(comp='cartesian', nSnaps=2), dict(comp='total'), dict(comp='spheric_complex')):
        ref = oracle.calculate_spectrum(args, tr, dt, **kw)
        for variant in (None, always):
            rad, cnt = emu.run(args, tr, dt, kind='recur', variant=variant, **kw)
            for key, r in ref['radiation'].items():
                assert max(rel_errors(rad[key], r)) < 1e-10, (kw, variant, key)
            assert cnt[0] == ref['passed']
    tr, dt, info = cases.betatron_tracks(3, seed=0)
    args = cases.betatron_args(info, grid=(100, 3, 2))
    ref = oracle.calculate_spectrum(args, tr, dt, comp='cartesian')
    for variant in (None, always):
        rad, cnt = emu.run(args, tr, dt, kind='recur', comp='cartesian', variant=variant)
        for key, r in ref['radiation'].items():
            assert max(rel_errors(rad[key], r)) < 1e-11, (variant, key)
        assert cnt[0] == ref['passed']
    tr, dt, info = cases.undulator_tracks(1, near=True)
    args = cases.undulator_args(info, near=True, grid=(64, 4, 2), L_scr=2.0)
    args['grid'][0] = (1.0, 400.0)          # high frequencies: the guard cuts the upper nodes
    ref = oracle.calculate_spectrum(args, tr, dt, L_screen=2.0)
    for variant in (None, always):
        rad, cnt = emu.run(args, tr, dt, kind='recur', L_screen=2.0, variant=variant)
        assert max(rel_errors(rad['total'], ref['radiation']['total'])) < 1e-10, variant
        assert cnt[0] == ref['passed']
