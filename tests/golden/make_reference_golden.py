"""Generates tests/golden/reference_cases.npz and reference_kat.json by running the UNMODIFIED reference
(/root/reference/synchrad: calc.py + utils.py + kernel_farfield.cl / kernel_nearfield.cl) in this container
through oracle/run_reference.py (pyopencl/mako/h5py stand-ins from oracle/clshim, the kernels compiled for the
host with g++ -O2 -ffp-contract=off).  The reference cannot travel to the GPU box, the vectors can:

    python tests/golden/make_reference_golden.py            # needs /root/reference; output is committed

Cases = the oracle's small cases (make_golden.small_cases) + option/dtype variants of the reference's public
call + both file layouts + the reference's own two test scripts (tests/test_undulator_analytic*.py) with the
RNG seeded, stored as summary values.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402
from golden.make_golden import small_cases  # noqa: E402


def extra_cases():
    """Variants of the public call beyond small_cases(): name -> (args, tracks, dt, kwargs)."""
    out = {}
    tr, dt, info = cases.undulator_tracks(4, seed=5)
    for t, w in zip(tr, (0.5, 2.0, 1.25, 3.0)):
        t[6] = w
    out['opt_weights_mean'] = (cases.undulator_args(info, grid=(48, 4, 3)), tr, dt, dict(weights_normalize='mean'))
    out['opt_weights_max'] = (cases.undulator_args(info, grid=(48, 4, 3)), tr, dt, dict(weights_normalize='max'))
    out['opt_weights_ones_npmax'] = (cases.undulator_args(info, grid=(48, 4, 3)), tr, dt,
                                     dict(weights_normalize='ones', Np_max=3))
    out['opt_snaps_per_track_range'] = (cases.undulator_args(info, grid=(48, 4, 3)), tr, dt, dict(nSnaps=4))
    tr7 = [t[:7] for t in tr]                                            # 7-element tracks (calc.py:582-584)
    out['opt_seven_element_tracks'] = (cases.undulator_args(info, grid=(40, 3, 3)), tr7, dt,
                                       dict(comp='cartesian', it_range=(0, 1200), nSnaps=2))
    trs = [t[:7] + [s] for t, s in zip(tr, (0, 100, 333, 1700))]         # it_start beyond some snapshots
    out['opt_it_start_late'] = (cases.undulator_args(info, grid=(40, 3, 3)), trs, dt,
                                dict(it_range=(0, 2000), nSnaps=5))
    out['opt_it_range_short'] = (cases.undulator_args(info, grid=(40, 3, 3)), trs, dt,
                                 dict(it_range=(50, 900), nSnaps=3, comp='spheric'))
    out['opt_single_node_axes'] = (cases.undulator_args(info, grid=(33, 1, 1)), tr, dt, dict(comp='cartesian'))
    # single precision, the reference's literal fp32 path, plain and with native functions
    out['float_far_total'] = (cases.undulator_args(info, grid=(64, 4, 4), dtype='float'), tr, dt, {})
    out['float_far_cartesian_complex'] = (cases.undulator_args(info, grid=(48, 4, 3), dtype='float'), tr, dt,
                                          dict(comp='cartesian_complex', sigma_particle=1e-5))
    a = cases.undulator_args(info, grid=(64, 4, 4), dtype='float')
    a['native'] = True
    out['float_native_far_total'] = (a, tr, dt, {})
    trn, dtn, infon = cases.undulator_tracks(2, near=True, seed=7)
    out['float_near_total'] = (cases.undulator_args(infon, near=True, grid=(32, 5, 3), dtype='float'), trn, dtn,
                               dict(L_screen=1e5))
    out['near_cartesian_snaps'] = (cases.undulator_args(infon, near=True, grid=(40, 5, 3)), trn, dtn,
                                   dict(comp='cartesian', L_screen=1e5, nSnaps=2))
    trb, dtb, infob = cases.betatron_tracks(5, seed=3, samples_per_osc=32)
    out['betatron_si_cartesian'] = (cases.betatron_args(infob, grid=(128, 5, 4)), trb, dtb, dict(comp='cartesian'))
    return out


POST = [('get_energy', dict(lambda0_um=1)), ('get_energy', dict(phot_num=True)),
        ('get_energy_spectrum', dict(lambda0_um=0.8)), ('get_full_spectrum', dict(normalize_to_weights=True))]
# spot maps (utils.py:104-158), stored as '<case>/spot<i>'; K0REL: k0 = omega_min + K0REL * (omega_max - omega_min)
K0REL = 0.37
POST_SPOT = [('get_spot', dict(lambda0_um=1)), ('get_spot', dict(k0=K0REL, phot_num=True)),
             ('get_spot_cartesian', dict(bins=(24, 18), lambda0_um=1)),
             ('get_spot_cartesian', dict(k0=K0REL, th_part=0.5, bins=(16, 16)))]


def spot_requests(args):
    """POST_SPOT with k0 placed inside the case's omega range (single-node axes have no k0 slice: the
    reference indexes omega[i + 1])."""
    lo, hi = args['grid'][0]
    n_w = args['grid'][-1][0]
    out = []
    for meth, kw in POST_SPOT:
        kw = dict(kw)
        if 'k0' in kw:
            if n_w < 3 or 'wavelengthGrid' in args.get('Features', ()):   # descending omega: the reference's
                continue                                                    # index search runs off the axis
            kw['k0'] = lo + kw['k0'] * (hi - lo)
        out.append((meth, kw))
    return out


def file_flow_case(rr):
    """Both file layouts through the reference's own h5py calls: tracks file -> calculate_spectrum(file_tracks,
    file_spectrum) -> spectrum file -> SynchRad(file_spectrum=...)."""
    from synchrad_b200 import trackio
    tr, dt, info = cases.undulator_tracks(3, seed=9)
    tr = [t[:7] + [s] for t, s in zip(tr, (0, 3, 11))]
    args = cases.undulator_args(info, grid=(40, 4, 3))
    with tempfile.TemporaryDirectory() as tmp:
        ft, fs = os.path.join(tmp, 'tracks.h5'), os.path.join(tmp, 'spectrum.h5')
        trackio.write_tracks(ft, tr, dt, it_range=(0, 1600))
        res = rr.run(args, None, file_tracks=ft, file_spectrum=fs, comp='cartesian', nSnaps=2, Np_max=3)
        import types
        holder = types.SimpleNamespace()
        trackio.read_spectrum(fs, holder)
        stored = dict(radiation=holder.Data['radiation'], Args=holder.Args, snap_iterations=holder.snap_iterations)
    return args, tr, dt, res, stored


def c4_cases():
    """BASELINE configs[3] (spiral beam, tutorials/Spiral_Beam_Part1.ipynb) at a size the CPU reference finishes in
    seconds: a 200-particle spiral of the same recipe on a coarser grid, with the notebook's two calls (:255-274:
    incoherent `total`, coherent `cartesian_complex`) in the config's single precision and in double, plus the
    coherent call with a particle form factor.  name -> (args, tracks, dt, kwargs)."""
    tr, dt, info = cases.spiral_tracks(200, seed=0)
    out = {}
    out['c4_float_total'] = (cases.spiral_args(info, grid=(128, 8, 8), dtype='float'), tr, dt, dict(Np_max=100))
    out['c4_double_total'] = (cases.spiral_args(info, grid=(128, 8, 8), dtype='double'), tr, dt, dict(Np_max=100))
    out['c4_double_coherent'] = (cases.spiral_args(info, grid=(128, 8, 8), dtype='double'), tr, dt,
                                 dict(comp='cartesian_complex'))
    out['c4_double_coherent_sigma'] = (cases.spiral_args(info, grid=(96, 6, 5), dtype='double'), tr, dt,
                                       dict(comp='cartesian_complex', sigma_particle=1e-10))
    out['c4_float_coherent'] = (cases.spiral_args(info, grid=(96, 6, 5), dtype='float'), tr, dt,
                                dict(comp='cartesian_complex'))
    return out


C4_POST = [('get_energy', dict(normalize_to_weights=True, lambda0_um=1e6)),      # Spiral_Beam_Part1.ipynb:283-284
           ('get_energy_spectrum', dict(normalize_to_weights=True))]


def main_c4():
    """Writes reference_c4.npz / reference_c4_meta.json (kept apart from reference_cases.npz so that adding the C4
    recipe does not rewrite the round-1 vectors)."""
    from oracle import run_reference as rr
    assert rr.available(), 'needs /root/reference'
    blobs, meta = {}, {}
    for name, (args, tracks, dt, kw) in c4_cases().items():
        res = rr.run(args, tracks, timeStep=dt, post=C4_POST, **kw)
        for key, arr in res['radiation'].items():
            blobs[f'{name}/{key}'] = arr
        for i, v in res['post'].items():
            blobs[f'{name}/post{i}'] = np.asarray(v)
        meta[name] = dict(total_weight=res['total_weight'], keys=list(res['radiation']))
        print(name, 'ok', {k: v.shape for k, v in res['radiation'].items()}, float(res['post'][0]))
    # "Enhancement due to coherency" as the notebook prints it (:283-287), from the reference's own utils.py
    meta['_coherent_gain_double'] = float(blobs['c4_double_coherent/post0'] / blobs['c4_double_total/post0'])
    meta['_device'] = res['device']
    print('coherent gain', meta['_coherent_gain_double'])
    np.savez_compressed(os.path.join(HERE, 'reference_c4.npz'), **blobs)
    with open(os.path.join(HERE, 'reference_c4_meta.json'), 'w') as f:
        json.dump(meta, f, indent=1)


def main():
    if '--c4' in sys.argv:
        return main_c4()
    from oracle import run_reference as rr
    assert rr.available(), 'needs /root/reference'
    blobs, meta = {}, {}
    allc = dict(small_cases())
    allc.update(extra_cases())
    for name, (args, tracks, dt, kw) in allc.items():
        spots = spot_requests(args) if args['grid'][-1][1] > 1 else []     # griddata needs a 2-D point set
        res = rr.run(args, tracks, timeStep=dt, post=POST + spots, **kw)
        for key, arr in res['radiation'].items():
            blobs[f'{name}/{key}'] = arr
        for i, v in res['post'].items():
            if i < len(POST):
                blobs[f'{name}/post{i}'] = v
            elif isinstance(v, tuple):             # get_spot_cartesian returns (map, extent): keep both
                blobs[f'{name}/spot{i - len(POST)}'] = np.asarray(v[0], dtype=np.double)
                blobs[f'{name}/spot{i - len(POST)}_extent'] = np.asarray(v[1], dtype=np.double)
            else:
                blobs[f'{name}/spot{i - len(POST)}'] = v
        blobs[f'{name}/snap_iterations'] = res['snap_iterations']
        for k, v in res['Args'].items():              # the reference's own Args after the call (a1, a6 of SURVEY §8)
            if isinstance(v, (np.ndarray, np.generic, int, float)) and not isinstance(v, bool):
                blobs[f'{name}/Args/{k}'] = np.asarray(v)
        meta[name] = dict(total_weight=res['total_weight'], keys=list(res['radiation']))
        print(name, 'ok', {k: v.shape for k, v in res['radiation'].items()})
    meta['_device'] = res['device']
    # file flow
    args, tr, dt, res, stored = file_flow_case(rr)
    for key, arr in res['radiation'].items():
        blobs[f'file_flow/{key}'] = arr
    meta['file_flow'] = dict(total_weight=res['total_weight'], keys=list(res['radiation']),
                             stored_keys=sorted(stored['radiation']), stored_args=sorted(stored['Args']),
                             stored_snaps=[int(v) for v in np.asarray(stored['snap_iterations']).ravel()])
    for key, arr in stored['radiation'].items():
        assert np.array_equal(arr, res['radiation'][key])
    np.savez_compressed(os.path.join(HERE, 'reference_cases.npz'), **blobs)
    print('wrote', len(blobs), 'arrays')

    with open(os.path.join(HERE, 'reference_cases_meta.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    if '--cases-only' in sys.argv:
        return
    # the reference's own test scripts (seeded): summary values of the full-size grids
    kat = {'_source': 'unmodified reference run in this container by tests/golden/make_reference_golden.py '
                      '(oracle/run_reference.py); tests/test_undulator_analytic.py and ..._near.py with '
                      'np.random seeded through cases.undulator_tracks(24, seed=0)'}
    for tag, near, Np in (('C1_far_double_24', False, 24), ('C2_near_double_2', True, 2)):
        tr, dt, info = cases.undulator_tracks(Np, near=near, seed=0)
        args = cases.undulator_args(info, near=near)
        kw = dict(L_screen=1e5) if near else {}
        res = rr.run(args, tr, timeStep=dt, comp='total', Np_max=Np, post=[('get_energy', dict(lambda0_um=1))], **kw)
        S = res['radiation']['total'][0]
        rs = np.random.RandomState(1)
        spots = [[int(rs.randint(n)) for n in S.shape] for _ in range(12)]
        from scipy.constants import c, hbar
        Et = cases.undulator_energy_theory(info, 2e6 * np.pi * hbar * c)
        E = float(res['post'][0])
        kat[tag] = dict(argmax=[int(i) for i in np.unravel_index(S.argmax(), S.shape)], max=float(S.max()),
                        sum=float(S.sum()), l2=float(np.linalg.norm(S)), energy_J=E,
                        deviation_percent=abs(E - Et) / Et * 100,
                        spots=[[s, float(S[tuple(s)])] for s in spots])
        print(tag, kat[tag]['energy_J'], kat[tag]['deviation_percent'])
    with open(os.path.join(HERE, 'reference_kat.json'), 'w') as f:
        json.dump(kat, f, indent=1)


if __name__ == '__main__':
    main()
