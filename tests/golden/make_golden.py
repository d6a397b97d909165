"""Generates tests/golden/small_cases.npz from the strict CPU oracle (run here, committed).

These vectors are ORACLE outputs (the oracle is pinned bit for bit to the reference, see
make_reference_golden.py / tests/test_reference_pin.py, whose reference_cases.npz holds the same cases
produced by the unmodified reference); they let the GPU parity tests run against stored arrays as well as
against a live oracle.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402


def small_cases():
    out = {}
    tr, dt, info = cases.undulator_tracks(3, seed=1)
    out['far_total'] = (cases.undulator_args(info, grid=(128, 6, 4)), tr, dt, dict(comp='total'))
    out['far_cartesian_snaps'] = (cases.undulator_args(info, grid=(96, 5, 3)), tr, dt,
                                  dict(comp='cartesian', nSnaps=3))
    out['far_cartesian_complex'] = (cases.undulator_args(info, grid=(64, 4, 4)), tr, dt,
                                    dict(comp='cartesian_complex', sigma_particle=2e-5))
    out['far_spheric'] = (cases.undulator_args(info, grid=(64, 4, 4)), tr, dt, dict(comp='spheric'))
    out['far_spheric_complex'] = (cases.undulator_args(info, grid=(64, 4, 4)), tr, dt,
                                  dict(comp='spheric_complex'))
    tr2 = [t[:7] + [s] for t, s in zip(tr, [0, 5, 17])]
    out['far_it_range'] = (cases.undulator_args(info, grid=(70, 5, 3)), tr2, dt,
                           dict(nSnaps=3, it_range=(0, 1500)))
    trn, dtn, infon = cases.undulator_tracks(2, near=True, seed=2)
    out['near_total'] = (cases.undulator_args(infon, near=True, grid=(128, 6, 4)), trn, dtn,
                         dict(comp='total', L_screen=1e5))
    out['near_cartesian_complex'] = (cases.undulator_args(infon, near=True, grid=(40, 6, 4)), trn, dtn,
                                     dict(comp='cartesian_complex', L_screen=1e5))
    trw, dtw, infow = cases.wiggler_tracks(6, 256)
    out['wiggler_cartesian'] = (cases.wiggler_args(infow, grid=(256, 6, 4)), trw, dtw, dict(comp='cartesian'))
    out['wiggler_loggrid'] = (cases.wiggler_args(infow, grid=(200, 6, 4), features=['logGrid']), trw, dtw, {})
    out['wiggler_wavelengthgrid'] = (cases.wiggler_args(infow, grid=(200, 6, 4), features=['wavelengthGrid']),
                                     trw, dtw, {})
    tr5, dt5 = cases.c5_tracks_numpy(3, 600)
    out['c5_small'] = (cases.c5_args(grid=(256, 4, 4)), tr5, dt5, {})
    return out


if __name__ == '__main__':
    from oracle import reference_path as rp
    rp.build()
    blobs = {}
    for name, (args, tracks, dt, kw) in small_cases().items():
        res = rp.calculate_spectrum(args, tracks, dt, **kw)
        for key, arr in res['radiation'].items():
            blobs[f'{name}/{key}'] = arr
    np.savez_compressed(os.path.join(HERE, 'small_cases.npz'), **blobs)
    print('wrote', len(blobs), 'arrays')
