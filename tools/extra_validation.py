"""Unusual shapes on the GPU vs the oracle: many omega chunks, many snapshots, one very long track, ranges past the track end."""
import contextlib, io, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import cases, fuzzcases
from oracle import reference_path as rp
from synchrad.calc import SynchRad

def gpu(args, tracks, dt, phasor='auto', **kw):
    a = dict(args); a['phasor'] = phasor
    with contextlib.redirect_stdout(io.StringIO()):
        c = SynchRad(a); c.calculate_spectrum([list(t) for t in tracks], timeStep=dt, verbose=False, **kw)
    return c

def check(name, args, tracks, dt, phasors=('auto', 'recur', 'direct'), **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        ref = rp.calculate_spectrum(args, tracks, dt, **kw)
    for ph in phasors:
        c = gpu(args, tracks, dt, phasor=ph, **kw)
        e = fuzzcases.vector_errors(c.Data['radiation'], ref['radiation'])
        print(f'{name:34s} {ph:7s} kernel={c.last_run["kernel"]:10s} err={e:.2e} passed_equal={c.last_run["passed_updates"] == ref["passed"]}')
        assert e < 1e-9

tr, dt = cases.c5_tracks_numpy(3, 500)
check('1024 omega nodes (4 chunks)', cases.c5_args(grid=(1024, 2, 2)), tr, dt)
check('1000 omega nodes ragged', cases.c5_args(grid=(1000, 2, 1)), tr, dt, comp='cartesian_complex')
check('100 snapshots', cases.c5_args(grid=(64, 2, 2)), tr, dt, nSnaps=100, comp='cartesian')
check('range past track end', cases.c5_args(grid=(64, 2, 2)), [t[:7] + [s] for t, s in zip(tr, (0, 100, 700))], dt, nSnaps=4, it_range=(0, 2000))
long1, dtl = cases.c5_tracks_numpy(1, 200000)
check('one 200k-step track', cases.c5_args(grid=(256, 2, 2)), long1, dtl, phasors=('auto', 'recur'))
a32 = cases.c5_args(grid=(600, 2, 2), dtype='float')
with contextlib.redirect_stdout(io.StringIO()):
    r64 = rp.calculate_spectrum(cases.c5_args(grid=(600, 2, 2)), tr, dt)
for ph in ('auto', 'recur', 'direct'):
    c = gpu(a32, tr, dt, phasor=ph)
    print('float 600 nodes', ph, c.last_run['kernel'], c.last_run['tile_width'], fuzzcases.vector_errors(c.Data['radiation'], r64['radiation']))
print('extra validation ok')
