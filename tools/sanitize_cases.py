"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import cases
from synchrad.calc import SynchRad

def run(args, tracks, dt, **kw):
    c = SynchRad(dict(args)); c.calculate_spectrum([list(t) for t in tracks], timeStep=dt, verbose=False, **kw)
    return float(sum(v.sum() for v in c.Data['radiation'].values()))

tr, dt, info = cases.undulator_tracks(3, seed=1)
short = [[c[:200] for c in t[:6]] + [t[6], s] for t, s in zip(tr, (0, 3, 9))]
out = []
for dtype in ('double', 'float'):
    for phasor in ('auto', 'direct'):
        a = cases.undulator_args(info, grid=(70, 3, 2), dtype=dtype); a['phasor'] = phasor
        for comp in ('total', 'cartesian_complex', 'spheric'):
            out.append(run(a, short, dt, comp=comp, nSnaps=2, it_range=(0, 190)))
        n = cases.undulator_args(info, near=True, grid=(40, 3, 2), dtype=dtype); n['phasor'] = phasor
        out.append(run(n, short, dt, comp='cartesian', L_screen=1e5))
        n2 = cases.undulator_args(info, near=True, grid=(40, 3, 2), L_scr=2.0, dtype=dtype); n2['phasor'] = phasor
        n2['grid'][0] = (1.0, 40.0)
        out.append(run(n2, short, dt, L_screen=2.0))
trw, dtw, infow = cases.wiggler_tracks(4, 100)
out.append(run(cases.wiggler_args(infow, grid=(300, 3, 2)), trw, dtw))
out.append(run(cases.wiggler_args(infow, grid=(100, 3, 2), features=['logGrid']), trw, dtw))
print('sanitize cases done', len(out), np.isfinite(out).all())
