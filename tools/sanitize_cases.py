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
# tensor-core pair kernel: 8-node tiles with 2 and 3 components, snapshots (layout transposes), partial path, scalar form
for grid, comp in (((256, 3, 2), 'total'), ((300, 3, 2), 'spheric_complex')):
    a = cases.undulator_args(info, grid=grid); a['phasor'] = 'pair'
    out.append(run(a, short, dt, comp=comp, nSnaps=3, it_range=(0, 190)))
w = cases.wiggler_args(infow, grid=(256, 3, 2)); w['phasor'] = 'pair'
out.append(run(w, trw, dtw, comp='cartesian', nSnaps=2))
a = cases.undulator_args(info, grid=(256, 3, 2)); a['phasor'] = 'pair_fma'
out.append(run(a, short, dt, nSnaps=2))
print('sanitize cases done', len(out), np.isfinite(out).all())
