"""One-off extended validation on a GPU box: many fuzz seeds + mid-size parity runs vs the oracle."""
import contextlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import cases, fuzzcases
from conftest import rel_errors
from oracle import reference_path as rp
from synchrad.calc import SynchRad

def gpu(args, tracks, dt, phasor='auto', **kw):
    a = dict(args); a['phasor'] = phasor
    with contextlib.redirect_stdout(io.StringIO()):
        c = SynchRad(a); c.calculate_spectrum([list(t) for t in tracks], timeStep=dt, verbose=False, **kw)
    return c

worst = 0.0; n = 0; t0 = time.time()
for seed in range(100, 100 + int(sys.argv[1]) if len(sys.argv) > 1 else 120):
    rs = np.random.RandomState(seed)
    for i in range(25):
        A, tracks, dt, kw = fuzzcases.rand_case(rs)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = rp.calculate_spectrum(A, tracks, dt, **kw)
        for phasor in (('auto',) if A.get('Features') else ('auto', 'direct')):
            c = gpu(A, tracks, dt, phasor=phasor, **kw)
            e = fuzzcases.vector_errors(c.Data['radiation'], ref['radiation'])
            worst = max(worst, e); n += 1
            if e > 1e-9:
                print('FAIL', seed, i, phasor, e, A['grid'], A.get('mode'), A.get('Features'), kw)
print(f'fuzz: {n} runs, worst whole-vector error {worst:.3e}, {time.time()-t0:.0f} s')

tr, dt = cases.c5_tracks_numpy(64, 2000)
A = cases.c5_args(grid=(256, 8, 8))
ref = rp.calculate_spectrum(A, tr, dt)
for phasor in ('auto', 'direct'):
    c = gpu(A, tr, dt, phasor=phasor)
    print('C5 mid-size', phasor, rel_errors(c.Data['radiation']['total'], ref['radiation']['total']),
          'passed equal', c.last_run['passed_updates'] == ref['passed'])
tb, dtb, infob = cases.betatron_tracks(64, seed=1)
Ab = cases.betatron_args(infob, grid=(256, 8, 8))
refb = rp.calculate_spectrum(Ab, tb, dtb, comp='cartesian')
cb = gpu(Ab, tb, dtb, comp='cartesian')
print('C3 recipe 64 e-', {k: rel_errors(cb.Data['radiation'][k], refb['radiation'][k]) for k in 'xyz'},
      'passed equal', cb.last_run['passed_updates'] == refb['passed'])
