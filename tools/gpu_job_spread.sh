mkdir -p gpurun_out
python - <<'PY' 2>&1 | tee gpurun_out/spread_check.txt
import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import cases
from conftest import rel_errors
from oracle import reference_path as rp
from synchrad.calc import SynchRad
def run(args, tr, dt, phasor, **kw):
    a = dict(args); a['phasor'] = phasor
    c = SynchRad(a); c.calculate_spectrum([list(t) for t in tr], timeStep=dt, verbose=False, **kw); return c
tr, dt = cases.c5_tracks_numpy(6, 700)
for grid in ((256, 4, 3), (200, 3, 2)):
    args = cases.c5_args(grid=grid)
    for kw in (dict(), dict(comp='cartesian', nSnaps=3), dict(comp='cartesian_complex', sigma_particle=1e-5)):
        ref = rp.calculate_spectrum(args, tr, dt, **kw)
        c = run(args, tr, dt, 'spread', **kw)
        e = max(max(rel_errors(c.Data['radiation'][k], ref['radiation'][k])) for k in ref['radiation'])
        print(grid, kw, 'kernel', c.last_run['kernel'], 'err', f'{e:.2e}', 'passed_equal', c.last_run['passed_updates'] == ref['passed'])
trw, dtw, infow = cases.wiggler_tracks(6, 256)
args = cases.wiggler_args(infow, grid=(256, 4, 3))
ref = rp.calculate_spectrum(args, trw, dtw, comp='cartesian')
c = run(args, trw, dtw, 'spread', comp='cartesian')
print('wiggler (guard-dominated)', f"{max(max(rel_errors(c.Data['radiation'][k], ref['radiation'][k])) for k in ref['radiation']):.2e}", c.last_run['passed_updates'] == ref['passed'])
PY
for ph in auto spread; do python tools/quick_perf.py 592 10000 double $ph 2 2>/dev/null | tee -a gpurun_out/spread_perf.txt; done
