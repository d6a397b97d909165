import os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import cases
from synchrad.calc import SynchRad
sys.stdout = sys.stderr
trb, dtb, infob = cases.betatron_tracks(1000, seed=0)
res = {}
for ph in ('auto', 'drec', 'recur', 'pair'):
    a = cases.betatron_args(infob); a['phasor'] = ph
    calc = SynchRad(dict(a))
    best = 1e9
    for r in range(3):
        calc.calculate_spectrum(trb, timeStep=dtb, verbose=False, comp='cartesian')
        if r: best = min(best, calc.last_run['integrate_ms'])
    res[ph] = calc.Data['radiation']['x'].copy()
    sys.__stdout__.write(f"C3 SI phasor={ph}: integrate_ms={best:.2f} kernel={calc.last_run['kernel']} tw={calc.last_run['tile_width']}\n")
for ph in ('drec', 'recur', 'pair'):
    sys.__stdout__.write(f"  {ph} vs auto: {np.abs(res[ph]-res['auto']).max()/np.abs(res['auto']).max():.2e}\n")
