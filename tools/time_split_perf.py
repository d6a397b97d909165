"""Time-axis split A/B on the single-electron BASELINE configs (development aid): integrate_ms with SRB_TIME_SPLIT=0
(particle chunks only) against the planner's choice.  usage: python tools/time_split_perf.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import cases
from synchrad.calc import SynchRad


def run(name, args, tracks, dt, **kw):
    out = {}
    for mode in ('0', None, '2', '4', '8'):
        if mode is None:
            os.environ.pop('SRB_TIME_SPLIT', None)
        else:
            os.environ['SRB_TIME_SPLIT'] = mode
        calc = SynchRad(dict(args))
        best = 1e9
        for r in range(4):
            calc.calculate_spectrum(tracks, timeStep=dt, verbose=False, **kw)
            if r:
                best = min(best, calc.last_run['integrate_ms'])
        lr = calc.last_run
        out[mode] = calc.Data['radiation']['total'].copy()
        sys.__stdout__.write(f"{name} SRB_TIME_SPLIT={mode}: integrate_ms={best:.3f} updates/s={lr['updates']/best*1e3:.3e} kernel={lr['kernel']} "
                             f"pc={lr['particle_chunks']} ts={lr['time_segments']} blocks={lr.get('grid_blocks')}\n")
    os.environ.pop('SRB_TIME_SPLIT', None)
    ref = out['0']
    for m, v in out.items():
        sys.__stdout__.write(f"   split={m} vs unsplit: {np.abs(v - ref).max() / np.abs(ref).max():.2e}\n")


sys.stdout = sys.stderr
tr1, dt, info = cases.undulator_tracks(1)
run('C1 single e-', cases.undulator_args(info), tr1, dt)
tr4, dt, info = cases.undulator_tracks(4, seed=0)
run('C1 x 4 e-', cases.undulator_args(info), tr4, dt)
a = cases.undulator_args(info); a['phasor'] = 'recur'
run('C1 single e- recur', a, tr1, dt)
trn1, dtn, infon = cases.undulator_tracks(1, near=True)
run('C2 single e- near', cases.undulator_args(infon, near=True), trn1, dtn, L_screen=1e5)
