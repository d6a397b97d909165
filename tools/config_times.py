"""End-to-end times of the BASELINE configs through the public API (development aid / profiles).
usage: python tools/config_times.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import cases
from synchrad.calc import SynchRad

def run(name, args, tracks, dt, reps=2, **kw):
    calc = SynchRad(dict(args))
    best = 1e9
    for r in range(reps + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        calc.calculate_spectrum(tracks, timeStep=dt, verbose=False, **kw)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        if r: best = min(best, t)
    lr = calc.last_run
    f = lr['passed_updates'] / max(lr['visited_updates'], 1)
    sys.__stdout__.write(f"{name}: updates={lr['updates']:.3e} e2e_s={best:.4f} updates/s={lr['updates']/best:.3e} "
                         f"integrate_ms={lr['integrate_ms']:.2f} kernel={lr['kernel']} tw={lr['tile_width']} pc={lr['particle_chunks']} guard_pass={f:.4f}\n")

sys.stdout = sys.stderr
tr1, dt, info = cases.undulator_tracks(1)
run('C1 (configs[0]) far undulator, single electron (128,32,32) double', cases.undulator_args(info), tr1, dt)
tr, dt, info = cases.undulator_tracks(24, seed=0)
run('C1 x 24 e- (the reference test script) double', cases.undulator_args(info), tr, dt)
a32 = cases.undulator_args(info, dtype='float'); a32['native'] = True
run('C1 x 24 e- float+native', a32, tr, dt)
trn1, dtn, infon = cases.undulator_tracks(1, near=True)
run('C2 (configs[1]) near undulator, single electron (128,256,32) double', cases.undulator_args(infon, near=True), trn1, dtn, L_screen=1e5)
trn, dtn, infon = cases.undulator_tracks(24, near=True, seed=0)
run('C2 x 24 e- near double', cases.undulator_args(infon, near=True), trn, dtn, L_screen=1e5)
trb, dtb, infob = cases.betatron_tracks(1000, seed=0)
run('C3 (configs[2]) betatron recipe (SI) 1e3 x 256 (256,32,32) cartesian double', cases.betatron_args(infob), trb, dtb, comp='cartesian')
trs, dts, infos = cases.spiral_tracks(10000, seed=0)
run('C4 (configs[3]) spiral beam 1e4 x 192 (512,64,64) float, mixed mode', cases.spiral_args(infos), trs, dts)
lit = cases.spiral_args(infos); lit['float_mode'] = 'literal'
run('C4 (configs[3]) spiral beam 1e4 x 192 (512,64,64) float, literal mode', lit, trs, dts)
run('C4 grid, double, coherent (cartesian_complex) as in the notebook', cases.spiral_args(infos, dtype='double'), trs, dts, comp='cartesian_complex')
