"""Where the host side of calculate_spectrum spends its time on the small BASELINE configs (development aid).
usage: python tools/host_overhead.py   (needs a GPU)"""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import cases
from synchrad.calc import SynchRad

out = sys.stdout
sys.stdout = sys.stderr


def profile(name, args, tracks, dt, reps, **kw):
    calc = SynchRad(dict(args))
    for _ in range(3):
        calc.calculate_spectrum(tracks, timeStep=dt, verbose=False, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        calc.calculate_spectrum(tracks, timeStep=dt, verbose=False, **kw)
    torch.cuda.synchronize()
    per = (time.perf_counter() - t0) / reps
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(reps):
        calc.calculate_spectrum(tracks, timeStep=dt, verbose=False, **kw)
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22)
    out.write(f'==== {name}: {per * 1e3:.3f} ms per call, integrate_ms {calc.last_run["integrate_ms"]:.3f}, {reps} calls profiled\n')
    out.write('\n'.join(l[:160] for l in s.getvalue().splitlines()[4:40]) + '\n')


tr1, dt, info = cases.undulator_tracks(1)
profile('C1 single electron', cases.undulator_args(info), tr1, dt, 200)
trb, dtb, infob = cases.betatron_tracks(1000, seed=0)
profile('C3 betatron 1e3 x 256 cartesian', cases.betatron_args(infob), trb, dtb, 20, comp='cartesian')
