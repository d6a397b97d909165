"""C3 / C1 through calculate_spectrum with the track list cut into 1, 2, 4, 8 batches (host packing of batch k+1 overlaps
the kernel on batch k): is pipelining worth it below the 256 MB threshold?  usage: python tools/c3_batches.py (needs a GPU)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import cases
from synchrad.calc import SynchRad

out = sys.stdout
sys.stdout = sys.stderr


def run(name, args, tracks, dt, nb, reps=8, **kw):
    a = dict(args)
    if nb > 1:
        a['max_batch_bytes'] = 96 * (sum(len(t[0]) for t in tracks) // nb + 1)
    calc = SynchRad(a)
    best = 1e9
    for r in range(reps + 2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        calc.calculate_spectrum(tracks, timeStep=dt, verbose=False, **kw)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        if r > 1: best = min(best, t)
    lr = calc.last_run
    out.write(f"{name}: batches={lr['batches']} best call {best * 1e3:.3f} ms, integrate_ms={lr['integrate_ms']:.2f} pack_s={lr['host_pack_s'] * 1e3:.2f} ms\n")
    return calc.Data['radiation']


tr1, dt, info = cases.undulator_tracks(1)
run('C1 single electron', cases.undulator_args(info), tr1, dt, 1, reps=50)
tr24, dt, info = cases.undulator_tracks(24, seed=0)
for nb in (1, 2, 4):
    run('C1 x 24', cases.undulator_args(info), tr24, dt, nb, reps=10)
trb, dtb, infob = cases.betatron_tracks(1000, seed=0)
ref = None
for nb in (1, 2, 4, 8):
    rad = run('C3 betatron 1e3 x 256 cartesian', cases.betatron_args(infob), trb, dtb, nb, comp='cartesian')
    if ref is None:
        ref = rad
    else:
        out.write('   max rel diff vs 1 batch: %.2e\n' % max(np.abs(rad[k] - ref[k]).max() / np.abs(ref[k]).max() for k in ref))
trs, dts, infos = cases.spiral_tracks(10000, seed=0)
for nb in (1, 4, 8):
    run('C4 spiral 1e4 x 192 float mixed', cases.spiral_args(infos), trs, dts, nb, reps=2)
