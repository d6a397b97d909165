"""Small runs for compute-sanitizer of the round-2 code paths: the warp-specialised pair kernel (mbarrier ring, TMA-staged
inputs, two-consumer flush hand-over), the corrected-recurrence kernel, the lane = step path of the recurrence kernels
(staging-area reduction), the rolled flush through the staging area, the time-axis split and the on-device choice among
three candidates."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import cases
from synchrad.calc import SynchRad

out = []


def run(args, tracks, dt, phasor='auto', **kw):
    a = dict(args); a['phasor'] = phasor
    c = SynchRad(a)
    c.calculate_spectrum([list(t) for t in tracks], timeStep=dt, verbose=False, **kw)
    out.append(float(sum(np.abs(v).sum() for v in c.Data['radiation'].values())))
    return c.last_run


tr, dt, info = cases.undulator_tracks(3, seed=1)
short = [[c[:330] for c in t[:6]] + [t[6], s] for t, s in zip(tr, (0, 3, 9))]
# warp-specialised pair kernel: 256- and 128-node grids, snapshots (flush hand-over between the two consumer warps)
for grid in ((256, 3, 2), (128, 2, 3)):
    lr = run(cases.undulator_args(info, grid=grid), short, dt, nSnaps=3, it_range=(0, 320))
    assert lr['kernel'] == 'pair', lr
# time-axis split on the same kernel and on the recurrence kernel
os.environ['SRB_TIME_SPLIT'] = '3'
for ph in ('auto', 'recur', 'direct'):
    lr = run(cases.undulator_args(info, grid=(128, 2, 3)), short, dt, phasor=ph, nSnaps=2, comp='cartesian_complex')
    assert lr['time_segments'] == 3, lr
del os.environ['SRB_TIME_SPLIT']
# corrected recurrence: near field at large L, and far field chosen on the device (tracks far from the origin)
trn, dtn, infon = cases.undulator_tracks(2, near=True, seed=2)
shortn = [[c[:330] for c in t[:6]] + [t[6]] for t in trn]
lr = run(cases.undulator_args(infon, near=True, grid=(64, 6, 3)), shortn, dtn, L_screen=1e5)
assert lr['kernel'] == 'drec', lr
far = [[t[0], t[1], t[2] + 10.0] + list(t[3:]) for t in short]
lr = run(cases.undulator_args(info, grid=(128, 3, 2)), far, dt)
assert lr['kernel'] == 'drec', lr
# guard-dominated SI-unit recipe: lane = step path + rolled flush of 16-node tiles, double and mixed fp32
trb, dtb, infob = cases.betatron_tracks(6, seed=0)
for dtype in ('double', 'float'):
    lr = run(cases.betatron_args(infob, grid=(256, 3, 2), dtype=dtype) if 'dtype' in cases.betatron_args.__code__.co_varnames
             else dict(cases.betatron_args(infob, grid=(256, 3, 2)), dtype=dtype), trb, dtb, comp='cartesian')
    assert lr['kernel'] == 'recurrence', lr
print('sanitize cases3 done', len(out), np.isfinite(out).all())
