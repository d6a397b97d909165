"""Small runs for compute-sanitizer of what tools/sanitize_cases.py does not reach: the literal fp32 kernels
(srb_literal.cuh) and the on-device energy-spectrum integrals (k_energy_spectrum, both layouts, far and near)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import cases
from synchrad.calc import SynchRad
from synchrad_b200 import engine

tr, dt, info = cases.undulator_tracks(3, seed=1)
short = [[c[:200] for c in t[:6]] + [t[6], s] for t, s in zip(tr, (0, 3, 9))]
out = []
for near, comp, grid in ((False, 'total', (70, 5, 3)), (False, 'cartesian_complex', (33, 4, 2)), (True, 'cartesian', (40, 6, 3))):
    for dtype in ('double', 'float'):
        a = cases.undulator_args(info, near=near, grid=grid, dtype=dtype)
        if dtype == 'float':
            a['float_mode'] = 'literal'
        kw = dict(L_screen=1e5) if near else {}
        c = SynchRad(dict(a))
        c.calculate_spectrum([list(t) for t in short], timeStep=dt, verbose=False, comp=comp, nSnaps=2, it_range=(0, 190), **kw)
        out.append(float(sum(v.sum() for v in c.Data['radiation'].values())))
        for it in (0, -1):
            out.append(float(c.get_energy_spectrum(lambda0_um=1, iteration=it, on_device=True).sum()))
        n_w, n_2, n_p = grid
        dev = [torch.as_tensor(np.ascontiguousarray(v.swapaxes(-1, -3)), device='cuda:0') for v in c.Data['radiation'].values()]
        r = engine.energy_spectrum(c.Args['mode'], dev, comp.endswith('complex'), 2, n_w, n_2, n_p, -1,
                                   c.Args['radius'] if near else c.Args['theta'], float(c.Args['dph']), layout=0)
        out.append(float(r.sum().item()))
print('sanitize cases2 done', len(out), np.isfinite(out).all())
