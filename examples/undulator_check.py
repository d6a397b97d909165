"""Analytic undulator check, far and near field, double then float — the same physics and the same
acceptance criterion as the reference's tests/test_undulator_analytic*.py (integrated energy vs
(7 pi/24)/137 K0^2 (1+K0^2/2) N_periods per electron), run through the drop-in API.

    python examples/undulator_check.py [Np]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases                                   # noqa: E402  (track recipe shared with the tests)
from synchrad.calc import SynchRad             # noqa: E402
from synchrad.utils import J_in_um             # noqa: E402

Np = int(sys.argv[1]) if len(sys.argv) > 1 else 24
for near in (False, True):
    tracks, dt, info = cases.undulator_tracks(Np, near=near, seed=0)
    calc_input = cases.undulator_args(info, near=near)
    del calc_input['dtype']
    kw = dict(L_screen=1e5) if near else {}
    theory = cases.undulator_energy_theory(info, J_in_um)
    calc = SynchRad(calc_input)
    for label in ('double precision', 'single precision + native'):
        t0 = time.time()
        calc.calculate_spectrum(tracks.copy(), timeStep=dt, comp='total', Np_max=Np, **kw)
        if calc.rank == 0:
            dev = abs(calc.get_energy(lambda0_um=1) - theory) / theory
            print('{:s}field, {:s}: {:d} particle(s) in {:.3g} s; deviation from analytic estimate {:.2f}%  [{}]'
                  .format(calc.Args['mode'], label, Np, time.time() - t0, 100 * dev, calc.last_run['kernel']))
        calc.Args['dtype'] = 'float'
        calc.Args['native'] = True
        calc._init_args(calc.Args)
        calc._init_data()
        calc._compile_kernels()
