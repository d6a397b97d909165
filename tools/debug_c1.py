"""debug aid: C1 (single electron, 128x32x32) on the GPU against the oracle, a few repetitions"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import cases
from oracle import reference_path as rp
from synchrad.calc import SynchRad
sys.stdout = sys.stderr
Np = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tracks, dt, info = cases.undulator_tracks(Np, seed=0 if Np > 1 else None)
args = cases.undulator_args(info)
ref = rp.calculate_spectrum(args, tracks, dt)['radiation']['total']
for rep in range(4):
    calc = SynchRad(dict(args))
    calc.calculate_spectrum([list(t) for t in tracks], timeStep=dt, verbose=False)
    got = calc.Data['radiation']['total']
    d = np.abs(got - ref)
    bad = np.argwhere(d > 1e-9 * ref.max())
    sys.__stdout__.write(f"lib={os.environ.get('SYNCHRAD_B200_LIB','default')[-20:]} notma={os.environ.get('SRB_WS_NOTMA')} rep{rep}: max rel {d.max()/ref.max():.3e} bad nodes {len(bad)} "
                         f"first bad (w,th,ph) {bad[:3].tolist()} nan={np.isnan(got).sum()} kernel={calc.last_run['kernel']} blocks={calc.last_run['grid_blocks']}\n")
