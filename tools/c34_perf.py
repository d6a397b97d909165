"""C3 / C4 recipes (guard-dominated BASELINE configs): integrate_ms of the library selected by SYNCHRAD_B200_LIB
(development aid).  usage: python tools/c34_perf.py [c3] [c4] [c4d]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import cases
from synchrad.calc import SynchRad
which = sys.argv[1:] or ['c3', 'c4']
sys.stdout = sys.stderr
tag = f"lib={os.path.basename(os.environ.get('SYNCHRAD_B200_LIB', 'default'))} carve={os.environ.get('SRB_CARVEOUT', '1')}"


def run(name, args, tracks, dt, **kw):
    calc = SynchRad(dict(args))
    best = 1e9
    for r in range(3):
        calc.calculate_spectrum(tracks, timeStep=dt, verbose=False, **kw)
        if r:
            best = min(best, calc.last_run['integrate_ms'])
    lr = calc.last_run
    key = sorted(calc.Data['radiation'])[0]
    sys.__stdout__.write(f"{tag} {name}: integrate_ms={best:.2f} kernel={lr['kernel']} tw={lr['tile_width']} pc={lr['particle_chunks']} "
                         f"checksum={float(calc.Data['radiation'][key].sum()):.12e}\n")


if 'c3' in which:
    trb, dtb, infob = cases.betatron_tracks(1000, seed=0)
    run('C3', cases.betatron_args(infob), trb, dtb, comp='cartesian')
if 'c4' in which:
    trs, dts, infos = cases.spiral_tracks(3000, seed=0)
    run('C4 mixed fp32 (3000 p)', cases.spiral_args(infos), trs, dts)
if 'c4d' in which:
    trs, dts, infos = cases.spiral_tracks(3000, seed=0)
    run('C4 double coherent (3000 p)', cases.spiral_args(infos, dtype='double'), trs, dts, comp='cartesian_complex')
