"""Device-resident throughput of the hot path on the C5 recipe (development aid).
usage: python tools/quick_perf.py [particles] [steps] [dtype] [phasor] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synchrad.calc import SynchRad
from synchrad_b200 import engine, synthetic

n_p = int(sys.argv[1]) if len(sys.argv) > 1 else 592
n_s = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
dtype = sys.argv[3] if len(sys.argv) > 3 else 'double'
phasor = sys.argv[4] if len(sys.argv) > 4 else 'auto'
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
grid = tuple(int(v) for v in os.environ.get('GRID', '256,32,32').split(','))
sys.stdout = sys.stderr
args = synthetic.c5_args(grid, dtype=dtype); args['ctx'] = [0, 0]; args['phasor'] = phasor
mode = os.environ.get('MODE', 'far'); comp = os.environ.get('COMP', 'total')
L = float(os.environ.get('LSCREEN', '1e5'))
if mode == 'near':
    args['mode'] = 'near'
    args['grid'][1] = (0.0, L * 0.03)
    args['L_screen'] = L
if os.environ.get('NATIVE'):
    args['native'] = True
calc = SynchRad(args)
if mode == 'near':
    calc.Args['L_screen'] = L
calc.Args['timeStep'] = calc.dtype(synthetic.C5_DT)
batch = synthetic.c5_batch(n_p, n_s, device='cuda:0')
upd = n_p * (n_s - 1) * int(np.prod(grid))
best = 1e30
for r in range(reps + 1):
    res = engine.integrate(calc.Args, calc.dtype, calc._grid, None, comp, 1, phasor=phasor, native=bool(os.environ.get('NATIVE')),
                           device_tracks=batch, timing=True, timeStep=synthetic.C5_DT)
    if r:
        best = min(best, res.elapsed_ms)
i = res.info
sys.__stdout__.write(f"mode={mode} comp={comp} native={bool(os.environ.get('NATIVE'))} grid={grid} lib={os.environ.get('SYNCHRAD_B200_LIB','default')} {dtype} {phasor} kind={res.kind} tw={i.tile_width} "
                     f"pc={i.n_particle_chunks} blocks={i.grid_blocks} thr={i.block_threads} smem={i.smem_bytes} "
                     f"ms={best:.2f} updates/s={upd / best * 1e3:.4e} checksum={float(res.spectra[0].sum()):.10e}\n")
