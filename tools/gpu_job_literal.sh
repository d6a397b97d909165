# literal-fp32 kernel: parity (reference vectors, C4) + throughput on the C5 recipe and the C4 config
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 -k "literal or c4 or float or reference_vectors" 2>&1 | tail -3
python - <<'PY' 2>&1 | grep -v WARN | grep -v "^$"
import sys, os; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from synchrad.calc import SynchRad
from synchrad_b200 import engine, synthetic
for native in (False,):
    args = synthetic.c5_args((256, 32, 32), dtype='float'); args['ctx'] = [0, 0]; args['float_mode'] = 'literal'
    calc = SynchRad(args); calc.Args['timeStep'] = calc.dtype(synthetic.C5_DT)
    batch = synthetic.c5_batch(592, 10000, device='cuda:0')
    best = 1e30
    for r in range(3):
        res = engine.integrate(calc.Args, calc.dtype, calc._grid, None, 'total', 1, device_tracks=batch, timing=True, timeStep=synthetic.C5_DT)
        if r: best = min(best, res.elapsed_ms)
    upd = 592 * 9999 * 262144
    print(f'literal C5 probe: {best:.1f} ms  {upd / best * 1e3:.4e} updates/s kind={res.kind} tw={res.info.tile_width} blocks={res.info.grid_blocks}')
PY
