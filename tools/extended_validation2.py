"""Round-2 extended validation on a GPU box: random small problems (tests/fuzzcases.py) through every kernel the planner
can reach -- auto (incl. the on-device choice), direct, recurrence (incl. the lane = step path), corrected recurrence --
and with the time-axis split forced, against the strict oracle.  usage: python tools/extended_validation2.py [seeds]"""
import contextlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import fuzzcases
from oracle import reference_path as rp
from synchrad.calc import SynchRad
from synchrad_b200 import host


def gpu(args, tracks, dt, phasor='auto', **kw):
    a = dict(args); a['phasor'] = phasor
    with contextlib.redirect_stdout(io.StringIO()):
        c = SynchRad(a); c.calculate_spectrum([list(t) for t in tracks], timeStep=dt, verbose=False, **kw)
    return c


worst = {}; n = 0; fails = 0; kinds = {}; t0 = time.time()
for seed in range(500, 500 + (int(sys.argv[1]) if len(sys.argv) > 1 else 20)):
    rs = np.random.RandomState(seed)
    for i in range(25):
        A, tracks, dt, kw = fuzzcases.rand_case(rs)
        with contextlib.redirect_stdout(io.StringIO()):
            ref = rp.calculate_spectrum(A, tracks, dt, **kw)
        Ai, dtype = host.init_args(dict(A))
        uniform = host.omega_is_uniform(Ai)
        double = dtype is np.double
        runs = [('auto', None), ('direct', None), ('auto', '3')]
        if uniform:
            runs += [('recur', None), ('recur', '2')]
            if double:
                runs += [('drec', None)]
        for phasor, split in runs:
            if split is None:
                os.environ.pop('SRB_TIME_SPLIT', None)
            else:
                os.environ['SRB_TIME_SPLIT'] = split
            try:
                c = gpu(A, tracks, dt, phasor=phasor, **kw)
            except RuntimeError as err:          # the corrected recurrence refuses omega * L beyond 3e10 rad
                assert phasor == 'drec' and 'first-order' in str(err), err
                kinds['drec refused'] = kinds.get('drec refused', 0) + 1
                continue
            e = fuzzcases.vector_errors(c.Data['radiation'], ref['radiation'])
            tol = 1e-9 if double else 1e-4
            key = (phasor, split, 'f64' if double else 'f32')
            worst[key] = max(worst.get(key, 0.0), e); n += 1
            kinds[c.last_run['kernel']] = kinds.get(c.last_run['kernel'], 0) + 1
            if not (e <= tol):
                fails += 1
                print('FAIL', seed, i, key, e, A['grid'], A.get('mode'), A.get('Features'), kw, c.last_run['kernel'])
        # single precision: the mixed mode against the fp64 oracle, the literal mode against the oracle's fp32 restatement
        # of the reference (both 1e-4, north_star's single-precision tolerance)
        os.environ.pop('SRB_TIME_SPLIT', None)
        if seed % 2 == 0:
            A32 = dict(A); A32['dtype'] = 'float'
            with contextlib.redirect_stdout(io.StringIO()):
                ref32 = rp.calculate_spectrum(A32, tracks, dt, **kw)
            for mode, want in (('mixed', ref), ('literal', ref32)):
                Am = dict(A32); Am['float_mode'] = mode
                c = gpu(Am, tracks, dt, **kw)
                e = fuzzcases.vector_errors(c.Data['radiation'], want['radiation'])
                key = ('auto', mode, 'f32')
                worst[key] = max(worst.get(key, 0.0), e); n += 1
                kinds[c.last_run['kernel']] = kinds.get(c.last_run['kernel'], 0) + 1
                if not (e <= 1e-4):
                    fails += 1
                    print('FAIL', seed, i, key, e, A['grid'], A.get('mode'), A.get('Features'), kw, c.last_run['kernel'])
os.environ.pop('SRB_TIME_SPLIT', None)
print(f'fuzz: {n} runs, {fails} failures, {time.time() - t0:.0f} s; kernels that ran: {kinds}')
for k in sorted(worst, key=str):
    print(f'  phasor={k[0]:7s} time_split={k[1]} {k[2]}: worst whole-vector error {worst[k]:.3e}')
