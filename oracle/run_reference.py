"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

Runs the UNMODIFIED reference (`/root/reference/synchrad`: calc.py, utils.py, kernel_*.cl) in this image.
The reference needs pyopencl, mako and h5py, none of which is installed; `oracle/clshim/` provides stand-ins
that compile the reference's own kernel sources for the host (see clshim/pyopencl/__init__.py).  Everything
else -- argument handling, axes, tables, the per-particle launch loop, snapshots, the axis swap, the HDF5
writer calls, utils.py post-processing -- is the reference's code, imported from where it lies.

Because the repo ships its own `synchrad` alias package, the reference is imported in a separate process whose
sys.path has /root/reference first and the repo root absent:

    from oracle import run_reference
    res = run_reference.run(args, tracks, timeStep=dt, comp='total', ...)        # -> dict like reference_path's

/root/reference does not exist on the GPU box: only tests/golden/make_reference_golden.py (run here, output
committed) and the `not gpu` tests (skipped when the reference is absent) call this.
"""
import os
import pickle
import subprocess
import sys
import tempfile

REFERENCE_ROOT = os.environ.get('SYNCHRAD_REFERENCE', '/root/reference')
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def available():
    return os.path.exists(os.path.join(REFERENCE_ROOT, 'synchrad', 'calc.py'))


def run(args, tracks=None, cxxflags=None, threads=None, post=(), backend='clshim', **kw):
    """One `SynchRad(args).calculate_spectrum(tracks, **kw)` of the reference.  `post`: names of Utilities
    methods to evaluate afterwards, each as (method, kwargs).  Returns dict(radiation, total_weight, Args,
    snap_iterations, post)."""
    if not available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    with tempfile.TemporaryDirectory() as tmp:
        req, out = os.path.join(tmp, 'req.pkl'), os.path.join(tmp, 'out.pkl')
        with open(req, 'wb') as f:
            pickle.dump(dict(args=args, tracks=tracks, kw=kw, post=list(post)), f)
        env = dict(os.environ)
        env.pop('PYTHONPATH', None)
        if cxxflags is not None:
            env['CLSHIM_CXXFLAGS'] = cxxflags
        if threads is not None:
            env['OMP_NUM_THREADS'] = str(threads)
        env['SYNCHRAD_REFERENCE_BACKEND'] = backend
        r = subprocess.run([sys.executable, os.path.abspath(__file__), req, out], env=env, cwd=tmp,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('reference run failed:\n' + r.stdout[-2000:] + r.stderr[-4000:])
        with open(out, 'rb') as f:
            res = pickle.load(f)
        res['log'] = r.stdout
        return res


def run_many(requests, cxxflags=None, threads=None, backend='clshim'):
    """Several runs in ONE child process: `requests` is a list of dict(args=, tracks=, kw=, post=[...]); returns
    the list of results (an exception raised by the reference for one request is returned as
    dict(error=repr) in its slot)."""
    if not available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    with tempfile.TemporaryDirectory() as tmp:
        req, out = os.path.join(tmp, 'req.pkl'), os.path.join(tmp, 'out.pkl')
        with open(req, 'wb') as f:
            pickle.dump(dict(many=[dict(args=r['args'], tracks=r.get('tracks'), kw=r.get('kw', {}),
                                        post=list(r.get('post', ()))) for r in requests]), f)
        env = dict(os.environ)
        env.pop('PYTHONPATH', None)
        if cxxflags is not None:
            env['CLSHIM_CXXFLAGS'] = cxxflags
        if threads is not None:
            env['OMP_NUM_THREADS'] = str(threads)
        env['SYNCHRAD_REFERENCE_BACKEND'] = backend
        r = subprocess.run([sys.executable, os.path.abspath(__file__), req, out], env=env, cwd=tmp,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('reference run failed:\n' + r.stdout[-2000:] + r.stderr[-4000:])
        with open(out, 'rb') as f:
            return pickle.load(f)


def run_script(path, backend='clshim', threads=None, seed=None):
    """Execute one of the reference's own scripts (e.g. /root/reference/tests/test_undulator_analytic.py) UNMODIFIED
    in a child process, on the stand-ins (`backend='clshim'`: the reference's kernels compiled for the host) or on
    the product's PyOpenCL-signature binding with emulated kernels (`backend='compat_emu'`).  `seed` seeds
    numpy's global RNG before the script starts (the scripts draw their energy spread unseeded).  Returns stdout."""
    if not available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    env = dict(os.environ)
    env.pop('PYTHONPATH', None)
    env['SYNCHRAD_REFERENCE_BACKEND'] = backend
    if threads is not None:
        env['OMP_NUM_THREADS'] = str(threads)
    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--script', os.path.abspath(path),
                            '' if seed is None else str(int(seed))], env=env, cwd=tmp, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('reference script failed:\n' + r.stdout[-2000:] + r.stderr[-4000:])
    return r.stdout


def _intern(v):
    """calc.py compares option strings with `is`; literals are interned, unpickled strings are not."""
    if isinstance(v, str):
        return sys.intern(v)
    if isinstance(v, dict):
        return {_intern(k): _intern(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_intern(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_intern(x) for x in v)
    return v


def _child_setup():
    import warnings
    warnings.filterwarnings('ignore', category=SyntaxWarning)
    bad = {os.path.abspath(p or os.getcwd()) for p in (_REPO, _HERE)}
    backend = os.environ.get('SYNCHRAD_REFERENCE_BACKEND', 'clshim')
    sys.path[:] = [REFERENCE_ROOT] + [p for p in sys.path if os.path.abspath(p or os.getcwd()) not in bad] \
        + [os.path.join(_HERE, 'clshim')]            # appended: real pyopencl/mako/h5py win when installed
    if backend == 'compat_emu':
        # the product's reference-side binding (synchrad_b200/compat: pyopencl facade over the C ABI) under the
        # unmodified reference, with the CPU emulation of the kernels (tests/emu) standing in for
        # srb_integrate_host -- checks the binding's marshalling in a container without a GPU.  /root/reference
        # stays first (the repo's `synchrad` alias must not shadow it), the repo root last (synchrad_b200 imports).
        sys.path.insert(1, os.path.join(_REPO, 'synchrad_b200', 'compat'))
        sys.path.append(_REPO)
        import ctypes
        import pyopencl
        assert 'compat' in pyopencl.__file__, pyopencl.__file__
        emu = ctypes.CDLL(os.path.join(_REPO, 'tests', 'emu', 'libsrb_emu.so'))
        emu.srb_emu_integrate.restype = ctypes.c_int

        def emu_integrate_host(grid, tracks, spectra_ptrs, n_spectra, device):
            tw = next((o for o in (2, 4, 8) if 32 * o >= grid.nOmega), 8)
            cnt = (ctypes.c_ulonglong * 2)(0, 0)
            rc = emu.srb_emu_integrate(ctypes.byref(grid), ctypes.byref(tracks), spectra_ptrs, n_spectra, 0, tw,
                                       ctypes.c_uint32(1), cnt, ctypes.c_int(1), ctypes.c_uint32(1))
            if rc != 0:
                raise RuntimeError('emulator has no such configuration')
        pyopencl._integrate_host = emu_integrate_host


def _child(req_path, out_path):
    _child_setup()
    import numpy as np
    with open(req_path, 'rb') as f:
        req = _intern(pickle.load(f))
    from synchrad.calc import SynchRad
    import synchrad
    assert os.path.abspath(synchrad.__path__[0]).startswith(os.path.abspath(REFERENCE_ROOT)), synchrad.__path__

    def one(req):
        args = dict(req['args'])
        args.setdefault('ctx', [0, 0])
        calc = SynchRad(args)
        kw = dict(req['kw'])
        kw.setdefault('verbose', False)
        tracks = req['tracks']
        if tracks is not None:
            tracks = [list(t) for t in tracks]
            calc.calculate_spectrum(particleTracks=tracks, **kw)
        else:
            calc.calculate_spectrum(**kw)
        snaps = calc.snap_iterations.get() if hasattr(calc.snap_iterations, 'get') else calc.snap_iterations
        keep = {k: v for k, v in calc.Args.items() if k not in ('grid', 'ctx')}
        post = {}
        for i, (meth, pkw) in enumerate(req['post']):
            r = getattr(calc, meth)(**pkw)
            post[i] = tuple(np.asarray(x) for x in r) if isinstance(r, tuple) else np.asarray(r)
        return dict(radiation=calc.Data['radiation'], total_weight=float(calc.total_weight), Args=keep,
                    snap_iterations=np.asarray(snaps), post=post,
                    device=f'{calc.dev_type} {calc.dev_name} / {calc.ocl_version}')

    if 'many' in req:
        res = []
        for r in req['many']:
            try:
                res.append(one(r))
            except Exception as exc:                      # reported per request, like the reference would raise
                res.append(dict(error=f'{type(exc).__name__}: {exc}'))
    else:
        res = one(req)
    with open(out_path, 'wb') as f:
        pickle.dump(res, f)


if __name__ == '__main__':
    if sys.argv[1] == '--script':
        _child_setup()
        import runpy
        if sys.argv[3]:
            import numpy
            numpy.random.seed(int(sys.argv[3]))
        runpy.run_path(sys.argv[2], run_name='__main__')
    else:
        _child(sys.argv[1], sys.argv[2])
