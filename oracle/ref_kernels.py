"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

oracle/_ref: the reference's OWN kernels as host libraries.  `build()` (run in the build container, where
/root/reference exists) renders kernel_farfield.cl / kernel_nearfield.cl the way calc.py:605-624 does
(`${my_dtype}`, `${f_native}`), wraps them with the prelude oracle/clshim/cl_shim.hpp and an OpenMP NDRange loop
(clshim/pyopencl: `Program.translation_unit`) and compiles them with g++ from where the sources lie -- piped on
stdin, only the binaries and a manifest of the kernel signatures are written, into oracle/_ref/ (git-ignored,
shipped to the GPU box with the snapshot).  Two flavours:

    strict : -O2 -ffp-contract=off            the checker (what tests/golden/reference_cases.npz was made with)
    fast   : -O3 -march=x86-64-v3 -ffp-contract=fast   the CPU baseline of bench.py (`cpu_baseline.kind = "reference"`)

`load()` needs only oracle/_ref (no reference sources), so `bench.py --impl reference` and the GPU-box tests can
drive the reference's kernels; the per-particle launch loop around them is oracle/reference_path.py.
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, '_ref')
SHIM_DIR = os.path.join(_HERE, 'clshim')
REFERENCE_ROOT = os.environ.get('SYNCHRAD_REFERENCE', '/root/reference')
FLAVOURS = {'strict': '-O2 -ffp-contract=off -fno-fast-math',
            'fast': '-O3 -march=x86-64-v3 -ffp-contract=fast'}
SOURCES = {'far': 'kernel_farfield.cl', 'near': 'kernel_nearfield.cl'}
KERNEL_OF_COMP = {'total': 'total', 'cartesian': 'cartesian_comps', 'cartesian_complex': 'cartesian_comps_complex',
                  'spheric': 'spheric_comps', 'spheric_complex': 'spheric_comps_complex'}     # calc.py:324-353


def _shim_pyopencl():
    """The stand-in module, imported under a private name so that a real pyopencl is never shadowed."""
    import importlib.util
    name = '_clshim_pyopencl'
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(SHIM_DIR, 'pyopencl', '__init__.py'),
                                                  submodule_search_locations=[os.path.join(SHIM_DIR, 'pyopencl')])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _lib_name(mode, dtype, flavour):
    return f'libref_{mode}_{dtype}_{flavour}.so'


def reference_present():
    return all(os.path.exists(os.path.join(REFERENCE_ROOT, 'synchrad', s)) for s in SOURCES.values())


def available(flavour='fast'):
    return os.path.exists(os.path.join(REF_DIR, 'manifest.json')) and all(
        os.path.exists(os.path.join(REF_DIR, _lib_name(m, d, flavour))) for m in SOURCES for d in ('double', 'float'))


def build(force=False):
    """Compile the reference's kernel sources into oracle/_ref (no-op without /root/reference)."""
    if not reference_present():
        return False
    if available('fast') and available('strict') and not force:
        return True
    cl = _shim_pyopencl()
    import re
    os.makedirs(REF_DIR, exist_ok=True)
    manifest = {'_about': 'kernel signatures of the reference kernels compiled into this directory by '
                          'oracle/ref_kernels.py (types and parameter names only)', 'flavours': FLAVOURS}
    for mode, fname in SOURCES.items():
        with open(os.path.join(REFERENCE_ROOT, 'synchrad', fname)) as f:
            text = f.read()
        for dtype in ('double', 'float'):
            src = re.sub(r'\$\{\s*(\w+)\s*\}', lambda m: {'my_dtype': dtype, 'f_native': ''}[m.group(1)], text)
            kernels, tu = cl.Program(None, src).translation_unit()
            manifest[f'{mode}_{dtype}'] = {k: [list(p) for p in sig] for k, sig in kernels.items()}
            for flavour, flags in FLAVOURS.items():
                out = os.path.join(REF_DIR, _lib_name(mode, dtype, flavour))
                cmd = ['g++', '-std=c++17', *flags.split(), '-fopenmp', '-fPIC', '-shared', '-Wno-unknown-pragmas',
                       '-I', SHIM_DIR, '-x', 'c++', '-', '-o', out]
                r = subprocess.run(cmd, input=tu.encode(), capture_output=True)
                if r.returncode != 0:
                    raise RuntimeError(f'reference kernels ({fname}, {dtype}, {flavour}) failed to build:\n'
                                       + r.stderr.decode()[-3000:])
    with open(os.path.join(REF_DIR, 'manifest.json'), 'w') as f:
        json.dump(manifest, f, indent=1)
    return True


_loaded = {}


def load(mode, dtype, flavour='fast'):
    """The compiled reference program for (mode, dtype): an object with one callable per kernel,
    `prog.total(queue, (global,), (local,), *args)` like a built pyopencl Program."""
    key = (mode, dtype, flavour)
    if key not in _loaded:
        if not available(flavour):
            raise RuntimeError('oracle/_ref is missing: run __graft_entry__.build() where /root/reference exists')
        cl = _shim_pyopencl()
        with open(os.path.join(REF_DIR, 'manifest.json')) as f:
            sigs = json.load(f)[f'{mode}_{dtype}']
        kernels = {k: [tuple(p) for p in sig] for k, sig in sigs.items()}
        _loaded[key] = cl._Built(ctypes.CDLL(os.path.join(REF_DIR, _lib_name(mode, dtype, flavour))), kernels)
    return _loaded[key]


def process_track(prog, mode, comp, spectra, arrs, wp, it_start, it_end, tables, L_screen, grid_nums, dt, nSnaps,
                  snaps, form_factor, dtype, wgs=32, Buffer=None):
    """One kernel launch for one particle: the argument list of calc.py:292-353 (`_process_track`).
    `Buffer`: the buffer wrapper of the pyopencl look-alike `prog` came from (default: oracle/clshim's)."""
    B = Buffer if Buffer is not None else _shim_pyopencl().Buffer
    n_nodes = int(np.prod(grid_nums))
    if n_nodes <= wgs:                                         # calc.py:640-646 (`_get_wgs`, CPU device: WGS = 32)
        lsz, gsz = n_nodes, n_nodes
    else:
        lsz, gsz = wgs, int(np.ceil(1. * n_nodes / wgs)) * wgs
    args = [B(a) for a in arrs] + [dtype(wp), np.uint32(it_start), np.uint32(it_end), np.uint32(arrs[0].size)]
    if mode == 'far':
        args += [B(tables[k]) for k in ('omega', 'axisA', 'axisB', 'sinPhi', 'cosPhi')]
    else:
        args += [B(tables[k]) for k in ('omega', 'axisA', 'sinPhi', 'cosPhi')] + [dtype(L_screen)]
    args += [np.uint32(v) for v in grid_nums] + [dtype(dt), np.uint32(nSnaps), B(snaps)]
    if comp in ('cartesian_complex', 'spheric_complex'):
        args += [B(form_factor)]
    kern = getattr(prog, KERNEL_OF_COMP[comp], None)
    if kern is None:
        raise AttributeError(f'{mode}-field program has no kernel {KERNEL_OF_COMP[comp]!r}')
    kern(None, (gsz,), (lsz,), *[B(s) for s in spectra], *args)
