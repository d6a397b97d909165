"""TEST INFRASTRUCTURE ONLY — drives the single-warp CPU emulation of the device code
(tests/emu/emu.cpp) through the SAME host packing the product uses, so kernel logic can be
checked against the oracle without a GPU.  Never imported by the product."""
import ctypes
import os
import subprocess

import numpy as np

from synchrad_b200 import _lib, host

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libsrb_emu.so')
_SRC = [os.path.join(_HERE, 'emu.cpp'),
        os.path.join(_HERE, '..', '..', 'synchrad_b200', 'csrc', 'srb_core.cuh'),
        os.path.join(_HERE, '..', '..', 'synchrad_b200', 'csrc', 'srb_pair.cuh'),
        os.path.join(_HERE, '..', '..', 'synchrad_b200', 'csrc', 'srb_literal.cuh'),
        os.path.join(_HERE, '..', '..', 'synchrad_b200', 'csrc', 'srb_ws.cuh'),
        os.path.join(_HERE, '..', '..', 'synchrad_b200', 'csrc', 'srb_drec.cuh')]


def build(so=None, defines=()):
    so = so or _SO
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in _SRC):
        return
    subprocess.check_call(['/usr/bin/g++', '-std=c++17', '-O2', '-mfma', '-ffp-contract=off', '-fopenmp',
                           *[f'-D{d}' for d in defines], '-shared', '-fPIC', '-o', so, _SRC[0]])


def run(Args, tracks, timeStep, comp='total', L_screen=None, it_range=None, nSnaps=1,
        sigma_particle=0, kind='recur', tw=None, nPC=1, prepass=True, nTS=1, variant=None):
    """Returns (radiation dict in host layout, counters).  variant = (name, defines): a separately built emulator."""
    A = dict(Args)
    A, dtype = host.init_args(A)
    A['sigma_particle'] = dtype(sigma_particle)
    if A['mode'] == 'near':
        A['L_screen'] = L_screen
    A['timeStep'] = dtype(timeStep)
    literal_pre = A.get('float_mode') == 'literal' and A.get('dtype', 'double') != 'double'
    weights = [float(np.float32(t[6])) if literal_pre else t[6] for t in tracks]
    pk = host.pack_tracks(tracks, weights, np.double, it_range, nSnaps)
    spectra, cnt = run_packed(A, dtype, pk, timeStep, comp, nSnaps, kind=kind, tw=tw, nPC=nPC, prepass=prepass, nTS=nTS,
                              variant=variant)
    rad = {k: np.ascontiguousarray(s.swapaxes(-1, -3)) for k, s in zip(host.COMP_KEYS[comp], spectra)}
    return rad, cnt


def run_packed(A, dtype, pk, timeStep, comp, nSnaps, kind='recur', tw=None, nPC=1, prepass=True, nTS=1, variant=None,
               spectra=None):
    """The emulated srb_integrate on tracks already in the C-ABI layout (host.PackedTracks); `A` are initialised Args
    with sigma_particle / L_screen set.  Returns (spectra in the DEVICE layout (nSnaps, nPhi, nAxis2, nOmega), accumulated
    into `spectra` when given like the C ABI does; (passed, visited))."""
    so = _SO if variant is None else os.path.join(_HERE, f'libsrb_emu_{variant[0]}.so')
    build(so, () if variant is None else variant[1])
    lib = ctypes.CDLL(so)
    L_screen = A.get('L_screen')
    T = host.grid_tables(A)
    n_w, n_2, n_p = (int(v) for v in A['gridNodeNums'])
    g = _lib.srb_grid()
    g.mode, g.comp = _lib.MODE[A['mode']], _lib.COMP[comp]
    literal = host.float_mode(A) == 'literal'
    g.dtype = 0 if dtype is np.double else (2 if literal else 1)
    g.omega_uniform = 1 if host.omega_is_uniform(A) else 0
    g.nOmega, g.nAxis2, g.nPhi, g.nSnaps = n_w, n_2, n_p, nSnaps
    g.omega = T['omega'].ctypes.data
    g.sinPhi, g.cosPhi = T['sinPhi'].ctypes.data, T['cosPhi'].ctypes.data
    if A['mode'] == 'far':
        g.sinTheta, g.cosTheta = T['sinTheta'].ctypes.data, T['cosTheta'].ctypes.data
    else:
        g.radius = T['radius'].ctypes.data
        g.L_screen = float(np.float32(L_screen)) if literal else float(L_screen)
    ff = host.form_factor(A)
    g.formFactor = ff.ctypes.data
    g.dt = float(np.float32(timeStep)) if literal else float(timeStep)
    g.omega_first_host, g.omega_last_host = float(T['omega'][0]), float(T['omega'][-1])
    t = _lib.srb_tracks()
    t.nTracks = pk.n
    for nm, a in zip(('x', 'y', 'z', 'ux', 'uy', 'uz'), pk.coords):
        setattr(t, nm, a.ctypes.data)
    t.offsets, t.w = pk.offsets.ctypes.data, pk.w.ctypes.data
    t.itStart, t.itEnd, t.itSnaps = pk.itStart.ctypes.data, pk.itEnd.ctypes.data, pk.itSnaps.ctypes.data
    t.itSnapsStride, t.totalSteps_host = pk.snapStride, pk.total
    keys = host.COMP_KEYS[comp]
    if spectra is None:
        spectra = [np.zeros((nSnaps, n_p, n_2, n_w)) for _ in keys]
    sp = (ctypes.c_void_p * len(keys))(*[s.ctypes.data for s in spectra])
    kind_i = {'direct': 0, 'recur': 1, 'pair': 3, 'pair_fma': 4, 'drec': 5, 'pair_ws': 6}[kind]
    if literal:
        kind_i, tw = 0, None   # the C side switches to the literal kind; tile widths of the direct layout
    if tw is None:
        tiles = 16 if kind_i == 1 else 32
        opts = ([4, 8, 16] if A['mode'] == 'far' else [2, 4, 8]) if kind_i == 1 else ([4, 8] if kind_i == 4 else ([2, 4, 8, 16] if kind_i == 3 and not comp.startswith('spheric') else ([4, 8] if kind_i == 6 else [2, 4, 8])))   # as make_plan (srb_api.cu)
        tw = next((o for o in opts if tiles * o >= n_w), opts[-1])
    cnt = (ctypes.c_ulonglong * 2)(0, 0)
    lib.srb_emu_integrate.restype = ctypes.c_int
    rc = lib.srb_emu_integrate(ctypes.byref(g), ctypes.byref(t), sp, len(keys), kind_i, int(tw),
                               ctypes.c_uint32(nPC), cnt, ctypes.c_int(1 if prepass else 0), ctypes.c_uint32(nTS))
    assert rc == 0, 'emulator has no such configuration'
    return spectra, (cnt[0], cnt[1])
