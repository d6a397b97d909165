"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

Host-side restatement of the reference's `SynchRad.calculate_spectrum` flow, driving the
CPU kernels of `oracle_kernels.cpp` one particle at a time exactly as the reference drives
its OpenCL kernels.  Citations are file:line in /root/reference/synchrad/.

    grid / axes ................. calc.py:355-451   (`_init_args`)
    device tables ............... calc.py:486-512   (`_init_data`: 2*pi*omega in dtype, sin/cos)
    spectrum shapes / keys ...... calc.py:453-484   (`_init_raditaion`, FormFactor)
    snapshot iterations ......... calc.py:626-630   (`_set_snap_iterations`)
    per-track marshalling ....... calc.py:292-322, 579-603
    track selection / weights ... calc.py:230-265   (Np_max, [rank::size], weights_normalize)
    D2H + axis swap ............. calc.py:573-577
    MPI sum ..................... calc.py:560-571   (emulated by summing rank results)
    post-processing ............. utils.py:23-102   (`get_full_spectrum`/`get_energy`)

Parity pin: PINNED -- this flow plus oracle_kernels.cpp reproduce the outputs of the unmodified reference
(run in the build container through oracle/run_reference.py, stored in tests/golden/reference_cases.npz) bit for
bit, in double and single precision (tests/test_reference_pin.py).  `lib='ref_strict'|'ref_fast'` drives the
reference's own kernels from oracle/_ref (oracle/ref_kernels.py) with the same launch loop.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

COMP_CODES = {'total': 0, 'cartesian': 1, 'cartesian_complex': 2, 'spheric': 3,
              'spheric_complex': 4}
COMP_KEYS = {
    'total': ['total'],
    'cartesian': ['x', 'y', 'z'],
    'cartesian_complex': ['xre', 'xim', 'yre', 'yim', 'zre', 'zim'],
    'spheric': ['r', 'theta', 'phi'],
    'spheric_complex': ['rre', 'rim', 'thetare', 'thetaim', 'phire', 'phiim'],
}
CTYPES = {'double': 0, 'float': 1, 'longdouble': 2}

from scipy.constants import alpha as alpha_fs, c as _c, hbar as _hbar   # utils.py:3-4

_trapz = getattr(np, 'trapezoid', None) or np.trapz

J_in_um = 2e6 * np.pi * _hbar * _c  # utils.py:16


def build(force=False):
    """Compile the oracle libraries with the committed Makefile (g++, OpenMP)."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n))
                            for n in ('liboracle_strict.so', 'liboracle_fast.so'))
    if need:
        subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))


_libs = {}


def set_threads(n):
    """OpenMP threads of the oracle kernels (torchrun exports OMP_NUM_THREADS=1 and libgomp reads it
    when first loaded, usually by torch; this sets the runtime value afterwards)."""
    try:
        ctypes.CDLL('libgomp.so.1').omp_set_num_threads(int(n))
    except OSError:
        pass


def _lib(kind='strict'):
    if kind not in _libs:
        path = os.path.join(_HERE, f'liboracle_{kind}.so')
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.srb_oracle_particle.restype = ctypes.c_int
        lib.srb_oracle_particle.argtypes = (
            [ctypes.c_int] * 3 + [ctypes.POINTER(ctypes.c_void_p)] + [ctypes.c_void_p] * 6
            + [ctypes.c_double] + [ctypes.c_uint32] * 3 + [ctypes.c_void_p] * 5
            + [ctypes.c_double] + [ctypes.c_uint32] * 3 + [ctypes.c_double, ctypes.c_uint32,
                                                            ctypes.c_void_p, ctypes.c_void_p,
                                                            ctypes.POINTER(ctypes.c_longlong)])
        _libs[kind] = lib
    return _libs[kind]


# --------------------------------------------------------------------------- host restatement
def build_args(Args):
    """calc.py:355-451 — fills defaults and the spectral axes into a copy of Args."""
    A = dict(Args)
    A.setdefault('mode', 'far')
    A.setdefault('dtype', 'double')
    A.setdefault('ctx', None)
    A.setdefault('Features', [])
    dtype = np.double if A['dtype'] == 'double' else np.single
    A['gridNodeNums'] = A['grid'][-1]
    A['numGridNodes'] = int(np.prod(A['gridNodeNums']))
    No = A['gridNodeNums'][0]
    w_lo, w_hi = A['grid'][0]
    omega = np.linspace(w_lo, w_hi, No)            # np.r_[a:b:N*1j] is an inclusive linspace
    for feature in A['Features']:
        if feature == 'wavelengthGrid':
            A['wavelengths'] = np.linspace(1. / w_hi, 1. / w_lo, No)
            omega = 1. / A['wavelengths']
            break
        elif feature == 'logGrid':
            d_log_w = np.log(w_hi / w_lo) / (No - 1.0)
            omega = w_lo * np.exp(d_log_w * np.arange(No))
            break
    A['omega'] = omega.astype(dtype)
    A['dw'] = np.abs(omega[1:] - omega[:-1]) if No > 1 else np.array([1.], dtype=dtype)
    N2, Np = A['gridNodeNums'][1:]
    a_lo, a_hi = A['grid'][1]
    p_lo, p_hi = A['grid'][2]
    ax2 = np.linspace(a_lo, a_hi, N2)
    phi = p_lo + (p_hi - p_lo) / Np * np.arange(Np)   # end point excluded, calc.py:414,437
    d2 = ax2[1] - ax2[0] if N2 > 1 else (dtype(1.) if A['mode'] == 'far' else 1.)
    A['dph'] = phi[1] - phi[0] if Np > 1 else (dtype(1.) if A['mode'] == 'far' else 1.)
    A['phi'] = phi.astype(dtype)
    if A['mode'] == 'far':
        A['dth'] = d2
        A['theta'] = ax2.astype(dtype)
    else:
        A['dr'] = d2
        A['radius'] = ax2.astype(dtype)
    A['dV'] = A['dw'] * d2 * A['dph']
    return A, dtype


def device_tables(A, dtype):
    """calc.py:486-512 — the arrays the kernels read (host copies)."""
    D = {'omega': np.ascontiguousarray(dtype(2 * np.pi) * A['omega'])}
    D['sinPhi'] = np.ascontiguousarray(np.sin(A['phi']))
    D['cosPhi'] = np.ascontiguousarray(np.cos(A['phi']))
    if A['mode'] == 'far':
        D['axisA'] = np.ascontiguousarray(np.sin(A['theta']))
        D['axisB'] = np.ascontiguousarray(np.cos(A['theta']))
    else:
        D['axisA'] = np.ascontiguousarray(A['radius'])
        D['axisB'] = None
    for k, v in D.items():
        assert v is None or v.dtype == dtype, (k, v.dtype)
    return D


def snap_iterations(it_range, nSnaps):
    """calc.py:626-630"""
    return np.ascontiguousarray(
        np.linspace(it_range[0], it_range[1], int(nSnaps) + 1, dtype=np.uint32)[1:])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def calculate_spectrum(Args, particleTracks, timeStep, comp='total', L_screen=None,
                       Np_max=None, it_range=None, nSnaps=1, sigma_particle=0,
                       weights_normalize=None, ranks=1, ctype=None, lib='strict'):
    """Restated `SynchRad.calculate_spectrum` for the list-of-tracks input (calc.py:101-272).

    `ranks` emulates an mpirun of that size: each rank takes particleTracks[:Np][r::ranks],
    normalises weights rank-locally (Q7) and the float64 host spectra are summed
    (calc.py:560-571).  Returns dict(radiation=..., total_weight=..., Args=..., passed=...,
    updates=..., snap_iterations=...).  `ctype` overrides the compute type ('longdouble' truth).
    """
    A, dtype = build_args(Args)
    ct = CTYPES[ctype] if ctype else CTYPES[A['dtype']]
    sdtype = np.float32 if ct == 1 else np.float64      # storage dtype of the kernel arrays
    near = A['mode'] == 'near'
    A['sigma_particle'] = dtype(sigma_particle)
    if near:
        if L_screen is None:
            raise ValueError('Define L_screen argument for near-field calculation')
        A['L_screen'] = L_screen
    if near and comp in ('spheric', 'spheric_complex'):
        raise AttributeError(f'near-field kernels have no {comp} variant (calc.py:342)')
    D = device_tables(A, dtype)
    if ct == 2:   # truth run: identical (dtype-rounded) tables, stored as double
        D = {k: (None if v is None else v.astype(np.float64)) for k, v in D.items()}
    nSnaps = int(nSnaps)
    A['comp'] = comp
    if near:
        A['theta'] = np.arctan2(A['radius'], A['L_screen'])
    A['timeStep'] = dtype(timeStep)
    ff = np.exp(dtype(-0.5) * (dtype(2 * np.pi) * A['omega'] * A['sigma_particle']) ** 2)
    ff = np.ascontiguousarray(ff.astype(sdtype))
    keys = COMP_KEYS[comp]
    No, N2, Nphi = A['gridNodeNums']
    shape_dev = (nSnaps, Nphi, N2, No)
    if it_range is not None:
        it_range = tuple(it_range)
    Np = len(particleTracks)
    if Np_max is not None:
        Np = min(Np_max, Np)
    ref_prog, ref_buffer = None, None
    if lib in ('ref_strict', 'ref_fast'):       # the reference's own kernels from oracle/_ref (ref_kernels.py)
        from . import ref_kernels
        assert ct != 2, 'no long-double build of the reference kernels'
        ref_prog = ref_kernels.load(A['mode'], A['dtype'], lib[4:])
    elif lib == 'compat':
        # the PRODUCT's reference-side binding (synchrad_b200/compat: pyopencl call signature -> C ABI -> GPU) driven
        # with the reference's own argument list, one launch per particle -- what the unmodified calc.py would do
        from . import ref_kernels
        from synchrad_b200.compat import pyopencl as compat_cl
        ref_prog = compat_cl.Program(None, '', mode=A['mode'], dtype=A['dtype']).build()
        ref_buffer = compat_cl.Buffer
    else:
        lib_ = _lib(lib)
    total = {k: np.zeros((nSnaps, No, N2, Nphi)) for k in keys}
    total_weight = 0.0
    passed = ctypes.c_longlong(0)
    updates = 0
    snaps = snap_iterations(it_range, nSnaps) if it_range is not None else None
    for r in range(ranks):
        mine = [list(t) for t in particleTracks[:Np][r::ranks]]
        if weights_normalize in ('mean', 'max'):
            ws = [t[6] for t in mine]
            norm = np.mean(ws) if weights_normalize == 'mean' else np.max(ws)
        spectra = [np.zeros(shape_dev, dtype=sdtype) for _ in keys]
        sp_ptrs = (ctypes.c_void_p * len(keys))(*[s.ctypes.data for s in spectra])
        for t in mine:
            if weights_normalize in ('mean', 'max'):
                t[6] = t[6] / norm
            elif weights_normalize == 'ones':
                t[6] = 1.0
            total_weight += t[6]
            it_start = np.uint32(t[7]) if len(t) == 8 else np.uint32(0)
            arrs = [np.ascontiguousarray(np.asarray(c).astype(dtype)).astype(sdtype)
                    for c in t[:6]]
            n = arrs[0].size
            if it_range is None:                      # calc.py:297-301
                it_start = np.uint32(0)
                rng = (0, n)
                snaps = snap_iterations(rng, nSnaps)
            else:
                rng = it_range
            if ref_prog is not None:
                ref_kernels.process_track(ref_prog, A['mode'], comp, spectra, arrs, t[6], it_start, rng[-1], D,
                                          A['L_screen'] if near else None, A['gridNodeNums'], A['timeStep'], nSnaps,
                                          snaps, ff, dtype, Buffer=ref_buffer)
                updates += max(0, min(n - 1, int(rng[-1]) - 1)) * A['numGridNodes']
                continue
            rc = lib_.srb_oracle_particle(
                1 if near else 0, COMP_CODES[comp], ct, sp_ptrs,
                *[_ptr(a) for a in arrs], float(dtype(t[6])), int(it_start), int(rng[-1]), n,
                _ptr(D['omega']), _ptr(D['axisA']), _ptr(D['axisB']), _ptr(D['sinPhi']),
                _ptr(D['cosPhi']), float(dtype(A['L_screen'])) if near else 0.0,
                No, N2, Nphi, float(A['timeStep']), nSnaps, _ptr(snaps), _ptr(ff),
                ctypes.byref(passed))
            if rc != 0:
                raise RuntimeError('oracle kernel rejected its arguments')
            updates += max(0, min(n - 1, int(rng[-1]) - 1)) * A['numGridNodes']
        for k, s in zip(keys, spectra):               # calc.py:573-577 then :560-571
            total[k] += np.ascontiguousarray(s.swapaxes(-1, -3), dtype=np.double)
    return dict(radiation=total, total_weight=total_weight, Args=A, passed=None if ref_prog is not None else passed.value,
                updates=updates, snap_iterations=snaps)


# --------------------------------------------------------------------------- utils.py restated
def get_full_spectrum(res, lambda0_um=None, comp='total', iteration=-1, phot_num=False,
                      normalize_to_weights=False):
    """utils.py:23-73"""
    A, rad = res['Args'], res['radiation']
    val = 0.0
    if A['comp'].split('_')[-1] == 'complex':
        if comp == 'total':
            for k in rad:
                val = val + rad[k][iteration].astype(np.double) ** 2
        else:
            val = rad[comp + 're'][iteration] + 1j * rad[comp + 'im'][iteration]
    else:
        if comp == 'total':
            for k in rad:
                val = val + rad[k][iteration].astype(np.double)
        else:
            val = val + rad[comp][iteration].astype(np.double)
    if A['mode'] == 'far':
        val = alpha_fs / (4 * np.pi ** 2) * val
    else:
        val = alpha_fs * np.pi / 4 * val / (2 * np.pi) ** 2
    if normalize_to_weights:
        val = val / res['total_weight']
    if phot_num:
        val = val / A['omega'][:, None, None]
    elif lambda0_um is not None:
        val = val * (J_in_um / lambda0_um)
    return val


def get_energy_spectrum(res, **kw):
    """utils.py:75-93"""
    A = res['Args']
    val = get_full_spectrum(res, **kw)
    if A['mode'] == 'far':
        th = 0.5 * (A['theta'][1:] + A['theta'][:-1])
        v = 0.5 * (val[:, 1:, :] + val[:, :-1, :])
        return A['dph'] * _trapz(v * np.sin(th)[None, :, None], th, axis=1).sum(-1)
    r = A['radius']
    return A['dph'] * _trapz(val * r[None, :, None], r, axis=1).sum(-1)


def get_energy(res, **kw):
    """utils.py:95-102"""
    return _trapz(get_energy_spectrum(res, **kw), res['Args']['omega'])
