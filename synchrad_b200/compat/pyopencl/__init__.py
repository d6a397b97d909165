"""`pyopencl` as the reference uses it (calc.py:3-5, 513-558, 605-624, 324-353), backed by libsynchrad_b200.so.
See synchrad_b200/compat/__init__.py."""
import ctypes
import re

import numpy as np

VERSION_TEXT = 'synchrad_b200 compat layer'

_COMP_OF_KERNEL = {'total': 'total', 'cartesian_comps': 'cartesian', 'cartesian_comps_complex': 'cartesian_complex',
                   'spheric_comps': 'spheric', 'spheric_comps_complex': 'spheric_complex'}
_N_SPECTRA = {'total': 1, 'cartesian': 3, 'cartesian_complex': 6, 'spheric': 3, 'spheric_complex': 6}


class device_type:
    GPU = 4

    @staticmethod
    def to_string(value):
        return {4: 'GPU'}.get(value, 'UNKNOWN')


class _Platform:
    vendor = 'NVIDIA CUDA (synchrad_b200)'
    name = 'synchrad_b200'


class Device:
    type = device_type.GPU
    platform = _Platform()
    opencl_c_version = 'ahead-of-time sm_100a kernels (libsynchrad_b200.so)'
    max_work_group_size = 1024

    def __init__(self, index=0):
        self.index = index
        self.name = f'CUDA device {index}'


class Context:
    def __init__(self, index=0):
        self.devices = [Device(index)]


def create_some_context(interactive=None, answers=None):
    """`answers=[platform, device]` (calc.py:516-523) selects the CUDA device; anything else -> device 0."""
    index = int(answers[1]) if answers is not None and len(answers) > 1 else 0
    return Context(index)


class CommandQueue:
    def __init__(self, context, device=None):
        self.context = context
        self.device = device or context.devices[0]

    def finish(self):
        pass


class Buffer:
    """What `Array.data` hands to a kernel call: a host array the library reads / accumulates into."""
    def __init__(self, ndarray):
        self.ndarray = ndarray


def _integrate_host(grid, tracks, spectra_ptrs, n_spectra, device):
    """The one call into the product library; tests substitute the CPU emulation of the kernels here."""
    from synchrad_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.srb_integrate_host(ctypes.byref(grid), ctypes.byref(tracks), spectra_ptrs, n_spectra, None, device))


class _Kernel:
    """One `__kernel` of kernel_farfield.cl / kernel_nearfield.cl as a callable with PyOpenCL's call signature:
    kernel(queue, global_size, local_size, spectra..., x, y, z, ux, uy, uz, wp, itStart, itEnd, nSteps,
           omega, [sinTheta, cosTheta | radius], sinPhi, cosPhi, [L], nOmega, nAxis2, nPhi, dt, nSnaps, itSnaps
           [, FormFactor])                                     (kernel_farfield.cl:6-28, kernel_nearfield.cl:5-27)"""

    def __init__(self, program, name):
        self.program, self.name = program, name
        self.comp = _COMP_OF_KERNEL[name]

    def __call__(self, queue, global_size, local_size, *args):
        from synchrad_b200 import _lib
        P = self.program
        n_sp = _N_SPECTRA[self.comp]
        far = P.mode == 'far'
        # spectra + (6 track arrays, wp, itStart, itEnd, nSteps) + 5 axis arguments (far: omega, sin/cos theta, sin/cos
        # phi; near: omega, radius, sin/cos phi, L) + 3 node counts + (dt, nSnaps, itSnaps) [+ FormFactor]
        n_expected = n_sp + 10 + 5 + 3 + 3 + (1 if self.comp.endswith('complex') else 0)
        if len(args) != n_expected:
            raise TypeError(f'{self.name}: {n_expected} kernel arguments expected, {len(args)} given')
        a = list(args)
        spectra = [b.ndarray for b in a[:n_sp]]
        x, y, z, ux, uy, uz = (np.ascontiguousarray(b.ndarray, dtype=np.float64) for b in a[n_sp:n_sp + 6])
        wp, it_start, it_end, n_steps = a[n_sp + 6:n_sp + 10]
        k = n_sp + 10
        f64 = lambda b: np.ascontiguousarray(b.ndarray, dtype=np.float64)      # noqa: E731
        if far:
            omega, sin_t, cos_t, sin_p, cos_p = (f64(b) for b in a[k:k + 5])
            L = 0.0
            k += 5
        else:
            omega, radius, sin_p, cos_p = (f64(b) for b in a[k:k + 4])
            L = float(a[k + 4])
            k += 5
        n_w, n_2, n_p = (int(v) for v in a[k:k + 3])
        dt, n_snaps, snaps = float(a[k + 3]), int(a[k + 4]), a[k + 5].ndarray
        ff = f64(a[k + 6]) if self.comp.endswith('complex') else None
        if int(n_steps) != x.size:
            raise ValueError(f'{self.name}: nSteps does not match the track length')
        if spectra[0].size != n_snaps * n_w * n_2 * n_p:
            raise ValueError(f'{self.name}: spectrum buffer does not match nSnaps x grid')

        g = _lib.srb_grid()
        g.mode, g.comp = _lib.MODE[P.mode], _lib.COMP[self.comp]
        g.dtype = _lib.DTYPE['double' if P.dtype == 'double' else 'float_literal']
        g.native, g.phasor = 0, _lib.PHASOR['auto']
        g.nOmega, g.nAxis2, g.nPhi, g.nSnaps = n_w, n_2, n_p, n_snaps
        d_om = np.diff(omega)
        g.omega_uniform = int(n_w >= 2 and bool(np.all(d_om > 0)) and
                              float(np.abs(d_om - d_om.mean()).max()) <= 1e-9 * abs(float(d_om.mean())))
        g.omega, g.sinPhi, g.cosPhi = omega.ctypes.data, sin_p.ctypes.data, cos_p.ctypes.data
        if far:
            g.sinTheta, g.cosTheta = sin_t.ctypes.data, cos_t.ctypes.data
        else:
            g.radius, g.L_screen = radius.ctypes.data, L
        if ff is not None:
            g.formFactor = ff.ctypes.data
        g.dt = dt
        g.omega_first_host, g.omega_last_host = float(omega[0]), float(omega[-1])

        offsets = np.array([0, x.size], dtype=np.uint64)
        w = np.array([float(wp)], dtype=np.float64)
        its = np.array([int(it_start)], dtype=np.uint32)
        ite = np.array([int(it_end)], dtype=np.uint32)
        snaps = np.ascontiguousarray(snaps, dtype=np.uint32)
        t = _lib.srb_tracks()
        t.nTracks = 1
        for nm, arr in zip(('x', 'y', 'z', 'ux', 'uy', 'uz'), (x, y, z, ux, uy, uz)):
            setattr(t, nm, arr.ctypes.data)
        t.offsets, t.w, t.itStart, t.itEnd, t.itSnaps = (offsets.ctypes.data, w.ctypes.data, its.ctypes.data,
                                                         ite.ctypes.data, snaps.ctypes.data)
        t.itSnapsStride, t.totalSteps_host = 0, x.size

        # the library accumulates into float64 buffers in the reference's device layout; the reference's own arrays
        # may be float32 (dtype='float'): add in their dtype like `spectrum[...] +=` does (kernel_farfield.cl:102)
        tmp = [np.zeros(s.size, dtype=np.float64) for s in spectra]
        ptrs = (ctypes.c_void_p * n_sp)(*[b.ctypes.data for b in tmp])
        _integrate_host(g, t, ptrs, n_sp, queue.device.index if queue is not None else 0)
        for s, b in zip(spectra, tmp):
            s.reshape(-1)[...] += b.astype(s.dtype)


class _Built:
    def __init__(self, program):
        for name in program.kernel_names:
            setattr(self, name, _Kernel(program, name))


class Program:
    """`cl.Program(ctx, src)`: the source is only inspected for what the reference templated into it (calc.py:605-624):
    which kernel file it is (near-field kernels take `distanceToScreen`) and the compute type."""

    def __init__(self, context, src, mode=None, dtype=None):
        names = re.findall(r'__kernel\s+void\s+(\w+)\s*\(', src or '')
        self.kernel_names = [n for n in names if n in _COMP_OF_KERNEL] or \
            (list(_COMP_OF_KERNEL) if (mode or 'far') == 'far' else list(_COMP_OF_KERNEL)[:3])
        unknown = [n for n in names if n not in _COMP_OF_KERNEL]
        if unknown:
            raise NotImplementedError(f'synchrad_b200 compat: no such kernel in the library: {unknown}')
        self.mode = mode or ('near' if 'distanceToScreen' in (src or '') else 'far')
        if dtype is None:
            dtype = 'float' if re.search(r'__global\s+float\s*\*\s*spectrum', src or '') else 'double'
        self.dtype = dtype
        self.context = context

    def build(self, options=None):
        return _Built(self)
