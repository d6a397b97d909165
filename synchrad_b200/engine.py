"""Device plumbing of the hot path: torch tensors as device buffers, ctypes into the C ABI.

Replaces the PyOpenCL context/queue/array layer and the per-particle launch loop of the
reference (calc.py:257-267, 292-353, 513-558, 573-603).  PyTorch is used for device memory,
streams and `torch.distributed` only; all arithmetic happens in libsynchrad_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import host

_TORCH = {np.float64: torch.float64, np.float32: torch.float32, np.uint64: torch.int64,
          np.uint32: torch.int32, np.double: torch.float64, np.single: torch.float32}


def _tt(np_dt):
    return _TORCH[np.dtype(np_dt).type]


def require_cuda(device_index):
    if not torch.cuda.is_available():
        raise RuntimeError('synchrad_b200 needs a CUDA device (B200, sm_100a); there is no CPU '
                           'fallback. Use Args["ctx"]=False for an analysis-only object.')
    n = torch.cuda.device_count()
    if not (0 <= device_index < n):
        raise RuntimeError(f'CUDA device {device_index} requested but only {n} visible')
    return torch.device('cuda', device_index)


class PinnedAlloc:
    """alloc(shape, dtype) -> NumPy view of a pinned torch tensor (keeps the tensors)."""

    def __init__(self, pin=True):
        self.pin = pin and torch.cuda.is_available()
        self.tensors = []

    def __call__(self, shape, np_dt):
        t = torch.empty(shape, dtype=_tt(np_dt), pin_memory=self.pin)
        self.tensors.append(t)
        a = t.numpy()
        return a.view(np_dt) if a.dtype != np.dtype(np_dt) else a


class DeviceGrid:
    """The kernel-side tables of one SynchRad object (the `self.Data[...]` of calc.py:486-512)."""

    def __init__(self, Args, dtype, device):
        self.device = device
        self.dtype = dtype
        self.host = host.grid_tables(Args)
        self.dev = {k: torch.from_numpy(v).to(device) for k, v in self.host.items()}
        self.uniform = host.omega_is_uniform(Args)


class Result:
    __slots__ = ('spectra', 'counters', 'info', 'updates', 'elapsed_ms', 'events', '_keep', '_kind_dev')

    @property
    def kind(self):
        """SRB_KIND_* of the kernel that ran.  With phasor='auto' and two eligible kernels the library chooses on the
        device without a host round trip (include/synchrad_b200.h: SRB_KIND_ON_DEVICE); reading it here synchronises."""
        k = int(self.info.kind)
        return int(self._kind_dev.item()) if k < 0 and self._kind_dev is not None else k


def integrate(Args, dtype, grid, packed, comp, nSnaps, native=False, phasor='auto',
              counters=True, device_tracks=None, timing=False, timeStep=None,
              max_scratch_bytes=None, spectra=None, counters_into=None, upload_stream=None):
    """Run the hot path for the packed tracks of this rank.

    Returns Result with `spectra`: list of float64 device tensors (nSnaps, nPhi, nAxis2, nOmega).
    `spectra` (optional): tensors of a previous call to accumulate into (the C ABI is `+=`), which is
    how track sets larger than the device memory are processed batch by batch.
    `device_tracks` (optional) are tracks already resident on the device: a dict with the
    PackedTracks field names holding torch tensors (used by bench.py's device-resident leg).
    `timing`: True -> elapsed_ms is filled (synchronises); 'events' -> only the CUDA events are recorded (`events`), the
    caller reads them after its own synchronisation.  `upload_stream`: stream for the H2D copies of `packed` (the compute
    stream waits for them), so that they overlap the kernel of a previous call.
    """
    lib = _lib.load()
    dev = grid.device
    mode = Args['mode']
    n_out = lib.srb_num_spectra(_lib.MODE[mode], _lib.COMP[comp])
    if n_out < 0:
        # the reference would fail with AttributeError at calc.py:342 (no such near kernel)
        raise AttributeError(f"no {comp!r} kernel in {mode}-field mode")
    n_w, n_2, n_p = (int(v) for v in Args['gridNodeNums'])
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        if device_tracks is None:
            d = {}
            names = ('x', 'y', 'z', 'ux', 'uy', 'uz')
            with torch.cuda.stream(upload_stream if upload_stream is not None else stream):
                for nm, a in zip(names, packed.coords):
                    d[nm] = torch.from_numpy(a.view(a.dtype)).to(dev, non_blocking=True)
                d['offsets'] = torch.from_numpy(packed.offsets.view(np.int64)).to(dev, non_blocking=True)
                d['w'] = torch.from_numpy(packed.w).to(dev, non_blocking=True)
                for nm in ('itStart', 'itEnd', 'itSnaps'):
                    d[nm] = torch.from_numpy(getattr(packed, nm).view(np.int32)).to(dev, non_blocking=True)
            if upload_stream is not None:
                stream.wait_stream(upload_stream)
                for v in d.values():
                    v.record_stream(stream)            # allocated on the upload stream, consumed on the compute stream
            n_tracks, total, stride = packed.n, packed.total, packed.snapStride
        else:
            d = device_tracks
            n_tracks, total, stride = d['n'], d['total'], d['snapStride']
        ff = None
        if comp == 'cartesian_complex' and mode == 'far':
            ff = torch.from_numpy(host.form_factor(Args)).to(dev)

        g = _lib.srb_grid()
        literal = host.float_mode(Args) == 'literal'
        g.mode, g.comp = _lib.MODE[mode], _lib.COMP[comp]
        g.dtype = _lib.DTYPE['double' if dtype is np.double else ('float_literal' if literal else 'float')]
        g.native = 1 if native else 0
        g.phasor = _lib.PHASOR[phasor]
        g.omega_uniform = 1 if grid.uniform else 0
        g.nOmega, g.nAxis2, g.nPhi, g.nSnaps = n_w, n_2, n_p, int(nSnaps)
        T = grid.dev
        g.omega = T['omega'].data_ptr()
        g.sinPhi, g.cosPhi = T['sinPhi'].data_ptr(), T['cosPhi'].data_ptr()
        if mode == 'far':
            g.sinTheta, g.cosTheta = T['sinTheta'].data_ptr(), T['cosTheta'].data_ptr()
        else:
            g.radius = T['radius'].data_ptr()
            g.L_screen = float(np.float32(Args['L_screen'])) if literal else float(Args['L_screen'])
        g.formFactor = ff.data_ptr() if ff is not None else None
        g.dt = float(Args['timeStep']) if (timeStep is None or dtype is np.double or literal) else float(timeStep)
        g.omega_first_host = float(grid.host['omega'][0])
        g.omega_last_host = float(grid.host['omega'][-1])

        t = _lib.srb_tracks()
        t.nTracks = n_tracks
        for nm in ('x', 'y', 'z', 'ux', 'uy', 'uz', 'offsets', 'w', 'itStart', 'itEnd', 'itSnaps'):
            setattr(t, nm, d[nm].data_ptr())
        t.itSnapsStride = stride
        t.totalSteps_host = total

        if spectra is None:
            spectra = [torch.zeros((int(nSnaps), n_p, n_2, n_w), dtype=torch.float64, device=dev)
                       for _ in range(n_out)]
        sp = (ctypes.c_void_p * n_out)(*[s.data_ptr() for s in spectra])
        nbytes = lib.srb_scratch_bytes(ctypes.byref(g), ctypes.byref(t)) if n_tracks else 0
        scratch = None
        if max_scratch_bytes is not None:
            nbytes = min(nbytes, int(max_scratch_bytes))
        if nbytes:
            # scratch = [pre-pass planes][private partial spectra]; a smaller buffer is legal and only
            # reduces parallelism / disables the pre-pass, so back off instead of failing on a full GPU
            for frac in (1.0, 0.25, 0.0):
                try:
                    n_try = int(nbytes * frac)
                    scratch = torch.empty(n_try, dtype=torch.uint8, device=dev) if n_try else None
                    nbytes = n_try
                    break
                except torch.cuda.OutOfMemoryError:
                    torch.cuda.empty_cache()
        cnt = torch.zeros(4, dtype=torch.int64, device=dev) if counters else None
        if timing:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        torch.cuda.nvtx.range_push(f'srb_integrate[{mode},{comp},{n_tracks} tracks]')   # profiler ranges
        rc = lib.srb_integrate(
            ctypes.byref(g), ctypes.byref(t), sp, n_out,
            scratch.data_ptr() if scratch is not None else None, nbytes,
            cnt.data_ptr() if cnt is not None else None, ctypes.c_void_p(stream.cuda_stream))
        torch.cuda.nvtx.range_pop()
        _lib.check(rc)
        res = Result()
        res.elapsed_ms, res.events = None, None
        if timing:
            e1.record(stream)
            res.events = (e0, e1)
            if timing != 'events':
                e1.synchronize()
                res.elapsed_ms = e0.elapsed_time(e1)
        info = _lib.srb_launch_info()
        lib.srb_last_launch(ctypes.byref(info))
        res._kind_dev = cnt[2:3].clone() if cnt is not None else None
        if cnt is not None:
            cnt = cnt[:2]                       # [passed, visited]; slot 2 is the on-device kernel choice (see Result.kind)
        if counters_into is not None and cnt is not None:
            counters_into += cnt
            cnt = counters_into
        res.spectra, res.counters, res.info = spectra, cnt, info
        res.updates = None
        # keep inputs alive until the stream has consumed them
        res._keep = (d, ff, scratch)
        return res


def energy_spectrum(mode, spectra, coherent, nSnaps, n_w, n_2, n_p, iteration, axis2, dphi, layout=1):
    """Angle integrals of utils.py:75-95 on the device (srb_energy_spectrum): returns float64[n_w] on the device.
    `spectra`: float64 device tensors, layout 1 = (nSnaps, n_w, n_2, n_p), 0 = (nSnaps, n_p, n_2, n_w)."""
    lib = _lib.load()
    dev = spectra[0].device
    it = int(iteration) % int(nSnaps)
    ax = torch.as_tensor(np.ascontiguousarray(axis2, dtype=np.float64), device=dev)
    out = torch.empty(n_w, dtype=torch.float64, device=dev)
    ptrs = (ctypes.c_void_p * len(spectra))(*[s.data_ptr() for s in spectra])
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        _lib.check(lib.srb_energy_spectrum(0 if mode == 'far' else 1, int(layout), ptrs, len(spectra), int(bool(coherent)),
                                           n_w, n_2, n_p, int(nSnaps), it, ax.data_ptr(), float(dphi), out.data_ptr(),
                                           ctypes.c_void_p(stream.cuda_stream)))
    return out


def to_host_layout(spectra, nSnaps, n_w, n_2, n_p):
    """Device-side `swapaxes(-1,-3)` (calc.py:573-577): (nSnaps,nPhi,nA2,nOmega) ->
    contiguous (nSnaps,nOmega,nA2,nPhi), still on the device."""
    lib = _lib.load()
    out = []
    for s in spectra:
        dst = torch.empty((int(nSnaps), n_w, n_2, n_p), dtype=torch.float64, device=s.device)
        with torch.cuda.device(s.device):
            stream = torch.cuda.current_stream(s.device)
            _lib.check(lib.srb_swap_axes(s.data_ptr(), dst.data_ptr(), int(nSnaps), n_w, n_2, n_p,
                                         ctypes.c_void_p(stream.cuda_stream)))
        out.append(dst)
    return out
