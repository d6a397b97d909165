"""`SynchRad` — drop-in for the reference class of the same name (synchrad/calc.py:21-666) on the
spectral-integration path.

Same constructor, same `calculate_spectrum(...)` keyword arguments and defaults
(calc.py:29, :101-107), same result containers (`Data['radiation'][key]` float64
`(nSnaps, nOmega, nTheta|nR, nPhi)`, `total_weight`, `snap_iterations`, `Args[...]`), same
private hooks the reference's tests call (`_init_args`, `_init_data`, `_compile_kernels`,
tests/test_undulator_analytic.py:98-100).  What changed underneath:

    PyOpenCL context / queue ............ a CUDA device + the current torch stream
    Mako templating + JIT ............... ahead-of-time sm_100a kernels (libsynchrad_b200.so)
    one H2D + launch per particle ....... all tracks of the rank packed once, ONE launch
    mpi4py split + Reduce ............... torch.distributed: tracks[rank::size], one NCCL reduce

Documented deviations: string options are compared with `==` (the reference
uses `is`); `dtype='single'` is accepted as an alias of `'float'`; `ctx=None` selects the
current CUDA device instead of prompting on stdin; `native` is honoured only when truthy and
dtype is float (Q9); the cross-particle sum is carried in fp64 (Q5); `dtype='float'` is mixed precision
by default (see host.grid_tables), `Args['float_mode']='literal'` selects the all-fp32 reproduction of
the reference's single-precision kernels.
"""
import os

import numpy as np

from . import host
from .utils import Utilities


def _dist(auto_init=False):
    """torch.distributed, if a process group is up (the MPI.COMM_WORLD of calc.py:86-92).

    The reference picks up MPI implicitly when launched under mpirun; the equivalent here is a
    torchrun launch (RANK / WORLD_SIZE / MASTER_* in the environment): with `ctx='mpi'` and no
    process group yet, one NCCL group is created, one rank per GPU (LOCAL_RANK)."""
    try:
        import torch.distributed as dist
    except ImportError:          # pragma: no cover
        return None
    if not dist.is_available():
        return None
    if not dist.is_initialized() and auto_init and int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch
        local = int(os.environ.get('LOCAL_RANK', '0'))
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        else:                    # host-only use (analysis objects, CPU tests)
            dist.init_process_group('gloo')
    return dist if dist.is_initialized() else None


class SynchRad(Utilities):
    """Spectral-integration calculator; see the module docstring and the reference docstrings
    (calc.py:30-84, :108-167) for the meaning of `Args` and of the keyword arguments."""

    def __init__(self, Args={}, file_spectrum=None):
        dist = _dist(auto_init=isinstance(Args, dict) and Args.get('ctx') == 'mpi')
        if dist is not None:
            self.comm = dist
            self.rank = dist.get_rank()
            self.size = dist.get_world_size()
        else:
            self.comm = None
            self.rank = 0
            self.size = 1
        self.last_run = None
        if file_spectrum is None:
            self._init_args(Args)
            self._init_comm()
            self._init_data()
            self._compile_kernels()
        else:
            self._read_args(file_spectrum)

    # ------------------------------------------------------------------ configuration
    def _init_args(self, Args):
        self.Args, self.dtype = host.init_args(Args)

    def _init_comm(self):
        """Device selection (replaces the OpenCL context choice of calc.py:513-558)."""
        ctx = self.Args['ctx']
        self.device = None
        if ctx is False:
            self.dev_type, self.dev_name, self.plat_name, self.ocl_version = \
                'Starting without', '', 'None', 'None'
        else:
            import torch
            from . import engine
            if ctx is None:
                index = torch.cuda.current_device() if torch.cuda.is_available() else 0
            elif isinstance(ctx, str) and ctx == 'mpi':
                n = max(torch.cuda.device_count(), 1)
                index = int(os.environ.get('LOCAL_RANK', self.rank % n))
            elif isinstance(ctx, (list, tuple)) and len(ctx) == 2:
                index = int(ctx[1])          # [platform, device] -> device index
            elif isinstance(ctx, int):
                index = ctx
            else:
                raise ValueError(f"ctx must be None, False, 'mpi' or [platform, device]; got {ctx!r}")
            self.device = engine.require_cuda(index)
            self.dev_type = 'GPU'
            self.dev_name = torch.cuda.get_device_name(self.device)
            self.plat_name = 'NVIDIA CUDA'
            cc = torch.cuda.get_device_capability(self.device)
            self.ocl_version = f'CUDA {torch.version.cuda}, sm_{cc[0]}{cc[1]}'
        msg = '  {} device: {}'.format(self.dev_type, self.dev_name)
        if self.size > 1:
            msgs = [None] * self.size
            self.comm.all_gather_object(msgs, msg)
        else:
            msgs = [msg]
        if self.rank == 0:
            print('Running on {} devices'.format(self.size))
            for s in msgs:
                print(s)
            print('Platform: {}\nCompiler: {}'.format(self.plat_name, self.ocl_version))

    def _init_data(self):
        self.Data = {}
        self._grid = None
        if self.plat_name == 'None':
            return
        from . import engine
        self._grid = engine.DeviceGrid(self.Args, self.dtype, self.device)
        self.Data.update(self._grid.dev)      # omega (x 2 pi), sin/cos tables, as device tensors

    def _compile_kernels(self):
        """Kernels are compiled ahead of time; this loads the library and fixes the variant the
        reference would have templated (`my_dtype`, `f_native`; calc.py:605-624)."""
        if self.plat_name == 'None':
            return
        from . import _lib
        _lib.load()
        self._native = bool(self.Args.get('native', False)) and self.dtype is np.single \
            and host.float_mode(self.Args) == 'mixed'
        self._phasor = self.Args.get('phasor', 'auto')    # extension: 'auto' | 'direct' | 'recur' | 'pair' | 'pair_fma'
        if self.Args.get('native', False) and self.rank == 0:
            if not self._native:
                print("NOTE: 'native' is honoured for dtype='float' (mixed mode) only (SURVEY Q9); ignored here")
            elif self._phasor != 'direct' and host.omega_is_uniform(self.Args):
                print("NOTE: 'native' selects the MUFU sincos of the per-node (phasor='direct') kernel; on a uniform "
                      "omega grid the planner uses the pair/recurrence kernels, which evaluate no per-node sincos")

    def _set_snap_iterations(self, it_range, nSnaps):
        self.snap_iterations = host.snap_iterations(it_range, nSnaps)

    # ------------------------------------------------------------------ the hot path
    def calculate_spectrum(self, particleTracks=[], file_tracks=None, timeStep=None,
                           comp='total', L_screen=None, Np_max=None, it_range=None, nSnaps=1,
                           sigma_particle=0, weights_normalize=None, file_spectrum=None,
                           verbose=True):
        if self.plat_name == 'None':
            raise RuntimeError('this SynchRad object was created without a device (ctx=False)')
        from . import engine
        import time
        import torch
        track_source = None
        t_call = time.perf_counter()
        t_open = t_pack = 0.0

        if comp not in host.COMP_KEYS:
            raise ValueError(f'unknown comp {comp!r}')
        self.Args['sigma_particle'] = self.dtype(sigma_particle)
        if self.Args['mode'] == 'near':
            if L_screen is not None:
                self.Args['L_screen'] = L_screen
            elif 'L_screen' not in self.Args:
                raise ValueError('Define L_screen argument for near-field calculation')
        nSnaps = int(nSnaps)
        if nSnaps < 1:
            raise ValueError('nSnaps must be >= 1')
        self.Args['comp'] = comp
        if self.Args['mode'] == 'near':
            self.Args['theta'] = np.arctan2(self.Args['radius'], self.Args['L_screen'])
        if timeStep is not None:
            self.Args['timeStep'] = self.dtype(timeStep)
            self._timeStep64 = float(timeStep)
        if it_range is not None:
            it_range = tuple(int(v) for v in it_range)

        if file_tracks is not None:
            from . import trackio
            t0 = time.perf_counter()
            trackio.read_seconds = 0.0
            cdt, file_range, n_file = trackio.read_header(file_tracks)
            self.Args['timeStep'] = self.dtype(cdt)
            self._timeStep64 = float(cdt)
            if it_range is None:
                if file_range is not None:
                    it_range = tuple(int(v) for v in file_range)
                    if self.rank == 0 and verbose:
                        print('it_range from the input file will be used')
                elif self.rank == 0 and verbose:
                    print('Separate it_range for each track will be used')
            partition = self.Args.get('partition', 'round_robin')
            lengths = None
            if partition == 'balanced' and self.size > 1:     # needs every track's length: headers of the first Np tracks
                n_sel = int(n_file) if Np_max is None else min(int(Np_max), int(n_file))
                with trackio.TrackSource(file_tracks, range(n_sel)) as all_tracks:
                    lengths = [t.n for t in all_tracks.tracks]
            index = host.select_tracks(int(n_file), Np_max, self.rank, self.size, lengths, partition)
            # headers only: the samples go from the file straight into the pinned SoA buffers when the
            # batches are packed below (the reference holds every local track in host RAM, calc.py:210-216)
            track_source = trackio.TrackSource(file_tracks, index)
            particleTracks = track_source.tracks
            t_open = time.perf_counter() - t0
            if self.rank == 0 and verbose:
                print('Tracks are loaded')
        elif isinstance(particleTracks, host.PackedTracks):
            pass      # extension: this rank's tracks already in the C-ABI layout (host.pack_tracks)
        else:
            if it_range is None and self.rank == 0 and verbose:
                print('Separate it_range for each track will be used')
            partition = self.Args.get('partition', 'round_robin')
            lengths = [host.track_length(t) for t in particleTracks] if partition == 'balanced' and self.size > 1 else None
            index = host.select_tracks(len(particleTracks), Np_max, self.rank, self.size, lengths, partition)
            particleTracks = [particleTracks[i] for i in index]
        if 'timeStep' not in self.Args:
            raise ValueError('timeStep is required (c*dt in the units of the coordinates)')
        if it_range is not None:
            self._set_snap_iterations(it_range, nSnaps)

        dt64 = getattr(self, '_timeStep64', float(self.Args['timeStep']))
        run = dict(native=self._native, phasor=self._phasor, timing='events', timeStep=dt64)
        if isinstance(particleTracks, host.PackedTracks):
            packed = particleTracks
            if weights_normalize is not None or Np_max is not None:
                raise ValueError('pre-packed tracks: normalise weights / select tracks before packing')
            if (packed.snapStride == 0) != (it_range is not None) or packed.itSnaps.shape[-1] != nSnaps:
                raise ValueError('pre-packed tracks were packed for a different it_range / nSnaps')
            self.total_weight = float(np.sum(packed.w[:packed.n]))
            batches = [packed]
        else:
            weights = host.normalized_weights([t[6] for t in particleTracks], weights_normalize)
            for t, w in zip(particleTracks, weights):       # the reference mutates the track list
                if weights_normalize in ('mean', 'max', 'ones') and isinstance(t, list):
                    t[6] = float(w)
            self.total_weight = float(np.sum(weights)) if len(weights) else 0.0
            # Track sets larger than the device are integrated batch by batch into the same spectra
            # (the reference streams one track at a time, calc.py:257-267).  Per sample: 48 B of
            # coordinates + 48 B of pre-pass planes.
            lengths = [host.track_length(t) for t in particleTracks]
            budget = self.Args.get('max_batch_bytes')
            if budget is None:
                # the device bound only matters for sets of GBs; cudaMemGetInfo costs ~1.3 ms, 6 x the kernel time of a
                # single-electron call, so it is asked only then
                need = 96 * int(sum(lengths))
                free = torch.cuda.mem_get_info(self.device)[0] if need > (1 << 30) else (4 << 30)
                # large sets also go in batches of ~256 MB so that host packing / file reading overlaps the kernel
                spans = host.pipelined_batches(lengths, max(int(0.6 * free) // 96, 1))
            else:
                spans = host.split_batches(lengths, max(int(budget) // 96, 1))
            batches = spans
        res, h2d, upd, ms = None, 0, 0, 0.0
        done_tracks = []
        # Track sets larger than the device: batch k+1 is packed on the host and uploaded on a second stream while batch
        # k is being integrated (srb_integrate never synchronises; the reference's loop at calc.py:257-267 is serial)
        upload = torch.cuda.Stream(self.device) if len(batches) > 1 else None
        timers = []
        bar = None
        if len(batches) > 1 and self.rank == 0 and verbose:
            try:                                 # the reference walks its per-particle loop under tqdm (calc.py:252-253)
                from tqdm import tqdm
                bar = tqdm(total=len(particleTracks))
            except ImportError:
                pass
        try:
            for b in batches:
                if len(timers) >= 2:
                    # the host runs at most two batches ahead of the device: bounds the tracks in flight on the device (and
                    # in pinned memory) for sets larger than the device, and keeps the progress bar honest
                    timers[-2][1].synchronize()
                    if bar is not None:
                        bar.update(done_tracks[len(timers) - 2])
                if isinstance(b, tuple):
                    t0 = time.perf_counter()
                    alloc = engine.PinnedAlloc()
                    packed = host.pack_tracks(particleTracks[b[0]:b[1]], weights[b[0]:b[1]], np.double, it_range,
                                              nSnaps, alloc, lengths=lengths[b[0]:b[1]])
                    t_pack += time.perf_counter() - t0
                res = engine.integrate(self.Args, self.dtype, self._grid, packed, comp, nSnaps,
                                       spectra=None if res is None else res.spectra,
                                       counters_into=None if res is None else res.counters, upload_stream=upload, **run)
                timers.append(res.events)
                done_tracks.append(int(packed.n))
                h2d += int(sum(a.nbytes for a in packed.coords) + packed.offsets.nbytes + packed.w.nbytes
                           + packed.itStart.nbytes + packed.itEnd.nbytes + packed.itSnaps.nbytes)
                upd += int(packed.updates_per_node)
            torch.cuda.synchronize(self.device)
            ms = float(sum(e0.elapsed_time(e1) for e0, e1 in timers))
            if bar is not None:
                bar.update(sum(done_tracks[max(len(timers) - 2, 0):]))
        finally:
            if bar is not None:
                bar.close()
            if track_source is not None:
                track_source.close()
        if it_range is None and packed.n:
            self.snap_iterations = np.array(packed.itSnaps[packed.n - 1])   # last track's, as in the reference
        elif it_range is None:
            self.snap_iterations = np.zeros(nSnaps, dtype=np.uint32)
        n_w, n_2, n_p = (int(v) for v in self.Args['gridNodeNums'])
        dev_out = engine.to_host_layout(res.spectra, nSnaps, n_w, n_2, n_p)
        keys = host.COMP_KEYS[comp]

        cnt = res.counters
        if self.size > 1:                       # replaces _gather_result_mpi (calc.py:560-571)
            from .dist import reduce_to_root
            # the update count travels with the guard counters, so that root reports all three for the whole job
            cnt = torch.cat([cnt, torch.tensor([upd], dtype=cnt.dtype, device=cnt.device)])
            dev_out, self.total_weight, cnt = reduce_to_root(self.comm, dev_out, self.total_weight, cnt)
            upd = int(cnt[2].item())
        self.Data['radiation'] = {k: d.cpu().numpy() for k, d in zip(keys, dev_out)}
        # calc.py:475-480 keeps the form-factor table in Data as well (there a device array; here the host table)
        self.Data['FormFactor'] = host.form_factor(self.Args)
        # the summed result stays on the device of rank 0 for the on-device post-processing (utils.py, on_device=True)
        self._dev_radiation = dict(zip(keys, dev_out)) if self.rank == 0 else None
        c = cnt.cpu().numpy()
        self.last_run = {
            'passed_updates': int(c[0]), 'visited_updates': int(c[1]),
            'updates': upd * int(self.Args['numGridNodes']), 'batches': len(batches),
            'kernel': {0: 'direct', 1: 'recurrence', 2: 'literal', 3: 'pair', 4: 'pair_fma', 5: 'drec'}[res.kind], 'integrate_ms': ms,
            'tile_width': int(res.info.tile_width), 'particle_chunks': int(res.info.n_particle_chunks),
            'time_segments': int(res.info.n_time_segments),
            'grid_blocks': int(res.info.grid_blocks), 'kernels_launched': int(res.info.kernels_launched) + len(keys),
            'h2d_bytes': h2d,
            'd2h_bytes': int(sum(v.nbytes for v in self.Data['radiation'].values())),
            # host-side seconds of this call: opening the tracks file (headers of every track), packing the tracks
            # into the pinned SoA buffers (of which reading the samples out of the file), everything else up to here
            'file_open_s': t_open, 'host_pack_s': t_pack,
            'file_read_s': float(trackio.read_seconds) if file_tracks is not None else 0.0,
            'total_s': time.perf_counter() - t_call,
        }

        if file_spectrum is not None and self.rank == 0:
            from . import trackio
            trackio.write_spectrum(file_spectrum, self)
            print(f'Spectrum is saved to {file_spectrum}')

    # ------------------------------------------------------------------ analysis-only objects
    def _read_args(self, file_spectrum):
        if self.rank == 0:
            from . import trackio
            trackio.read_spectrum(file_spectrum, self)
