#!/usr/bin/env python
"""Benchmark of the spectral-integration hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype double|float] [--phasor auto|direct|recur|pair|pair_fma]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's kernels compiled for the host cores (oracle/_ref)
    python bench.py --scaling strong --workload c3|c4 [--gpus N]   # a FIXED BASELINE config split over N GPUs

Workload (config.workload): BASELINE.json configs[4] "synthetic PIC-scale tracks 10^5 particles x
10^4 steps, 256x32x32 grid, sharded over 8 GPUs" — each GPU integrates its shard of 12 500
particles x 10^4 steps (SURVEY §8d C5 recipe, generated on the device from a seed), so N = 8 is the
whole of C5 and the scaling is weak (fixed work per GPU, no data-path collective except one NCCL
reduce of the 2 MiB spectrum).  A "step" is one full pass of the hot path over that batch.

metric: updates/s, update = one (particle, time step, spectral node) inner iteration,
updates = sum_p (n_p - 1) * nOmega * nTheta * nPhi (kernel_farfield.cl:59-63).

Legs of the default run, all in the ONE JSON line:
    value ........ fp64, tracks resident in HBM, CUDA events, max over ranks            (headline kernel number)
    e2e .......... the same shard through SynchRad.calculate_spectrum with pinned HOST buffers in, host spectrum out
    parity ....... the CPU arm's sample particles (strict build of the reference's own kernels) integrated on the GPU
                   at the full bench shape, same phasor: max/l2 relative error; the run FAILS above 1e-9
    fp32 ......... mixed-precision (default float mode) and literal (reference-parity float mode) throughput + rooflines
    e2e_api ...... calculate_spectrum(file_tracks=..., file_spectrum=...) (tutorials/PIC/compute_spectrum.py:16-18) on
                   a bounded sample, host-side seconds broken out
    e2e_list ..... calculate_spectrum(particleTracks=<Python list of NumPy tracks>) on the whole shard: the reference's
                   in-memory call signature, packing / upload / kernel pipelined batch by batch
    cpu_baseline . the reference's kernels (g++ -O3, scalar libm, OpenMP) on the host cores, bounded sample
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'particle*step*spectral-pt updates/s'
ALG_SLOTS = {'double': 30.0, 'float': 30.0, 'native': 9.0}   # algorithmic issue slots per update
GRID = (256, 32, 32)
KIND_NAMES = {0: 'direct', 1: 'recurrence', 2: 'literal', 3: 'pair', 4: 'pair, DFMA', 5: 'corrected recurrence'}


def kernel_label(info, fp64, kind=None):
    """k_integrate<...> as reported by srb_last_launch (kind: Result.kind, the on-device choice when phasor='auto'): the
    fp64 pair kernel runs on DMMA when tile width x components % 8 == 0"""
    kind, tw, nc = int(info.kind if kind is None else kind), int(info.tile_width), int(info.n_components)
    name = 'pair, DMMA' if (kind == 3 and fp64 and (tw * nc) % 8 == 0) else KIND_NAMES[kind]
    return 'k_integrate<%s, tile %d>' % (name, tw)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------ CPU arm (oracle/: checker + baseline)
def cpu_arm():
    """('reference', 'ref_fast') when oracle/_ref holds the reference's own kernels compiled for the host
    (oracle/ref_kernels.py, built where /root/reference exists and shipped with the snapshot), else the
    oracle port ('port', 'fast')."""
    from oracle import ref_kernels
    return ('reference', 'ref_fast') if ref_kernels.available('fast') else ('port', 'fast')


CPU_ARM_TEXT = {'reference': "reference kernels (the reference's own kernel_farfield.cl compiled for the host: oracle/_ref, g++ -O3 "
                             "-march=x86-64-v3 -ffp-contract=fast, scalar libm sin/cos, OpenMP over work-items; not pocl), "
                             "one launch per particle as calc.py does",
                'port': 'oracle fast build (restated kernels, g++ -O3 AVX2, scalar libm, OpenMP), one call per particle'}


def cpu_sample_tracks(n_particles, n_steps):
    """The CPU arm's sample of the workload: particles of the same synthetic recipe (seed 4321), generated on the host."""
    from synchrad_b200 import synthetic
    batch = synthetic.c5_batch(n_particles, n_steps, seed=4321, device='cpu')
    return batch, synthetic.batch_to_track_list(batch)


def cpu_run(tracks, lib, threads=None):
    """Runs the reference's kernels on the host cores (oracle/_ref; the oracle's restated kernels when oracle/_ref is
    absent), OpenMP over grid nodes, one particle per launch as the reference launches.  Returns (result, seconds)."""
    from oracle import reference_path as rp
    from synchrad_b200 import synthetic
    rp.build()
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core (libgomp reads the
    # variable when the oracle library is first loaded)
    os.environ['OMP_NUM_THREADS'] = str(threads or os.cpu_count())
    rp.set_threads(threads or os.cpu_count())
    args = synthetic.c5_args(GRID)
    args['ctx'] = False
    t0 = time.perf_counter()
    res = rp.calculate_spectrum(args, tracks, synthetic.C5_DT, lib=lib)
    return res, time.perf_counter() - t0


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path on the host cores: its OpenCL C
    kernels compiled for the host (oracle/_ref, see oracle/ref_kernels.py; there is no OpenCL runtime in the
    image) behind the per-particle launch loop of calc.py as restated in oracle/reference_path.py.  Falls back
    to the oracle port (oracle/oracle_kernels.cpp, -O3 AVX2 OpenMP) only when oracle/_ref is absent."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    cores = os.cpu_count()
    kind, lib = cpu_arm()
    n_steps = a.track_steps
    _, tracks = cpu_sample_tracks(a.ref_particles, n_steps)
    times, upd, rates = [], 0, []
    for i in range(a.warmup + a.steps):
        res, dt = cpu_run(tracks, lib)
        if i >= a.warmup:
            times.append(dt); upd += res['updates']; rates.append(res['updates'] / dt)
    tot = sum(times)
    val = upd / tot
    sample = (f'{a.ref_particles} particle(s) x {n_steps} samples x {GRID[0]}x{GRID[1]}x{GRID[2]} nodes per step '
              f'({a.steps} timed steps = {a.steps * a.ref_particles} particle passes; per-step rate min {min(rates):.4g} / '
              f'max {max(rates):.4g}); ' + CPU_ARM_TEXT[kind])
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'updates/s', 'n_gpus': a.gpus,
        'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': 1e3 * tot / max(a.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': 'C5 shard recipe (synthetic PIC-scale tracks, 256x32x32 grid, far, total), '
                               'bounded sample: ' + sample, 'grid': list(GRID), 'track_steps': n_steps},
        'cpu_baseline': {'value': val, 'unit': 'updates/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': val, 'unit': 'updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ ncu-derived evidence, keyed
def csrc_hash():
    """sha256 over the CUDA sources + the C header: the key of profiles/*_ncu_headline.json."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, 'synchrad_b200', 'csrc')
    for f in sorted(os.listdir(d)):
        if f.endswith(('.cu', '.cuh')):
            h.update(f.encode()); h.update(open(os.path.join(d, f), 'rb').read())
    h.update(open(os.path.join(ROOT, 'include', 'synchrad_b200.h'), 'rb').read())
    return h.hexdigest()[:16]


def ncu_evidence(kernel_name, n_p, n_s):
    """DRAM traffic and pipe-busy fractions of ONE launch of the headline kernel from a tracked ncu capture
    (profiles/r02_ncu_headline.json, written by tools/ncu_headline.py from the .ncu-rep of the same command).  Used only
    when the capture was taken with exactly these sources (csrc hash), this kernel and this shard; else None + why."""
    path = os.path.join(ROOT, 'profiles', 'r02_ncu_headline.json')
    if not os.path.exists(path):
        return None, 'no profiles/r02_ncu_headline.json'
    rec = json.load(open(path))
    want = csrc_hash()
    if rec.get('csrc_sha') != want:
        return None, f"profiles/r02_ncu_headline.json was captured with csrc {rec.get('csrc_sha')}, this build is {want}: not reported"
    if rec.get('kernel') != kernel_name or rec.get('particles') != n_p or rec.get('track_steps') != n_s:
        return None, 'capture is of another kernel / shard: not reported'
    return rec, 'profiles/r02_ncu_headline.json (csrc %s)' % want


# ------------------------------------------------------------------------------------ product arm
def run_product(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    # keep stdout for the ONE JSON line: Python prints go to stderr, and so does anything C libraries
    # write to file descriptor 1 (NCCL prints its version banner there)
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    from synchrad.calc import SynchRad
    from synchrad_b200 import _lib, engine, host, synthetic

    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def pipe_peak(which):
        p = ctypes.c_double()
        _lib.check(lib.srb_pipe_peak(which, ctypes.byref(p)))
        return p.value

    if a.scaling == 'strong':
        return run_strong(a, rank, local, world, dev, json_fd, barrier, max_over_ranks)

    def make_calc(dtype, float_mode=None, phasor='auto'):
        args = synthetic.c5_args(GRID, dtype=dtype)
        args['ctx'] = [0, local]
        args['phasor'] = phasor
        if float_mode:
            args['float_mode'] = float_mode
        c = SynchRad(args)
        c.Args['timeStep'] = c.dtype(synthetic.C5_DT)
        return c

    calc = make_calc(a.dtype, phasor=a.phasor)
    n_p, n_s = a.particles_per_gpu, a.track_steps
    batch = synthetic.c5_batch(n_p, n_s, seed=1234 + rank, device=dev)
    nodes = int(np.prod(GRID))
    updates_rank = n_p * (n_s - 1) * nodes
    nbytes_tracks = 6 * 8 * n_p * n_s

    def sub_batch(b, n):
        """the first n particles of a device/host batch (views, no copies)"""
        if n == b['n']:
            return b
        out = {k: b[k][:n * n_s] for k in ('x', 'y', 'z', 'ux', 'uy', 'uz')}
        out.update(offsets=b['offsets'][:n + 1], w=b['w'][:n], itStart=b['itStart'][:n], itEnd=b['itEnd'][:n],
                   itSnaps=b['itSnaps'][:n], n=n, total=n * n_s, snapStride=b['snapStride'])
        return out

    def integrate(c, tracks, timed, phasor):
        return engine.integrate(c.Args, c.dtype, c._grid, None, 'total', 1, native=False, phasor=phasor,
                                device_tracks=tracks, timing=timed, timeStep=synthetic.C5_DT)

    kernel_ms, launches, info, counters = [], 0, None, None

    def step(timed):
        nonlocal launches, info, counters
        res = integrate(calc, batch, timed, a.phasor)
        if world > 1:
            dist.reduce(res.spectra[0], dst=0, op=dist.ReduceOp.SUM)   # the one exchange of the path
        if timed:
            kernel_ms.append(res.elapsed_ms)
            launches += int(res.info.kernels_launched)
        info, counters = res.info, res.counters
        return res

    # ---------------- device-resident leg: `value`
    for _ in range(a.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(torch.cuda.current_device() if 'CUDA_VISIBLE_DEVICES' not in os.environ else local)
    if rank == 0:
        sampler.start()
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        last = step(True)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / a.steps
    value = world * updates_rank / (ms_step * 1e-3)
    cnt = counters.cpu().numpy()
    guard_pass = float(cnt[0]) / max(float(cnt[1]), 1.0)
    checksum = float(last.spectra[0].sum().item())
    kind_run = last.kind
    del last

    # ---------------- fp32 block (rank 0 of a single-GPU run): both float modes on the same recipe
    fp32 = None
    if world == 1 and not a.no_fp32 and a.dtype == 'double':
        fp32 = {}
        ffma_peak, mufu_peak = pipe_peak(1), pipe_peak(2)
        for mode, frac in (('mixed', 1.0), ('literal', a.fp32_literal_fraction)):
            c32 = make_calc('float', float_mode=mode)
            np32 = max(1, int(round(n_p * frac)))
            b32 = sub_batch(batch, np32)
            integrate(c32, b32, False, 'auto')                       # warm-up
            torch.cuda.synchronize(dev)
            ms = [integrate(c32, b32, True, 'auto').elapsed_ms for _ in range(a.fp32_steps)]
            res32 = integrate(c32, b32, True, 'auto')
            k_ms32 = (sum(ms) + res32.elapsed_ms) / (len(ms) + 1)
            upd32 = np32 * (n_s - 1) * nodes
            rate = upd32 / (k_ms32 * 1e-3)
            fp32[mode] = {
                'value': rate, 'unit': 'updates/s', 'kernel': kernel_label(res32.info, False, res32.kind),
                'what': {'mixed': "dtype='float' default: fp64 tracks/tables/per-step work, fp32 per-omega phasor + accumulate "
                                  '(judged against the fp64 oracle, <= 1e-4)',
                         'literal': "dtype='float', float_mode='literal': every operation of the reference kernels in fp32 "
                                    "(matches the reference's fp32 output <= 1e-4)"}[mode],
                'particles': np32, 'track_steps': n_s, 'ms_per_launch': k_ms32, 'launches_timed': len(ms) + 1,
                'spectrum_checksum': float(res32.spectra[0].sum().item()),
                'roofline': {'bound': 'fp32_pipe', 'algorithmic_slots_per_update': ALG_SLOTS['float'],
                             'achieved': rate * ALG_SLOTS['float'] / 1e12, 'peak': ffma_peak / 1e12,
                             'unit': 'Tslot/s (FFMA-pipe lane issue slots)', 'frac': rate * ALG_SLOTS['float'] / ffma_peak,
                             'peak_source': 'srb_pipe_peak FFMA micro-kernel, this run', 'mufu_peak_Tops': mufu_peak / 1e12},
            }
            del res32, c32, b32
        torch.cuda.empty_cache()

    # ---------------- parity leg: the CPU arm's sample particles on the GPU at the full bench shape
    parity, cpu = None, None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import ref_kernels
        kind, fast_lib = cpu_arm()
        strict_lib = 'ref_strict' if ref_kernels.available('strict') else 'strict'
        cbatch, ctracks = cpu_sample_tracks(a.cpu_particles, n_s)
        ref_fast, dt_fast = cpu_run(ctracks, fast_lib)
        cpu = {'value': ref_fast['updates'] / dt_fast, 'unit': 'updates/s', 'cores': os.cpu_count(), 'kind': kind,
               'sample': f'{a.cpu_particles} particle(s) x {n_s} samples x {GRID[0]}x{GRID[1]}x{GRID[2]} nodes '
                         f"({ref_fast['updates']:.3g} updates, {dt_fast:.1f} s) of the same synthetic recipe; " + CPU_ARM_TEXT[kind]}
        npar = min(a.parity_particles, a.cpu_particles)
        ref, dt_strict = cpu_run(ctracks[:npar], strict_lib)
        # guard decisions: the reference's kernels do not count them; the oracle port (bit-identical to the reference,
        # tests/test_reference_pin.py) does
        port, dt_port = cpu_run(ctracks[:npar], 'strict')
        dbatch = sub_batch({k: (v.to(dev) if hasattr(v, 'to') else v) for k, v in cbatch.items()}, npar)
        resp = integrate(calc, dbatch, False, a.phasor)
        got = engine.to_host_layout(resp.spectra, 1, *GRID)[0].cpu().numpy()
        want = ref['radiation']['total']
        pcnt = resp.counters.cpu().numpy()
        max_rel = float(np.abs(got - want).max() / np.abs(want).max())
        l2_rel = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        tol = 1e-9 if a.dtype == 'double' else 1e-4
        parity = {'max_rel': max_rel, 'l2_rel': l2_rel, 'tolerance': tol, 'ok': bool(max(max_rel, l2_rel) <= tol),
                  'passed_equal': bool(int(pcnt[0]) == int(port['passed'])), 'passed_updates_gpu': int(pcnt[0]),
                  'checker_port_bit_identical_to_reference_kernels': bool(np.array_equal(port['radiation']['total'], want)),
                  'checker': strict_lib + (' (the reference\'s own kernels, g++ -O2 -ffp-contract=off)' if strict_lib == 'ref_strict'
                                           else ' (oracle port, strict build)'),
                  'kernel': kernel_label(resp.info, a.dtype == 'double', resp.kind),
                  'sample': f'{npar} particle(s) x {n_s} samples x {GRID[0]}x{GRID[1]}x{GRID[2]} nodes = the CPU arm\'s first '
                            f'particles, full bench shape, phasor={a.phasor}; checker {dt_strict:.1f} s'}
        del resp, dbatch

    # ---------------- end-to-end leg: host buffers in, host spectrum out, every step
    pk = host.PackedTracks()
    alloc = engine.PinnedAlloc()
    pk.n, pk.total, pk.snapStride = n_p, n_p * n_s, 1
    pk.coords = []
    for k in ('x', 'y', 'z', 'ux', 'uy', 'uz'):
        h = alloc((pk.total,), np.float64)
        h[:] = batch[k].cpu().numpy()
        pk.coords.append(h)
    pk.offsets = alloc((n_p + 1,), np.uint64); pk.offsets[:] = batch['offsets'].cpu().numpy()
    pk.w = alloc((n_p,), np.float64); pk.w[:] = 1.0
    pk.itStart = alloc((n_p,), np.uint32); pk.itStart[:] = 0
    pk.itEnd = alloc((n_p,), np.uint32); pk.itEnd[:] = n_s
    pk.itSnaps = alloc((n_p, 1), np.uint32); pk.itSnaps[:] = n_s
    pk.updates_per_node = n_p * (n_s - 1)
    del batch
    torch.cuda.empty_cache()
    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    calc.calculate_spectrum(pk, timeStep=synthetic.C5_DT, comp='total', verbose=False)   # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        calc.calculate_spectrum(pk, timeStep=synthetic.C5_DT, comp='total', verbose=False)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    e2e_value = world * updates_rank / e2e_s
    h2d, d2h = calc.last_run['h2d_bytes'], calc.last_run['d2h_bytes']
    e2e_checksum = float(calc.Data['radiation']['total'].sum()) if rank == 0 else None
    launches_e2e = calc.last_run['kernels_launched']

    # ---------------- API leg: the reference's named call, tracks file in -> spectrum file out (bounded sample)
    e2e_api = None
    if rank == 0 and world == 1 and a.api_particles > 0:
        try:
            from synchrad_b200 import trackio
            na = min(a.api_particles, n_p)
            tr_list = [[pk.coords[c][i * n_s:(i + 1) * n_s] for c in range(6)] + [1.0, 0] for i in range(na)]
            with tempfile.TemporaryDirectory(dir=a.tmpdir) as tmp:
                ft, fs = os.path.join(tmp, 'tracks.h5'), os.path.join(tmp, 'spectrum.h5')
                tw0 = time.perf_counter()
                trackio.write_tracks(ft, tr_list, cdt=synthetic.C5_DT)
                write_s = time.perf_counter() - tw0
                fsize = os.path.getsize(ft)
                capi = make_calc(a.dtype, phasor=a.phasor)
                capi.calculate_spectrum(file_tracks=ft, file_spectrum=fs, comp='total', verbose=False)   # warm-up
                torch.cuda.synchronize(dev)
                runs = []
                for _ in range(a.api_steps):
                    t0 = time.perf_counter()
                    capi.calculate_spectrum(file_tracks=ft, file_spectrum=fs, comp='total', verbose=False)
                    torch.cuda.synchronize(dev)
                    runs.append((time.perf_counter() - t0, dict(capi.last_run)))
                api_s = sum(r[0] for r in runs) / len(runs)
                lr = runs[-1][1]
                upd_api = na * (n_s - 1) * nodes
                e2e_api = {
                    'value': upd_api / api_s, 'unit': 'updates/s', 's_per_call': api_s, 'calls_timed': len(runs),
                    'path': "SynchRad(calc_input).calculate_spectrum(file_tracks=<HDF5 tracks file>, file_spectrum=<HDF5 out>) "
                            '(tutorials/PIC/compute_spectrum.py:16-18); page cache warm',
                    'particles': na, 'track_steps': n_s, 'tracks_file_bytes': fsize, 'hdf5_backend': trackio.BACKEND,
                    'seconds': {'file_open_headers': lr['file_open_s'], 'host_pack_total': lr['host_pack_s'],
                                'of_which_file_read': lr['file_read_s'], 'gpu_integrate': lr['integrate_ms'] * 1e-3,
                                'call_total': runs[-1][0]},
                    'h2d_bytes_per_call': lr['h2d_bytes'], 'd2h_bytes_per_call': lr['d2h_bytes'],
                    'pipelined_batches': lr['batches'],     # file reading + packing of batch k+1 overlaps the kernel on batch k
                    'spectrum_file_bytes': os.path.getsize(fs), 'tracks_file_write_s_untimed': write_s,
                }
            del tr_list
        except Exception as exc:                                       # an auxiliary leg must not cost the headline line
            e2e_api = {'error': repr(exc)[:400]}

    # ---------------- list leg: the reference's in-memory call signature (a Python list of per-particle NumPy arrays,
    # tests/test_undulator_analytic.py:69) on the WHOLE shard: packing into pinned memory, upload and kernel are pipelined
    # batch by batch inside calculate_spectrum (host.pipelined_batches)
    e2e_list = None
    if rank == 0 and world == 1 and a.list_steps > 0:
        try:
            tr_all = [[pk.coords[c][i * n_s:(i + 1) * n_s] for c in range(6)] + [1.0] for i in range(n_p)]
            clist = make_calc(a.dtype, phasor=a.phasor)
            secs = []
            for _ in range(a.list_steps + 1):                      # the first call is the warm-up
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                clist.calculate_spectrum(tr_all, timeStep=synthetic.C5_DT, comp='total', verbose=False)
                torch.cuda.synchronize(dev)
                secs.append(time.perf_counter() - t0)
            list_s = sum(secs[1:]) / len(secs[1:])
            lr = clist.last_run
            chk = float(clist.Data['radiation']['total'].sum())
            e2e_list = {
                'value': updates_rank / list_s, 'unit': 'updates/s', 's_per_call': list_s, 'calls_timed': len(secs) - 1,
                'path': 'SynchRad(calc_input).calculate_spectrum(particleTracks=<list of [x, y, z, ux, uy, uz, w] NumPy '
                        'arrays>, timeStep=...) -> host float64 spectrum; the whole shard',
                'particles': n_p, 'track_steps': n_s, 'pipelined_batches': lr['batches'],
                'seconds': {'host_pack_total': lr['host_pack_s'], 'gpu_integrate': lr['integrate_ms'] * 1e-3,
                            'call_total': secs[-1]},
                'h2d_bytes_per_call': lr['h2d_bytes'], 'd2h_bytes_per_call': lr['d2h_bytes'],
                'spectrum_checksum': chk,
                'checksum_rel_diff_vs_e2e': abs(chk - e2e_checksum) / abs(e2e_checksum) if e2e_checksum else None,
            }
            del tr_all, clist
        except Exception as exc:                                   # an auxiliary leg must not cost the headline line
            e2e_list = {'error': repr(exc)[:400]}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel
    peak = pipe_peak(0 if a.dtype == 'double' else 1)
    k_ms = sum(kernel_ms) / len(kernel_ms)
    slots_alg = ALG_SLOTS[a.dtype]
    achieved = updates_rank * slots_alg / (k_ms * 1e-3)          # algorithmic slots/s of one launch
    issued = None
    tw, nc = int(info.tile_width), int(info.n_components)
    if kind_run == 1:      # recurrence: per lane and step (TW-2) chain + NC*TW accumulate + 2 seed ops, for TW half-updates
        issued = 2.0 * ((tw - 2) + nc * tw + 2) / tw
    elif kind_run in (3, 4):   # pair: per lane and step 4 (X = Y*Z) + 4*NC*TW/2 accumulate FMAs, for TW updates;
        issued = (4 + 2 * nc * tw) / tw   # fp64 with TW*NC % 8 == 0: the accumulate FMAs are issued as DMMA.8x8x4 (256 each)
    kernel_name = kernel_label(info, a.dtype == 'double', kind_run)
    kname = kernel_name[len('k_integrate<'):].rsplit(',', 1)[0]
    ev, ev_src = ncu_evidence(kernel_name, n_p, n_s)
    roofline = {
        'bound': 'fp64_pipe' if a.dtype == 'double' else 'fp32_pipe',
        'achieved': achieved / 1e12, 'peak': peak / 1e12, 'unit': 'Tslot/s (FMA-pipe lane issue slots)',
        'frac': achieved / peak,
        'peak_source': 'srb_pipe_peak micro-kernel measured in this run on this GPU (no fp64 figure in '
                       'MEASURED_PEAKS.json); nominal 148 SM x 64 x 1.965 GHz = 18.6 Tslot/s fp64',
        'algorithmic_slots_per_update': slots_alg,
        'issued_main_loop_slots_per_update': issued,
        'frac_issued_main_loop': (updates_rank * issued / (k_ms * 1e-3) / peak) if issued else None,
        'kernel_ms_per_launch': k_ms,
        'kernel': kernel_name,
        'traffic': ev['dram_bytes_per_launch'] if ev else None,
        'ncu': ({k: ev[k] for k in ev if k not in ('csrc_sha',)} if ev else None),
        'ncu_source': ev_src,
        'hbm_algorithmic_bytes_per_launch': nbytes_tracks,
        'hbm_gbs_algorithmic': nbytes_tracks / (k_ms * 1e-3) / 1e9,
    }
    # HBM side, for completeness (the path is FP64-pipe bound): against the driver-measured copy bandwidth
    try:
        mp = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        roofline['hbm_peak_gbs'] = mp['hbm_gbs']
        roofline['hbm_peak_source'] = 'MEASURED_PEAKS.json (driver-written copy bandwidth)'
    except Exception:
        roofline['hbm_peak_gbs'] = 7700.0
        roofline['hbm_peak_source'] = 'B200_PROFILING.md fallback (no MEASURED_PEAKS.json)'
    roofline['hbm_frac_algorithmic'] = roofline['hbm_gbs_algorithmic'] / roofline['hbm_peak_gbs']
    if ev:
        roofline['hbm_gbs_traffic'] = ev['dram_bytes_per_launch'] / (k_ms * 1e-3) / 1e9
        roofline['hbm_frac_traffic'] = roofline['hbm_gbs_traffic'] / roofline['hbm_peak_gbs']
    line = {
        'metric': METRIC, 'value': value, 'unit': 'updates/s', 'n_gpus': world, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64' if a.dtype == 'double' else 'f32 (per-omega) / f64 (per-step)',
        'data': 'synthetic',
        'config': {
            'workload': f'C5 (BASELINE configs[4]) per-GPU shard: {n_p} particles x {n_s} samples per GPU, '
                        f'{GRID[0]}x{GRID[1]}x{GRID[2]} (omega,theta,phi) far-field, comp=total; '
                        f'N=8 is the whole 10^5-particle C5',
            'grid': list(GRID), 'particles_per_gpu': n_p, 'track_steps': n_s, 'updates_per_step': world * updates_rank,
            'phasor': kname, 'tile_width': int(info.tile_width),
            'particle_chunks': int(info.n_particle_chunks), 'grid_blocks': int(info.grid_blocks),
            'block_threads': int(info.block_threads),
            'l2_policy': f'inputs larger than L2 ({nbytes_tracks / 1e9:.1f} GB of tracks per GPU vs 126 MB)',
            'guard_pass_fraction': guard_pass, 'spectrum_checksum': checksum, 'csrc_sha': csrc_hash(),
        },
        'e2e': {'value': e2e_value, 'unit': 'updates/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'steps': e2e_steps, 's_per_step': e2e_s, 'spectrum_checksum': e2e_checksum,
                'path': 'SynchRad.calculate_spectrum(pinned host tracks in C-ABI layout) -> host float64 spectrum'},
        'gpu_launches': launches,
        'gpu_launches_e2e_per_step': launches_e2e,
        'clocks': clocks,
        'roofline': roofline,
    }
    if parity is not None:
        line['parity'] = parity
    if fp32 is not None:
        line['fp32'] = fp32
    if e2e_api is not None:
        line['e2e_api'] = e2e_api
    if e2e_list is not None:
        line['e2e_list'] = e2e_list
    if cpu is not None:
        line['cpu_baseline'] = cpu
    os.write(json_fd, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity['ok']:
        sys.stderr.write('bench.py: PARITY FAILURE %r\n' % (parity,))
        sys.exit(3)


def run_strong(a, rank, local, world, dev, json_fd, barrier, max_over_ranks):
    """--scaling strong: a FIXED BASELINE config (C3: betatron ensemble 10^3 electrons, 256x32x32, double, cartesian;
    C4: spiral beam 10^4 particles, 512x64x64, single) split over the N GPUs the way the reference splits over MPI ranks
    (tracks[rank::size], one reduce to rank 0; calc.py:212,236,560-571), through SynchRad(ctx='mpi').  value = the
    config's updates / the slowest rank's wall time of calculate_spectrum (host lists in, reduced host spectrum out)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import cases
    from synchrad.calc import SynchRad
    if a.workload == 'c3':
        tracks, dt, info = cases.betatron_tracks(1000, seed=0)
        args = cases.betatron_args(info, grid=(256, 32, 32))
        kw = dict(comp='cartesian')
        name = 'C3 (BASELINE configs[2]): betatron ensemble 10^3 electrons x 256 samples, 256x32x32, double, cartesian'
    else:
        tracks, dt, info = cases.spiral_tracks(10000, seed=0)
        args = cases.spiral_args(info)
        kw = dict(comp='total')
        name = 'C4 (BASELINE configs[3]): spiral beam 10^4 particles x 192 samples, 512x64x64, single precision, total'
    args['ctx'] = 'mpi' if world > 1 else [0, local]
    calc = SynchRad(dict(args))
    updates = sum(len(t[0]) - 1 for t in tracks) * int(np.prod(args['grid'][-1]))
    times, kms = [], []
    for i in range(a.warmup + a.steps):
        barrier()
        t0 = time.perf_counter()
        calc.calculate_spectrum(tracks, timeStep=dt, verbose=False, **kw)
        torch.cuda.synchronize(dev)
        t = max_over_ranks(time.perf_counter() - t0)
        k = max_over_ranks(calc.last_run['integrate_ms'])
        if i >= a.warmup:
            times.append(t); kms.append(k)
    s = sum(times) / len(times)
    if rank == 0:
        key = 'total' if a.workload == 'c4' else 'x'
        line = {'metric': METRIC, 'value': updates / s, 'unit': 'updates/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
                'ms_per_step': 1e3 * s, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'f64' if a.workload == 'c3' else 'f32 (per-omega) / f64 (per-step)', 'data': 'synthetic',
                'config': {'workload': name + f'; fixed problem split over {world} GPU(s), tracks[rank::size] + one NCCL reduce',
                           'updates_per_step': updates, 'kernel': calc.last_run['kernel'],
                           'guard_pass_fraction': calc.last_run['passed_updates'] / max(calc.last_run['visited_updates'], 1),
                           'spectrum_checksum': float(calc.Data['radiation'][key].sum())},
                'kernel_ms_max_over_ranks': sum(kms) / len(kms),
                'e2e': {'value': updates / s, 'unit': 'updates/s', 'h2d_bytes_per_step': calc.last_run['h2d_bytes'],
                        'd2h_bytes_per_step': calc.last_run['d2h_bytes'],
                        'path': 'SynchRad.calculate_spectrum(list of tracks) incl. Python-side packing, H2D, kernel, reduce, D2H'},
                'gpu_launches': calc.last_run['kernels_launched'] * a.steps}
        os.write(json_fd, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=2)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    p.add_argument('--dtype', default='double', choices=['double', 'float'])
    p.add_argument('--phasor', default='auto', choices=['auto', 'direct', 'recur', 'pair', 'pair_fma'])
    p.add_argument('--scaling', default='weak', choices=['weak', 'strong'])
    p.add_argument('--workload', default='c3', choices=['c3', 'c4'], help='--scaling strong: which fixed BASELINE config')
    p.add_argument('--particles-per-gpu', type=int, default=12500)
    p.add_argument('--track-steps', type=int, default=10000)
    p.add_argument('--e2e-steps', type=int, default=2)
    p.add_argument('--cpu-particles', type=int, default=8, help='CPU baseline sample (reference kernels, fast build)')
    p.add_argument('--parity-particles', type=int, default=2, help='of those, checked on the GPU against the strict build')
    p.add_argument('--ref-particles', type=int, default=4, help='--impl reference: particles per step (about 4 s each on 16 cores)')
    p.add_argument('--fp32-steps', type=int, default=1, help='timed launches per float mode on top of one (besides the warm-up)')
    p.add_argument('--fp32-literal-fraction', type=float, default=0.2,
                   help='share of the shard the literal-fp32 leg integrates (same recipe, linear in particles)')
    p.add_argument('--api-particles', type=int, default=1250, help='tracks in the file of the e2e_api leg (0 = skip)')
    p.add_argument('--api-steps', type=int, default=2)
    p.add_argument('--list-steps', type=int, default=1, help='timed calls of the list-of-tracks leg on the whole shard (0 = skip)')
    p.add_argument('--tmpdir', default=None)
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--no-fp32', action='store_true')
    a = p.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_product(a)


if __name__ == '__main__':
    main()
