#!/usr/bin/env python
"""Benchmark of the spectral-integration hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype double|float] [--phasor auto|direct|recur|pair|pair_fma]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's kernels compiled for the host cores (oracle/_ref)

Workload (config.workload): BASELINE.json configs[4] "synthetic PIC-scale tracks 10^5 particles x
10^4 steps, 256x32x32 grid, sharded over 8 GPUs" — each GPU integrates its shard of 12 500
particles x 10^4 steps (SURVEY §8d C5 recipe, generated on the device from a seed), so N = 8 is the
whole of C5 and the scaling is weak (fixed work per GPU, no data-path collective except one NCCL
reduce of the 2 MiB spectrum).  A "step" is one full pass of the hot path over that batch.

metric: updates/s, update = one (particle, time step, spectral node) inner iteration,
updates = sum_p (n_p - 1) * nOmega * nTheta * nPhi (kernel_farfield.cl:59-63).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'particle*step*spectral-pt updates/s'
ALG_SLOTS = {'double': 30.0, 'float': 30.0}      # algorithmic issue slots per update
GRID = (256, 32, 32)


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


def cpu_arm():
    """('reference', 'ref_fast') when oracle/_ref holds the reference's own kernels compiled for the host
    (oracle/ref_kernels.py, built where /root/reference exists and shipped with the snapshot), else the
    oracle port ('port', 'fast')."""
    from oracle import ref_kernels
    return ('reference', 'ref_fast') if ref_kernels.available('fast') else ('port', 'fast')


CPU_ARM_TEXT = {'reference': "the reference's own kernel_farfield.cl compiled for the host cores (oracle/_ref: g++ -O3 "
                             "-march=x86-64-v3 -ffp-contract=fast, OpenMP over work-items), one launch per particle as calc.py does",
                'port': 'oracle fast build (restated kernels, -O3 AVX2 OpenMP), one call per particle'}


def cpu_port_rate(n_particles, n_steps, threads=None):
    """Times the reference's kernels on the host cores (oracle/_ref; the oracle's fast build of the restated
    kernels when oracle/_ref is absent), OpenMP over grid nodes, one particle per launch as the reference
    launches, on a bounded sample of the workload."""
    from oracle import reference_path as rp
    from synchrad_b200 import synthetic
    rp.build()
    lib = cpu_arm()[1]
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core (libgomp reads the
    # variable when the oracle library is first loaded)
    os.environ['OMP_NUM_THREADS'] = str(threads or os.cpu_count())
    rp.set_threads(threads or os.cpu_count())
    batch = synthetic.c5_batch(n_particles, n_steps, seed=4321, device='cpu')
    tracks = synthetic.batch_to_track_list(batch)
    args = synthetic.c5_args(GRID)
    args['ctx'] = False
    t0 = time.perf_counter()
    res = rp.calculate_spectrum(args, tracks, synthetic.C5_DT, lib=lib)
    dt = time.perf_counter() - t0
    return res['updates'] / dt, res['updates'], dt


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path on the host cores: its OpenCL C
    kernels compiled for the host (oracle/_ref, see oracle/ref_kernels.py; there is no OpenCL runtime in the
    image) behind the per-particle launch loop of calc.py as restated in oracle/reference_path.py.  Falls back
    to the oracle port (oracle/oracle_kernels.cpp, -O3 AVX2 OpenMP) only when oracle/_ref is absent."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    cores = os.cpu_count()
    kind = cpu_arm()[0]
    n_steps = a.track_steps
    times, upd = [], 0
    for i in range(a.warmup + a.steps):
        rate, u, dt = cpu_port_rate(a.ref_particles, n_steps)
        if i >= a.warmup:
            times.append(dt); upd += u
    tot = sum(times)
    val = upd / tot
    sample = (f'{a.ref_particles} particle(s) x {n_steps} samples x {GRID[0]}x{GRID[1]}x{GRID[2]} nodes per step; '
              + CPU_ARM_TEXT[kind])
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
    args = synthetic.c5_args(GRID, dtype=a.dtype)
    args['ctx'] = [0, local]
    args['phasor'] = a.phasor
    calc = SynchRad(args)
    calc.Args['timeStep'] = calc.dtype(synthetic.C5_DT)
    n_p, n_s = a.particles_per_gpu, a.track_steps
    batch = synthetic.c5_batch(n_p, n_s, seed=1234 + rank, device=dev)
    updates_rank = n_p * (n_s - 1) * int(np.prod(GRID))
    nbytes_tracks = 6 * 8 * n_p * n_s

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

    kernel_ms, launches, info, counters = [], 0, None, None

    def step(timed):
        nonlocal launches, info, counters
        res = engine.integrate(calc.Args, calc.dtype, calc._grid, None, 'total', 1, native=False,
                               phasor=a.phasor, device_tracks=batch, timing=timed,
                               timeStep=synthetic.C5_DT)
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

    # ---------------- roofline denominators measured on this device, CPU baseline on this host
    peak = ctypes.c_double()
    which = 0 if a.dtype == 'double' else 1
    _lib.check(lib.srb_pipe_peak(which, ctypes.byref(peak)))
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        rate, upd, dt = cpu_port_rate(a.cpu_particles, n_s)
        cpu = {'value': rate, 'unit': 'updates/s', 'cores': os.cpu_count(), 'kind': cpu_arm()[0],
               'sample': f'{a.cpu_particles} particle(s) x {n_s} samples x {GRID[0]}x{GRID[1]}x{GRID[2]} nodes '
                         f'({upd:.3g} updates, {dt:.1f} s) of the same synthetic recipe; ' + CPU_ARM_TEXT[cpu_arm()[0]]}
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    k_ms = sum(kernel_ms) / len(kernel_ms)
    # DRAM bytes of ONE launch of this exact configuration, from an ncu capture of the same kernel
    # (dram__bytes_read.sum + dram__bytes_write.sum of the full-size launch)
    default_cfg = (a.dtype == 'double' and a.phasor == 'auto' and n_p == 12500 and n_s == 10000)
    traffic = 29002931968 + 649213952 if default_cfg else None   # profiles/r01_ncu_full_size_launch_metrics_dmma.csv
    slots_alg = ALG_SLOTS[a.dtype]
    achieved = updates_rank * slots_alg / (k_ms * 1e-3)          # algorithmic slots/s of one launch
    issued = None
    tw, nc = int(info.tile_width), int(info.n_components)
    if int(info.kind) == 1:      # recurrence: per lane and step (TW-2) chain + NC*TW accumulate + 2 seed ops, for TW half-updates
        issued = 2.0 * ((tw - 2) + nc * tw + 2) / tw
    elif int(info.kind) == 3:    # pair: per lane and step 4 (X = Y*Z) + 4*NC*TW/2 accumulate FMAs, for TW updates;
        issued = (4 + 2 * nc * tw) / tw   # fp64 with TW*NC % 8 == 0: the accumulate FMAs are issued as DMMA.8x8x4 (256 each)
    elif int(info.kind) == 4:    # pair kernel kept on the scalar pipe
        issued = (4 + 2 * nc * tw) / tw
    mma = int(info.kind) == 3 and a.dtype == 'double' and (tw * nc) % 8 == 0
    kname = {0: 'direct', 1: 'recurrence', 2: 'literal', 3: 'pair, DMMA' if mma else 'pair', 4: 'pair, DFMA'}[int(info.kind)]
    roofline = {
        'bound': 'fp64_pipe' if a.dtype == 'double' else 'fp32_pipe',
        'achieved': achieved / 1e12, 'peak': peak.value / 1e12, 'unit': 'Tslot/s (FMA-pipe lane issue slots)',
        'frac': achieved / peak.value,
        'peak_source': 'srb_pipe_peak micro-kernel measured in this run on this GPU (no fp64 figure in '
                       'MEASURED_PEAKS.json); nominal 148 SM x 64 x 1.965 GHz = 18.6 Tslot/s fp64',
        'algorithmic_slots_per_update': slots_alg,
        'issued_main_loop_slots_per_update': issued,
        'frac_issued_main_loop': (updates_rank * issued / (k_ms * 1e-3) / peak.value) if issued else None,
        'kernel_ms_per_launch': k_ms,
        'kernel': 'k_integrate<%s, tile %d>' % (kname, info.tile_width),
        'traffic': traffic,
        # full-size ncu pass (profiles/r01_ncu_full_size_launch_metrics_dmma.csv): 5.12e11 DMMA.8x8x4 warp-instructions
        # x 16 cycles / (592 sub-partitions x 2.324e10 cycles) = 0.595 of the FP64 units' time in DMMA, plus
        # sm__pipe_fp64_cycles_active = 0.185 for the DFMA/DMUL stream (X = Y*Z and the prep phase)
        'fp64_units_busy_ncu': {'dmma': 0.595, 'dfma_pipe': 0.185} if default_cfg else None,
        'hbm_algorithmic_bytes_per_launch': nbytes_tracks,
        'hbm_gbs_algorithmic': nbytes_tracks / (k_ms * 1e-3) / 1e9,
    }
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
            'l2_policy': f'inputs larger than L2 ({nbytes_tracks / 1e9:.1f} GB of tracks per GPU vs 126 MB)',
            'guard_pass_fraction': guard_pass, 'spectrum_checksum': checksum,
        },
        'e2e': {'value': e2e_value, 'unit': 'updates/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'steps': e2e_steps, 's_per_step': e2e_s, 'spectrum_checksum': e2e_checksum,
                'path': 'SynchRad.calculate_spectrum(pinned host tracks in C-ABI layout) -> host float64 spectrum'},
        'gpu_launches': launches,
        'gpu_launches_e2e_per_step': launches_e2e,
        'clocks': clocks,
        'roofline': roofline,
    }
    if cpu is not None:
        line['cpu_baseline'] = cpu
    os.write(json_fd, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=2)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    p.add_argument('--dtype', default='double', choices=['double', 'float'])
    p.add_argument('--phasor', default='auto', choices=['auto', 'direct', 'recur', 'pair', 'pair_fma'])
    p.add_argument('--particles-per-gpu', type=int, default=12500)
    p.add_argument('--track-steps', type=int, default=10000)
    p.add_argument('--e2e-steps', type=int, default=2)
    p.add_argument('--cpu-particles', type=int, default=2)
    p.add_argument('--ref-particles', type=int, default=1)
    p.add_argument('--no-cpu-baseline', action='store_true')
    a = p.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_product(a)


if __name__ == '__main__':
    main()
