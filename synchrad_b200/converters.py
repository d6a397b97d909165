"""Producers of the tracks file the hot path reads (the data format on the input side of the path).

Mirrors the callable surface of the reference's converters.py -- `tracksFromOPMD` (:19-128), `tracksFromVSIM`
(:230-293), `split_track_by_nans` (:295-352) -- and the two module-level helpers of utils.py, `read_tracks`
(:218-240) and `get_Larmor` (:242-266): same names, arguments, file layout and printed messages.  The implementation
is array-at-a-time NumPy (run boundaries of the NaN mask instead of per-sample Python lists) and writes through
trackio's HDF5 backend (h5py when installed, else the bundled h5lite).  `tracksFromOPMD` takes the openPMD-viewer
objects by duck typing (`ts.iterations`, `ts.t`, `ts.iterate`, `ts.get_particle`; `pt.species`, `pt.selected_pid`,
`pt.N_selected`, `pt.__init__`), exactly the members the reference touches, so openPMD-viewer itself is not imported.

Not mirrored: `tracksFromOPMD_old` and its numba helpers `record_particles_step/_first` (converters.py:130-228,
354-393) -- the reference's own version stops at `np.int` (:166-167, removed in NumPy 1.24) and is superseded by
`tracksFromOPMD`; calling it here raises NotImplementedError with that explanation.
"""
import numpy as np

from . import trackio

c = 299792458.0           # scipy.constants.c (exact SI value), converters.py:2

_VARS = ('x', 'y', 'z', 'ux', 'uy', 'uz', 'w')


def split_track_by_nans(x, y, z, ux, uy, uz, w):
    """Cut one particle's time series at the samples where the weight is NaN (the particle is absent from that
    iteration): list of [x, y, z, ux, uy, uz, w_first, it_first] per uninterrupted run (converters.py:295-352)."""
    w = np.asarray(w)
    present = ~np.isnan(w.astype(np.double, copy=False))
    edge = np.diff(np.concatenate(([0], present.view(np.int8), [0])))
    first, last = np.flatnonzero(edge == 1), np.flatnonzero(edge == -1)
    cols = [np.asarray(v) for v in (x, y, z, ux, uy, uz)]
    return [[v[a:b].copy() for v in cols] + [w[a], int(a)] for a, b in zip(first, last)]


class _TrackFileWriter:
    """tracks/<i>/... + misc/... in the converters' layout, shared by the converters below."""

    def __init__(self, fname):
        self.f = trackio._h5.File(fname, 'w')
        self.n, self.it_lo, self.it_hi = 0, np.inf, 0

    def add(self, x, y, z, ux, uy, uz, w, it_start):
        g = f'tracks/{self.n:d}/'
        for name, v in zip(_VARS[:6], (x, y, z, ux, uy, uz)):
            self.f[g + name] = np.ascontiguousarray(v, dtype=np.double)
        self.f[g + 'w'] = np.double(w)
        self.f[g + 'it_start'] = np.int64(it_start)
        self.it_lo = min(self.it_lo, int(it_start))
        self.it_hi = max(self.it_hi, int(it_start) + int(np.size(x)))
        self.n += 1

    def close(self, cdt, it_lo=None, cdt_array=None):
        f = self.f
        f['misc/cdt'] = np.double(cdt)
        if cdt_array is not None:
            f['misc/cdt_array'] = np.asarray(cdt_array, dtype=np.double)
        f['misc/N_particles'] = np.int64(self.n)
        lo = self.it_lo if it_lo is None else it_lo
        # with no track written the reference stores [inf, 0] as float64; keep the dtype rule: ints when finite
        rng = np.array([lo, self.it_hi])
        f['misc/it_range'] = rng.astype(np.int64) if np.all(np.isfinite(rng)) else rng.astype(np.double)
        f['misc/propagation_direction'] = 'z'
        f.close()


def tracks_from_series(fname, series, t, z_is_xi=False, shortest_track=8):
    """The writer half of `tracksFromOPMD` (converters.py:83-128) on plain arrays.

    `series`: dict of the seven variables, each (N_particles, N_iterations) with NaN where a particle is absent;
    `t`: times of the iterations in seconds.  Tracks longer than `shortest_track` samples are written; returns
    their number.  With `z_is_xi` the co-moving coordinate is turned into z by adding c*t of the track's own
    iterations (the reference adds the whole `t` array, :103, which only broadcasts for full-length tracks; for
    those the two agree)."""
    t = np.asarray(t, dtype=np.double)
    out = _TrackFileWriter(fname)
    cols = [np.asarray(series[v]) for v in _VARS]
    for ip in range(cols[0].shape[0]):
        for x, y, z, ux, uy, uz, w, it0 in split_track_by_nans(*[col[ip] for col in cols]):
            if x.size > shortest_track:
                if z_is_xi:
                    z = z + c * t[it0:it0 + z.size]
                out.add(x, y, z, ux, uy, uz, w, it0)
    out.close(cdt=(t[1] - t[0]) * c, cdt_array=(t[1:] - t[:-1]) * c)
    return out.n


def tracksFromOPMD(ts, pt, ref_iteration, fname='./tracks.h5', Np_select=None, dNp=1, sample_selection='random',
                   Nit_min=None, Nit_max=None, z_is_xi=False, shortest_track=8):
    """openPMD time series + ParticleTracker -> tracks file (converters.py:19-128)."""
    all_pid = pt.selected_pid.copy()
    selected_pid = all_pid
    if Np_select is not None:
        if all_pid.size < Np_select:
            Np_select = all_pid.size
            print(f"Selected sample of {Np_select} tracks it too large. ",
                  f"Only {all_pid.size} tracks are available in ParticleTracker")
        if sample_selection == 'random':
            selected_pid = np.random.choice(all_pid, size=Np_select)
        elif sample_selection == 'sequential':
            selected_pid = all_pid[:Np_select]
        else:
            # the reference prints this and then fails on an unbound name (:41-43); fail with the same words
            raise ValueError(f"Selected sampling method '{sample_selection}' is not available.")
    if dNp > 1:
        selected_pid = selected_pid[::dNp]

    pt.__init__(ts, species=pt.species, iteration=ref_iteration, select=selected_pid, preserve_particle_index=True)

    iterations = np.asarray(ts.iterations).copy()
    t = np.asarray(ts.t, dtype=np.double).copy()
    keep = np.ones(iterations.size, dtype=bool)
    if Nit_min is not None:
        keep &= iterations >= Nit_min
    if Nit_max is not None:
        keep &= iterations <= Nit_max
    t = t[keep]            # as in the reference the iteration window sets cdt only; `ts.iterate` walks every iteration

    per_iteration = ts.iterate(ts.get_particle, select=pt, var_list=list(_VARS), species=pt.species)
    n_sel = int(pt.N_selected)
    series = {}
    lengths = [len(v) for v in per_iteration[0]]
    for name, rows in zip(_VARS, per_iteration):
        a = np.full((len(rows), n_sel), np.nan)
        for it, (row, n) in enumerate(zip(rows, lengths)):
            if n == n_sel:                 # iterations with an inconsistent particle count count as "absent" (:72-76)
                a[it] = row
        series[name] = a.T
    if z_is_xi and t.size != series['x'].shape[1]:
        raise ValueError('z_is_xi needs the times of every iteration: do not combine it with Nit_min / Nit_max')
    tracks_from_series(fname, series, t, z_is_xi=z_is_xi, shortest_track=shortest_track)


def tracksFromVSIM(file_vsim, file_synchrad, cdt, length_unit=1, dNit=1, dNp=None, Np_select=None, verbose=True):
    """VSim `tracks` dataset (N_t, N_p, 6) with columns (z, y, x, uz, uy, ux) in SI -> tracks file
    (converters.py:230-293): axes swapped for z-propagation, samples with x <= 0 (out of the box) dropped,
    lengths divided by `length_unit`, velocities by c, unit weights."""
    dt = cdt / length_unit
    src = trackio._h5.File(file_vsim, 'r')
    try:
        data = np.asarray(src['tracks'][()])
    finally:
        src.close()
    ip_indices = np.arange(data.shape[1])
    if dNp is not None:
        ip_indices = ip_indices[::dNp]
    if Np_select is not None:
        ip_indices = ip_indices[:Np_select]
    out = _TrackFileWriter(file_synchrad)
    it_lo = np.inf
    for ip in ip_indices:
        z, y, x, uz, uy, ux = data[::dNit, ip, :].T
        inside = np.flatnonzero(x > 0)
        it_start = inside[0]               # IndexError for a particle that never enters the box, as in the reference
        it_lo = min(it_lo, int(it_start))  # taken before the length cut (:268-269)
        if inside.size > 8:
            # the kept samples are stored back to back even when they are not consecutive iterations (:266)
            out.add(x[inside] / length_unit, y[inside] / length_unit, z[inside] / length_unit,
                    ux[inside] / c, uy[inside] / c, uz[inside] / c, 1.0, it_start)
    out.close(cdt=dt * dNit, it_lo=it_lo)
    if verbose:
        print(f'written {out.n} tracks to {file_synchrad}')


def read_tracks(filename, N_particles=None, dt_step=1):
    """Tracks file -> rectangular arrays (utils.py:218-240): every track cut to the shortest one, every `dt_step`-th
    sample kept.  Returns x, y, z, ux, uy, uz (N_particles, n), w (N_particles,), dt in seconds."""
    f = trackio._h5.File(filename, 'r')
    try:
        dt = f['misc/cdt'][()] * dt_step / c
        if N_particles is None:
            N_particles = int(f['misc/N_particles'][()])
        groups = [f[f'tracks/{ip}'] for ip in range(N_particles)]
        n_min = min(int(np.prod(g['x'].shape)) for g in groups)
        cols = []
        buf = np.empty(n_min)
        for name in _VARS[:6]:
            a = np.empty((N_particles, len(range(0, n_min, dt_step))))
            for ip, g in enumerate(groups):
                if int(np.prod(g[name].shape)) == n_min:
                    g[name].read_direct(buf)
                    a[ip] = buf[::dt_step]
                else:
                    a[ip] = np.asarray(g[name][()])[:n_min:dt_step]
            cols.append(a)
        w = np.ascontiguousarray([g['w'][()] for g in groups])
    finally:
        f.close()
    return (*cols, w, dt)


def get_Larmor(x, y, z, ux, uy, uz, dt):
    """Instantaneous Larmor power in watts along the tracks (utils.py:242-266): the Lienard formula
    P = 2e^2/(3c) gamma^6 (|dbeta/dt|^2 - |beta x dbeta/dt|^2) evaluated in CGS and converted (erg/s -> W)."""
    u = np.stack([np.asarray(ux, dtype=np.double), np.asarray(uy, dtype=np.double), np.asarray(uz, dtype=np.double)])
    gamma = np.sqrt(1.0 + u[0] ** 2 + u[1] ** 2 + u[2] ** 2)
    bx, by, bz = u / gamma
    dbx, dby, dbz = (np.gradient(b, dt, axis=-1) for b in (bx, by, bz))
    e_cgs = 4.8032047e-10
    c_cgs = c * 1e2
    power = 2 * e_cgs ** 2 / 3 / c_cgs * gamma ** 6 * (
        dbx ** 2 + dby ** 2 + dbz ** 2
        - (by * dbz - bz * dby) ** 2
        - (bz * dbx - bx * dbz) ** 2
        - (bx * dby - by * dbx) ** 2)
    power *= 1e-7
    return power


def tracksFromOPMD_old(*args, **kwargs):
    raise NotImplementedError(
        "tracksFromOPMD_old is not mirrored: the reference's own version fails on NumPy >= 1.24 (np.int, "
        'converters.py:166-167) and is superseded by tracksFromOPMD')


def __getattr__(name):
    if name in ('record_particles_step', 'record_particles_first'):
        raise NotImplementedError(f'{name} is a numba helper of tracksFromOPMD_old (converters.py:354-393), which is not '
                                  'mirrored: use tracksFromOPMD')
    raise AttributeError(name)


__all__ = ['tracksFromOPMD', 'tracksFromOPMD_old', 'tracksFromVSIM', 'split_track_by_nans', 'tracks_from_series',
           'read_tracks', 'get_Larmor']
