"""HDF5 track / spectrum files of the path, in the reference's layouts.

tracks file (written by the reference's converters, converters.py:102-127; read at calc.py:186-219):
    tracks/<i>/{x,y,z,ux,uy,uz}  float64 [n_i]      tracks/<i>/w, tracks/<i>/it_start  scalars
    misc/cdt  scalar   misc/N_particles  scalar   misc/it_range  int[2] (optional)
    misc/propagation_direction  string (optional, not used by the path)
spectrum file (calc.py:274-290 written, :648-666 read):
    radiation/<key>  float64 (nSnaps, nOmega, nTheta|nR, nPhi)
    Args/<k>  for every Args key except 'grid' and 'ctx'      snap_iterations  uint32[nSnaps]
    total_weight  scalar

h5py is used when it is importable; otherwise the bundled minimal reader/writer (h5lite) handles
exactly these layouts.
"""
import time as _time

import numpy as np

try:                                   # pragma: no cover - not installed in the build image
    import h5py as _h5
    BACKEND = 'h5py'
except ImportError:
    from . import h5lite as _h5
    BACKEND = 'h5lite'

_COMPS = ('x', 'y', 'z', 'ux', 'uy', 'uz')

# seconds spent reading samples out of tracks files (FileTrack.read_into); reset and read by calc.calculate_spectrum
read_seconds = 0.0


def read_header(path):
    """(cdt, it_range or None, N_particles)  — calc.py:187-205"""
    f = _h5.File(path, 'r')
    try:
        cdt = float(f['misc/cdt'][()])
        rng = np.asarray(f['misc/it_range'][()], dtype=np.double) if 'it_range' in f['misc'].keys() else None
        # a converter that wrote no track leaves [inf, 0] (converters.py:86-87,124): no usable range
        rng = tuple(int(v) for v in rng) if rng is not None and np.all(np.isfinite(rng)) else None
        n = int(f['misc/N_particles'][()])
    finally:
        f.close()
    return cdt, rng, n


def read_tracks(path, indices):
    """List of [x,y,z,ux,uy,uz,w,it_start] for the given track indices — calc.py:210-216"""
    f = _h5.File(path, 'r')
    out = []
    try:
        for ip in indices:
            g = f[f'tracks/{int(ip):d}']
            tr = [np.asarray(g[c][()], dtype=np.double) for c in _COMPS]
            tr.append(float(g['w'][()]))
            tr.append(int(g['it_start'][()]) if 'it_start' in g.keys() else 0)
            out.append(tr)
    finally:
        f.close()
    return out


class FileTrack:
    """One track of an open tracks file, read on demand: behaves like the reference's 8-element track list
    (`t[0..5]` coordinate arrays, `t[6]` weight, `t[7]` it_start) but holds no sample data; `read_into(c, dest)`
    fills a slice of the packed (pinned) SoA buffer straight from the file -- SURVEY §8f-3."""
    __slots__ = ('_g', 'n', '_w', 'it_start')

    def __init__(self, group):
        self._g = group
        shape = group['x'].shape
        self.n = int(shape[0]) if len(shape) else 1
        self._w = float(group['w'][()])
        self.it_start = int(group['it_start'][()]) if 'it_start' in group.keys() else 0

    def __len__(self):
        return 8

    def __getitem__(self, k):
        if k < 0:
            k += 8
        if k < 6:
            return np.asarray(self._g[_COMPS[k]][()], dtype=np.double)
        if k == 6:
            return self._w
        if k == 7:
            return self.it_start
        raise IndexError(k)

    def read_into(self, c, dest):
        global read_seconds
        t0 = _time.perf_counter()
        self._g[_COMPS[c]].read_direct(dest)
        read_seconds += _time.perf_counter() - t0


class TrackSource:
    """The tracks `indices` of a tracks file as FileTrack objects; keep it open until they are packed."""

    def __init__(self, path, indices):
        self._f = _h5.File(path, 'r')
        try:
            self.tracks = [FileTrack(self._f[f'tracks/{int(ip):d}']) for ip in indices]
        except Exception:
            self._f.close()
            raise

    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def write_tracks(path, tracks, cdt, it_range=None):
    """Write a tracks file in the converters' layout (converters.py:102-127)."""
    f = _h5.File(path, 'w')
    try:
        lo, hi = None, None
        for i, t in enumerate(tracks):
            for c, a in zip(_COMPS, t[:6]):
                f[f'tracks/{i:d}/{c}'] = np.asarray(a, dtype=np.double)
            f[f'tracks/{i:d}/w'] = np.double(t[6])
            its = int(t[7]) if len(t) == 8 else 0
            f[f'tracks/{i:d}/it_start'] = np.int64(its)
            n = np.asarray(t[0]).size
            lo = its if lo is None else min(lo, its)
            hi = its + n if hi is None else max(hi, its + n)
        f['misc/cdt'] = np.double(cdt)
        f['misc/N_particles'] = np.int64(len(tracks))
        if it_range is True and lo is not None:
            it_range = (lo, hi)
        if it_range not in (None, False, True):
            f['misc/it_range'] = np.asarray(it_range, dtype=np.int64)
        f['misc/propagation_direction'] = 'z'
    finally:
        f.close()


def write_spectrum(path, calc):
    """calc.py:274-290"""
    f = _h5.File(path, 'w')
    try:
        for key, arr in calc.Data['radiation'].items():
            f['radiation/' + key] = np.asarray(arr, dtype=np.double)
        for key, val in calc.Args.items():
            if key in ('grid', 'ctx'):
                continue
            if isinstance(val, (list, tuple)) and len(val) == 0:
                val = np.zeros((0,), dtype=np.double)            # what h5py stores for []
            elif isinstance(val, (list, tuple)) and all(isinstance(v, str) for v in val):
                val = np.array([v.encode() for v in val])        # Features: fixed-length strings
            elif isinstance(val, bool):
                val = np.uint8(val)
            f['Args/' + key] = val
        f['snap_iterations'] = np.asarray(calc.snap_iterations, dtype=np.uint32)
        f['total_weight'] = np.double(calc.total_weight)
    finally:
        f.close()


def read_spectrum(path, calc):
    """calc.py:648-666: fills calc.Data['radiation'], calc.Args, snap_iterations, total_weight."""
    calc.Data = {'radiation': {}}
    calc.Args = {}
    f = _h5.File(path, 'r')
    try:
        for key in f['radiation'].keys():
            calc.Data['radiation'][key] = f['radiation/' + key][()]
        for key in f['Args'].keys():
            val = f['Args/' + key][()]
            if isinstance(val, bytes):
                val = val.decode()
            elif isinstance(val, np.ndarray) and val.dtype.kind in 'SO':
                val = [v.decode() if isinstance(v, bytes) else v for v in val.tolist()]
            calc.Args[key] = val
        calc.snap_iterations = f['snap_iterations'][()]
        calc.total_weight = f['total_weight'][()]
    finally:
        f.close()
    dt = calc.Args.get('dtype', 'double')
    calc.dtype = np.double if dt == 'double' else np.single
