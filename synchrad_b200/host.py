"""Host-side (NumPy) logic of the spectral-integration path: the reference's `calc_input`
grid/dtype semantics and the packing of particle tracks into the batched SoA layout the C ABI
takes.  Mirrors, behaviourally, these parts of /root/reference/synchrad/calc.py:

    init_args ............. `_init_args`            calc.py:355-451
    grid_tables ........... `_init_data`            calc.py:486-512
    form_factor ........... `_init_raditaion`       calc.py:473-480
    snap_iterations ....... `_set_snap_iterations`  calc.py:626-630
    select_tracks ......... particle split          calc.py:204-212, 230-236
    pack_tracks ........... `_track_to_device` + the per-track scalars of `_process_track`
                            calc.py:579-603, 292-307 — for ALL tracks at once
"""
import numpy as np

COMP_KEYS = {
    'total': ['total'],
    'cartesian': ['x', 'y', 'z'],
    'cartesian_complex': ['xre', 'xim', 'yre', 'yim', 'zre', 'zim'],
    'spheric': ['r', 'theta', 'phi'],
    'spheric_complex': ['rre', 'rim', 'thetare', 'thetaim', 'phire', 'phiim'],
}


def np_dtype(name):
    # the reference accepts 'double' and 'float' (calc.py:364-367); its docstring says "single"
    if name == 'double':
        return np.double
    if name in ('float', 'single'):
        return np.single
    raise ValueError(f"dtype must be 'double' or 'float', got {name!r}")


def init_args(Args):
    """Fill defaults and build the spectral axes in place (returns (Args, numpy dtype))."""
    if 'grid' not in Args:
        raise KeyError("calc_input needs a 'grid' entry")
    Args.setdefault('mode', 'far')
    Args.setdefault('dtype', 'double')
    if Args['mode'] not in ('far', 'near'):
        raise ValueError(f"mode must be 'far' or 'near', got {Args['mode']!r}")
    dtype = np_dtype(Args['dtype'])
    if dtype is np.single:
        if Args['mode'] == 'far':
            print('WARNING: Chosen single precision should be used with care '
                  'for the farfield calculations\n')
        else:
            print('WARNING: Chosen single precision is not recommended '
                  'for the nearfield calculations\n')
    Args.setdefault('ctx', None)
    Args.setdefault('Features', [])

    nodes = Args['grid'][-1]
    Args['gridNodeNums'] = nodes
    Args['numGridNodes'] = int(np.prod(nodes))
    n_w, n_2, n_p = (int(v) for v in nodes)

    w_lo, w_hi = Args['grid'][0]
    omega = np.linspace(w_lo, w_hi, n_w)
    for feature in Args['Features']:          # first matching feature wins
        if feature == 'wavelengthGrid':
            Args['wavelengths'] = np.linspace(1. / w_hi, 1. / w_lo, n_w)
            omega = 1. / Args['wavelengths']
            break
        if feature == 'logGrid':
            step = np.log(w_hi / w_lo) / (n_w - 1.0)
            omega = w_lo * np.exp(step * np.arange(n_w))
            break
    Args['omega'] = omega.astype(dtype)
    Args['dw'] = np.abs(np.diff(omega)) if n_w > 1 else np.array([1.], dtype=dtype)

    lo2, hi2 = Args['grid'][1]
    p_lo, p_hi = Args['grid'][2]
    axis2 = np.linspace(lo2, hi2, n_2)
    phi = p_lo + (p_hi - p_lo) / n_p * np.arange(n_p)      # end point excluded
    far = Args['mode'] == 'far'
    unit = dtype(1.) if far else 1.
    d2 = axis2[1] - axis2[0] if n_2 > 1 else unit
    Args['dph'] = phi[1] - phi[0] if n_p > 1 else unit
    Args['phi'] = phi.astype(dtype)
    if far:
        Args['dth'] = d2
        Args['theta'] = axis2.astype(dtype)
    else:
        Args['dr'] = d2
        Args['radius'] = axis2.astype(dtype)
    Args['dV'] = Args['dw'] * d2 * Args['dph']
    return Args, dtype


def axes64(Args):
    """The spectral axes in float64, whatever Args['dtype'] is.  For dtype='double' these are
    exactly Args['omega'|'theta'|'radius'|'phi']; for 'float' they are the axes BEFORE the
    reference's cast to float32 (recomputed from Args['grid']), because the device side keeps
    everything that enters tau = t - n.r in fp64 (see grid_tables)."""
    if np_dtype(Args['dtype']) is np.double:
        keys = ('omega', 'phi') + (('theta',) if Args['mode'] == 'far' else ('radius',))
        return {k: np.asarray(Args[k], dtype=np.double) for k in keys}
    tmp = {k: Args[k] for k in ('grid', 'mode', 'Features') if k in Args}
    tmp['dtype'] = 'double'
    tmp, _ = init_args(tmp)
    return axes64(tmp)


def omega_is_uniform(Args):
    """True when the omega axis is the default ascending linspace (phasor recurrence allowed)."""
    feats = Args.get('Features', [])
    return not any(f in ('wavelengthGrid', 'logGrid') for f in feats) and \
        int(Args['gridNodeNums'][0]) >= 2 and Args['grid'][0][1] > Args['grid'][0][0]


def float_mode(Args):
    """'mixed' (default) or 'literal' for dtype='float'; always 'double' semantics otherwise.
    Extension key `Args['float_mode']`, see grid_tables and csrc/srb_literal.cuh."""
    if np_dtype(Args['dtype']) is np.double:
        return None
    mode = Args.get('float_mode', 'mixed')
    if mode not in ('mixed', 'literal'):
        raise ValueError(f"float_mode must be 'mixed' or 'literal', got {mode!r}")
    return mode


def literal_tables(Args):
    """float_mode='literal': exactly the float32 arrays `_init_data` uploads (calc.py:494-512), stored
    in float64 buffers (every value is float32-representable)."""
    f = np.float32
    T = {'omega': f(2 * np.pi) * Args['omega'].astype(f),
         'sinPhi': np.sin(Args['phi'].astype(f)), 'cosPhi': np.cos(Args['phi'].astype(f))}
    if Args['mode'] == 'far':
        T['sinTheta'] = np.sin(Args['theta'].astype(f))
        T['cosTheta'] = np.cos(Args['theta'].astype(f))
    else:
        T['radius'] = Args['radius'].astype(f)
    assert all(v.dtype == np.float32 for v in T.values())
    return {k: np.ascontiguousarray(v.astype(np.float64)) for k, v in T.items()}


def grid_tables(Args, dtype=None):
    """The float64 arrays the kernels read: omega pre-multiplied by 2*pi (calc.py:494-495),
    sin/cos of the angular axes (calc.py:498-512).

    dtype='double': bit-identical to what the reference uploads.  dtype='float': the reference
    would upload float32 tables (and float32 tracks, calc.py:585-597); a float32 n.r alone
    perturbs the phase by ~0.4 rad on the undulator test (fp32 spectrum 5.7 % of max
    away from fp64), so this implementation keeps tables, tracks and the per-(direction, step)
    work in fp64 and runs only the per-omega phasor/accumulate arithmetic in fp32."""
    if float_mode(Args) == 'literal':
        return literal_tables(Args)
    ax = axes64(Args)
    T = {'omega': np.ascontiguousarray(np.double(2 * np.pi) * ax['omega']),
         'sinPhi': np.ascontiguousarray(np.sin(ax['phi'])),
         'cosPhi': np.ascontiguousarray(np.cos(ax['phi']))}
    if Args['mode'] == 'far':
        T['sinTheta'] = np.ascontiguousarray(np.sin(ax['theta']))
        T['cosTheta'] = np.ascontiguousarray(np.cos(ax['theta']))
    else:
        T['radius'] = np.ascontiguousarray(ax['radius'])
    return T


def form_factor(Args, dtype=None):
    """Gaussian particle form factor exp(-(2 pi omega sigma)^2 / 2) (calc.py:475-478), float64."""
    if float_mode(Args) == 'literal':      # float32 arithmetic as in the reference
        f = np.float32
        e = f(-0.5) * (f(2 * np.pi) * Args['omega'].astype(f) * f(Args['sigma_particle'])) ** 2
        return np.ascontiguousarray(np.exp(e).astype(np.float64))
    om = axes64(Args)['omega']
    e = -0.5 * (2 * np.pi * om * float(Args['sigma_particle'])) ** 2
    return np.ascontiguousarray(np.exp(e))


def snap_iterations(it_range, nSnaps):
    return np.ascontiguousarray(
        np.linspace(it_range[0], it_range[1], int(nSnaps) + 1, dtype=np.uint32)[1:])


def select_tracks(n_available, Np_max, rank, size, lengths=None, partition='round_robin'):
    """Indices of the tracks this rank integrates, out of the first min(Np_max, N) tracks.

    'round_robin' (default): tracks[rank::size], the reference's split (calc.py:204-212, 230-236).
    'balanced' (extension): contiguous slices with equal work sum(n_p - 1) per rank, for track sets whose
    lengths vary a lot (needs `lengths`, the sample counts of those tracks).  The spectrum is a sum over particles, so
    the partition changes the result only through the summation order; rank-local weight normalisation (Q7) follows
    the partition, as it follows the reference's."""
    n = n_available if Np_max is None else min(int(Np_max), n_available)
    if partition == 'round_robin' or size == 1:
        return np.arange(n)[rank::size]
    if partition != 'balanced':
        raise ValueError(f"partition must be 'round_robin' or 'balanced', got {partition!r}")
    work = np.maximum(np.asarray(lengths[:n], dtype=np.int64) - 1, 0)
    csum = np.concatenate(([0], np.cumsum(work)))
    total = int(csum[-1])
    # slice boundaries at the track starts nearest to the ideal cut points r * total / size
    bounds = [0]
    for r in range(1, size):
        target = (total * r) / size
        i = int(np.searchsorted(csum, target, side='left'))
        if i > 0 and abs(csum[i - 1] - target) <= abs(csum[min(i, n)] - target):
            i -= 1
        bounds.append(min(max(i, bounds[-1]), n))
    bounds.append(n)
    return np.arange(bounds[rank], bounds[rank + 1])


def normalized_weights(weights, mode):
    """weights_normalize of calc.py:241-263, on the RANK-LOCAL list (Q7)."""
    w = np.asarray(weights, dtype=np.double)
    if mode == 'mean':
        return w / np.mean(w) if w.size else w
    if mode == 'max':
        return w / np.max(w) if w.size else w
    if mode == 'ones':
        return np.ones_like(w)
    return w


class PackedTracks:
    """All tracks of a rank in the C-ABI layout (host arrays; `alloc` decides where they live)."""
    __slots__ = ('n', 'coords', 'offsets', 'w', 'itStart', 'itEnd', 'itSnaps', 'snapStride',
                 'total', 'updates_per_node')


def track_length(t):
    """Samples of a track without touching its data when it is a lazy file track (trackio.FileTrack)."""
    n = getattr(t, 'n', None)
    return int(n) if n is not None else int(np.asarray(t[0]).size)


def pack_tracks(tracks, weights, dtype, it_range, nSnaps, alloc=None, lengths=None):
    """Concatenate tracks into SoA arrays of `dtype` (the product always packs float64: see
    grid_tables for why the reference's astype(float32) is not reproduced in 'float' mode).

    tracks : list of [x, y, z, ux, uy, uz, w(, it_start)] (calc.py:110-116)
    weights: per-track weights after normalisation
    it_range None -> per track it_start=0, it_range=(0, n), own snapshot row (calc.py:297-301)
    alloc(shape, dtype) -> writable ndarray (e.g. a view of a pinned torch tensor)
    lengths: the tracks' sample counts when the caller already has them
    """
    if alloc is None:
        alloc = lambda shape, dt: np.empty(shape, dtype=dt)
    n = len(tracks)
    lens = np.fromiter((track_length(t) for t in tracks) if lengths is None else lengths, dtype=np.uint64, count=n)
    P = PackedTracks()
    P.n = n
    P.offsets = alloc((n + 1,), np.uint64)
    P.offsets[0] = 0
    np.cumsum(lens, out=P.offsets[1:])
    P.total = int(P.offsets[n])
    P.coords = [alloc((max(P.total, 1),), dtype) for _ in range(6)]
    P.w = alloc((max(n, 1),), dtype)
    P.itStart = alloc((max(n, 1),), np.uint32)
    P.itEnd = alloc((max(n, 1),), np.uint32)
    nSnaps = int(nSnaps)
    if it_range is None:
        P.snapStride = nSnaps
        P.itSnaps = alloc((max(n, 1), nSnaps), np.uint32)
    else:
        P.snapStride = 0
        P.itSnaps = alloc((nSnaps,), np.uint32)
        P.itSnaps[:] = snap_iterations(it_range, nSnaps)
    in_memory = n > 0 and all(getattr(t, 'read_into', None) is None for t in tracks)
    if in_memory:
        # track lists held in memory: one concatenation per coordinate (astype(dtype) happens inside) instead of
        # 6 slice assignments per track -- the reference converts and uploads array by array (calc.py:579-603)
        for c in range(6):
            cols = [t[c] for t in tracks]
            cols = [a if type(a) is np.ndarray and a.ndim == 1 else np.asarray(a).reshape(-1) for a in cols]
            sizes = np.fromiter((a.size for a in cols), dtype=np.uint64, count=n)
            if not np.array_equal(sizes, lens):
                raise ValueError(f'track {int(np.flatnonzero(sizes != lens)[0])}: coordinate arrays differ in length')
            if P.total:
                np.concatenate(cols, out=P.coords[c][:P.total], casting='unsafe')
    else:
        for i, t in enumerate(tracks):
            o, e = int(P.offsets[i]), int(P.offsets[i + 1])
            direct = getattr(t, 'read_into', None)  # lazy file track: file -> packed buffer, no intermediate array
            for c in range(6):
                if direct is not None:
                    try:
                        direct(c, P.coords[c][o:e])
                    except ValueError as exc:
                        raise ValueError(f'track {i}: coordinate arrays differ in length ({exc})') from None
                    continue
                a = np.asarray(t[c])
                if a.size != e - o:
                    raise ValueError(f'track {i}: coordinate arrays differ in length')
                P.coords[c][o:e] = a        # astype(dtype) happens in the assignment
    if n:
        P.w[:n] = np.asarray(weights, dtype=np.double)[:n]
    m = lens.astype(np.int64)
    if it_range is None:
        if n:
            P.itStart[:n] = 0
            P.itEnd[:n] = m
            for length in np.unique(m):     # one snapshot row per distinct track length (calc.py:297-301, 626-630)
                P.itSnaps[:n][m == length] = snap_iterations((0, int(length)), nSnaps)
        end = m
    else:
        if n:
            starts = np.fromiter((t[7] if len(t) == 8 else 0 for t in tracks), dtype=np.int64, count=n)
            if starts.min() < 0 or starts.max() > 0xFFFFFFFF:
                raise OverflowError(f'track {int(np.flatnonzero((starts < 0) | (starts > 0xFFFFFFFF))[0])}: '
                                    'it_start does not fit uint32 (calc.py:599)')
            P.itStart[:n] = starts
            P.itEnd[:n] = it_range[-1]
        end = np.full(n, int(it_range[-1]), dtype=np.int64)
    upd = int(np.maximum(0, np.minimum(m - 1, end - 1)).sum()) if n else 0
    P.updates_per_node = upd
    return P


def split_batches(lengths, max_steps):
    """Contiguous batches of tracks with at most `max_steps` samples each (a single longer track is
    its own batch).  Returns a list of (first, last_exclusive) index pairs."""
    out, start, acc = [], 0, 0
    for i, n in enumerate(lengths):
        n = int(n)
        if i > start and acc + n > max_steps:
            out.append((start, i))
            start, acc = i, 0
        acc += n
    if start < len(lengths) or not out:
        out.append((start, len(lengths)))
    return out


PIPELINE_BATCH_BYTES = 256 << 20      # host bytes of coordinates (48 per sample) per pipelined batch: upper bound
PIPELINE_MIN_BATCH_BYTES = 16 << 20   # and lower bound (smaller batches cost more kernel efficiency than packing hides)


def pipelined_batches(lengths, device_steps, batch_bytes=None):
    """split_batches for a track list that is still to be packed: sets above 24 MB of coordinates are cut into equal
    batches -- a quarter of the set each, at least 16 MB, at most `batch_bytes` (256 MB) -- so that packing (or reading
    from the tracks file) batch k+1 into pinned memory overlaps the integration of batch k and the pinned buffers are
    recycled instead of growing with the set (the reference's loop, calc.py:257-267, converts, uploads and launches one
    track at a time, serially).  Measured (tools/c3_batches.py, profiles/r02_host_overhead.txt): C4 (92 MB) 1499 ->
    1457 ms in 4 batches, the 600 MB file leg of bench.py 1.39 -> 1.24 s in 4; C3 (12 MB) is slower in 2 (35.9 -> 36.5 ms)
    and stays whole.  `device_steps` bounds a batch by what fits on the device."""
    total = int(sum(int(n) for n in lengths))
    hi = int(PIPELINE_BATCH_BYTES if batch_bytes is None else batch_bytes)
    lo = min(PIPELINE_MIN_BATCH_BYTES, hi)
    per_batch = max(min(hi, max(48 * total // 4, lo)) // 48, 1)
    steps = int(device_steps)
    if total * 2 > per_batch * 3:
        n_b = -(-total // per_batch)
        if -(-total // n_b) * 5 < steps * 4:          # well inside the device bound: balanced shares of whole tracks
            cum = np.cumsum(np.asarray(lengths, dtype=np.int64))
            cuts = np.searchsorted(cum, total * np.arange(1, n_b) / n_b, side='left') + 1
            cuts = np.unique(np.clip(cuts, 1, len(lengths) - 1)) if len(lengths) > 1 else np.zeros(0, dtype=np.int64)
            edges = [0] + [int(c) for c in cuts] + [len(lengths)]
            spans = [(a, b) for a, b in zip(edges, edges[1:]) if b > a]
            bounds = np.concatenate(([0], cum))
            if all(bounds[b] - bounds[a] <= steps for a, b in spans):    # ragged shares can exceed the average
                return spans
        steps = min(steps, -(-total // n_b))
    return split_batches(lengths, max(steps, 1))

