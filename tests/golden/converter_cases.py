"""Inputs of the converter / helper parity cases, shared by the generator that runs the UNMODIFIED reference
(make_converter_golden.py) and by tests/test_converters.py.  Pure NumPy, deterministic."""
import numpy as np

C = 299792458.0


class FakeTimeSeries:
    """The members of openPMD-viewer's OpenPMDTimeSeries that converters.py:19-128 touches."""

    def __init__(self, data, iterations, t):
        self.data, self.iterations, self.t = data, np.asarray(iterations), np.asarray(t, dtype=np.double)

    def get_particle(self, var_list=None, select=None, species=None, iteration=None, **kw):
        k = int(np.flatnonzero(self.iterations == iteration)[0])
        rows = self.data[k]                                  # dict var -> (N_all,) with NaN = absent; or short list
        if isinstance(rows, str):                            # an iteration whose particle count is inconsistent
            return [np.zeros(1) for _ in var_list]
        idx = select.index
        return [rows[v][idx] for v in var_list]

    def iterate(self, called_method, *args, **kwargs):
        out = [called_method(*args, iteration=it, **kwargs) for it in self.iterations]
        return tuple([r[k] for r in out] for k in range(len(out[0])))


class FakeTracker:
    """ParticleTracker stand-in: `selected_pid`, `species`, `N_selected`, re-initialised with `select=` pids."""

    def __init__(self, ts, species=None, iteration=None, select=None, preserve_particle_index=False):
        self.species = species
        if select is None:
            select = np.arange(ts.n_all)
        self.selected_pid = np.asarray(select)
        self.index = np.asarray(select, dtype=np.int64)
        self.N_selected = self.selected_pid.size
        self.init_kwargs = dict(iteration=iteration, preserve_particle_index=preserve_particle_index)


def opmd_series(n_all=12, n_it=40, seed=3, drop=True):
    """A synthetic run: particles enter / leave (NaN), one iteration with a wrong particle count."""
    rng = np.random.default_rng(seed)
    iterations = 100 + 20 * np.arange(n_it)
    t = 1e-15 * (iterations + 0.25 * np.sin(np.arange(n_it)))        # slightly non-uniform: cdt_array differs from cdt
    base = {v: rng.normal(size=(n_it, n_all)) for v in ('x', 'y', 'z', 'ux', 'uy', 'uz')}
    base['w'] = np.tile(rng.uniform(1, 2, n_all), (n_it, 1))
    if drop:
        for ip in range(n_all):
            a, b = sorted(rng.integers(0, n_it, 2))
            if ip % 3 == 0:
                for v in base:
                    base[v][:a, ip] = np.nan                  # enters late
            elif ip % 3 == 1:
                for v in base:
                    base[v][b:, ip] = np.nan                  # leaves early
            if ip % 4 == 2:
                for v in base:
                    base[v][a:a + 2, ip] = np.nan             # a gap: two pieces
    data = [{v: base[v][k] for v in base} for k in range(n_it)]
    if drop:
        data[22] = 'inconsistent'
    ts = FakeTimeSeries(data, iterations, t)
    ts.n_all = n_all
    return ts


OPMD_CASES = {
    'all': dict(),
    'sequential_5': dict(Np_select=5, sample_selection='sequential'),
    'random_6_seed': dict(Np_select=6, sample_selection='random'),
    'too_many': dict(Np_select=50, sample_selection='sequential'),
    'every_third': dict(dNp=3),
    'from_200': dict(Nit_min=200),      # both bounds at once: IndexError in the reference (converters.py:56-60)
    'to_700': dict(Nit_max=700),
    'short_12': dict(shortest_track=12),
}


def vsim_array(n_t=60, n_p=9, seed=5):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=(n_t, n_p, 6))
    a[:, :, 2] = np.abs(a[:, :, 2]) + 0.1                  # column 2 is x after the axis swap: inside the box
    for ip in range(n_p):
        k = int(rng.integers(0, n_t // 2))
        a[:k, ip, 2] = -1.0                                  # not yet in the box
        if ip == 4:
            a[:n_t - 5, ip, 2] = 0.0                         # a track that is too short
        if ip == 6:
            a[k + 10:k + 13, ip, 2] = -2.0                   # leaves and re-enters: samples are concatenated
    a[:, :, 3:] *= C
    return a


VSIM_CASES = {
    'plain': dict(cdt=0.05, verbose=False),
    'units': dict(cdt=0.05, length_unit=1e-6, dNit=2, verbose=False),
    'subset': dict(cdt=0.1, dNp=2, Np_select=3, verbose=False),
}


def nan_series(seed=11, n=50):
    rng = np.random.default_rng(seed)
    cols = [rng.normal(size=n) for _ in range(6)]
    w = rng.uniform(1, 2, n)
    out = {}
    for name, holes in (('none', []), ('lead', [0, 1]), ('tail', [n - 1]), ('mid', [10, 11, 30]), ('single', [5]),
                        ('all', list(range(n))), ('alternate', list(range(0, n, 2)))):
        ww = w.copy()
        ww[holes] = np.nan
        out[name] = cols + [ww]
    return out


def helper_tracks(seed=2):
    """Ragged tracks for read_tracks / get_Larmor."""
    rng = np.random.default_rng(seed)
    tracks = []
    for n in (64, 50, 71, 50, 90):
        t = np.arange(n) * 0.1
        ux, uy = 2 * np.cos(t + rng.uniform()), np.sin(0.5 * t)
        uz = np.sqrt(100.0 + rng.uniform() - ux ** 2 - uy ** 2)
        g = np.sqrt(1 + ux ** 2 + uy ** 2 + uz ** 2)
        tracks.append([np.cumsum(ux / g) * 0.1, np.cumsum(uy / g) * 0.1, np.cumsum(uz / g) * 0.1, ux, uy, uz,
                       float(rng.uniform(0.5, 2)), int(rng.integers(0, 9))])
    return tracks
