"""Host-side semantics of calc_input and track packing (calc.py:355-451, 579-630)."""
import numpy as np
import pytest

from synchrad_b200 import host


def test_grid_axes_far_defaults():
    A, dt = host.init_args({'grid': [(1.0, 3.0), (0.0, 0.2), (0.0, 2 * np.pi), (5, 3, 4)]})
    assert dt is np.double and A['mode'] == 'far' and A['dtype'] == 'double' and A['ctx'] is None
    np.testing.assert_allclose(A['omega'], [1, 1.5, 2, 2.5, 3])
    np.testing.assert_allclose(A['theta'], [0, 0.1, 0.2])
    np.testing.assert_allclose(A['phi'], 2 * np.pi / 4 * np.arange(4))     # end point excluded
    np.testing.assert_allclose(A['dw'], 0.5)
    assert A['numGridNodes'] == 60 and tuple(A['gridNodeNums']) == (5, 3, 4)
    np.testing.assert_allclose(A['dV'], 0.5 * 0.1 * (np.pi / 2))
    assert host.omega_is_uniform(A)


def test_grid_features_and_near():
    A, _ = host.init_args({'grid': [(1.0, 4.0), (0.0, 2.0), (0.0, 1.0), (4, 3, 1)], 'mode': 'near',
                           'Features': ['wavelengthGrid']})
    np.testing.assert_allclose(A['omega'], 1 / np.linspace(0.25, 1, 4))      # descending omega
    assert A['dph'] == 1.0 and 'radius' in A and 'theta' not in A and A['dr'] == 1.0
    assert not host.omega_is_uniform(A)
    B, _ = host.init_args({'grid': [(1.0, 8.0), (0.0, 2.0), (0.0, 1.0), (4, 3, 2)], 'Features': ['logGrid']})
    np.testing.assert_allclose(B['omega'], [1, 2, 4, 8])
    assert not host.omega_is_uniform(B)


def test_float_dtype_and_alias(capsys):
    A, dt = host.init_args({'grid': [(1.0, 3.0), (0.0, 0.2), (0.0, 1.0), (5, 3, 4)], 'dtype': 'float'})
    assert dt is np.single and A['omega'].dtype == np.float32 and A['dw'].dtype == np.float64
    assert 'WARNING' in capsys.readouterr().out
    _, dt2 = host.init_args({'grid': [(1.0, 3.0), (0.0, 0.2), (0.0, 1.0), (5, 3, 4)], 'dtype': 'single'})
    assert dt2 is np.single
    with pytest.raises(ValueError):
        host.init_args({'grid': [(1.0, 3.0), (0.0, 0.2), (0.0, 1.0), (5, 3, 4)], 'dtype': 'half'})


def test_tables_are_two_pi_omega():
    A, dt = host.init_args({'grid': [(1.0, 3.0), (0.0, 0.2), (0.0, 1.0), (5, 3, 4)]})
    T = host.grid_tables(A)
    assert all(v.dtype == np.float64 for v in T.values())
    np.testing.assert_array_equal(T['omega'], np.double(2 * np.pi) * A['omega'])    # calc.py:494-495
    np.testing.assert_array_equal(T['sinTheta'], np.sin(A['theta']))
    # 'float': user-visible axes are float32 like the reference's, device tables stay float64
    # and are built from the un-rounded axes (mixed precision, see host.grid_tables)
    F, _ = host.init_args({'grid': [(1.0, 3.0), (0.0, 0.2), (0.0, 1.0), (5, 3, 4)], 'dtype': 'float'})
    assert F['omega'].dtype == np.float32
    np.testing.assert_array_equal(host.grid_tables(F)['omega'], T['omega'])


def test_snap_iterations():
    np.testing.assert_array_equal(host.snap_iterations((0, 10), 1), [10])
    np.testing.assert_array_equal(host.snap_iterations((0, 10), 4), [2, 5, 7, 10])   # floor of 2.5, 7.5
    assert host.snap_iterations((3, 11), 2).dtype == np.uint32


def test_select_tracks_round_robin():
    np.testing.assert_array_equal(host.select_tracks(10, None, 1, 4), [1, 5, 9])
    np.testing.assert_array_equal(host.select_tracks(10, 6, 0, 4), [0, 4])
    np.testing.assert_array_equal(host.select_tracks(3, 100, 3, 4), [])


def test_weights_normalize_is_rank_local():
    np.testing.assert_allclose(host.normalized_weights([1, 3], 'mean'), [0.5, 1.5])
    np.testing.assert_allclose(host.normalized_weights([1, 4], 'max'), [0.25, 1.0])
    np.testing.assert_allclose(host.normalized_weights([1, 4], 'ones'), [1, 1])
    np.testing.assert_allclose(host.normalized_weights([1, 4], None), [1, 4])


def _tracks():
    rs = np.random.RandomState(0)
    return [[rs.rand(n) for _ in range(6)] + [w, s] for n, w, s in ((5, 1.0, 0), (9, 2.5, 3), (1, 0.5, 7))]


def test_pack_tracks_per_track_ranges():
    tr = _tracks()
    P = host.pack_tracks(tr, [t[6] for t in tr], np.float32, None, 2)
    np.testing.assert_array_equal(P.offsets, [0, 5, 14, 15])
    assert P.coords[0].dtype == np.float32 and P.total == 15
    np.testing.assert_array_equal(P.coords[3][5:14], tr[1][3].astype(np.float32))
    np.testing.assert_array_equal(P.itStart[:3], [0, 0, 0])            # it_start forced to 0 (calc.py:299)
    np.testing.assert_array_equal(P.itEnd[:3], [5, 9, 1])
    assert P.snapStride == 2
    np.testing.assert_array_equal(P.itSnaps[1], host.snap_iterations((0, 9), 2))
    assert P.updates_per_node == 4 + 8 + 0


def test_pack_tracks_global_range_and_seven_element_tracks():
    tr = _tracks()
    tr[0] = tr[0][:7]                                                    # it_start defaults to 0
    P = host.pack_tracks(tr, [1, 1, 1], np.float64, (2, 8), 3)
    np.testing.assert_array_equal(P.itStart[:3], [0, 3, 7])
    np.testing.assert_array_equal(P.itEnd[:3], [8, 8, 8])
    assert P.snapStride == 0 and P.itSnaps.shape == (3,)
    assert P.updates_per_node == 4 + 7 + 0


def test_pack_tracks_empty_and_ragged():
    P = host.pack_tracks([], [], np.float64, None, 1)
    assert P.n == 0 and P.total == 0
    bad = _tracks()
    bad[0][2] = np.zeros(3)
    with pytest.raises(ValueError):
        host.pack_tracks(bad, [1, 1, 1], np.float64, None, 1)


def test_split_batches():
    assert host.split_batches([], 100) == [(0, 0)]
    assert host.split_batches([10, 10, 10], 100) == [(0, 3)]
    assert host.split_batches([60, 60, 60], 100) == [(0, 1), (1, 2), (2, 3)]
    assert host.split_batches([30, 30, 30, 30, 500, 5, 5], 100) == [(0, 3), (3, 4), (4, 5), (5, 7)]
    assert host.split_batches([500], 100) == [(0, 1)]


def test_pipelined_batches():
    def ok(spans, n):
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] and a[1] > a[0] for a, b in zip(spans, spans[1:]))
        return spans
    # small sets (below 24 MB of coordinates): one batch; the device bound still applies
    assert host.pipelined_batches([10, 10, 10], 10 ** 9) == [(0, 3)]
    assert host.pipelined_batches([60, 60, 60], 100) == [(0, 1), (1, 2), (2, 3)]
    assert host.pipelined_batches([256] * 1000, 10 ** 9) == [(0, 1000)]                       # C3: 12 MB
    # above: a quarter of the set per batch, balanced over whole tracks ...
    assert ok(host.pipelined_batches([192] * 10000, 10 ** 9), 10000) == [(0, 2500), (2500, 5000), (5000, 7500), (7500, 10000)]
    spans = ok(host.pipelined_batches([10000] * 1250, 10 ** 9), 1250)                         # bench.py's file leg: 600 MB
    assert len(spans) == 4 and {b - a for a, b in spans} <= {312, 313}
    # ... at most 256 MB each
    spans = ok(host.pipelined_batches([10000] * 12500, 10 ** 9), 12500)
    assert len(spans) == 23 and max(b - a for a, b in spans) <= 544
    # explicit batch size (tests): up to 1.5 x stays whole, above it equal shares
    assert host.pipelined_batches([100] * 14, 10 ** 9, batch_bytes=48 * 1000) == [(0, 14)]
    spans = ok(host.pipelined_batches([100] * 64, 10 ** 9, batch_bytes=48 * 1000), 64)
    assert len(spans) == 7 and {b - a for a, b in spans} <= {9, 10}
    # ragged lengths: every track exactly once, shares within one track of each other
    L = list(np.random.default_rng(0).integers(10, 100000, 3000))
    spans = ok(host.pipelined_batches(L, 10 ** 12), len(L))
    sizes = [sum(L[a:b]) for a, b in spans]
    assert max(sizes) - min(sizes) <= 2 * max(L)
    # a single long track cannot be cut
    assert host.pipelined_batches([10 ** 7], 10 ** 12) == [(0, 1)]
    # the device bound wins when it is the smaller one
    assert len(host.pipelined_batches([100] * 64, 300, batch_bytes=48 * 1000)) == 22


def test_balanced_partition_covers_every_track_once():
    """Extension Args['partition']='balanced' (SURVEY §8e): contiguous slices of near-equal work sum(n_p - 1); the
    default stays the reference's tracks[rank::size]."""
    from synchrad_b200 import host
    rs = np.random.RandomState(3)
    for size in (1, 2, 3, 8):
        lengths = rs.randint(2, 5000, size=57).tolist()
        parts = [host.select_tracks(len(lengths), 50, r, size, lengths, 'balanced') for r in range(size)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(50))          # Np_max respected, no overlap
        assert all(np.all(np.diff(p) == 1) for p in parts if len(p) > 1)          # contiguous
        work = [sum(lengths[i] - 1 for i in p) for p in parts]
        rr = [sum(lengths[i] - 1 for i in host.select_tracks(len(lengths), 50, r, size)) for r in range(size)]
        assert max(work) <= max(rr) + max(lengths)            # never much worse than round robin, exact cover above
    np.testing.assert_array_equal(host.select_tracks(10, 7, 1, 3), [1, 4])       # reference split unchanged
    with pytest.raises(ValueError):
        host.select_tracks(5, None, 0, 2, [3] * 5, 'nope')


def _pack_naive(tracks, weights, it_range, nSnaps):
    """Track-at-a-time packing as the reference does it (calc.py:579-603, 292-307): the yardstick for pack_tracks."""
    lens = [int(np.asarray(t[0]).size) for t in tracks]
    coords = [np.concatenate([np.asarray(t[c], dtype=np.double).reshape(-1) for t in tracks]) if tracks else np.zeros(0)
              for c in range(6)]
    if it_range is None:
        starts, ends = [0] * len(tracks), lens
        snaps = [host.snap_iterations((0, m), nSnaps) for m in lens]
    else:
        starts = [t[7] if len(t) == 8 else 0 for t in tracks]
        ends = [it_range[-1]] * len(tracks)
        snaps = host.snap_iterations(it_range, nSnaps)
    upd = sum(max(0, min(m - 1, e - 1)) for m, e in zip(lens, ends))
    return lens, coords, starts, ends, snaps, upd


def test_pack_tracks_equals_track_at_a_time_packing():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(1, 40), min_size=0, max_size=12), st.integers(1, 4), st.booleans(), st.integers(0, 2 ** 31))
    def check(lens, nSnaps, with_range, seed):
        rng = np.random.default_rng(seed)
        kinds = (np.float64, np.float32, np.int64, list)
        tracks = []
        for m in lens:
            cols = []
            for c in range(6):
                a = rng.normal(size=m) * 10
                k = kinds[int(rng.integers(len(kinds)))]
                cols.append([float(v) for v in a] if k is list else a.astype(k))
            t = cols + [float(rng.uniform(0.5, 2))]
            if rng.integers(2):
                t.append(int(rng.integers(0, 9)))
            tracks.append(t)
        it_range = (1, 30) if with_range else None
        w = [t[6] for t in tracks]
        P = host.pack_tracks(tracks, w, np.double, it_range, nSnaps)
        n = len(tracks)
        L, coords, starts, ends, snaps, upd = _pack_naive(tracks, w, it_range, nSnaps)
        assert P.n == n and P.total == sum(L) and list(np.diff(P.offsets[:n + 1])) == L
        for c in range(6):
            assert np.array_equal(P.coords[c][:P.total], coords[c])
        assert list(P.w[:n]) == w and list(P.itStart[:n]) == starts and list(P.itEnd[:n]) == ends
        if it_range is None:
            assert P.snapStride == nSnaps and all(np.array_equal(P.itSnaps[i], snaps[i]) for i in range(n))
        else:
            assert P.snapStride == 0 and np.array_equal(P.itSnaps, snaps)
        assert P.updates_per_node == upd
        # the same tracks, lengths handed in (what calc.py does), and as lazy file-style tracks mixed with lists
        P2 = host.pack_tracks(tracks, w, np.double, it_range, nSnaps, lengths=L)
        assert all(np.array_equal(P2.coords[c][:P.total], coords[c]) for c in range(6))

        class Lazy:
            def __init__(self, t):
                self.t, self.n = t, int(np.asarray(t[0]).size)

            def __len__(self):
                return len(self.t)

            def __getitem__(self, k):
                return self.t[k]

            def read_into(self, c, dest):
                dest[...] = np.asarray(self.t[c], dtype=np.double)
        mixed = [Lazy(t) if i % 2 else t for i, t in enumerate(tracks)]
        P3 = host.pack_tracks(mixed, w, np.double, it_range, nSnaps)
        assert all(np.array_equal(P3.coords[c][:P.total], coords[c]) for c in range(6))
        assert list(P3.itStart[:n]) == starts and P3.updates_per_node == upd

    check()


def test_pipelined_batches_properties():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.lists(st.integers(0, 5000), min_size=0, max_size=60), st.integers(1, 200000), st.integers(48, 48 * 20000))
    def check(lengths, device_steps, batch_bytes):
        spans = host.pipelined_batches(lengths, device_steps, batch_bytes=batch_bytes)
        n = len(lengths)
        assert spans[0][0] == 0 and spans[-1][1] == n                       # every track exactly once, in order
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert all(b > a for a, b in spans) or n == 0
        for a, b in spans:                                                  # the device bound holds unless ONE track breaks it
            assert sum(lengths[a:b]) <= device_steps or b - a == 1

    check()
