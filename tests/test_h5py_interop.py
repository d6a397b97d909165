"""Cross-checks of the bundled HDF5 reader/writer (synchrad_b200/h5lite.py) against libhdf5 through h5py.

h5py is not installed in the build image, so these tests SKIP there (see README, "HDF5"); they run wherever h5py is
importable and are the check the round-1 review asked for: files h5lite writes must open in h5py with the reference's
own access pattern, and files h5py writes in the reference's layouts (default `libver`, i.e. the classic format) must
read back through h5lite -- vlen strings, scalar datasets, groups with more members than one symbol-table node holds."""
import numpy as np
import pytest

h5py = pytest.importorskip('h5py')

from synchrad_b200 import h5lite  # noqa: E402


def _tracks(n_tracks, n=37, seed=0):
    rng = np.random.default_rng(seed)
    return [[rng.normal(size=n + i) for _ in range(6)] + [float(rng.uniform(0.5, 2)), int(i % 5)] for i in range(n_tracks)]


def test_h5py_reads_what_h5lite_writes(tmp_path):
    fn = str(tmp_path / 'lite.h5')
    tracks = _tracks(40)                                   # > 8 members: several symbol-table nodes under one B-tree
    f = h5lite.File(fn, 'w')
    for i, t in enumerate(tracks):
        for c, a in zip(('x', 'y', 'z', 'ux', 'uy', 'uz'), t[:6]):
            f[f'tracks/{i:d}/{c}'] = a
        f[f'tracks/{i:d}/w'] = np.double(t[6])
        f[f'tracks/{i:d}/it_start'] = np.int64(t[7])
    f['misc/cdt'] = np.double(0.25)
    f['misc/N_particles'] = np.int64(len(tracks))
    f['misc/it_range'] = np.array([0, 99], dtype=np.int64)
    f['misc/propagation_direction'] = 'z'
    f['Args/Features'] = np.array([b'wavelengthGrid', b'logGrid'])
    f['Args/empty'] = np.zeros((0,), dtype=np.double)
    f['radiation/total'] = np.arange(2 * 3 * 4 * 5, dtype=np.double).reshape(2, 3, 4, 5)
    f['snap_iterations'] = np.array([5, 10], dtype=np.uint32)
    f.close()
    with h5py.File(fn, 'r') as g:                          # the reference's access pattern (calc.py:186-219, 648-666)
        assert g['misc/cdt'][()] == 0.25 and g['misc/N_particles'][()] == 40
        assert 'it_range' in g['misc'].keys() and list(g['misc/it_range'][()]) == [0, 99]
        assert sorted(g['tracks'].keys(), key=int) == [str(i) for i in range(40)]
        for i, t in enumerate(tracks):
            for c, a in zip(('x', 'y', 'z', 'ux', 'uy', 'uz'), t[:6]):
                np.testing.assert_array_equal(g[f'tracks/{i}/{c}'][()], a)
            assert g[f'tracks/{i}/w'][()] == t[6] and g[f'tracks/{i}/it_start'][()] == t[7]
        val = g['misc/propagation_direction'][()]
        assert (val.decode() if isinstance(val, bytes) else val) == 'z'
        assert [v.decode() for v in g['Args/Features'][()]] == ['wavelengthGrid', 'logGrid']
        assert g['Args/empty'][()].shape == (0,)
        np.testing.assert_array_equal(g['radiation/total'][()], np.arange(120.).reshape(2, 3, 4, 5))
        assert g['snap_iterations'].dtype == np.uint32 and list(g['snap_iterations'][()]) == [5, 10]


def test_h5lite_reads_what_h5py_writes(tmp_path):
    fn = str(tmp_path / 'py.h5')
    tracks = _tracks(40, seed=1)
    with h5py.File(fn, 'w') as g:                          # as converters.py:86-127 and calc.py:274-290 write
        for i, t in enumerate(tracks):
            for c, a in zip(('x', 'y', 'z', 'ux', 'uy', 'uz'), t[:6]):
                g[f'tracks/{i:d}/{c}'] = a
            g[f'tracks/{i:d}/w'] = t[6]
            g[f'tracks/{i:d}/it_start'] = t[7]
        g['misc/cdt'] = 0.25
        g['misc/cdt_array'] = np.full(9, 0.25)
        g['misc/N_particles'] = len(tracks)
        g['misc/it_range'] = np.array([0, 99])
        g['misc/propagation_direction'] = 'z'               # variable-length string
        g['Args/mode'] = 'far'
        g['Args/Features'] = []
        g['Args/gridNodeNums'] = (4, 3, 2)
        g['Args/native'] = False
        g['radiation/total'] = np.arange(24.).reshape(1, 4, 3, 2)
        g['snap_iterations'] = np.array([7], dtype=np.uint32)
        g['total_weight'] = 40.0
    f = h5lite.File(fn, 'r')
    try:
        assert float(f['misc/cdt'][()]) == 0.25 and int(f['misc/N_particles'][()]) == 40
        assert 'it_range' in f['misc'].keys() and [int(v) for v in f['misc/it_range'][()]] == [0, 99]
        assert sorted(f['tracks'].keys(), key=int) == [str(i) for i in range(40)]
        buf = np.empty(tracks[3][0].size)
        f['tracks/3/x'].read_direct(buf)
        np.testing.assert_array_equal(buf, tracks[3][0])
        for i, t in enumerate(tracks):
            assert f[f'tracks/{i}/x'].shape == t[0].shape
            np.testing.assert_array_equal(f[f'tracks/{i}/uz'][()], t[5])
            assert float(f[f'tracks/{i}/w'][()]) == t[6] and int(f[f'tracks/{i}/it_start'][()]) == t[7]
        val = f['misc/propagation_direction'][()]
        assert (val.decode() if isinstance(val, bytes) else val) == 'z'
        mode = f['Args/mode'][()]
        assert (mode.decode() if isinstance(mode, bytes) else mode) == 'far'
        assert np.asarray(f['Args/Features'][()]).size == 0
        assert [int(v) for v in f['Args/gridNodeNums'][()]] == [4, 3, 2]
        assert not bool(f['Args/native'][()])
        np.testing.assert_array_equal(f['radiation/total'][()], np.arange(24.).reshape(1, 4, 3, 2))
        assert float(f['total_weight'][()]) == 40.0
    finally:
        f.close()


def test_product_round_trip_across_backends(tmp_path):
    """trackio (whatever backend it picked) writes; both h5py and h5lite read the same tracks back."""
    from synchrad_b200 import trackio
    fn = str(tmp_path / 't.h5')
    tracks = _tracks(12, seed=2)
    trackio.write_tracks(fn, tracks, cdt=0.5, it_range=True)
    for mod in (h5py, h5lite):
        f = mod.File(fn, 'r')
        try:
            assert float(f['misc/cdt'][()]) == 0.5
            for i, t in enumerate(tracks):
                np.testing.assert_array_equal(np.asarray(f[f'tracks/{i}/y'][()]), t[1])
        finally:
            f.close()
