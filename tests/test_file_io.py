"""HDF5 layouts of the path through the bundled minimal HDF5 implementation
(h5py is not available in the image): byte-level format checks, round trips, large groups."""
import struct

import numpy as np
import pytest

from synchrad_b200 import h5lite, trackio


def test_format_landmarks(tmp_path):
    p = str(tmp_path / 'a.h5')
    f = h5lite.File(p, 'w')
    f['g/x'] = np.arange(5, dtype=np.float64)
    f['s'] = 'far'
    f.close()
    raw = open(p, 'rb').read()
    assert raw[:8] == b'\x89HDF\r\n\x1a\n' and raw[8] == 0          # superblock v0
    assert raw[13] == 8 and raw[14] == 8                               # 8-byte offsets / lengths
    eof, = struct.unpack_from('<Q', raw, 40)
    assert eof == len(raw)
    root_ohdr, = struct.unpack_from('<Q', raw, 64)
    assert raw[root_ohdr] == 1                                          # v1 object header
    assert raw.count(b'SNOD') == 2 and raw.count(b'TREE') == 2 and raw.count(b'HEAP') == 2
    assert np.arange(5, dtype='<f8').tobytes() in raw


def test_round_trip_types(tmp_path):
    p = str(tmp_path / 'b.h5')
    vals = {
        'f64': np.linspace(0, 1, 7), 'f32': np.arange(3, dtype=np.float32), 'u32': np.array([1, 2, 3], np.uint32),
        'i64s': np.int64(-5), 'f64s': np.double(2.5), 'empty': np.zeros((0,)), 'cube': np.arange(24.).reshape(2, 3, 4),
        'deep/er/still': np.uint8(7), 'text': 'cartesian_complex', 'strs': np.array([b'logGrid', b'x']),
    }
    f = h5lite.File(p, 'w')
    for k, v in vals.items():
        f[k] = v
    f.close()
    g = h5lite.File(p, 'r')
    assert sorted(g.keys()) == sorted({k.split('/')[0] for k in vals})
    for k, v in vals.items():
        got = g[k][()]
        if isinstance(v, str):
            assert got == v.encode()
        else:
            np.testing.assert_array_equal(got, v)
            assert np.asarray(got).dtype == np.asarray(v).dtype
    assert 'er' in g['deep'] and 'nope' not in g['deep']
    with pytest.raises(KeyError):
        g['missing/x']
    g.close()


def test_large_group_multi_level_btree(tmp_path):
    """2000 tracks -> tracks group with 2000 children: SNOD leaves, two B-tree levels."""
    p = str(tmp_path / 'c.h5')
    rs = np.random.RandomState(0)
    tracks = [[rs.rand(3 + i % 5) for _ in range(6)] + [1.0 + i, i % 7] for i in range(2000)]
    trackio.write_tracks(p, tracks, cdt=0.125, it_range=True)
    cdt, rng, n = trackio.read_header(p)
    assert cdt == 0.125 and n == 2000 and rng == (0, 6 + 7)
    idx = [0, 1, 999, 1234, 1999]
    back = trackio.read_tracks(p, idx)
    for i, t in zip(idx, back):
        for c in range(6):
            np.testing.assert_array_equal(t[c], tracks[i][c])
        assert t[6] == tracks[i][6] and t[7] == tracks[i][7]
    raw = open(p, 'rb').read()
    assert raw.count(b'TREE') > 2000                                   # one per group + internal nodes


def test_reader_handles_continuation_and_compact_and_vlen(tmp_path):
    """Hand-assembled file exercising reader paths the writer never produces: object-header
    continuation block, compact layout, version-2 dataspace and a variable-length string in a global
    heap (what h5py emits for `f['Args/mode'] = 'far'`)."""
    from synchrad_b200.h5lite import _Writer, _msg, _pad8, _space_msg, _dtype_msg, UNDEF
    p = str(tmp_path / 'd.h5')
    w = _Writer(p)
    w.pos = 96
    # global heap with one object "near"
    gcol = bytearray(b'GCOL' + struct.pack('<B3xQ', 1, 4096))
    gcol += struct.pack('<HHIQ', 1, 1, 0, 4) + _pad8(b'near')
    gcol += struct.pack('<HHIQ', 0, 0, 0, 4096 - len(gcol) - 16)
    gcol += b'\x00' * (4096 - len(gcol))
    ga = w.alloc(4096); w.put(ga, bytes(gcol))
    # dataset 1: vlen string scalar, compact layout, header split by a continuation message
    vl_type = struct.pack('<BBBBI', 0x19, 0x01, 0x01, 0, 16) + struct.pack('<BBBBI', 0x13, 0x10, 0, 0, 1)
    rec = struct.pack('<IQI', 4, ga, 1)
    part2 = _msg(0x0008, struct.pack('<BBH', 3, 0, len(rec)) + rec)
    ca = w.alloc(len(part2)); w.put(ca, part2)
    part1 = _msg(0x0001, struct.pack('<BBBB', 2, 0, 0, 0)) + _msg(0x0003, vl_type) + _msg(0x0010, struct.pack('<QQ', ca, len(part2)))
    hdr = struct.pack('<BBHII4x', 1, 0, 4, 1, len(part1)) + part1
    d1 = w.alloc(len(hdr)); w.put(d1, hdr)
    # dataset 2: compact float array
    data = np.array([1.5, -2.0, 3.25]).tobytes()
    m = _msg(0x0001, _space_msg((3,))) + _msg(0x0003, _dtype_msg(np.float64)) + _msg(0x0008, struct.pack('<BBH', 3, 0, len(data)) + data)
    hdr = struct.pack('<BBHII4x', 1, 0, 3, 1, len(m)) + m
    d2 = w.alloc(len(hdr)); w.put(d2, hdr)
    root, btree, heap = w._write_group({'mode': d1, 'vals': d2})
    sb = b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack('<QQQQ', 0, UNDEF, w.pos, UNDEF) + struct.pack('<QQII', 0, root, 1, 0) + struct.pack('<QQ', btree, heap)
    w.put(0, sb); w.f.truncate(w.pos); w.f.close()
    f = h5lite.File(p, 'r')
    assert f['mode'][()] == b'near'
    np.testing.assert_array_equal(f['vals'][()], [1.5, -2.0, 3.25])
    f.close()


def test_unsupported_files_fail_loudly(tmp_path):
    p = str(tmp_path / 'e.h5')
    open(p, 'wb').write(b'not hdf5 at all' * 100)
    with pytest.raises(IOError):
        h5lite.File(p, 'r')
    sb = bytearray(b'\x89HDF\r\n\x1a\n' + b'\x02' + b'\x00' * 87)       # superblock v2 (libver='latest')
    open(p, 'wb').write(bytes(sb))
    with pytest.raises(NotImplementedError):
        h5lite.File(p, 'r')


class _FakeCalc:
    pass


def test_spectrum_file_layout_round_trip(tmp_path):
    """radiation/<key>, Args/<k> (all keys but grid/ctx), snap_iterations, total_weight — calc.py:274-290."""
    from synchrad_b200 import host
    A, dt = host.init_args({'grid': [(1.0, 3.0), (0.0, 0.2), (0.0, 2 * np.pi), (5, 3, 4)], 'ctx': [0, 0],
                            'Features': ['logGrid'], 'native': True})
    A.update(sigma_particle=dt(0.0), timeStep=dt(0.01), comp='cartesian')
    c = _FakeCalc()
    c.Args, c.dtype = A, dt
    c.Data = {'radiation': {k: np.random.RandomState(i).rand(2, 5, 3, 4) for i, k in enumerate('xyz')}}
    c.snap_iterations = np.array([5, 10], dtype=np.uint32)
    c.total_weight = 7.5
    p = str(tmp_path / 'spectrum.h5')
    trackio.write_spectrum(p, c)
    f = h5lite.File(p, 'r')
    assert sorted(f.keys()) == ['Args', 'radiation', 'snap_iterations', 'total_weight']
    assert 'grid' not in f['Args'].keys() and 'ctx' not in f['Args'].keys()
    for k in ('mode', 'dtype', 'gridNodeNums', 'numGridNodes', 'Features', 'omega', 'dw', 'dth', 'dph', 'theta',
              'phi', 'dV', 'sigma_particle', 'timeStep', 'comp', 'native'):
        assert k in f['Args'].keys(), k
    assert f['radiation/x'][()].shape == (2, 5, 3, 4) and f['radiation/x'][()].dtype == np.float64
    f.close()
    d = _FakeCalc()
    trackio.read_spectrum(p, d)
    assert d.Args['mode'] == 'far' and d.Args['comp'] == 'cartesian' and d.Args['Features'] == ['logGrid']
    np.testing.assert_array_equal(d.Args['omega'], A['omega'])
    np.testing.assert_array_equal(d.snap_iterations, c.snap_iterations)
    assert d.total_weight == 7.5 and d.dtype is np.double
    for k in 'xyz':
        np.testing.assert_array_equal(d.Data['radiation'][k], c.Data['radiation'][k])


def test_analysis_only_object_from_spectrum_file(tmp_path):
    """SynchRad(file_spectrum=...) needs no device (calc.py:98-99) and its Utilities work."""
    from synchrad.calc import SynchRad
    from synchrad_b200 import host
    A, dt = host.init_args({'grid': [(1.0, 3.0), (0.0, 0.2), (0.0, 2 * np.pi), (6, 4, 4)]})
    A.update(sigma_particle=0.0, timeStep=0.01, comp='total')
    c = _FakeCalc()
    c.Args, c.dtype = A, dt
    c.Data = {'radiation': {'total': np.ones((1, 6, 4, 4))}}
    c.snap_iterations = np.array([10], dtype=np.uint32)
    c.total_weight = 1.0
    p = str(tmp_path / 's.h5')
    trackio.write_spectrum(p, c)
    calc = SynchRad(file_spectrum=p)
    assert calc.Args['mode'] == 'far' and calc.get_full_spectrum().shape == (6, 4, 4)
    assert calc.get_energy() > 0


def test_lazy_file_tracks_pack_straight_from_the_file(tmp_path):
    """`file_tracks=` goes from the file into the packed SoA buffers without per-track arrays
    (trackio.TrackSource / FileTrack + host.pack_tracks), and gives exactly what the eager reader gives."""
    from synchrad_b200 import host
    rs = np.random.RandomState(3)
    lens = [1, 2, 57, 300, 31]
    tracks = [[rs.randn(n) for _ in range(6)] + [float(rs.uniform(0.5, 2)), int(rs.randint(0, 9))] for n in lens]
    path = str(tmp_path / 'tracks.h5')
    trackio.write_tracks(path, tracks, 0.02, it_range=(0, 250))
    index = [4, 0, 2, 3]
    eager = trackio.read_tracks(path, index)
    with trackio.TrackSource(path, index) as src:
        lazy = src.tracks
        assert [host.track_length(t) for t in lazy] == [lens[i] for i in index]
        assert [len(t) for t in lazy] == [8] * 4
        for a, b in zip(lazy, eager):
            assert a[6] == b[6] and a[7] == b[7] and a[-1] == b[7]
            np.testing.assert_array_equal(a[2], b[2])
        for it_range in (None, (0, 250)):
            w = [t[6] for t in eager]
            pe = host.pack_tracks(eager, w, np.double, it_range, 3)
            pl = host.pack_tracks(lazy, w, np.double, it_range, 3)
            assert pl.total == pe.total == sum(lens[i] for i in index)
            for ce, cl in zip(pe.coords, pl.coords):
                np.testing.assert_array_equal(ce, cl)
            for name in ('offsets', 'w', 'itStart', 'itEnd', 'itSnaps'):
                np.testing.assert_array_equal(getattr(pe, name), getattr(pl, name))
            assert pl.updates_per_node == pe.updates_per_node
    # a closed source cannot be read any more; a destination of the wrong size is refused
    with trackio.TrackSource(path, [3]) as src:
        with pytest.raises(ValueError):
            src.tracks[0].read_into(0, np.empty(7))
        dest = np.empty(300)
        src.tracks[0].read_into(5, dest)
        np.testing.assert_array_equal(dest, tracks[3][5])


def test_read_direct_converts_other_stored_types(tmp_path):
    path = str(tmp_path / 'f32.h5')
    f = h5lite.File(path, 'w')
    f['a/x'] = np.arange(10, dtype=np.float32) / 3
    f['a/i'] = np.arange(6, dtype=np.int64).reshape(2, 3)
    f.close()
    f = h5lite.File(path, 'r')
    assert f['a/x'].shape == (10,) and f['a/i'].shape == (2, 3)
    dest = np.empty(10)
    f['a/x'].read_direct(dest)
    np.testing.assert_array_equal(dest, (np.arange(10, dtype=np.float32) / 3).astype(np.double))
    di = np.empty((2, 3), dtype=np.int64)
    f['a/i'].read_direct(di)
    np.testing.assert_array_equal(di, np.arange(6).reshape(2, 3))
    f.close()


def test_group_listing_is_cached_for_many_tracks(tmp_path):
    """Opening 2000 tracks must not re-walk the `tracks` group B-tree per access (it did: 2.7 ms per track)."""
    import time
    rs = np.random.RandomState(0)
    tracks = [[rs.randn(4) for _ in range(6)] + [1.0, 0] for _ in range(2000)]
    path = str(tmp_path / 'many.h5')
    trackio.write_tracks(path, tracks, 0.01)
    t0 = time.perf_counter()
    got = trackio.read_tracks(path, range(2000))
    assert time.perf_counter() - t0 < 5.0
    np.testing.assert_array_equal(got[1999][3], tracks[1999][3])
