"""Tracks-file producers and helpers (synchrad_b200/converters.py) against the UNMODIFIED reference's converters.py /
utils.py on the same inputs: tests/golden/converter_cases.npz holds what the reference returned and wrote
(tests/golden/make_converter_golden.py); when /root/reference is present the reference is run live as well."""
import os

import numpy as np
import pytest

from golden import converter_cases as cc
from golden.make_converter_golden import flatten, reference_outputs
from oracle import run_reference
from synchrad_b200 import converters, trackio

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'converter_cases.npz')


def dump(path):
    out = {}
    f = trackio._h5.File(path, 'r')

    def walk(node, prefix):
        for k in node.keys():
            try:
                node[k].keys()
                walk(node[k], prefix + k + '/')
            except KeyError:
                out[prefix + k] = np.asarray(node[k][()])
    walk(f, '')
    f.close()
    return out


def product_outputs(tmp):
    res = {}
    for name, kw in cc.OPMD_CASES.items():
        ts = cc.opmd_series()
        pt = cc.FakeTracker(ts, species='electrons')
        fn = os.path.join(tmp, f'opmd_{name}.h5')
        np.random.seed(1234)
        converters.tracksFromOPMD(ts, pt, ref_iteration=100, fname=fn, **kw)
        assert pt.init_kwargs == dict(iteration=100, preserve_particle_index=True)
        res[f'opmd/{name}'] = dump(fn)
    vs = os.path.join(tmp, 'vsim.h5')
    f = trackio._h5.File(vs, 'w')
    f['tracks'] = cc.vsim_array()
    f.close()
    for name, kw in cc.VSIM_CASES.items():
        fn = os.path.join(tmp, f'vsim_{name}.h5')
        converters.tracksFromVSIM(vs, fn, **kw)
        res[f'vsim/{name}'] = dump(fn)
    for name, cols in cc.nan_series().items():
        res[f'split/{name}'] = [[np.asarray(v) for v in p] for p in converters.split_track_by_nans(*cols)]
    fn = os.path.join(tmp, 'helper.h5')
    trackio.write_tracks(fn, cc.helper_tracks(), cdt=0.1)
    for name, kw in (('all', {}), ('first3', dict(N_particles=3)), ('step4', dict(dt_step=4))):
        out = converters.read_tracks(fn, **kw)
        res[f'read/{name}'] = [np.asarray(v) for v in out]
        res[f'larmor/{name}'] = np.asarray(converters.get_Larmor(*out[:6], out[7]))
    return res


def assert_same(got, want):
    assert sorted(got) == sorted(want)
    for k in sorted(want):
        g, w = np.asarray(got[k]), np.asarray(want[k])
        assert g.shape == w.shape, (k, g.shape, w.shape)
        if w.dtype.kind in 'SUO':
            assert [str(v) for v in np.atleast_1d(g).tolist()] == [str(v) for v in np.atleast_1d(w).tolist()], k
        else:
            assert g.dtype.kind == w.dtype.kind, (k, g.dtype, w.dtype)
            if k.startswith('larmor/'):
                np.testing.assert_allclose(g, w, rtol=1e-13, atol=0, err_msg=k)   # same formula, other association
            else:
                assert np.array_equal(g, w, equal_nan=True), k


@pytest.fixture(scope='module')
def product(tmp_path_factory):
    return flatten(product_outputs(str(tmp_path_factory.mktemp('conv'))))


def test_converters_equal_the_stored_reference_outputs(product):
    stored = np.load(GOLD, allow_pickle=False)
    assert_same(product, {k: stored[k] for k in stored.files})
    # the cases do exercise the branches: pieces cut at gaps, short tracks dropped, concatenated VSim samples
    assert int(stored['opmd/all//misc/N_particles']) > cc.opmd_series().n_all - 4
    assert int(stored['opmd/short_12//misc/N_particles']) < int(stored['opmd/all//misc/N_particles'])
    assert int(stored['split/mid//n']) == 3 and int(stored['split/all//n']) == 0
    assert int(stored['vsim/plain//misc/N_particles']) == 8


@pytest.mark.skipif(not run_reference.available(), reason='needs /root/reference (build container)')
def test_converters_equal_the_live_reference(product):
    assert_same(product, flatten(reference_outputs()))


def test_converter_edge_cases(tmp_path):
    # both iteration bounds at once (IndexError in the reference, converters.py:56-60): the window sets cdt
    ts = cc.opmd_series(drop=False)
    pt = cc.FakeTracker(ts, species='e')
    fn = str(tmp_path / 'w.h5')
    converters.tracksFromOPMD(ts, pt, 100, fname=fn, Nit_min=200, Nit_max=700)
    d = dump(fn)
    assert int(d['misc/N_particles']) == ts.n_all
    k = int(np.flatnonzero(ts.iterations >= 200)[0])
    assert float(d['misc/cdt']) == (ts.t[k + 1] - ts.t[k]) * cc.C
    # z_is_xi on full-length tracks: z + c t (converters.py:102-103)
    converters.tracksFromOPMD(ts, pt, 100, fname=fn, z_is_xi=True)
    d2 = dump(fn)
    want = np.array([ts.data[i]['z'][0] for i in range(len(ts.data))]) + cc.C * ts.t
    assert np.array_equal(d2['tracks/0/z'], want)
    with pytest.raises(ValueError):
        converters.tracksFromOPMD(ts, pt, 100, fname=fn, z_is_xi=True, Nit_min=200)
    with pytest.raises(ValueError):
        converters.tracksFromOPMD(ts, pt, 100, fname=fn, Np_select=3, sample_selection='best')
    with pytest.raises(NotImplementedError):
        converters.tracksFromOPMD_old(ts, pt, 100)
    # nothing survives the length cut: an empty, still readable file
    converters.tracksFromOPMD(ts, pt, 100, fname=fn, shortest_track=10 ** 6)
    assert trackio.read_header(fn)[2] == 0


def test_reference_import_paths():
    """README of the reference: `from synchrad.utils import tracksFromOPMD`; utils.py:8 re-exports converters.py."""
    import synchrad.converters as sc
    import synchrad.utils as su
    for name in ('tracksFromOPMD', 'tracksFromVSIM', 'split_track_by_nans', 'read_tracks', 'get_Larmor', 'J_in_um'):
        assert callable(getattr(su, name)) or name == 'J_in_um'
    assert sc.tracksFromOPMD is converters.tracksFromOPMD
    with pytest.raises(NotImplementedError):
        from synchrad.utils import record_particles_step  # noqa: F401
    with pytest.raises(ImportError):
        from synchrad.utils import no_such_name  # noqa: F401


def test_converted_file_feeds_the_path(tmp_path):
    """What tracksFromOPMD writes is what calculate_spectrum(file_tracks=...) reads (calc.py:186-219): header, ragged
    tracks with their it_start, packed into the C-ABI layout."""
    from synchrad_b200 import host
    ts = cc.opmd_series()
    pt = cc.FakeTracker(ts, species='e')
    fn = str(tmp_path / 't.h5')
    converters.tracksFromOPMD(ts, pt, 100, fname=fn)
    cdt, rng, n = trackio.read_header(fn)
    assert cdt == (ts.t[1] - ts.t[0]) * cc.C and n > 0 and rng[0] >= 0 and rng[1] <= len(ts.data)
    with trackio.TrackSource(fn, range(n)) as src:
        packed = host.pack_tracks(src.tracks, [t[6] for t in src.tracks], np.double, rng, 2, None)
        lists = trackio.read_tracks(fn, range(n))
    assert packed.n == n and packed.total == sum(t[0].size for t in lists)
    off = packed.offsets
    for i, t in enumerate(lists):
        assert np.array_equal(packed.coords[0][off[i]:off[i + 1]], t[0])
        assert packed.itStart[i] == t[7]
