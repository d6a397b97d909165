"""Child of make_converter_golden.py: runs inside oracle/run_reference.run_script, i.e. with /root/reference first on
sys.path (so `synchrad` IS the unmodified reference) and the h5py stand-in of oracle/clshim.  Calls the reference's
converters.py / utils.py functions on the shared cases and pickles what they returned / wrote."""
import importlib.util
import os
import pickle
import tempfile

import numpy as np

here = os.environ['CONVERTER_CASES_DIR']
spec = importlib.util.spec_from_file_location('converter_cases', os.path.join(here, 'converter_cases.py'))
cc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cc)

import synchrad.utils as ru            # the reference (utils.py:8 pulls in converters.py)
import h5py

assert '/reference/' in ru.__file__, ru.__file__


def dump(path):
    """Every dataset of an HDF5 file as {path: array}."""
    out = {}
    f = h5py.File(path, 'r')

    def walk(node, prefix):
        for k in node.keys():
            try:
                sub = node[k]
                sub.keys()
                walk(sub, prefix + k + '/')
            except KeyError:
                out[prefix + k] = np.asarray(node[k][()])
    walk(f, '')
    f.close()
    return out


res = {}
with tempfile.TemporaryDirectory() as tmp:
    for name, kw in cc.OPMD_CASES.items():
        ts = cc.opmd_series()
        pt = cc.FakeTracker(ts, species='electrons')
        fn = os.path.join(tmp, f'opmd_{name}.h5')
        np.random.seed(1234)
        ru.tracksFromOPMD(ts, pt, ref_iteration=100, fname=fn, **kw)
        res[f'opmd/{name}'] = dump(fn)
    vs = os.path.join(tmp, 'vsim.h5')
    f = h5py.File(vs, 'w')
    f['tracks'] = cc.vsim_array()
    f.close()
    for name, kw in cc.VSIM_CASES.items():
        fn = os.path.join(tmp, f'vsim_{name}.h5')
        ru.tracksFromVSIM(vs, fn, **kw)
        res[f'vsim/{name}'] = dump(fn)
    for name, cols in cc.nan_series().items():
        pieces = ru.split_track_by_nans(*cols)
        res[f'split/{name}'] = [[np.asarray(v) for v in p] for p in pieces]
    # read_tracks / get_Larmor on a tracks file in the converters' layout
    fn = os.path.join(tmp, 'helper.h5')
    f = h5py.File(fn, 'w')
    tr = cc.helper_tracks()
    for i, t in enumerate(tr):
        for cname, a in zip(('x', 'y', 'z', 'ux', 'uy', 'uz'), t[:6]):
            f[f'tracks/{i}/{cname}'] = a
        f[f'tracks/{i}/w'] = t[6]
        f[f'tracks/{i}/it_start'] = t[7]
    f['misc/cdt'] = 0.1
    f['misc/N_particles'] = len(tr)
    f.close()
    for name, kw in (('all', {}), ('first3', dict(N_particles=3)), ('step4', dict(dt_step=4))):
        out = ru.read_tracks(fn, **kw)
        res[f'read/{name}'] = [np.asarray(v) for v in out]
        res[f'larmor/{name}'] = np.asarray(ru.get_Larmor(*out[:6], out[7]))

with open(os.environ['CONVERTER_GOLDEN_OUT'], 'wb') as fo:
    pickle.dump(res, fo)
