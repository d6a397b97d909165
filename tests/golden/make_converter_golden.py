"""tests/golden/converter_cases.npz: outputs of the UNMODIFIED reference's converters.py (tracksFromOPMD,
tracksFromVSIM, split_track_by_nans) and utils.py helpers (read_tracks, get_Larmor) on tests/golden/converter_cases.py.

    python tests/golden/make_converter_golden.py          (build container only: needs /root/reference)

The reference's modules are imported from where they lie (oracle/run_reference.run_script: /root/reference first on
sys.path); h5py is the stand-in of oracle/clshim (the repo's h5lite), openPMD-viewer's objects are the duck-typed
fakes of converter_cases.py."""
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import run_reference  # noqa: E402


def reference_outputs():
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, 'out.pkl')
        os.environ['CONVERTER_CASES_DIR'] = HERE
        os.environ['CONVERTER_GOLDEN_OUT'] = out
        run_reference.run_script(os.path.join(HERE, '_converter_reference_child.py'))
        with open(out, 'rb') as f:
            return pickle.load(f)


def flatten(res):
    flat = {}
    for key, val in res.items():
        if isinstance(val, dict):
            for k, a in val.items():
                flat[f'{key}//{k}'] = a
        elif isinstance(val, list) and val and isinstance(val[0], list):
            flat[f'{key}//n'] = np.int64(len(val))
            for i, piece in enumerate(val):
                for j, a in enumerate(piece):
                    flat[f'{key}//{i}/{j}'] = a
        elif isinstance(val, list):
            flat[f'{key}//n'] = np.int64(len(val))
            for j, a in enumerate(val):
                flat[f'{key}//{j}'] = a
        else:
            flat[key] = val
    return flat


if __name__ == '__main__':
    flat = flatten(reference_outputs())
    np.savez_compressed(os.path.join(HERE, 'converter_cases.npz'), **flat)
    print(len(flat), 'arrays ->', os.path.join(HERE, 'converter_cases.npz'))
