"""`from synchrad.utils import J_in_um` — the reference's import path (utils.py:16-19)."""
from synchrad_b200.utils import (Utilities, J_in_um, r_e, omega_1m, energy_1m_eV,  # noqa: F401
                                 alpha_fs)

# The reference's `utils` also re-exports converters.py (tracksFromOPMD, tracksFromVSIM, ...) and defines
# read_tracks / get_Larmor: host-side file conversion outside the spectral-integration path this package
# replaces (out of scope).  Fail with a clear message instead of a bare ImportError.
_OUT_OF_SCOPE = ('tracksFromOPMD', 'tracksFromOPMD_old', 'tracksFromVSIM', 'split_track_by_nans',
                 'record_particles_step', 'record_particles_first', 'read_tracks', 'get_Larmor')


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise NotImplementedError(
            f'synchrad.utils.{name} belongs to the reference\'s track converters / helpers, which synchrad_b200 '
            'does not replace (only the spectral-integration path is). Use the reference package to write the tracks '
            'file; this package reads that layout (synchrad_b200/trackio.py).')
    raise AttributeError(name)
