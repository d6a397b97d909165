"""`from synchrad.utils import J_in_um, tracksFromOPMD, read_tracks` -- the reference's import path: utils.py:8
re-exports converters.py and :16-19, :218-266 define the constants and the two module-level helpers."""
from synchrad_b200.utils import (Utilities, J_in_um, r_e, omega_1m, energy_1m_eV,  # noqa: F401
                                 alpha_fs)
from synchrad_b200.converters import (tracksFromOPMD, tracksFromOPMD_old, tracksFromVSIM,  # noqa: F401
                                      split_track_by_nans, tracks_from_series, read_tracks, get_Larmor)


def __getattr__(name):          # the helpers of the unmirrored tracksFromOPMD_old: a clear message instead of ImportError
    import synchrad_b200.converters as _c
    return getattr(_c, name)
