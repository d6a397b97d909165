"""`from synchrad.converters import tracksFromOPMD` -- the reference's import path (converters.py)."""
from synchrad_b200.converters import *  # noqa: F401,F403
from synchrad_b200.converters import __all__  # noqa: F401


def __getattr__(name):          # the helpers of the unmirrored tracksFromOPMD_old: a clear message instead of ImportError
    import synchrad_b200.converters as _c
    return getattr(_c, name)
