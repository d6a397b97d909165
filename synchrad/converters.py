"""`from synchrad.converters import tracksFromOPMD` -- the reference's import path (converters.py)."""
from synchrad_b200.converters import *  # noqa: F401,F403
from synchrad_b200.converters import __all__  # noqa: F401
