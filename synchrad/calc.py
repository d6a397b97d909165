"""`from synchrad.calc import SynchRad` — the reference's import path (calc.py:21)."""
from synchrad_b200.calc import SynchRad  # noqa: F401
