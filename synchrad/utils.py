"""`from synchrad.utils import J_in_um` — the reference's import path (utils.py:16-19)."""
from synchrad_b200.utils import (Utilities, J_in_um, r_e, omega_1m, energy_1m_eV,  # noqa: F401
                                 alpha_fs)
