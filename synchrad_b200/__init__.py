"""synchrad_b200 — B200-native (sm_100a) spectral-integration path of SynchRad.

Public surface: `synchrad_b200.calc.SynchRad` (also importable as `synchrad.calc.SynchRad`,
the reference's import path) and the C ABI in include/synchrad_b200.h.
"""
__version__ = '0.1.0'
