"""ORACLE — TEST INFRASTRUCTURE ONLY.  `h5py.File` for the reference host code when h5py is not installed: the
repo's minimal HDF5 reader/writer (synchrad_b200/h5lite.py), loaded by path so that the repo's own `synchrad`
alias package never shadows the reference's."""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'synchrad_b200', 'h5lite.py')
_spec = importlib.util.spec_from_file_location('_clshim_h5lite', _path)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
File = _mod.File
