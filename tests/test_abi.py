"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares
(no compute calls here)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, 'include', '*.h')):
        text = re.sub(r'/\*.*?\*/', '', open(h).read(), flags=re.S)
        names |= set(re.findall(r'\b(srb_[a-z_0-9]+)\s*\(', text))
    return names


def test_library_exports_every_declared_symbol():
    from synchrad_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = declared_symbols()
    assert decl, 'no declarations found'
    assert decl == set(_lib.SYMBOLS), (decl ^ set(_lib.SYMBOLS))
    for name in decl:
        assert hasattr(lib, name), name


def test_struct_layouts_match_header():
    """ctypes mirrors of srb_grid / srb_tracks: field order and count must match the header."""
    from synchrad_b200 import _lib
    text = open(os.path.join(ROOT, 'include', 'synchrad_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    for cname, cls in (('srb_grid', _lib.srb_grid), ('srb_tracks', _lib.srb_tracks),
                       ('srb_launch_info', _lib.srb_launch_info)):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), text, flags=re.S).group(1)
        fields = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(','):
                fields.append(re.findall(r'[A-Za-z_0-9]+', part)[-1])
        assert fields == [f[0] for f in cls._fields_], (cname, fields)


def test_enum_constants_match_header():
    """SRB_PHASOR_* / SRB_DTYPE_* values in the header are the ones the ctypes layer sends."""
    from synchrad_b200 import _lib
    text = open(os.path.join(ROOT, 'include', 'synchrad_b200.h')).read()
    defs = {k: int(v) for k, v in re.findall(r'#define\s+(SRB_[A-Z0-9_]+)\s+(\d+)', text)}
    assert {k: defs['SRB_PHASOR_' + k.upper()] for k in _lib.PHASOR} == _lib.PHASOR
    assert len([k for k in defs if k.startswith('SRB_PHASOR_')]) == len(_lib.PHASOR)
    assert defs['SRB_DTYPE_F64'] == 0 and defs['SRB_DTYPE_F32'] == 1 and defs['SRB_DTYPE_F32_LITERAL'] == 2


def test_host_only_entry_points():
    from synchrad_b200 import _lib
    lib = _lib.load()
    assert lib.srb_version() == 3      # ABI 3: srb_launch_info.n_time_segments (2: counters[4], on-device kernel choice)
    assert lib.srb_num_spectra(0, 0) == 1 and lib.srb_num_spectra(0, 1) == 3
    assert lib.srb_num_spectra(0, 2) == 6 and lib.srb_num_spectra(0, 4) == 6
    assert lib.srb_num_spectra(1, 3) < 0 and lib.srb_num_spectra(1, 4) < 0   # no near spheric kernels
    assert lib.srb_num_spectra(2, 0) < 0 and lib.srb_num_spectra(0, 9) < 0


def test_errors_are_reported_not_thrown():
    from synchrad_b200 import _lib
    lib = _lib.load()
    g, t = _lib.srb_grid(), _lib.srb_tracks()
    g.mode = 7
    rc = lib.srb_integrate(ctypes.byref(g), ctypes.byref(t), None, 0, None, 0, None, None)
    assert rc != 0 and b'mode' in lib.srb_last_error()


def test_product_has_no_cpu_fallback():
    """The package must not import anything under oracle/ or tests/emu, and constructing a
    calculator without CUDA must raise (never compute on the CPU)."""
    import torch
    pkg = os.path.join(ROOT, 'synchrad_b200')
    for path in glob.glob(os.path.join(pkg, '*.py')) + glob.glob(os.path.join(ROOT, 'synchrad', '*.py')):
        src = open(path).read()
        assert 'oracle' not in src and 'emu' not in src, path
    if not torch.cuda.is_available():
        from synchrad.calc import SynchRad
        with pytest.raises(RuntimeError, match='no CPU'):
            SynchRad({'grid': [(1, 2), (0, 1), (0, 1), (4, 2, 2)]})
