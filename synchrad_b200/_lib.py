"""ctypes binding of libsynchrad_b200.so (include/synchrad_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `synchrad_b200.build.build()`.
There is NO fallback: if the shared object is missing or a symbol is absent this module
raises, and every compute entry point of the package goes through it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SYNCHRAD_B200_LIB') or os.path.join(_HERE, 'csrc', 'libsynchrad_b200.so')

MODE = {'far': 0, 'near': 1}
COMP = {'total': 0, 'cartesian': 1, 'cartesian_complex': 2, 'spheric': 3, 'spheric_complex': 4}
DTYPE = {'double': 0, 'float': 1, 'float_literal': 2}
PHASOR = {'auto': 0, 'direct': 1, 'recur': 2, 'pair': 3, 'pair_fma': 4, 'drec': 5}

# every symbol include/synchrad_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = ('srb_version', 'srb_last_error', 'srb_num_spectra', 'srb_scratch_bytes',
           'srb_integrate', 'srb_integrate_host', 'srb_swap_axes', 'srb_last_launch',
           'srb_pipe_peak', 'srb_energy_spectrum')


class srb_grid(ctypes.Structure):
    _fields_ = [
        ('mode', ctypes.c_int32), ('comp', ctypes.c_int32), ('dtype', ctypes.c_int32),
        ('native', ctypes.c_int32), ('phasor', ctypes.c_int32), ('omega_uniform', ctypes.c_int32),
        ('nOmega', ctypes.c_uint32), ('nAxis2', ctypes.c_uint32), ('nPhi', ctypes.c_uint32),
        ('nSnaps', ctypes.c_uint32),
        ('omega', ctypes.c_void_p), ('sinTheta', ctypes.c_void_p), ('cosTheta', ctypes.c_void_p),
        ('radius', ctypes.c_void_p), ('sinPhi', ctypes.c_void_p), ('cosPhi', ctypes.c_void_p),
        ('formFactor', ctypes.c_void_p),
        ('L_screen', ctypes.c_double), ('dt', ctypes.c_double),
        ('omega_first_host', ctypes.c_double), ('omega_last_host', ctypes.c_double),
    ]


class srb_tracks(ctypes.Structure):
    _fields_ = [
        ('nTracks', ctypes.c_uint32),
        ('x', ctypes.c_void_p), ('y', ctypes.c_void_p), ('z', ctypes.c_void_p),
        ('ux', ctypes.c_void_p), ('uy', ctypes.c_void_p), ('uz', ctypes.c_void_p),
        ('offsets', ctypes.c_void_p), ('w', ctypes.c_void_p),
        ('itStart', ctypes.c_void_p), ('itEnd', ctypes.c_void_p), ('itSnaps', ctypes.c_void_p),
        ('itSnapsStride', ctypes.c_uint32),
        ('totalSteps_host', ctypes.c_uint64),
    ]


class srb_launch_info(ctypes.Structure):
    _fields_ = [
        ('kind', ctypes.c_int32), ('tile_width', ctypes.c_int32),
        ('chunk_nodes', ctypes.c_uint32), ('n_chunks', ctypes.c_uint32),
        ('n_virtual_dirs', ctypes.c_uint32), ('n_particle_chunks', ctypes.c_uint32),
        ('grid_blocks', ctypes.c_uint32), ('block_threads', ctypes.c_uint32),
        ('smem_bytes', ctypes.c_uint32), ('kernels_launched', ctypes.c_uint32),
        ('n_components', ctypes.c_uint32),
        ('n_time_segments', ctypes.c_uint32),
    ]


_lib = None


def load():
    """Load the CUDA library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; '
            'g.build()"` (nvcc, sm_100a). synchrad_b200 has no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name in SYMBOLS:
        if not hasattr(lib, name):
            raise RuntimeError(f'{LIB_PATH} does not export {name}')
    P = ctypes.POINTER
    lib.srb_version.restype = ctypes.c_int
    lib.srb_last_error.restype = ctypes.c_char_p
    lib.srb_num_spectra.restype = ctypes.c_int
    lib.srb_num_spectra.argtypes = [ctypes.c_int, ctypes.c_int]
    lib.srb_scratch_bytes.restype = ctypes.c_size_t
    lib.srb_scratch_bytes.argtypes = [P(srb_grid), P(srb_tracks)]
    lib.srb_integrate.restype = ctypes.c_int
    lib.srb_integrate.argtypes = [P(srb_grid), P(srb_tracks), P(ctypes.c_void_p), ctypes.c_int,
                                  ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                  ctypes.c_void_p]
    lib.srb_integrate_host.restype = ctypes.c_int
    lib.srb_integrate_host.argtypes = [P(srb_grid), P(srb_tracks), P(ctypes.c_void_p),
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.srb_swap_axes.restype = ctypes.c_int
    lib.srb_swap_axes.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_uint32] * 4 \
        + [ctypes.c_void_p]
    lib.srb_energy_spectrum.restype = ctypes.c_int
    lib.srb_energy_spectrum.argtypes = [ctypes.c_int, ctypes.c_int, P(ctypes.c_void_p), ctypes.c_int, ctypes.c_int] \
        + [ctypes.c_uint32] * 5 + [ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    lib.srb_pipe_peak.restype = ctypes.c_int
    lib.srb_pipe_peak.argtypes = [ctypes.c_int, P(ctypes.c_double)]
    lib.srb_last_launch.restype = ctypes.c_int
    lib.srb_last_launch.argtypes = [P(srb_launch_info)]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().srb_last_error().decode(errors='replace')
        raise RuntimeError(f'synchrad_b200: {msg}')
