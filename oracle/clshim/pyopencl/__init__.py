"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

A stand-in for the `pyopencl` module, just large enough for the UNMODIFIED reference host code
(/root/reference/synchrad/calc.py) to run in an image that has no OpenCL runtime:

    cl.create_some_context / cl.CommandQueue / cl.device_type   calc.py:4, 513-558
    cl.Program(ctx, src).build() and kernel calls               calc.py:624, 324-353
    pyopencl.array.to_device / zeros / Array.get / .data        calc.py:479-512, 573-603, 630

`Program.build()` does what an OpenCL CPU driver does: it compiles the kernel source it is handed -- the
reference's own kernel_farfield.cl / kernel_nearfield.cl as rendered by calc.py:605-624 -- for the host
(g++ with the prelude cl_shim.hpp, source piped on stdin, nothing of it is written to disk) and runs the
NDRange as an OpenMP loop over work-items.  The arithmetic policy (dot order, rsqrt, libm sin/cos, no FMA
contraction by default) is stated in cl_shim.hpp; CLSHIM_CXXFLAGS overrides the optimisation flags.
Binaries are cached by content hash under oracle/_ref/ (git-ignored).
"""
import ctypes
import hashlib
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM_DIR = os.path.dirname(_HERE)
_CACHE = os.environ.get('CLSHIM_CACHE', os.path.join(os.path.dirname(_SHIM_DIR), '_ref'))
DEFAULT_CXXFLAGS = '-O2 -ffp-contract=off -fno-fast-math'

VERSION_TEXT = 'clshim (g++ host execution of OpenCL C)'


class device_type:
    CPU = 2

    @staticmethod
    def to_string(value):
        return {2: 'CPU'}.get(value, 'UNKNOWN')


class _Platform:
    vendor = 'clshim'
    name = 'clshim host platform'


class Device:
    type = device_type.CPU
    name = 'host cores via g++/OpenMP'
    platform = _Platform()
    opencl_c_version = 'OpenCL C (clshim: g++ -std=c++17 + cl_shim.hpp)'
    max_work_group_size = 1024


class Context:
    def __init__(self):
        self.devices = [Device()]


def create_some_context(interactive=None, answers=None):
    return Context()


class CommandQueue:
    def __init__(self, context, device=None):
        self.context = context
        self.device = device or context.devices[0]

    def finish(self):
        pass


class Buffer:
    """What `Array.data` hands to a kernel call."""
    def __init__(self, ndarray):
        self.ndarray = ndarray


_C_TYPES = {'double': (ctypes.c_double, np.float64), 'float': (ctypes.c_float, np.float32),
            'uint': (ctypes.c_uint32, np.uint32), 'int': (ctypes.c_int32, np.int32)}
_KERNEL_RE = re.compile(r'__kernel\s+void\s+(\w+)\s*\((.*?)\)\s*\{', re.S)


def _parse_kernels(src):
    kernels = {}
    for name, params in _KERNEL_RE.findall(src):
        sig = []
        for p in params.split(','):
            toks = p.replace('*', ' * ').split()
            is_ptr = '*' in toks
            base = [t for t in toks if t in _C_TYPES]
            if len(base) != 1:
                raise NotImplementedError(f'clshim: cannot parse kernel parameter {p.strip()!r}')
            sig.append((base[0], is_ptr, toks[-1]))
        kernels[name] = sig
    return kernels


class _Kernel:
    def __init__(self, name, sig, fn):
        self.name, self.sig, self.fn = name, sig, fn
        fn.restype = None
        fn.argtypes = [ctypes.c_size_t] + [ctypes.c_void_p if ptr else _C_TYPES[t][0] for t, ptr, _ in sig]

    def __call__(self, queue, global_size, local_size, *args):
        if len(args) != len(self.sig):
            raise TypeError(f'{self.name}: {len(self.sig)} kernel arguments expected, {len(args)} given')
        if len(global_size) != 1:
            raise NotImplementedError('clshim: 1-D NDRange only')
        if local_size is not None and global_size[0] % local_size[0] != 0:
            raise ValueError('global size is not a multiple of the work-group size')   # CL_INVALID_WORK_GROUP_SIZE
        conv, keep = [], []
        for (t, ptr, pname), a in zip(self.sig, args):
            ct, nt = _C_TYPES[t]
            if ptr:
                if not isinstance(a, Buffer):
                    raise TypeError(f'{self.name}: argument {pname} must be a buffer')
                if a.ndarray.dtype != nt:
                    raise TypeError(f'{self.name}: buffer {pname} holds {a.ndarray.dtype}, kernel reads {t}')
                keep.append(a.ndarray)
                conv.append(a.ndarray.ctypes.data)
            else:
                # pyopencl packs a scalar by its own numpy dtype; a plain Python number is rejected there
                if not isinstance(a, np.generic):
                    raise TypeError(f'{self.name}: scalar {pname} must be a sized numpy scalar, got {type(a)}')
                if a.dtype.itemsize != np.dtype(nt).itemsize or a.dtype.kind != np.dtype(nt).kind:
                    raise TypeError(f'{self.name}: scalar {pname} is {a.dtype}, kernel takes {t}')
                conv.append(ct(a.item()))
        self.fn(int(global_size[0]), *conv)


class _Built:
    def __init__(self, lib, kernels):
        self._lib = lib
        for name, sig in kernels.items():
            setattr(self, name, _Kernel(name, sig, getattr(lib, 'clshim_launch_' + name)))


class Program:
    def __init__(self, context, src):
        self.context, self.src = context, src

    def translation_unit(self):
        kernels = _parse_kernels(self.src)
        out = ['#include "cl_shim.hpp"', 'namespace clshim { thread_local size_t global_id0; }', self.src]
        for name, sig in kernels.items():
            params = ', '.join(f'{t}{"*" if ptr else ""} {p}' for t, ptr, p in sig)
            call = ', '.join(p for _, _, p in sig)
            out.append(f'extern "C" void clshim_launch_{name}(size_t clshim_n, {params}) {{\n'
                       f'  _Pragma("omp parallel for schedule(static)")\n'
                       f'  for (size_t g = 0; g < clshim_n; g++) {{ clshim::global_id0 = g; {name}({call}); }}\n}}')
        return kernels, '\n'.join(out) + '\n'

    def build(self, options=None):
        kernels, tu = self.translation_unit()
        flags = os.environ.get('CLSHIM_CXXFLAGS', DEFAULT_CXXFLAGS).split()
        with open(os.path.join(_SHIM_DIR, 'cl_shim.hpp'), 'rb') as f:
            prelude = f.read()
        tag = hashlib.sha256(prelude + tu.encode() + ' '.join(flags).encode()).hexdigest()[:16]
        os.makedirs(_CACHE, exist_ok=True)
        so = os.path.join(_CACHE, f'clshim_{tag}.so')
        if not os.path.exists(so):
            tmp = f'{so}.{os.getpid()}.tmp'
            cmd = ['g++', '-std=c++17', *flags, '-fopenmp', '-fPIC', '-shared', '-Wno-unknown-pragmas',
                   '-I', _SHIM_DIR, '-x', 'c++', '-', '-o', tmp]
            r = subprocess.run(cmd, input=tu.encode(), capture_output=True)
            if r.returncode != 0:
                raise RuntimeError('clshim: kernel build failed\n' + r.stderr.decode()[-4000:])
            os.replace(tmp, so)
        return _Built(ctypes.CDLL(so), kernels)
