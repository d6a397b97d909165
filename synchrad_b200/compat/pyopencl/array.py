"""`pyopencl.array` as the reference uses it (`to_device`, `zeros`, `.data`, `.get()`, `.size`): host-memory arrays
that the library reads and accumulates into.  See synchrad_b200/compat/__init__.py."""
import numpy as np

from . import Buffer


class Array:
    def __init__(self, queue, ndarray):
        self.queue = queue
        self._a = ndarray
        self.data = Buffer(ndarray)

    shape = property(lambda self: self._a.shape)
    dtype = property(lambda self: self._a.dtype)
    size = property(lambda self: self._a.size)
    nbytes = property(lambda self: self._a.nbytes)

    def get(self):
        return self._a.copy()


def to_device(queue, ary):
    return Array(queue, np.array(ary, copy=True, order='C'))


def zeros(queue, shape, dtype):
    return Array(queue, np.zeros(shape, dtype=dtype))
