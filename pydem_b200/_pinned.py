"""Page-locked host buffers as NumPy arrays (pool), so the reference-facing calls move data at
full PCIe rate.  Buffers come from ``pdm_host_alloc`` (cudaHostAlloc) and return to the pool
when the array that wraps them is garbage collected."""
import ctypes as ct
import weakref

import numpy as np

from . import _lib

_POOL = {}       # nbytes -> [address, ...] free buffers
_MAX_POOLED_BYTES = 8 << 30
_pooled = 0


class _Owner(object):
    __slots__ = ("addr", "nbytes", "__weakref__")


def _release(addr, nbytes):
    global _pooled
    if _pooled + nbytes <= _MAX_POOLED_BYTES:
        _POOL.setdefault(nbytes, []).append(addr)
        _pooled += nbytes
    else:
        try:
            _lib.load().pdm_host_free(ct.c_void_p(addr))
        except Exception:
            pass


def empty(shape, dtype):
    """Uninitialised pinned array of the given shape/dtype."""
    global _pooled
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    free = _POOL.get(nbytes)
    if free:
        addr = free.pop()
        _pooled -= nbytes
    else:
        p = ct.c_void_p()
        _lib.check(_lib.load().pdm_host_alloc(max(nbytes, 1), ct.byref(p)))
        addr = p.value
    own = _Owner()
    own.addr, own.nbytes = addr, nbytes
    weakref.finalize(own, _release, addr, nbytes)
    buf = (ct.c_char * max(nbytes, 1)).from_address(addr)
    buf._owner = own  # keeps the owner alive as long as any view of the buffer lives
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def pinned_copy(a):
    out = empty(a.shape, a.dtype)
    out[...] = a
    return out
