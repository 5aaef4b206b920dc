"""Argument marshalling shared by the `renderer` and `ggx` module mirrors.

The reference's Cython signatures take typed, C-contiguous NumPy buffers
(`np.ndarray[float, ndim=2, mode='c']`, smoothed_transient/renderer.pyx:13-200): a wrong dtype / layout
raises ValueError there, a wrong shape raises AssertionError.  The same rules apply here; additionally a
torch CUDA tensor of the same dtype/shape is accepted wherever an array is, and is used in place (no copy);
the call is then ordered against torch's current stream on both sides (see _ffi._LibProxy), so tensors just written by
torch ops are read correctly and torch ops enqueued after the call see its outputs.
"""
import ctypes as C
import math
import numpy as np

_NP = {'f32': np.float32, 'i32': np.int32, 'f64': np.float64, 'u8': np.uint8}
_CT = {'f32': C.c_float, 'i32': C.c_int, 'f64': C.c_double, 'u8': C.c_uint8}


# stream of the torch CUDA tensors seen by as_pointer() since the last C-ABI call of this thread: _ffi's library proxy orders the
# context's stream against it (nlos_ctx_wait_stream before the call, nlos_ctx_signal_stream after) and clears it
import threading
tls = threading.local()


def _is_torch(a):
    return type(a).__module__.startswith('torch') and hasattr(a, 'data_ptr')


def as_pointer(a, kind, ndim, name, allow_none=False):
    """-> (ctypes pointer, shape tuple). Keeps no reference: the caller's array must outlive the call."""
    if a is None:
        if allow_none:
            return None, None
        raise ValueError('%s must not be None' % name)
    if _is_torch(a):
        import torch
        want = {'f32': torch.float32, 'i32': torch.int32, 'f64': torch.float64, 'u8': torch.uint8}[kind]
        if a.dtype != want:
            raise ValueError("Buffer dtype mismatch for %s: expected %s but got %s" % (name, want, a.dtype))
        if a.dim() != ndim:
            raise ValueError('Buffer has wrong number of dimensions for %s (expected %d, got %d)' % (name, ndim, a.dim()))
        if not a.is_contiguous():
            raise ValueError('%s: tensor is not C-contiguous' % name)
        if a.is_cuda:
            tls.cuda_stream = torch.cuda.current_stream(a.device).cuda_stream
        return C.cast(C.c_void_p(a.data_ptr()), C.POINTER(_CT[kind])), tuple(a.shape)
    if not isinstance(a, np.ndarray):
        raise TypeError('Argument %s has incorrect type (expected numpy.ndarray, got %s)' % (name, type(a).__name__))
    if a.dtype != _NP[kind]:
        raise ValueError("Buffer dtype mismatch for %s, expected '%s' but got '%s'" % (name, np.dtype(_NP[kind]).name, a.dtype.name))
    if a.ndim != ndim:
        raise ValueError('Buffer has wrong number of dimensions for %s (expected %d, got %d)' % (name, ndim, a.ndim))
    if not a.flags['C_CONTIGUOUS']:
        raise ValueError('%s: ndarray is not C-contiguous' % name)
    return a.ctypes.data_as(C.POINTER(_CT[kind])), tuple(a.shape)


def num_bins(lower_bound, upper_bound, resolution):
    """math.ceil((upper_bound - lower_bound)/resolution) with the operands typed C `float` (renderer.pyx:101)."""
    lo, hi, res = np.float32(lower_bound), np.float32(upper_bound), np.float32(resolution)
    return int(math.ceil(float(np.float32(np.float32(hi - lo) / res))))
