"""ctypes loader for libnlos_b200.so (the C ABI of include/nlos_b200.h).

The shared library is built in-tree by `make -C nlos_surface_optimization_b200/csrc` (or
`__graft_entry__.build()`).  There is no fallback: if the library is missing, or no B200-class GPU is
usable, importing works but creating a context raises NlosError loudly.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NLOS_B200_LIB', os.path.join(_HERE, 'libnlos_b200.so'))   # override: A/B builds of the kernels

NLOS_OK, NLOS_ERR_INVALID, NLOS_ERR_CUDA, NLOS_ERR_NOMEM = 0, 1, 2, 3


class NlosError(RuntimeError):
    pass


_lib = None
_lock = threading.Lock()

_f = C.POINTER(C.c_float)
_d = C.POINTER(C.c_double)
_i = C.POINTER(C.c_int)
_ctx = C.c_void_p

# every symbol declared in include/nlos_b200.h with its argument types (tests/test_abi_symbols.py
# cross-checks this table against the header)
SIGNATURES = {
    'nlos_ctx_create': (C.c_int, [C.c_int, C.POINTER(_ctx)]),
    'nlos_ctx_destroy': (None, [_ctx]),
    'nlos_last_error': (C.c_char_p, [_ctx]),
    'nlos_ctx_synchronize': (C.c_int, [_ctx]),
    'nlos_ctx_stream': (C.c_void_p, [_ctx]),
    'nlos_ctx_set_seed': (C.c_int, [_ctx, C.c_uint64]),
    'nlos_ctx_set_source_window': (C.c_int, [_ctx, C.c_int64, C.c_int64]),
    'nlos_ctx_set_option': (C.c_int, [_ctx, C.c_char_p, C.c_int64]),
    'nlos_ctx_get_work_counters': (C.c_int, [_ctx, C.POINTER(C.c_uint64)]),
    'nlos_ctx_wait_stream': (C.c_int, [_ctx, C.c_void_p]),
    'nlos_ctx_signal_stream': (C.c_int, [_ctx, C.c_void_p]),
    'nlos_ctx_get_timing': (C.c_int, [_ctx, _f]),
    'nlos_ctx_launch_count': (C.c_uint64, [_ctx]),
    'nlos_ctx_set_external_samples': (C.c_int, [_ctx, _f, C.c_int64]),
    'nlos_streamed_render_transient': (C.c_int, [_ctx, _f, C.c_int, _f, _f, C.c_int, _f, _f, _i, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                                 _d, _d, C.c_int, C.c_int, C.c_int]),
    'nlos_streamed_render_intensity': (C.c_int, [_ctx, _f, C.c_int, _f, _f, C.c_int, _f, _i, C.c_int, C.c_int, C.c_float, C.c_float, _d]),
    'nlos_streamed_render_gradient': (C.c_int, [_ctx, _d, _d, _f, C.c_int, _f, _f, C.c_int, _f, _i, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                                _d, _d, _d, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    'nlos_streamed_render_gradient_w_albedo': (C.c_int, [_ctx, _d, _d, _f, C.c_int, _f, _f, C.c_int, _f, _i, C.c_int, C.c_int, C.c_float, C.c_float,
                                                         C.c_float, _d, _d, _d, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    'nlos_streamed_render_gradient_albedo': (C.c_int, [_ctx, _d, _d, _f, C.c_int, _f, _f, C.c_int, _f, _i, C.c_int, C.c_int, C.c_float, C.c_float,
                                                       C.c_float, _d, _d, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _d]),
    'nlos_streamed_render_vertex_gradient': (C.c_int, [_ctx, C.c_int, _f, C.c_int, _f, _f, C.c_int, _i, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                                       _d, C.c_int, C.c_int, C.c_int]),
    'nlos_streamed_render_normal_smoothing': (C.c_int, [_ctx, _f, C.c_int, _i, C.c_int, _i, _d, _d]),
    'nlos_streamed_render_curvature_grad': (C.c_int, [_ctx, _f, C.c_int, _i, C.c_int, _d]),
    'nlos_ggx_streamed_render_transient': (C.c_int, [_ctx, _f, C.c_int, _f, _f, C.c_int, _f, _f, _i, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float,
                                                     C.c_float, _d, _d, C.c_int, C.c_int, C.c_int]),
    'nlos_ggx_streamed_render_intensity': (C.c_int, [_ctx, _f, C.c_int, _f, _f, C.c_int, _f, _i, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, _d]),
    'nlos_ggx_streamed_render_gradient': (C.c_int, [_ctx, _d, _d, _f, C.c_int, _f, _f, C.c_int, _f, _i, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float,
                                                    C.c_float, _d, _d, _d, C.c_int, C.c_int, C.c_int, C.c_int]),
    'nlos_ggx_streamed_render_gradient_alpha': (C.c_int, [_ctx, _d, _d, _f, C.c_int, _f, _f, C.c_int, _f, _i, C.c_int, C.c_float, C.c_int, C.c_float,
                                                          C.c_float, C.c_float, _d, _d, C.c_int, C.c_int, C.c_int, _d]),
    'nlos_jitter_streamed_render_transient': (C.c_int, [_ctx, _f, C.c_int, _f, _f, C.c_int, _f, _f, _i, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                                        _d, C.c_int, C.c_int, _d, _d, C.c_int]),
    'nlos_jitter_streamed_render_gradient': (C.c_int, [_ctx, _d, _d, _f, C.c_int, _f, _f, C.c_int, _f, _i, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                                       _d, _d, C.c_int, C.c_int, _d, _d, _d, C.c_int, C.c_int]),
    'nlos_sr_streamed_render_transient': (C.c_int, [_ctx, _f, C.c_int, _f, _f, C.c_int, _f, _f, _i, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _d, _d, C.c_int]),
    'nlos_sr_render_transient': (C.c_int, [_ctx, _f, _f, _f, C.c_int, _i, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _d, _d, C.c_int]),
    'nlos_sr_streamed_render_gradient': (C.c_int, [_ctx, _d, _f, C.c_int, _f, _f, C.c_int, _i, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int,
                                                   _d, _d, _d, C.c_int]),
    'nlos_embree3_tbb_line_intersection': (C.c_int, [_ctx, _f, _f, C.c_int, _f, C.c_int, _i, C.c_int, _f]),
    'nlos_embree3_tbb_short_line_intersection': (C.c_int, [_ctx, _f, _f, C.c_int, _f, C.c_int, _i, C.c_int, _f]),
    'nlos_barycentric_to_world': (C.c_int, [_ctx, _f, C.c_int, _i, C.c_int, _f, C.c_int, _f]),
    'nlos_debug_copy_visibility_words': (C.c_int, [_ctx, C.POINTER(C.c_uint32), C.c_int64, C.POINTER(C.c_int64)]),
    'nlos_debug_visibility': (C.c_int, [_ctx, _f, C.c_int, _f, C.c_int, _i, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]),
}


def load_library(path=None):
    """dlopen the C ABI and attach argtypes.  Raises NlosError if the library has not been built."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise NlosError('%s not found: build it with `make -C %s` (there is no CPU fallback)' % (p, os.path.join(_HERE, 'csrc')))
        lib = C.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError here = header/library mismatch, fail loudly
            fn.restype = res
            fn.argtypes = args
        if path is None:
            _lib = lib
        return lib


class _LibProxy(object):
    """The library as a Context sees it.  Entry points that received torch CUDA tensors (noted by _arrays.as_pointer in a
    thread-local) are ordered against torch's current stream: the context's stream waits for it before the call, and it waits
    for the context's stream after the call — the device-pointer path of the C ABI runs on the context's own non-blocking
    stream and would otherwise race with the caller's kernels on both sides."""

    def __init__(self, lib, owner):
        self._lib = lib; self._owner = owner; self._cache = {}

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name.startswith('nlos_ctx_') or name in ('nlos_last_error',):
            return fn
        w = self._cache.get(name)
        if w is None:
            from . import _arrays
            lib, owner = self._lib, self._owner

            def w(*args, _fn=fn):
                st = getattr(_arrays.tls, 'cuda_stream', None)
                _arrays.tls.cuda_stream = None
                if st is None:
                    return _fn(*args)
                lib.nlos_ctx_wait_stream(owner.handle, C.c_void_p(st))
                try:
                    return _fn(*args)
                finally:
                    lib.nlos_ctx_signal_stream(owner.handle, C.c_void_p(st))
            self._cache[name] = w
        return w


class Context(object):
    """Owns one nlos_ctx (stream + scratch buffers) on one GPU."""

    def __init__(self, device=0):
        self.lib = _LibProxy(load_library(), self)
        h = _ctx()
        rc = self.lib.nlos_ctx_create(int(device), C.byref(h))
        if rc != NLOS_OK:
            msg = self.lib.nlos_last_error(None)
            raise NlosError('nlos_ctx_create(device=%d) failed (%d): %s' % (device, rc, msg.decode() if msg else '?'))
        self.handle = h
        self.device = int(device)

    def check(self, rc, what):
        if rc != NLOS_OK:
            msg = self.lib.nlos_last_error(self.handle)
            raise NlosError('%s failed (%d): %s' % (what, rc, msg.decode() if msg else '?'))

    def set_seed(self, seed):
        self.check(self.lib.nlos_ctx_set_seed(self.handle, int(seed)), 'nlos_ctx_set_seed')

    def set_source_window(self, src_offset, num_sources_global):
        self.check(self.lib.nlos_ctx_set_source_window(self.handle, int(src_offset), int(num_sources_global)), 'nlos_ctx_set_source_window')

    def set_external_samples(self, stream):
        """TEST HOOK: draw the (S,T) pairs from `stream` (float32 array, or None to restore the generator); see include/nlos_b200.h."""
        if stream is None:
            self.check(self.lib.nlos_ctx_set_external_samples(self.handle, None, 0), 'nlos_ctx_set_external_samples')
            return
        import numpy as np
        a = np.ascontiguousarray(stream, dtype=np.float32)
        self.check(self.lib.nlos_ctx_set_external_samples(self.handle, a.ctypes.data_as(_f), a.size), 'nlos_ctx_set_external_samples')

    def set_option(self, key, value):
        self.check(self.lib.nlos_ctx_set_option(self.handle, key.encode(), int(value)), 'nlos_ctx_set_option(%s)' % key)

    def synchronize(self):
        self.check(self.lib.nlos_ctx_synchronize(self.handle), 'nlos_ctx_synchronize')

    def timing(self):
        buf = (C.c_float * 5)()
        self.check(self.lib.nlos_ctx_get_timing(self.handle, buf), 'nlos_ctx_get_timing')
        return dict(zip(('build_ms', 'forward_ms', 'residual_ms', 'gradient_ms', 'total_ms'), [float(x) for x in buf]))

    def work_counters(self):
        """MEASUREMENT: work counters of the last perspective-grid forward launch run with option count_work = 1."""
        buf = (C.c_uint64 * 8)()
        self.check(self.lib.nlos_ctx_get_work_counters(self.handle, buf), 'nlos_ctx_get_work_counters')
        keys = ('samples_generated', 'rays_traced', 'entry_words_scanned', 'cell_check_passes', 'visible_samples', 'sources_without_grid', 'grid_res', 'forward_algo')
        return dict(zip(keys, [int(x) for x in buf]))

    def visibility_words(self):
        """DEBUG: the visibility words of the last gradient call (see include/nlos_b200.h)."""
        import numpy as np
        n = C.c_int64(0)
        self.check(self.lib.nlos_debug_copy_visibility_words(self.handle, None, 0, C.byref(n)), 'nlos_debug_copy_visibility_words')
        out = np.zeros(n.value, dtype=np.uint32)
        if n.value:
            self.check(self.lib.nlos_debug_copy_visibility_words(self.handle, out.ctypes.data_as(C.POINTER(C.c_uint32)), n.value, None), 'nlos_debug_copy_visibility_words')
        return out

    def launch_count(self):
        return int(self.lib.nlos_ctx_launch_count(self.handle))

    @property
    def stream(self):
        return self.lib.nlos_ctx_stream(self.handle)

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.nlos_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=None):
    """One lazily created context per device; device defaults to LOCAL_RANK (one process per GPU) or 0."""
    if device is None:
        device = int(os.environ.get('NLOS_B200_DEVICE', os.environ.get('LOCAL_RANK', '0')))
    with _lock:
        ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _default_ctx[device] = ctx
    return ctx
