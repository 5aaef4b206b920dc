"""ctypes binding of the reference's OWN renderer, compiled unmodified from /root/reference into oracle/_ref/ by `make -C oracle ref`
(Embree / TBB tasking / MKL VSL / Boost.Random replaced by the stand-in headers of oracle/ref_shim).

TEST INFRASTRUCTURE ONLY — used by tests/test_reference_pin.py (when oracle/_ref exists) and tools/make_ref_fixtures.py (which writes
tests/golden/ref_*.npz, the committed reference outputs the oracle is pinned against on boxes without /root/reference).
The argument conventions are the reference's (in-place outputs, gradient accumulated), wrapped into functional form.

The reference draws its samples from per-thread Mersenne-Twister engines (rng_sse.cpp) seeded from a default-seeded mt19937
(sampler.cpp:25); with the stand-in's static schedule a run with a fixed OMP thread count is reproducible, but its sample stream is
unrelated to the oracle's counter-based one, so every comparison with it is statistical.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, '_ref')
_LIBS = {}
REFERENCE_ROOT = '/root/reference/transient_rendering_cython'


def available():
    return all(os.path.exists(os.path.join(_REF, 'libref_%s.so' % m)) for m in ('renderer', 'ggx', 'jitter', 'intersector', 'sr'))


def build():
    """Compile the reference modules (only possible where /root/reference exists); returns available()."""
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(['make', '-C', _HERE, '-s', 'ref'])
    return available()


def _lib(mod):
    if mod not in _LIBS:
        l = C.CDLL(os.path.join(_REF, 'libref_%s.so' % mod))
        for name in ('ref_streamed_render_gradient_albedo', 'ref_streamed_render_normal_smoothing', 'ref_ggx_streamed_render_gradient_alpha'):
            if hasattr(l, name):
                getattr(l, name).restype = C.c_double
        _LIBS[mod] = l
    return _LIBS[mod]


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _pf(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _pd(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _bins(lower, upper, resolution):
    import math
    return int(math.ceil(np.float32(np.float32(upper) - np.float32(lower)) / np.float32(resolution)))


def _scene(origin, normal, vertices, faces):
    return _f(origin), _f(normal), _f(vertices), np.ascontiguousarray(faces, dtype=np.int32)


def transient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, refine_scale=1, sigma_bin=1, vertex_normal=None,
              vertex_albedo=None, alpha=None):
    o, n, v, f = _scene(origin, normal, vertices, faces); vn, va = _f(vertex_normal), _f(vertex_albedo)
    L, B = o.shape[0], _bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B)
    if alpha is None:
        _lib('renderer').ref_streamed_render_transient(_pf(o), L, _pf(n), _pf(v), v.shape[0], _pf(vn), _pf(va), _pi(f), f.shape[0], int(num_sample),
                                                       C.c_float(lower), C.c_float(upper), C.c_float(resolution), _pd(T), _pd(pl), int(refine_scale), int(sigma_bin))
    else:
        _lib('ggx').ref_ggx_streamed_render_transient(_pf(o), L, _pf(n), _pf(v), v.shape[0], _pf(vn), _pf(va), _pi(f), f.shape[0], C.c_float(alpha), int(num_sample),
                                                      C.c_float(lower), C.c_float(upper), C.c_float(resolution), _pd(T), _pd(pl), int(refine_scale), int(sigma_bin))
    return T, pl


def intensity(origin, normal, vertices, faces, num_sample, lower, upper, vertex_normal=None, alpha=None):
    o, n, v, f = _scene(origin, normal, vertices, faces); vn = _f(vertex_normal)
    out = np.zeros(f.shape[0])
    if alpha is None:
        _lib('renderer').ref_streamed_render_intensity(_pf(o), o.shape[0], _pf(n), _pf(v), v.shape[0], _pf(vn), _pi(f), f.shape[0], int(num_sample), C.c_float(lower),
                                                       C.c_float(upper), _pd(out))
    else:
        _lib('ggx').ref_ggx_streamed_render_intensity(_pf(o), o.shape[0], _pf(n), _pf(v), v.shape[0], _pf(vn), _pi(f), f.shape[0], C.c_float(alpha), int(num_sample),
                                                      C.c_float(lower), C.c_float(upper), _pd(out))
    return out


def gradient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, data, weight, refine_scale, sigma_bin, testing_flag=1, loss_flag=0,
             vertex_normal=None, vertex_albedo=None, alpha=None):
    """-> (transient, gradient[V,3], pathlengths).  vertex_albedo selects streamed_render_gradient_w_albedo."""
    o, n, v, f = _scene(origin, normal, vertices, faces); vn, va = _f(vertex_normal), _f(vertex_albedo)
    data = np.ascontiguousarray(data, dtype=np.float64); weight = np.ascontiguousarray(weight, dtype=np.float64)
    L, B = o.shape[0], _bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    if alpha is not None:
        _lib('ggx').ref_ggx_streamed_render_gradient(_pd(data), _pd(weight), _pf(o), L, _pf(n), _pf(v), v.shape[0], _pf(vn), _pi(f), f.shape[0], C.c_float(alpha),
                                                     int(num_sample), C.c_float(lower), C.c_float(upper), C.c_float(resolution), _pd(T), _pd(pl), _pd(G),
                                                     int(refine_scale), int(sigma_bin), int(testing_flag))
    elif va is not None:
        _lib('renderer').ref_streamed_render_gradient_w_albedo(_pd(data), _pd(weight), _pf(o), L, _pf(n), _pf(v), v.shape[0], _pf(va), _pi(f), f.shape[0],
                                                               int(num_sample), C.c_float(lower), C.c_float(upper), C.c_float(resolution), _pd(T), _pd(pl), _pd(G),
                                                               int(refine_scale), int(sigma_bin), int(testing_flag), int(loss_flag))
    else:
        _lib('renderer').ref_streamed_render_gradient(_pd(data), _pd(weight), _pf(o), L, _pf(n), _pf(v), v.shape[0], _pf(vn), _pi(f), f.shape[0], int(num_sample),
                                                      C.c_float(lower), C.c_float(upper), C.c_float(resolution), _pd(T), _pd(pl), _pd(G), int(refine_scale),
                                                      int(sigma_bin), int(testing_flag), int(loss_flag))
    return T, G, pl


def gradient_albedo(origin, normal, vertices, faces, num_sample, lower, upper, resolution, data, weight, refine_scale, sigma_bin, vertex_albedo,
                    testing_flag=1, loss_flag=0):
    o, n, v, f = _scene(origin, normal, vertices, faces); va = _f(vertex_albedo)
    data = np.ascontiguousarray(data, dtype=np.float64); weight = np.ascontiguousarray(weight, dtype=np.float64)
    L, B = o.shape[0], _bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B)
    g = _lib('renderer').ref_streamed_render_gradient_albedo(_pd(data), _pd(weight), _pf(o), L, _pf(n), _pf(v), v.shape[0], _pf(va), _pi(f), f.shape[0], int(num_sample),
                                                             C.c_float(lower), C.c_float(upper), C.c_float(resolution), _pd(T), _pd(pl), int(refine_scale),
                                                             int(sigma_bin), int(testing_flag), int(loss_flag))
    return T, float(g)


def gradient_alpha(origin, normal, vertices, faces, num_sample, lower, upper, resolution, data, weight, refine_scale, sigma_bin, alpha, vertex_normal=None):
    o, n, v, f = _scene(origin, normal, vertices, faces); vn = _f(vertex_normal)
    data = np.ascontiguousarray(data, dtype=np.float64); weight = np.ascontiguousarray(weight, dtype=np.float64)
    L, B = o.shape[0], _bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B)
    g = _lib('ggx').ref_ggx_streamed_render_gradient_alpha(_pd(data), _pd(weight), _pf(o), L, _pf(n), _pf(v), v.shape[0], _pf(vn), _pi(f), f.shape[0], C.c_float(alpha),
                                                           int(num_sample), C.c_float(lower), C.c_float(upper), C.c_float(resolution), _pd(T), _pd(pl),
                                                           int(refine_scale), int(sigma_bin))
    return T, float(g)


def vertex_gradient(vertex_num, origin, normal, vertices, faces, num_sample, lower, upper, resolution, refine_scale, sigma_bin):
    o, n, v, f = _scene(origin, normal, vertices, faces)
    B = _bins(lower, upper, resolution)
    G = np.zeros((3, B))
    _lib('renderer').ref_streamed_render_vertex_gradient(int(vertex_num), _pf(o), o.shape[0], _pf(n), _pf(v), v.shape[0], _pi(f), f.shape[0], int(num_sample),
                                                         C.c_float(lower), C.c_float(upper), C.c_float(resolution), _pd(G), int(refine_scale), int(sigma_bin))
    return G


def normal_smoothing(vertices, faces, f_affinity):
    v = _f(vertices); f = np.ascontiguousarray(faces, dtype=np.int32); a = np.ascontiguousarray(f_affinity, dtype=np.int32)
    G = np.zeros((v.shape[0], 3))
    val = _lib('renderer').ref_streamed_render_normal_smoothing(_pf(v), v.shape[0], _pi(f), f.shape[0], _pi(a), _pd(G))
    return float(val), G


def curvature_grad(vertices, faces):
    v = _f(vertices); f = np.ascontiguousarray(faces, dtype=np.int32)
    G = np.zeros((v.shape[0], 3))
    _lib('renderer').ref_streamed_render_curvature_grad(_pf(v), v.shape[0], _pi(f), f.shape[0], _pd(G))
    return G


def jitter_transient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, jitter_weight, jitter_offset, vertex_normal=None, vertex_albedo=None):
    o, n, v, f = _scene(origin, normal, vertices, faces); vn, va = _f(vertex_normal), _f(vertex_albedo)
    jw = np.ascontiguousarray(jitter_weight, dtype=np.float64)
    L, B = o.shape[0], _bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B)
    _lib('jitter').ref_jitter_streamed_render_transient(_pf(o), L, _pf(n), _pf(v), v.shape[0], _pf(vn), _pf(va), _pi(f), f.shape[0], int(num_sample), C.c_float(lower),
                                                        C.c_float(upper), C.c_float(resolution), _pd(jw), int(jitter_offset), int(jw.shape[0]), _pd(T), _pd(pl))
    return T, pl


def jitter_gradient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, jitter_weight, jitter_grad, jitter_offset, data, weight,
                    testing_flag=1, vertex_normal=None):
    o, n, v, f = _scene(origin, normal, vertices, faces); vn = _f(vertex_normal)
    jw = np.ascontiguousarray(jitter_weight, dtype=np.float64); jg = np.ascontiguousarray(jitter_grad, dtype=np.float64)
    data = np.ascontiguousarray(data, dtype=np.float64); weight = np.ascontiguousarray(weight, dtype=np.float64)
    L, B = o.shape[0], _bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    _lib('jitter').ref_jitter_streamed_render_gradient(_pd(data), _pd(weight), _pf(o), L, _pf(n), _pf(v), v.shape[0], _pf(vn), _pi(f), f.shape[0], int(num_sample),
                                                       C.c_float(lower), C.c_float(upper), C.c_float(resolution), _pd(jw), _pd(jg), int(jitter_offset), int(jw.shape[0]),
                                                       _pd(T), _pd(pl), _pd(G), int(testing_flag))
    return T, G, pl


def intersect(origin, direction, vertices, faces, short=False):
    """embree3_tbb_line_intersection -> barycoord[N,3] = (primID,u,v) or (-1,.,.); short -> prim[N] (c_embree_intersector.cpp:19-70)."""
    o = _f(origin); d = _f(direction); v = _f(vertices); f = np.ascontiguousarray(faces, dtype=np.int32)
    out = np.zeros(o.shape[0] if short else (o.shape[0], 3), dtype=np.float32)
    fn = _lib('intersector').ref_embree3_tbb_short_line_intersection if short else _lib('intersector').ref_embree3_tbb_line_intersection
    fn(_pf(o), _pf(d), o.shape[0], _pf(v), v.shape[0], _pi(f), f.shape[0], _pf(out))
    return out


def bary_to_world(vertices, faces, bary):
    v = _f(vertices); f = np.ascontiguousarray(faces, dtype=np.int32); b = _f(bary)
    out = np.zeros((b.shape[0], 3), dtype=np.float32)
    _lib('intersector').ref_barycentric_to_world(_pf(v), _pi(f), _pf(b), b.shape[0], _pf(out))
    return out


def set_threads(n):
    """OpenMP thread count of the reference modules (they share one libgomp).  The reference's render_intensity adds into
    intensity[triangle] from every worker without synchronisation (smoothed_transient/transient_and_gradient.cpp:116, ggx/...:121): a data
    race under TBB and under the stand-in alike, so intensity is pinned with one thread."""
    C.CDLL('libgomp.so.1').omp_set_num_threads(int(n))


# ---- first-generation renderer (stratified_transient_raytracer/)
def sr_transient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, vertex_normal=None, vertex_albedo=None):
    o, n, v, f = _scene(np.reshape(origin, (-1, 3)), np.reshape(normal, (-1, 3)), vertices, faces); vn, va = _f(vertex_normal), _f(vertex_albedo)
    L, B = o.shape[0], _bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B)
    _lib('sr').ref_sr_streamed_render_transient(_pf(o), L, _pf(n), _pf(v), v.shape[0], _pf(vn), _pf(va), _pi(f), f.shape[0], int(num_sample), C.c_float(lower),
                                                C.c_float(upper), C.c_float(resolution), _pd(T), _pd(pl))
    return T, pl


def sr_render_transient(origin, normal, vertices, faces, num_sample, lower, upper, resolution):
    """single-origin render_transient (stratifiedTransientRenderer.cpp:132-218)."""
    o = _f(origin).reshape(3); n = _f(normal).reshape(3); v = _f(vertices); f = np.ascontiguousarray(faces, dtype=np.int32)
    B = _bins(lower, upper, resolution); T = np.zeros(B); pl = np.zeros(B)
    _lib('sr').ref_sr_render_transient(_pf(o), _pf(n), _pf(v), v.shape[0], _pi(f), f.shape[0], int(num_sample), C.c_float(lower), C.c_float(upper),
                                       C.c_float(resolution), _pd(T), _pd(pl))
    return T, pl


def sr_gradient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, w_width, data):
    o, n, v, f = _scene(origin, normal, vertices, faces)
    data = np.ascontiguousarray(data, dtype=np.float64)
    L, B = o.shape[0], _bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B); G = np.zeros((v.shape[0] + 2, 3))       # +2 rows: the reference writes one double past V*3 (index slip)
    _lib('sr').ref_sr_streamed_render_gradient(_pd(data), _pf(o), L, _pf(n), _pf(v), v.shape[0], _pi(f), f.shape[0], int(num_sample), C.c_float(lower),
                                               C.c_float(upper), C.c_float(resolution), int(w_width), _pd(T), _pd(pl), _pd(G))
    return T, G[:v.shape[0]], pl


def sampler_stream(n):
    """First n floats of the stream worker 0 draws in every render call (all reference modules share the sampler sources)."""
    out = np.zeros(int(n), dtype=np.float32)
    _lib('renderer').ref_sampler_stream(int(n), _pf(out))
    return out
