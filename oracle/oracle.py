"""ctypes binding of the CPU oracle (oracle/nlos_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.  PARITY: pinned to the reference's own code
(oracle/_ref, tests/test_reference_pin.py: same samples -> same numbers to float rounding): see the header of nlos_oracle.cpp.
"""
import ctypes as C
import math
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
DEFAULT_SEED = 5489        # boost::mt19937 default seed, the reference's built-in fixed seed (sampler.cpp:25)


def build(force=False):
    so = os.path.join(_HERE, 'libnlos_oracle.so')
    src = os.path.join(_HERE, 'nlos_oracle.cpp')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.nlos_oracle_normal_smoothing.restype = C.c_double
        _LIB.nlos_oracle_ggx.restype = C.c_float
    return _LIB


class Stats(C.Structure):
    _fields_ = [('rays', C.c_uint64), ('box_tests', C.c_uint64), ('tri_tests', C.c_uint64)]


def _f32(a, shape1=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def num_bins(lower, upper, resolution):
    """ceil((ub-lb)/res) in float32, as both sides of the reference boundary compute it (renderer.pyx:101, SSG.cpp:514)."""
    return int(math.ceil(np.float32(np.float32(upper) - np.float32(lower)) / np.float32(resolution)))


def transient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, refine_scale=1, sigma_bin=1,
              vertex_normal=None, vertex_albedo=None, alpha=-1.0, seed=DEFAULT_SEED, src_offset=0, brute=False,
              want_visibility=False, want_stats=False):
    origin = _f32(origin); normal = _f32(normal); vertices = _f32(vertices)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    vn = None if vertex_normal is None else _f32(vertex_normal)
    va = None if vertex_albedo is None else _f32(vertex_albedo)
    L, V, F = origin.shape[0], vertices.shape[0], faces.shape[0]
    B = num_bins(lower, upper, resolution)
    T = np.zeros((L, B), dtype=np.float64)
    pl = np.zeros(B, dtype=np.float64)
    spp = 1 + (num_sample - 1) // F
    vis = np.zeros((L, F, spp), dtype=np.uint8) if want_visibility else None
    st = Stats()
    rc = lib().nlos_oracle_transient(_p(origin, C.c_float), C.c_int64(L), _p(normal, C.c_float), _p(vertices, C.c_float), C.c_int(V),
                                     _p(vn, C.c_float), _p(va, C.c_float), _p(faces, C.c_int32), C.c_int(F), C.c_float(alpha),
                                     C.c_int(num_sample), C.c_float(lower), C.c_float(upper), C.c_float(resolution), C.c_int(B),
                                     _p(T, C.c_double), _p(pl, C.c_double), C.c_int(refine_scale), C.c_int(sigma_bin),
                                     C.c_uint64(seed), C.c_int64(src_offset), C.c_int(1 if brute else 0),
                                     _p(vis, C.c_uint8), C.byref(st) if want_stats else None)
    assert rc == 0
    out = [T, pl]
    if want_visibility:
        out.append(vis)
    if want_stats:
        out.append({'rays': st.rays, 'box_tests': st.box_tests, 'tri_tests': st.tri_tests})
    return tuple(out)


def gradient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, data, weight, refine_scale, sigma_bin,
             testing_flag=1, loss_flag=0, vertex_normal=None, vertex_albedo=None, alpha=-1.0, seed=DEFAULT_SEED, src_offset=0,
             brute=False, kind=0, gradient_inout=None):
    """kind 0: vertex gradient -> (transient, gradient[V,3], pathlengths); 1: albedo scalar; 2: alpha scalar -> (transient, g)."""
    origin = _f32(origin); normal = _f32(normal); vertices = _f32(vertices)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    vn = None if vertex_normal is None else _f32(vertex_normal)
    va = None if vertex_albedo is None else _f32(vertex_albedo)
    data = np.ascontiguousarray(data, dtype=np.float64); weight = np.ascontiguousarray(weight, dtype=np.float64)
    L, V, F = origin.shape[0], vertices.shape[0], faces.shape[0]
    B = num_bins(lower, upper, resolution)
    assert data.shape == (L, B) and weight.shape == (L, B)
    T = np.zeros((L, B), dtype=np.float64); pl = np.zeros(B, dtype=np.float64)
    G = np.zeros((V, 3), dtype=np.float64) if gradient_inout is None else gradient_inout
    scalar = C.c_double(0.0)
    rc = lib().nlos_oracle_gradient(_p(data, C.c_double), _p(weight, C.c_double), _p(origin, C.c_float), C.c_int64(L), _p(normal, C.c_float),
                                    _p(vertices, C.c_float), C.c_int(V), _p(vn, C.c_float), _p(va, C.c_float), _p(faces, C.c_int32), C.c_int(F),
                                    C.c_float(alpha), C.c_int(num_sample), C.c_float(lower), C.c_float(upper), C.c_float(resolution), C.c_int(B),
                                    _p(T, C.c_double), _p(pl, C.c_double), _p(G, C.c_double), C.c_int(refine_scale), C.c_int(sigma_bin),
                                    C.c_int(testing_flag), C.c_int(loss_flag), C.c_uint64(seed), C.c_int64(src_offset), C.c_int(1 if brute else 0),
                                    C.c_int(kind), C.byref(scalar))
    assert rc == 0
    if kind == 0:
        return T, G, pl
    return T, scalar.value


def intensity(origin, normal, vertices, faces, num_sample, lower, upper, alpha=-1.0, vertex_normal=None, seed=DEFAULT_SEED,
              src_offset=0, brute=False):
    origin = _f32(origin); normal = _f32(normal); vertices = _f32(vertices)
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    vn = None if vertex_normal is None else _f32(vertex_normal)
    L, V, F = origin.shape[0], vertices.shape[0], faces.shape[0]
    out = np.zeros(F, dtype=np.float64)
    rc = lib().nlos_oracle_intensity(_p(origin, C.c_float), C.c_int64(L), _p(normal, C.c_float), _p(vertices, C.c_float), C.c_int(V),
                                     _p(vn, C.c_float), _p(faces, C.c_int32), C.c_int(F), C.c_float(alpha), C.c_int(num_sample),
                                     C.c_float(lower), C.c_float(upper), _p(out, C.c_double), C.c_uint64(seed), C.c_int64(src_offset),
                                     C.c_int(1 if brute else 0))
    assert rc == 0
    return out


def vertex_gradient(vertex_num, origin, normal, vertices, faces, num_sample, lower, upper, resolution, refine_scale, sigma_bin,
                    seed=DEFAULT_SEED, brute=False):
    origin = _f32(origin)[:1]; normal = _f32(normal)[:1]; vertices = _f32(vertices)     # renderer.pyx:88 hard-codes measurement=1
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    V, F = vertices.shape[0], faces.shape[0]
    B = num_bins(lower, upper, resolution)
    G = np.zeros((B, 3), dtype=np.float64)
    rc = lib().nlos_oracle_vertex_gradient(C.c_int(vertex_num), _p(origin, C.c_float), C.c_int64(1), _p(normal, C.c_float), _p(vertices, C.c_float),
                                           C.c_int(V), _p(faces, C.c_int32), C.c_int(F), C.c_int(num_sample), C.c_float(lower), C.c_float(upper),
                                           C.c_float(resolution), C.c_int(B), _p(G, C.c_double), C.c_int(refine_scale), C.c_int(sigma_bin),
                                           C.c_uint64(seed), C.c_int(1 if brute else 0))
    assert rc == 0
    return G


def normal_smoothing(vertices, faces, f_affinity):
    vertices = _f32(vertices); faces = np.ascontiguousarray(faces, dtype=np.int32); aff = np.ascontiguousarray(f_affinity, dtype=np.int32)
    G = np.zeros((vertices.shape[0], 3), dtype=np.float64)
    val = lib().nlos_oracle_normal_smoothing(_p(vertices, C.c_float), C.c_int(vertices.shape[0]), _p(faces, C.c_int32), C.c_int(faces.shape[0]),
                                             _p(aff, C.c_int32), _p(G, C.c_double))
    return val, G


def curvature_grad(vertices, faces):
    vertices = _f32(vertices); faces = np.ascontiguousarray(faces, dtype=np.int32)
    G = np.zeros((vertices.shape[0], 3), dtype=np.float64)
    lib().nlos_oracle_curvature_grad(_p(vertices, C.c_float), C.c_int(vertices.shape[0]), _p(faces, C.c_int32), C.c_int(faces.shape[0]), _p(G, C.c_double))
    return G


def philox_st(seed, src, tri, k):
    S = C.c_float(); T = C.c_float()
    lib().nlos_oracle_philox(C.c_uint64(seed), C.c_int64(src), C.c_int(tri), C.c_int(k), C.byref(S), C.byref(T))
    return S.value, T.value


def isect(tri9, o, d):
    tri9 = _f32(tri9).reshape(9); o = _f32(o); d = _f32(d); out = np.zeros(3, dtype=np.float32)
    hit = lib().nlos_oracle_isect(_p(tri9, C.c_float), _p(o, C.c_float), _p(d, C.c_float), _p(out, C.c_float))
    return bool(hit), out


def ggx(which, alpha, n, w):
    n = _f32(n); w = _f32(w); dn = np.zeros(3, dtype=np.float32); dw = np.zeros(3, dtype=np.float32)
    val = lib().nlos_oracle_ggx(C.c_int(which), C.c_float(alpha), _p(n, C.c_float), _p(w, C.c_float), _p(dn, C.c_float), _p(dw, C.c_float))
    return (val if which < 2 else (dn, dw))


def taps(res, r, s):
    w = np.zeros(4 * r * s + 1, dtype=np.float64); s2 = C.c_double()
    lib().nlos_oracle_taps(C.c_float(res), C.c_int(r), C.c_int(s), _p(w, C.c_double), C.byref(s2))
    return w, s2.value


def threads():
    return lib().nlos_oracle_threads()


def intersect(origin, direction, vertices, faces, brute=False):
    """nearest hit per ray -> (barycoord[N,3] f32 = (primID,u,v) or (-1,0,0), prim[N] f32)"""
    origin = _f32(origin); direction = _f32(direction); vertices = _f32(vertices); faces = np.ascontiguousarray(faces, dtype=np.int32)
    N = origin.shape[0]
    out3 = np.zeros((N, 3), dtype=np.float32); out1 = np.zeros(N, dtype=np.float32)
    lib().nlos_oracle_intersect(_p(origin, C.c_float), _p(direction, C.c_float), C.c_int64(N), _p(vertices, C.c_float), C.c_int(vertices.shape[0]),
                                _p(faces, C.c_int32), C.c_int(faces.shape[0]), _p(out3, C.c_float), _p(out1, C.c_float), C.c_int(1 if brute else 0))
    return out3, out1


def bary_to_world(vertices, faces, bary):
    vertices = _f32(vertices); faces = np.ascontiguousarray(faces, dtype=np.int32); bary = _f32(bary)
    out = np.zeros((bary.shape[0], 3), dtype=np.float32)
    lib().nlos_oracle_bary_to_world(_p(vertices, C.c_float), _p(faces, C.c_int32), _p(bary, C.c_float), C.c_int64(bary.shape[0]), _p(out, C.c_float))
    return out


def jitter_transient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, jitter_weight, jitter_offset, vertex_normal=None,
                     vertex_albedo=None, seed=DEFAULT_SEED, src_offset=0, brute=False):
    origin = _f32(origin); normal = _f32(normal); vertices = _f32(vertices); faces = np.ascontiguousarray(faces, dtype=np.int32)
    vn = None if vertex_normal is None else _f32(vertex_normal); va = None if vertex_albedo is None else _f32(vertex_albedo)
    jw = np.ascontiguousarray(jitter_weight, dtype=np.float64).reshape(-1)
    L, V, F = origin.shape[0], vertices.shape[0], faces.shape[0]; B = num_bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B)
    lib().nlos_oracle_jitter_transient(_p(origin, C.c_float), C.c_int64(L), _p(normal, C.c_float), _p(vertices, C.c_float), C.c_int(V), _p(vn, C.c_float),
                                       _p(va, C.c_float), _p(faces, C.c_int32), C.c_int(F), C.c_int(num_sample), C.c_float(lower), C.c_float(upper),
                                       C.c_float(resolution), C.c_int(B), _p(jw, C.c_double), C.c_int(jitter_offset), C.c_int(jw.shape[0]),
                                       _p(T, C.c_double), _p(pl, C.c_double), C.c_uint64(seed), C.c_int64(src_offset), C.c_int(1 if brute else 0))
    return T, pl


def jitter_gradient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, jitter_weight, jitter_grad, jitter_offset, data, weight,
                    testing_flag=1, vertex_normal=None, seed=DEFAULT_SEED, src_offset=0, brute=False):
    origin = _f32(origin); normal = _f32(normal); vertices = _f32(vertices); faces = np.ascontiguousarray(faces, dtype=np.int32)
    vn = None if vertex_normal is None else _f32(vertex_normal)
    jw = np.ascontiguousarray(jitter_weight, dtype=np.float64).reshape(-1); jg = np.ascontiguousarray(jitter_grad, dtype=np.float64).reshape(-1)
    data = np.ascontiguousarray(data, dtype=np.float64); weight = np.ascontiguousarray(weight, dtype=np.float64)
    L, V, F = origin.shape[0], vertices.shape[0], faces.shape[0]; B = num_bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B); G = np.zeros((V, 3))
    lib().nlos_oracle_jitter_gradient(_p(data, C.c_double), _p(weight, C.c_double), _p(origin, C.c_float), C.c_int64(L), _p(normal, C.c_float),
                                      _p(vertices, C.c_float), C.c_int(V), _p(vn, C.c_float), _p(faces, C.c_int32), C.c_int(F), C.c_int(num_sample),
                                      C.c_float(lower), C.c_float(upper), C.c_float(resolution), C.c_int(B), _p(jw, C.c_double), _p(jg, C.c_double),
                                      C.c_int(jitter_offset), C.c_int(jw.shape[0]), _p(T, C.c_double), _p(pl, C.c_double), _p(G, C.c_double),
                                      C.c_int(testing_flag), C.c_uint64(seed), C.c_int64(src_offset), C.c_int(1 if brute else 0))
    return T, G, pl


def sr_transient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, vertex_normal=None, vertex_albedo=None, seed=DEFAULT_SEED,
                 src_offset=0, brute=False):
    """First-generation forward (stratified_transient_raytracer/): raw histogram, form factor NOT clamped."""
    origin = _f32(origin).reshape(-1, 3); normal = _f32(normal).reshape(-1, 3); vertices = _f32(vertices); faces = np.ascontiguousarray(faces, dtype=np.int32)
    vn = None if vertex_normal is None else _f32(vertex_normal); va = None if vertex_albedo is None else _f32(vertex_albedo)
    L, V, F = origin.shape[0], vertices.shape[0], faces.shape[0]; B = num_bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B)
    lib().nlos_oracle_sr_transient(_p(origin, C.c_float), C.c_int64(L), _p(normal, C.c_float), _p(vertices, C.c_float), C.c_int(V), _p(vn, C.c_float),
                                   _p(va, C.c_float), _p(faces, C.c_int32), C.c_int(F), C.c_int(num_sample), C.c_float(lower), C.c_float(upper),
                                   C.c_float(resolution), C.c_int(B), _p(T, C.c_double), _p(pl, C.c_double), C.c_uint64(seed), C.c_int64(src_offset),
                                   C.c_int(1 if brute else 0))
    return T, pl


def sr_gradient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, w_width, data, seed=DEFAULT_SEED, src_offset=0, brute=False,
                typos=False):
    """First-generation gradient: box-filtered residual, one tap, gn always.  typos=True reproduces the reference's index slips."""
    origin = _f32(origin); normal = _f32(normal); vertices = _f32(vertices); faces = np.ascontiguousarray(faces, dtype=np.int32)
    data = np.ascontiguousarray(data, dtype=np.float64)
    L, V, F = origin.shape[0], vertices.shape[0], faces.shape[0]; B = num_bins(lower, upper, resolution)
    T = np.zeros((L, B)); pl = np.zeros(B); G = np.zeros((V, 3))
    lib().nlos_oracle_sr_gradient(_p(data, C.c_double), _p(origin, C.c_float), C.c_int64(L), _p(normal, C.c_float), _p(vertices, C.c_float), C.c_int(V),
                                  _p(faces, C.c_int32), C.c_int(F), C.c_int(num_sample), C.c_float(lower), C.c_float(upper), C.c_float(resolution), C.c_int(B),
                                  C.c_int(w_width), _p(T, C.c_double), _p(pl, C.c_double), _p(G, C.c_double), C.c_uint64(seed), C.c_int64(src_offset),
                                  C.c_int(1 if brute else 0), C.c_int(1 if typos else 0))
    return T, G, pl


_EXT = None


def set_external_samples(stream):
    """Install (or, with None, remove) an external (S,T) sample stream — see g_ext_samples in nlos_oracle.cpp."""
    global _EXT
    if stream is None:
        _EXT = None
        lib().nlos_oracle_set_external_samples(None, C.c_int64(0))
        return
    _EXT = np.ascontiguousarray(stream, dtype=np.float32)
    lib().nlos_oracle_set_external_samples(_p(_EXT, C.c_float), C.c_int64(_EXT.size))


def set_threads(n):
    """OpenMP threads of the oracle (torchrun exports OMP_NUM_THREADS=1; the CPU arm of bench.py wants all host cores)."""
    lib().nlos_oracle_set_threads(int(n))


def canonical_counts(origin, vertices, faces, num_sample, seed=DEFAULT_SEED, src_offset=0):
    """box / triangle tests per path sample of the canonical traversal (Karras LBVH, one triangle per leaf, near child
    first, nearest hit) over every (source, triangle, k) sample of `origin`.  -> dict(rays, box_per_ray, tri_per_ray, visible_frac)."""
    o = _f32(origin); v = _f32(vertices); f = np.ascontiguousarray(faces, dtype=np.int32)
    out = (C.c_uint64 * 4)()
    lib().nlos_oracle_canonical_counts(_p(o, C.c_float), C.c_int64(o.shape[0]), _p(v, C.c_float), C.c_int(v.shape[0]), _p(f, C.c_int32), C.c_int(f.shape[0]),
                                       C.c_int(int(num_sample)), C.c_uint64(int(seed)), C.c_int64(int(src_offset)), out)
    r = max(int(out[0]), 1)
    return {'rays': int(out[0]), 'box_per_ray': out[1] / r, 'tri_per_ray': out[2] / r, 'visible_frac': out[3] / r}
