"""Drop-in mirror of the reference's Cython module `renderer` (smoothed_transient/renderer.pyx).

Same function names, same positional argument order, same in-place semantics (`transient`, `pathlengths`,
`intensity` filled; `gradient` accumulated into), same AssertionError on shape mismatches.  Every call goes
through the C ABI of include/nlos_b200.h into the sm_100a kernels; there is no CPU path.

Additions that do not change reference call sites: arrays may also be torch CUDA tensors (used in place,
no host copies), and an optional keyword `ctx=` selects an explicit _ffi.Context (default: one per device).
"""
from . import _ffi
from ._arrays import as_pointer, num_bins

__all__ = ['renderStreamedTransient', 'renderStreamedTransientShading', 'renderStreamedTransientwAlbedo',
           'renderStreamedGradient', 'renderStreamedShadingGradient', 'renderStreamedGradientWithAlbedo',
           'renderStreamedGradientAlbedo', 'renderStreamedTriangleIntensity', 'renderStreamedVertexGradient', 'renderStreamedNormalSmoothing',
           'renderStreamedCurvatureGradient']


def _common(origin, normal, vertices, faces):
    po, so = as_pointer(origin, 'f32', 2, 'origin')
    pn, sn = as_pointer(normal, 'f32', 2, 'normal')
    pv, sv = as_pointer(vertices, 'f32', 2, 'vertices')
    pf, sf = as_pointer(faces, 'i32', 2, 'faces')
    L = so[0]
    assert so[1] == 3, "origin needs to be Lx3"
    assert sn[0] == L, "normal needs to be Lx3"
    assert sn[1] == 3, "normal needs to be Lx3"
    assert sv[1] == 3, "vertices needs to be Vx3"
    assert sf[1] == 3, "faces needs to be Fx3"
    return po, pn, pv, pf, L, sv[0], sf[0]


def _bins(transient, pathlengths, L, lower_bound, upper_bound, resolution):
    B = num_bins(lower_bound, upper_bound, resolution)
    pt, st = as_pointer(transient, 'f64', 2, 'transient')
    pp, sp = as_pointer(pathlengths, 'f64', 1, 'pathlengths')
    msg = "transient dimension should  be LxB   (B = math.ceil((upper_bound-lower_bound)/resolution))"
    assert st[0] == L, msg
    assert st[1] == B, msg
    assert sp[0] == B, "pathlength dimension should be Bx1 (B = math.ceil((upper_bound-lower_bound)/resolution))"
    return pt, pp, B


def _data_weight(data, weight, L, B):
    pd, sd = as_pointer(data, 'f64', 2, 'data')
    pw, sw = as_pointer(weight, 'f64', 2, 'weight')
    msg = "data transient dimension should  be LxB   (B = math.ceil((upper_bound-lower_bound)/resolution))"
    assert sd[0] == L, msg
    assert sd[1] == B, msg
    assert sw[0] == L, "weighting should be LxB"
    assert sw[1] == B, "weighting should be LxB"
    return pd, pw


def _gradient(gradient, V):
    pg, sg = as_pointer(gradient, 'f64', 2, 'gradient')
    assert sg[0] == V, "gradient dimension should be Vx3"
    assert sg[1] == 3, "gradient dimension should be Vx3"
    return pg


def _transient(origin, normal, vertices, vertexNormal, albedo, faces, num_sample, lower_bound, upper_bound, resolution, transient,
               pathlengths, refine_scale, sigma_bin, ctx):
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pvn = pva = None
    if vertexNormal is not None:
        pvn, svn = as_pointer(vertexNormal, 'f32', 2, 'vertexNormal')
        assert svn[1] == 3, "vertex normal needs to be Vx3"
        assert V == svn[0], "vertex normal needs to be Vx3"
    if albedo is not None:
        pva, sva = as_pointer(albedo, 'f32', 1, 'albedo')
        assert V == sva[0], "albedo nees to be Vx1"
    pt, pp, B = _bins(transient, pathlengths, L, lower_bound, upper_bound, resolution)
    rc = cx.lib.nlos_streamed_render_transient(cx.handle, po, L, pn, pv, V, pvn, pva, pf, F, int(num_sample), float(lower_bound), float(upper_bound),
                                               float(resolution), pt, pp, int(refine_scale), int(sigma_bin), B)
    cx.check(rc, 'nlos_streamed_render_transient')


def renderStreamedTransient(origin, normal, vertices, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths,
                            refine_scale, sigma_bin, ctx=None):
    """renderer.pyx:175 -> streamed_render_transient(vertexNormal=NULL, vertexAlbedo=NULL)."""
    _transient(origin, normal, vertices, None, None, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths,
               refine_scale, sigma_bin, ctx)


def renderStreamedTransientShading(origin, normal, vertices, vertexNormal, faces, num_sample, lower_bound, upper_bound, resolution, transient,
                                   pathlengths, refine_scale, sigma_bin, ctx=None):
    """renderer.pyx:139 -> streamed_render_transient(vertexNormal, NULL)."""
    _transient(origin, normal, vertices, vertexNormal, None, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths,
               refine_scale, sigma_bin, ctx)


def renderStreamedTransientwAlbedo(origin, normal, vertices, albedo, faces, num_sample, lower_bound, upper_bound, resolution, transient,
                                   pathlengths, refine_scale, sigma_bin, ctx=None):
    """renderer.pyx:157 -> streamed_render_transient(NULL, albedo)."""
    _transient(origin, normal, vertices, None, albedo, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths,
               refine_scale, sigma_bin, ctx)


def renderStreamedGradient(origin, normal, vertices, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, gradient,
                           data, weight, refine_scale, sigma_bin, testing_flag, loss_flag, ctx=None):
    """renderer.pyx:94 -> streamed_render_gradient(vertexNormal=NULL)."""
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pt, pp, B = _bins(transient, pathlengths, L, lower_bound, upper_bound, resolution)
    pg = _gradient(gradient, V)
    pd, pw = _data_weight(data, weight, L, B)
    rc = cx.lib.nlos_streamed_render_gradient(cx.handle, pd, pw, po, L, pn, pv, V, None, pf, F, int(num_sample), float(lower_bound), float(upper_bound),
                                              float(resolution), pt, pp, pg, int(refine_scale), int(sigma_bin), int(testing_flag), int(loss_flag), B)
    cx.check(rc, 'nlos_streamed_render_gradient')


def renderStreamedShadingGradient(origin, normal, vertices, faces, vertexNormal, num_sample, lower_bound, upper_bound, resolution, transient,
                                  pathlengths, gradient, data, weight, refine_scale, sigma_bin, testing_flag, loss_flag, ctx=None):
    """renderer.pyx:116 -> streamed_render_gradient(vertexNormal)."""
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pvn, svn = as_pointer(vertexNormal, 'f32', 2, 'vertexNormal')
    assert svn[1] == 3, "vertex normal needs to be Vx3"
    assert V == svn[0], "vertex normal needs to be Vx3"
    pt, pp, B = _bins(transient, pathlengths, L, lower_bound, upper_bound, resolution)
    pg = _gradient(gradient, V)
    pd, pw = _data_weight(data, weight, L, B)
    rc = cx.lib.nlos_streamed_render_gradient(cx.handle, pd, pw, po, L, pn, pv, V, pvn, pf, F, int(num_sample), float(lower_bound), float(upper_bound),
                                              float(resolution), pt, pp, pg, int(refine_scale), int(sigma_bin), int(testing_flag), int(loss_flag), B)
    cx.check(rc, 'nlos_streamed_render_gradient')


def renderStreamedGradientWithAlbedo(origin, normal, vertices, faces, albedo, num_sample, lower_bound, upper_bound, resolution, transient,
                                     pathlengths, gradient, data, weight, refine_scale, sigma_bin, testing_flag, loss_flag, ctx=None):
    """renderer.pyx:55 -> streamed_render_gradient_w_albedo."""
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pva, sva = as_pointer(albedo, 'f32', 1, 'albedo')
    assert sva[0] == V, "albedo needs to be Vx1"
    pt, pp, B = _bins(transient, pathlengths, L, lower_bound, upper_bound, resolution)
    pg = _gradient(gradient, V)
    pd, pw = _data_weight(data, weight, L, B)
    rc = cx.lib.nlos_streamed_render_gradient_w_albedo(cx.handle, pd, pw, po, L, pn, pv, V, pva, pf, F, int(num_sample), float(lower_bound),
                                                       float(upper_bound), float(resolution), pt, pp, pg, int(refine_scale), int(sigma_bin),
                                                       int(testing_flag), int(loss_flag), B)
    cx.check(rc, 'nlos_streamed_render_gradient_w_albedo')


def renderStreamedGradientAlbedo(origin, normal, vertices, faces, albedo, num_sample, lower_bound, upper_bound, resolution, transient,
                                 pathlengths, data, weight, refine_scale, sigma_bin, testing_flag, loss_flag, ctx=None):
    """renderer.pyx:35 -> streamed_render_gradient_albedo; returns d loss / d albedo (float)."""
    import ctypes as C
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pva, sva = as_pointer(albedo, 'f32', 1, 'albedo')
    assert sva[0] == V, "albedo needs to be Vx1"
    pt, pp, B = _bins(transient, pathlengths, L, lower_bound, upper_bound, resolution)
    pd, pw = _data_weight(data, weight, L, B)
    out = C.c_double(0.0)
    rc = cx.lib.nlos_streamed_render_gradient_albedo(cx.handle, pd, pw, po, L, pn, pv, V, pva, pf, F, int(num_sample), float(lower_bound),
                                                     float(upper_bound), float(resolution), pt, pp, int(refine_scale), int(sigma_bin),
                                                     int(testing_flag), int(loss_flag), B, C.byref(out))
    cx.check(rc, 'nlos_streamed_render_gradient_albedo')
    return out.value


def renderStreamedTriangleIntensity(origin, normal, vertices, faces, num_sample, lower_bound, upper_bound, intensity, ctx=None):
    """renderer.pyx:191 -> streamed_render_intensity(vertexNormal=NULL)."""
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pi, si = as_pointer(intensity, 'f64', 1, 'intensity')
    assert si[0] == F, "intensity should be (F,)"
    rc = cx.lib.nlos_streamed_render_intensity(cx.handle, po, L, pn, pv, V, None, pf, F, int(num_sample), float(lower_bound), float(upper_bound), pi)
    cx.check(rc, 'nlos_streamed_render_intensity')


def renderStreamedVertexGradient(origin, normal, vertices, faces, num_sample, lower_bound, upper_bound, resolution, gradient, vertex_num,
                                 refine_scale, sigma_bin, ctx=None):
    """renderer.pyx:78 -> streamed_render_vertex_gradient with measurement = 1 (only origin[0] is used, renderer.pyx:88);
    gradient[B,3] receives the per-time-bin gradient of vertex `vertex_num` (always with the normal-variation term)."""
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    B = num_bins(lower_bound, upper_bound, resolution)
    pg, sg = as_pointer(gradient, 'f64', 2, 'gradient')
    assert sg[0] == B, "gradient dimension should be Vx3"
    assert sg[1] == 3, "gradient dimension should be Vx3"
    rc = cx.lib.nlos_streamed_render_vertex_gradient(cx.handle, int(vertex_num), po, 1, pn, pv, V, pf, F, int(num_sample), float(lower_bound),
                                                     float(upper_bound), float(resolution), pg, int(refine_scale), int(sigma_bin), B)
    cx.check(rc, 'nlos_streamed_render_vertex_gradient')


def renderStreamedNormalSmoothing(vertices, faces, f_affinity, gradient, ctx=None, value_out=None):
    """renderer.pyx:13 -> streamed_render_normal_smoothing; returns the regulariser value, fills gradient[V,3].
    Addition: `value_out` = a 1-element float64 torch CUDA tensor receives the value on the device (no host synchronisation; the
    tensor is returned) — used by the device-resident iteration (loop.DeviceIteration)."""
    import ctypes as C
    cx = ctx or _ffi.default_context()
    pv, sv = as_pointer(vertices, 'f32', 2, 'vertices')
    pf, sf = as_pointer(faces, 'i32', 2, 'faces')
    pa, sa = as_pointer(f_affinity, 'i32', 2, 'f_affinity')
    pg, sg = as_pointer(gradient, 'f64', 2, 'gradient')
    assert sv[1] == 3, "vertices needs to be Vx3"
    assert sf[1] == 3, "faces needs to be Fx3"
    assert sa[1] == 3, "face affinity needs to be Fx3"
    assert sa[0] == sf[0], "face affinity needs to be Fx3"
    assert sg[0] == sv[0], "gradient dimension should be Vx3"
    assert sg[1] == 3, "gradient dimension should be Vx3"
    if value_out is not None:
        po, so = as_pointer(value_out, 'f64', 1, 'value_out')
        assert so[0] == 1, "value_out needs one element"
        rc = cx.lib.nlos_streamed_render_normal_smoothing(cx.handle, pv, sv[0], pf, sf[0], pa, pg, po)
        cx.check(rc, 'nlos_streamed_render_normal_smoothing')
        return value_out
    out = C.c_double(0.0)
    rc = cx.lib.nlos_streamed_render_normal_smoothing(cx.handle, pv, sv[0], pf, sf[0], pa, pg, C.byref(out))
    cx.check(rc, 'nlos_streamed_render_normal_smoothing')
    return out.value


def renderStreamedCurvatureGradient(vertices, faces, gradient, ctx=None):
    """renderer.pyx:26 -> streamed_render_curvature_grad."""
    cx = ctx or _ffi.default_context()
    pv, sv = as_pointer(vertices, 'f32', 2, 'vertices')
    pf, sf = as_pointer(faces, 'i32', 2, 'faces')
    pg, sg = as_pointer(gradient, 'f64', 2, 'gradient')
    assert sv[1] == 3, "vertices needs to be Vx3"
    assert sf[1] == 3, "faces needs to be Fx3"
    assert sg[0] == sv[0], "gradient dimension should be Vx3"
    assert sg[1] == 3, "gradient dimension should be Vx3"
    rc = cx.lib.nlos_streamed_render_curvature_grad(cx.handle, pv, sv[0], pf, sf[0], pg)
    cx.check(rc, 'nlos_streamed_render_curvature_grad')
