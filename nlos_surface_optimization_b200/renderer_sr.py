"""Mirror of the reference's FIRST-GENERATION Cython module `renderer` (stratified_transient_raytracer/renderer.pyx) —
SURVEY.md 8f row N4.  It is the module the legacy drivers (transient_rendering_python/, `CY/rendering.py`) import under the
name `renderer`; it clashes with the second-generation module of the same name (smoothed_transient/), so here it is
`renderer_sr` (put it in sys.modules['renderer'] to run a legacy driver).

Differences from the second generation, all kept: no temporal smoothing, the forward form factor is not clamped
(stratifiedStreamedTransientRenderer.cpp:130-137), the gradient smooths the residual with a box filter of half-width `w_width`
applied twice (stratifiedStreamedGradientRenderer.cpp:447-458), always includes the normal-variation term and OVERWRITES
`gradient`.  Not kept: the index slips of the reference's gradient accumulation (:278, :290-291), which make its own output
unusable (its only caller also passes a wrongly shaped array, CY/rendering.py:31).
"""
from . import _ffi
from ._arrays import as_pointer, num_bins
from .renderer import _common, _bins, _gradient, renderStreamedCurvatureGradient  # noqa: F401  (renderer.pyx:13-19 is the same function)

__all__ = ['renderStreamedCurvatureGradient', 'renderStreamedGradient', 'renderStreamedTransientShading', 'renderStreamedTransientwAlbedo',
           'renderStreamedTransient', 'renderTransient']


def _transient(origin, normal, vertices, vertexNormal, albedo, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, ctx):
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
    rc = cx.lib.nlos_sr_streamed_render_transient(cx.handle, po, L, pn, pv, V, pvn, pva, pf, F, int(num_sample), float(lower_bound), float(upper_bound),
                                                  float(resolution), pt, pp, B)
    cx.check(rc, 'nlos_sr_streamed_render_transient')


def renderStreamedTransient(origin, normal, vertices, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, ctx=None):
    """renderer.pyx:76-88"""
    _transient(origin, normal, vertices, None, None, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, ctx)


def renderStreamedTransientShading(origin, normal, vertices, vertexNormal, faces, num_sample, lower_bound, upper_bound, resolution, transient,
                                   pathlengths, ctx=None):
    """renderer.pyx:36-52"""
    _transient(origin, normal, vertices, vertexNormal, None, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, ctx)


def renderStreamedTransientwAlbedo(origin, normal, vertices, albedo, faces, num_sample, lower_bound, upper_bound, resolution, transient,
                                   pathlengths, ctx=None):
    """renderer.pyx:56-71"""
    _transient(origin, normal, vertices, None, albedo, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, ctx)


def renderTransient(origin, normal, vertices, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, ctx=None):
    """renderer.pyx:93-102: one origin (origin[3], normal[3]) -> transient[numBins]"""
    cx = ctx or _ffi.default_context()
    po, so = as_pointer(origin, 'f32', 1, 'origin'); pn, sn = as_pointer(normal, 'f32', 1, 'normal')
    assert so[0] == 3, "origin needs to be 1x3"
    assert sn[0] == 3, "normal needs to be 1x3"
    pv, sv = as_pointer(vertices, 'f32', 2, 'vertices'); pf, sf = as_pointer(faces, 'i32', 2, 'faces')
    assert sv[1] == 3, "vertices needs to be Vx3"
    assert sf[1] == 3, "faces needs to be Fx3"
    B = num_bins(lower_bound, upper_bound, resolution)
    pt, st = as_pointer(transient, 'f64', 1, 'transient'); pp, sp = as_pointer(pathlengths, 'f64', 1, 'pathlengths')
    assert st[0] == B, "transient dimension should match number of bins = math.ceil((upper_bound-lower_bound)/resolution)"
    assert sp[0] == B, "pathlength dimension should match number of bins = math.ceil((upper_bound-lower_bound)/resolution)"
    rc = cx.lib.nlos_sr_render_transient(cx.handle, po, pn, pv, sv[0], pf, sf[0], int(num_sample), float(lower_bound), float(upper_bound), float(resolution),
                                         pt, pp, B)
    cx.check(rc, 'nlos_sr_render_transient')


def renderStreamedGradient(origin, normal, vertices, faces, num_sample, lower_bound, upper_bound, resolution, w_width, transient, pathlengths, gradient,
                           data, ctx=None):
    """renderer.pyx:22-34"""
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pt, pp, B = _bins(transient, pathlengths, L, lower_bound, upper_bound, resolution)
    pg = _gradient(gradient, V)
    pd, sd = as_pointer(data, 'f64', 2, 'data')
    msg = "data transient dimension should  be LxB   (B = math.ceil((upper_bound-lower_bound)/resolution))"
    assert sd[0] == L, msg
    assert sd[1] == B, msg
    rc = cx.lib.nlos_sr_streamed_render_gradient(cx.handle, pd, po, L, pn, pv, V, pf, F, int(num_sample), float(lower_bound), float(upper_bound),
                                                 float(resolution), int(w_width), pt, pp, pg, B)
    cx.check(rc, 'nlos_sr_streamed_render_gradient')
