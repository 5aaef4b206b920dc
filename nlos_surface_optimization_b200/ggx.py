"""Drop-in mirror of the reference's Cython module `ggx` (ggx/ggx.pyx): the `renderer` API with the GGX
roughness `alpha` inserted before `num_sample`, and no `loss_flag` (ggx.pyx:12-143)."""
import ctypes as C
from . import _ffi
from ._arrays import as_pointer
from .renderer import _common, _bins, _data_weight, _gradient

__all__ = ['renderStreamedTransient', 'renderStreamedTransientShading', 'renderStreamedTransientwAlbedo', 'renderStreamedGradient',
           'renderStreamedShadingGradient', 'renderStreamedGradientAlpha', 'renderStreamedTriangleIntensity']


def _transient(origin, normal, vertices, vertexNormal, albedo, faces, alpha, num_sample, lower_bound, upper_bound, resolution, transient,
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
    rc = cx.lib.nlos_ggx_streamed_render_transient(cx.handle, po, L, pn, pv, V, pvn, pva, pf, F, float(alpha), int(num_sample), float(lower_bound),
                                                   float(upper_bound), float(resolution), pt, pp, int(refine_scale), int(sigma_bin), B)
    cx.check(rc, 'nlos_ggx_streamed_render_transient')


def renderStreamedTransient(origin, normal, vertices, faces, alpha, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths,
                            refine_scale, sigma_bin, ctx=None):
    """ggx.pyx:118."""
    _transient(origin, normal, vertices, None, None, faces, alpha, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths,
               refine_scale, sigma_bin, ctx)


def renderStreamedTransientShading(origin, normal, vertices, vertexNormal, faces, alpha, num_sample, lower_bound, upper_bound, resolution,
                                   transient, pathlengths, refine_scale, sigma_bin, ctx=None):
    """ggx.pyx:82."""
    _transient(origin, normal, vertices, vertexNormal, None, faces, alpha, num_sample, lower_bound, upper_bound, resolution, transient,
               pathlengths, refine_scale, sigma_bin, ctx)


def renderStreamedTransientwAlbedo(origin, normal, vertices, albedo, faces, alpha, num_sample, lower_bound, upper_bound, resolution,
                                   transient, pathlengths, refine_scale, sigma_bin, ctx=None):
    """ggx.pyx:100."""
    _transient(origin, normal, vertices, None, albedo, faces, alpha, num_sample, lower_bound, upper_bound, resolution, transient,
               pathlengths, refine_scale, sigma_bin, ctx)


def _grad(origin, normal, vertices, faces, vertexNormal, alpha, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths,
          gradient, data, weight, refine_scale, sigma_bin, testing_flag, ctx):
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pvn = None
    if vertexNormal is not None:
        pvn, svn = as_pointer(vertexNormal, 'f32', 2, 'vertexNormal')
        assert svn[1] == 3, "vertex normal needs to be Vx3"
        assert V == svn[0], "vertex normal needs to be Vx3"
    pt, pp, B = _bins(transient, pathlengths, L, lower_bound, upper_bound, resolution)
    pg = _gradient(gradient, V)
    pd, pw = _data_weight(data, weight, L, B)
    rc = cx.lib.nlos_ggx_streamed_render_gradient(cx.handle, pd, pw, po, L, pn, pv, V, pvn, pf, F, float(alpha), int(num_sample), float(lower_bound),
                                                  float(upper_bound), float(resolution), pt, pp, pg, int(refine_scale), int(sigma_bin),
                                                  int(testing_flag), B)
    cx.check(rc, 'nlos_ggx_streamed_render_gradient')


def renderStreamedGradient(origin, normal, vertices, faces, alpha, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths,
                           gradient, data, weight, refine_scale, sigma_bin, testing_flag, ctx=None):
    """ggx.pyx:37."""
    _grad(origin, normal, vertices, faces, None, alpha, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, gradient,
          data, weight, refine_scale, sigma_bin, testing_flag, ctx)


def renderStreamedShadingGradient(origin, normal, vertices, faces, vertexNormal, alpha, num_sample, lower_bound, upper_bound, resolution,
                                  transient, pathlengths, gradient, data, weight, refine_scale, sigma_bin, testing_flag, ctx=None):
    """ggx.pyx:59."""
    _grad(origin, normal, vertices, faces, vertexNormal, alpha, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths,
          gradient, data, weight, refine_scale, sigma_bin, testing_flag, ctx)


def renderStreamedGradientAlpha(origin, normal, vertices, faces, alpha, num_sample, lower_bound, upper_bound, resolution, transient,
                                pathlengths, data, weight, refine_scale, sigma_bin, ctx=None):
    """ggx.pyx:12 -> streamed_render_gradient_alpha; returns d loss / d alpha (float)."""
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pt, pp, B = _bins(transient, pathlengths, L, lower_bound, upper_bound, resolution)
    pd, pw = _data_weight(data, weight, L, B)
    out = C.c_double(0.0)
    rc = cx.lib.nlos_ggx_streamed_render_gradient_alpha(cx.handle, pd, pw, po, L, pn, pv, V, None, pf, F, float(alpha), int(num_sample),
                                                        float(lower_bound), float(upper_bound), float(resolution), pt, pp, int(refine_scale),
                                                        int(sigma_bin), B, C.byref(out))
    cx.check(rc, 'nlos_ggx_streamed_render_gradient_alpha')
    return out.value


def renderStreamedTriangleIntensity(origin, normal, vertices, faces, alpha, num_sample, lower_bound, upper_bound, intensity, ctx=None):
    """ggx.pyx:134."""
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pi, si = as_pointer(intensity, 'f64', 1, 'intensity')
    assert si[0] == F, "intensity should be (F,)"
    rc = cx.lib.nlos_ggx_streamed_render_intensity(cx.handle, po, L, pn, pv, V, None, pf, F, float(alpha), int(num_sample), float(lower_bound),
                                                   float(upper_bound), pi)
    cx.check(rc, 'nlos_ggx_streamed_render_intensity')
