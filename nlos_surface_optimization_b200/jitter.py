"""Mirror of the reference's Cython module `jitter` (jitter/jitter.pyx) — SURVEY.md 8f row N3: the renderer with a
tabulated SPAD-jitter temporal kernel (`jitter_weight`, `jitter_grad`, `jitter_offset`; e.g. the 40-tap kernel of
jitter/jitter_info.mat) instead of the Gaussian.  Only the functions whose C++ side is live in the reference are
provided (jitter/stratifiedStreamed*Renderer.h keeps the other prototypes commented out).

`weight` / `jitter_weight` / `jitter_grad` are 2-D double arrays whose FIRST dimension is the kernel length and whose
data is read contiguously from element [0,0] (jitter.pyx:76, :152), i.e. shape (J, 1).
"""
from . import _ffi
from ._arrays import as_pointer
from .renderer import _common, _bins, _data_weight, _gradient

__all__ = ['renderStreamedTransient', 'renderStreamedTransientShading', 'renderStreamedTransientwAlbedo', 'renderStreamedGradient']


def _kernel(w, name):
    pw, sw = as_pointer(w, 'f64', 2, name)
    return pw, sw[0]


def _transient(origin, normal, vertices, vertexNormal, albedo, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths,
               weight, weight_offset, ctx):
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
    pw, J = _kernel(weight, 'weight')
    rc = cx.lib.nlos_jitter_streamed_render_transient(cx.handle, po, L, pn, pv, V, pvn, pva, pf, F, int(num_sample), float(lower_bound), float(upper_bound),
                                                      float(resolution), pw, int(weight_offset), J, pt, pp, B)
    cx.check(rc, 'nlos_jitter_streamed_render_transient')


def renderStreamedTransient(origin, normal, vertices, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, weight,
                            weight_offset, ctx=None):
    """jitter.pyx:140"""
    _transient(origin, normal, vertices, None, None, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, weight,
               weight_offset, ctx)


def renderStreamedTransientShading(origin, normal, vertices, vertexNormal, faces, num_sample, lower_bound, upper_bound, resolution, transient,
                                   pathlengths, weight, weight_offset, ctx=None):
    """jitter.pyx:104"""
    _transient(origin, normal, vertices, vertexNormal, None, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, weight,
               weight_offset, ctx)


def renderStreamedTransientwAlbedo(origin, normal, vertices, albedo, faces, num_sample, lower_bound, upper_bound, resolution, transient,
                                   pathlengths, weight, weight_offset, ctx=None):
    """jitter.pyx:122"""
    _transient(origin, normal, vertices, None, albedo, faces, num_sample, lower_bound, upper_bound, resolution, transient, pathlengths, weight,
               weight_offset, ctx)


def renderStreamedGradient(origin, normal, vertices, faces, num_sample, lower_bound, upper_bound, resolution, jitter_weight, jitter_grad,
                           jitter_offset, transient, pathlengths, gradient, data, weight, testing_flag, ctx=None):
    """jitter.pyx:59 -> streamed_render_gradient(vertexNormal=NULL, jitter_weight, jitter_grad, jitter_offset, len(jitter_weight))."""
    cx = ctx or _ffi.default_context()
    po, pn, pv, pf, L, V, F = _common(origin, normal, vertices, faces)
    pt, pp, B = _bins(transient, pathlengths, L, lower_bound, upper_bound, resolution)
    pg = _gradient(gradient, V)
    pd, pw = _data_weight(data, weight, L, B)
    pjw, J = _kernel(jitter_weight, 'jitter_weight')
    pjg, Jg = _kernel(jitter_grad, 'jitter_grad')
    assert J == Jg, "jitter_weight and jitter_grad need the same length"
    rc = cx.lib.nlos_jitter_streamed_render_gradient(cx.handle, pd, pw, po, L, pn, pv, V, None, pf, F, int(num_sample), float(lower_bound),
                                                     float(upper_bound), float(resolution), pjw, pjg, int(jitter_offset), J, pt, pp, pg,
                                                     int(testing_flag), B)
    cx.check(rc, 'nlos_jitter_streamed_render_gradient')
