"""Mirror of the reference's Cython module `embree_intersector` (embree_intersector/embree_intersector.pyx) on the
renderer's LBVH — SURVEY.md 8f row N2.  Used by the reference's space-carving projection and re-triangulation helpers
(exp_bunny/rendering.py:103-206); with this module they no longer need Embree.

Nearest hit = lexicographic minimum of (t, primitive index) over all float-valid Moeller-Trumbore hits, t in (0, inf),
no back-face culling; `direction` need not be normalised.  Outputs follow the reference: (primID, u, v) as float32 with
p = (1-u-v) v1 + u v2 + v v3, primID = -1 on a miss (u, v then left untouched).
"""
from . import _ffi
from ._arrays import as_pointer

__all__ = ['embree3_tbb_intersection', 'embree3_tbb_short_intersection', 'barycoord_to_world', 'PyMesh']


def _rays(origin, direction):
    po, so = as_pointer(origin, 'f32', 2, 'origin')
    pd, sd = as_pointer(direction, 'f32', 2, 'direction')
    assert so[0] == sd[0], "Origin and Direction need to be Nx3"
    assert so[1] == 3, "Origin needs to be Nx3"
    assert sd[1] == 3, "Direction needs to be Nx3"
    return po, pd, so[0]


def _mesh(v, f):
    pv, sv = as_pointer(v, 'f32', 2, 'v')
    pf, sf = as_pointer(f, 'i32', 2, 'f')
    assert sv[1] == 3, "vertex should be Vx3"
    assert sf[1] == 3, "face should be Fx3"
    return pv, sv[0], pf, sf[0]


def embree3_tbb_intersection(origin, direction, v, f, barycoord, ctx=None):
    """embree_intersector.pyx:92"""
    cx = ctx or _ffi.default_context()
    pv, V, pf, F = _mesh(v, f)
    po, pd, N = _rays(origin, direction)
    pb, sb = as_pointer(barycoord, 'f32', 2, 'barycoord')
    assert sb[0] == N, "barycoord needs to be Nx1 or Nx3"
    assert sb[1] == 3, "barycoord needs to be Nx3"
    cx.check(cx.lib.nlos_embree3_tbb_line_intersection(cx.handle, po, pd, N, pv, V, pf, F, pb), 'nlos_embree3_tbb_line_intersection')


def embree3_tbb_short_intersection(origin, direction, v, f, barycoord, ctx=None):
    """embree_intersector.pyx:81"""
    cx = ctx or _ffi.default_context()
    pv, V, pf, F = _mesh(v, f)
    po, pd, N = _rays(origin, direction)
    pb, sb = as_pointer(barycoord, 'f32', 1, 'barycoord')
    assert sb[0] == N, "barycoord needs to be Nx1"
    cx.check(cx.lib.nlos_embree3_tbb_short_line_intersection(cx.handle, po, pd, N, pv, V, pf, F, pb), 'nlos_embree3_tbb_short_line_intersection')


def barycoord_to_world(v, f, barycoord, intersection_p, ctx=None):
    """embree_intersector.pyx:69"""
    cx = ctx or _ffi.default_context()
    pv, V, pf, F = _mesh(v, f)
    pb, sb = as_pointer(barycoord, 'f32', 2, 'barycoord')
    pp, sp = as_pointer(intersection_p, 'f32', 2, 'intersection_p')
    assert sb[0] == sp[0], "barycoord and intersection_p should be Nx3"
    assert sb[1] == 3, "barycoord should be Nx3"
    assert sp[1] == 3, "intersection_p should be Nx3"
    cx.check(cx.lib.nlos_barycentric_to_world(cx.handle, pv, V, pf, F, pb, sb[0], pp), 'nlos_barycentric_to_world')


class PyMesh(object):
    """embree_intersector.pyx:8-65: a mesh that answers the same queries (the scene is rebuilt per query batch, like the
    free functions; the reference's Mesh class does the same, c_mesh.cpp)."""

    def __init__(self, v, f, ctx=None):
        _mesh(v, f)
        self.v, self.f, self.ctx = v, f, ctx
        self.vn = self.fn = self.face_area = None

    def test(self):
        """embree_intersector.pyx:14 / c_mesh.cpp:50-59: print the vertex and face tables."""
        import numpy as np
        print("vertices")
        for p in np.asarray(self.v):
            print("%f %f %f" % (p[0], p[1], p[2]))
        print("faces")
        for t in np.asarray(self.f):
            print("%d %d %d " % (t[0], t[1], t[2]))

    def embree3_tbb_intersection(self, origin, direction, barycoord):
        embree3_tbb_intersection(origin, direction, self.v, self.f, barycoord, ctx=self.ctx)

    def embree3_tbb_short_intersection(self, origin, direction, barycoord):
        embree3_tbb_short_intersection(origin, direction, self.v, self.f, barycoord, ctx=self.ctx)

    def barycoord_to_world(self, barycoord, intersection_p):
        barycoord_to_world(self.v, self.f, barycoord, intersection_p, ctx=self.ctx)

    def set_vn(self, vn):
        assert vn.shape[1] == 3, "vn needs to be #vertices x 3"
        assert vn.shape[0] == self.v.shape[0], "vn nees to be #vertices x 3"
        self.vn = vn

    def set_fn_and_face_area(self, fn, area):
        assert fn.shape[1] == 3, "fn needs to be #face x 3"
        assert fn.shape[0] == self.f.shape[0], "fn needs to be #face x 3"
        assert area.shape[0] == self.f.shape[0], "barycoord needs to be #face x 1"
        self.fn, self.face_area = fn, area
