"""nlos_surface_optimization_b200 — B200-native differentiable confocal transient renderer.

Drop-in for ONE path of cmu-ci-lab/nlos_surface_optimization: the `renderer` / `ggx` Cython modules
(forward transient + vertex / albedo / roughness gradients).  See DESIGN.md and INTEGRATION.md.

    from nlos_surface_optimization_b200 import renderer, ggx          # reference-signature modules
    from nlos_surface_optimization_b200 import rendering               # exp_bunny/rendering.py facade
"""
from . import _ffi
from ._ffi import Context, NlosError, default_context

__all__ = ['Context', 'NlosError', 'default_context', 'renderer', 'ggx', 'rendering', 'scenes', 'debug_visibility']
__version__ = '0.1.0'


def __getattr__(name):
    # lazy submodules keep `import nlos_surface_optimization_b200` cheap and free of side effects
    if name in ('renderer', 'ggx', 'rendering', 'scenes', 'dist', 'embree_intersector', 'jitter', 'renderer_sr'):
        import importlib
        return importlib.import_module('.' + name, __name__)
    raise AttributeError(name)


def debug_visibility(origin, vertices, faces, num_sample, ctx=None):
    """Per-sample geometric visibility [L,F,spp] (uint8) and traversal counters, through nlos_debug_visibility."""
    import ctypes as C
    import numpy as np
    from ._arrays import as_pointer
    cx = ctx or default_context()
    po, so = as_pointer(origin, 'f32', 2, 'origin')
    pv, sv = as_pointer(vertices, 'f32', 2, 'vertices')
    pf, sf = as_pointer(faces, 'i32', 2, 'faces')
    L, V, F = so[0], sv[0], sf[0]
    spp = max(1, 1 + (int(num_sample) - 1) // F)
    vis = np.zeros((L, F, spp), dtype=np.uint8)
    cnt = (C.c_uint64 * 3)()
    rc = cx.lib.nlos_debug_visibility(cx.handle, po, L, pv, V, pf, F, int(num_sample), vis.ctypes.data_as(C.POINTER(C.c_uint8)), cnt)
    cx.check(rc, 'nlos_debug_visibility')
    return vis, {'rays': int(cnt[0]), 'box_tests': int(cnt[1]), 'tri_tests': int(cnt[2])}
