"""Facade with the names and return values of the reference's exp_bunny/rendering.py:219-308 — the thin
callers of the renderer boundary that every optimisation loop uses (exp_bunny/test.py:161-170).

The functions that sit directly on the renderer / intersector boundary are mirrored (plus the small NumPy loss helpers the
loops call next to them); the remeshing helpers of the reference file (:32-180, CGAL / El Topo / pyigl) stay in the reference
and are untouched.  `mesh` and
`opt` are the reference's ad-hoc objects (attributes v, f, vn, alpha, albedo, f_affinity / lighting,
lighting_normal, sample_num, max_distance_bin, distance_resolution, bin_refine_resolution, sigma_bin,
testing_flag, loss_flag, alpha_flag, albedo_flag, jitter, normal).
"""
import numpy as np

from . import renderer, ggx

__all__ = ['create_weighting_function', 'inverseShadingRendering', 'inverseRenderingAlpha', 'inverseRenderingAlbedo', 'inverseRendering',
           'removeTriangle', 'forwardRendering', 'renderStreamedNormalSmoothing', 'renderStreamedCurvatureGradient', 'vertex_gradient',
           'space_carving_projection', 'face_normal_and_area', 'evaluate_loss_with_normal_smoothness', 'evaluate_loss_with_curvature']


def _bounds(opt):
    return 0, opt.max_distance_bin * opt.distance_resolution, opt.distance_resolution


def create_weighting_function(data, gamma=1):
    """Per-bin loss weights (role of exp_bunny/rendering.py:208-217): (data / max + 0.1) ** gamma, rescaled to mean 1.
    The two-step rescale (divide by the sum, then multiply by the element count) keeps the reference's rounding."""
    w = np.power(data / np.max(data) + 0.1, gamma)
    w = w / np.sum(w)
    return w * (data.shape[0] * data.shape[1])


def _per_vertex_normal(mesh):
    vn = np.empty(mesh.v.shape, dtype=np.float32, order='C')
    try:
        import cgal_api                                   # the reference's CGAL wrapper, if the user has it
        cgal_api.per_vertex_normal(mesh.v, mesh.f, vn)
    except ImportError:
        from .scenes import vertex_normals                # host stand-in (area-weighted), CGAL is out of scope
        vn[:] = vertex_normals(mesh.v, mesh.f)
    return vn


def inverseShadingRendering(mesh, data, weight, opt):
    """:219-229"""
    mesh.vn = _per_vertex_normal(mesh)
    L = opt.lighting.shape[0]
    transient = np.zeros((L, opt.max_distance_bin), dtype=np.double, order='C')
    pathlengths = np.zeros(opt.max_distance_bin, dtype=np.double, order='C')
    gradient = np.zeros(mesh.v.shape, dtype=np.double, order='C')
    lo, hi, res = _bounds(opt)
    renderer.renderStreamedShadingGradient(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, mesh.vn, opt.sample_num, lo, hi, res, transient,
                                           pathlengths, gradient, data, weight, opt.bin_refine_resolution, opt.sigma_bin, opt.testing_flag,
                                           getattr(opt, 'loss_flag', 0))
    return transient, gradient, pathlengths


def inverseRenderingAlpha(mesh, data, weight, opt):
    """:232-238"""
    L = opt.lighting.shape[0]
    transient = np.zeros((L, opt.max_distance_bin), dtype=np.double, order='C')
    pathlengths = np.zeros(opt.max_distance_bin, dtype=np.double, order='C')
    lo, hi, res = _bounds(opt)
    g = ggx.renderStreamedGradientAlpha(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, mesh.alpha, opt.sample_num, lo, hi, res, transient,
                                        pathlengths, data, weight, opt.bin_refine_resolution, opt.sigma_bin)
    return transient, g


def inverseRenderingAlbedo(mesh, data, weight, opt):
    """:241-250"""
    L = opt.lighting.shape[0]
    transient = np.zeros((L, opt.max_distance_bin), dtype=np.double, order='C')
    pathlengths = np.zeros(opt.max_distance_bin, dtype=np.double, order='C')
    albedo = np.ones(mesh.v.shape[0], dtype=np.float32, order='C') * mesh.albedo
    lo, hi, res = _bounds(opt)
    g = renderer.renderStreamedGradientAlbedo(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, albedo.astype(np.float32), opt.sample_num, lo, hi, res,
                                              transient, pathlengths, data, weight, opt.bin_refine_resolution, opt.sigma_bin, opt.testing_flag,
                                              getattr(opt, 'loss_flag', 0))
    return transient, g


def inverseRendering(mesh, data, weight, opt, ctx=None):
    """:252-269 — THE hot entry point: forward transient + vertex gradient.  (`ctx` is an addition: an explicit _ffi.Context.)"""
    L = opt.lighting.shape[0]
    transient = np.zeros((L, opt.max_distance_bin), dtype=np.double, order='C')
    pathlengths = np.zeros(opt.max_distance_bin, dtype=np.double, order='C')
    gradient = np.zeros(mesh.v.shape, dtype=np.double, order='C')
    lo, hi, res = _bounds(opt)
    if getattr(opt, 'alpha_flag', False):
        ggx.renderStreamedGradient(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, mesh.alpha, opt.sample_num, lo, hi, res, transient, pathlengths,
                                   gradient, data, weight, opt.bin_refine_resolution, opt.sigma_bin, opt.testing_flag)
    elif getattr(opt, 'jitter', False):
        from . import jitter
        jitter.renderStreamedGradient(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, opt.sample_num, lo, hi, res, opt.jitter_weight,
                                      opt.jitter_grad, opt.jitter_offset, transient, pathlengths, gradient, data, weight, opt.testing_flag)
    elif getattr(opt, 'albedo_flag', False):
        albedo = (np.ones(mesh.v.shape[0], dtype=np.float32, order='C') * mesh.albedo).astype(np.float32)
        renderer.renderStreamedGradientWithAlbedo(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, albedo, opt.sample_num, lo, hi, res, transient,
                                                  pathlengths, gradient, data, weight, opt.bin_refine_resolution, opt.sigma_bin, opt.testing_flag,
                                                  getattr(opt, 'loss_flag', 0))
    else:
        renderer.renderStreamedGradient(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, opt.sample_num, lo, hi, res, transient, pathlengths,
                                        gradient, data, weight, opt.bin_refine_resolution, opt.sigma_bin, opt.testing_flag, getattr(opt, 'loss_flag', 0), ctx=ctx)
    return transient, gradient, pathlengths


def removeTriangle(mesh, opt, ctx=None):
    """:271-278 — drops faces that no wall point ever sees (unless all three neighbours exist)."""
    intensity = np.zeros(mesh.f.shape[0], dtype=np.double, order='C')
    lo, hi, _ = _bounds(opt)
    renderer.renderStreamedTriangleIntensity(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, opt.sample_num, lo, hi, intensity, ctx=ctx)
    threshold = 0
    keep_face = np.logical_or((intensity > threshold), np.sum(mesh.f_affinity < 0, axis=1) == 0)
    print('remove #face:%d' % (mesh.f.shape[0] - np.sum(keep_face)))
    mesh.f = np.ascontiguousarray(mesh.f[keep_face, :])


def forwardRendering(mesh, opt):
    """:280-297 (the reference's *Shading branches omit refine_scale/sigma_bin — stale; (1,1) is passed here)."""
    L = opt.lighting.shape[0]
    transient = np.zeros((L, opt.max_distance_bin), dtype=np.double, order='C')
    pathlengths = np.zeros(opt.max_distance_bin, dtype=np.double, order='C')
    lo, hi, res = _bounds(opt)
    fn = getattr(opt, 'normal', 'fn') == 'fn'
    if getattr(opt, 'alpha_flag', False):
        if fn:
            ggx.renderStreamedTransient(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, mesh.alpha, opt.sample_num, lo, hi, res, transient, pathlengths, 1, 1)
        else:
            ggx.renderStreamedTransientShading(opt.lighting, opt.lighting_normal, mesh.v, mesh.vn, mesh.f, mesh.alpha, opt.sample_num, lo, hi, res, transient, pathlengths, 1, 1)
    else:
        if fn:
            renderer.renderStreamedTransient(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, opt.sample_num, lo, hi, res, transient, pathlengths, 1, 1)
        else:
            renderer.renderStreamedTransientShading(opt.lighting, opt.lighting_normal, mesh.v, mesh.vn, mesh.f, opt.sample_num, lo, hi, res, transient, pathlengths, 1, 1)
    return transient, pathlengths


def renderStreamedNormalSmoothing(mesh, ctx=None):
    """:299-302"""
    gradient = np.zeros(mesh.v.shape, dtype=np.double, order='C')
    val = renderer.renderStreamedNormalSmoothing(mesh.v, mesh.f, mesh.f_affinity, gradient, ctx=ctx)
    return val, gradient


def renderStreamedCurvatureGradient(mesh):
    """:304-307"""
    gradient = np.zeros(mesh.v.shape, dtype=np.double, order='C')
    renderer.renderStreamedCurvatureGradient(mesh.v, mesh.f, gradient)
    return gradient


def evaluate_loss_with_normal_smoothness(gt_transient, weight, transient, smoothing_val, mesh, render_opt):
    """:360-367 (pure NumPy)."""
    difference = transient - gt_transient
    difference *= np.sqrt(weight)
    L1 = np.linalg.norm(difference) ** 2 / difference.shape[0]
    L2 = render_opt.smooth_weight * smoothing_val
    return L1 + L2, L1


def vertex_gradient(mesh, vertex_num, opt):
    """:26-30 per-bin gradient of one vertex (debug / figure helper)."""
    gradient = np.zeros((opt.max_distance_bin, 3), dtype=np.double, order='C')
    lo, hi, res = _bounds(opt)
    renderer.renderStreamedVertexGradient(opt.lighting, opt.lighting_normal, mesh.v, mesh.f, opt.sample_num, lo, hi, res, gradient, vertex_num,
                                          opt.bin_refine_resolution, opt.sigma_bin)
    return gradient


def space_carving_projection(v, space_carving_mesh):
    """:193-206 push vertices that lie in front of the space-carving hull back onto it: one +z ray per vertex from the wall plane
    against the hull (embree_intersector on the LBVH), in place on v[:, 2]."""
    from . import embree_intersector
    direction = np.tile(np.array([0, 0, 1], dtype=np.float32), (v.shape[0], 1))
    barycoord = np.zeros((v.shape[0], 3), dtype=np.float32, order='C')
    foot = np.array(v, dtype=np.float32, order='C')
    foot[:, 2] = 0
    embree_intersector.embree3_tbb_intersection(foot, direction, space_carving_mesh.v, space_carving_mesh.f, barycoord)
    intersection_p = np.zeros((v.shape[0], 3), dtype=np.float32, order='C')
    embree_intersector.barycoord_to_world(space_carving_mesh.v, space_carving_mesh.f, barycoord, intersection_p)
    hit = barycoord[:, 0] >= 0
    v[hit, 2] = np.maximum(intersection_p[hit, 2], v[hit, 2])


def face_normal_and_area(v, f):
    """:310-318 (pure NumPy; the epsilon keeps degenerate faces finite, as in the reference)."""
    import sys
    n = np.cross(v[f[:, 1], :] - v[f[:, 0], :], v[f[:, 2], :] - v[f[:, 0], :], axis=1)
    n = n + sys.float_info.epsilon
    d = np.linalg.norm(n, axis=1)
    return n / d[:, None], d / 2


def evaluate_loss_with_curvature(gt_transient, weight, transient, mesh, render_opt):
    """:369-380 (pure NumPy): weighted L2 + smooth_weight * total surface area."""
    difference = (transient - gt_transient) * np.sqrt(weight)
    L1 = np.linalg.norm(difference) ** 2 / difference.shape[0]
    total_area = float(np.sum(face_normal_and_area(mesh.v, mesh.f)[1]))
    return L1 + render_opt.smooth_weight * total_area, L1, total_area
