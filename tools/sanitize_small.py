"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel family once."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import renderer, ggx, scenes
ctx = nb.Context(0)
o, n = scenes.wall_grid(3)
v, f = scenes.icosphere(2, 0.1, (0.01, -0.02, 0.45), noise=0.03, seed=1)
vn = scenes.vertex_normals(v, f); alb = np.ones(v.shape[0], np.float32)
ns, lb, ub, res, B = 3 * f.shape[0], 0.0, 1.44, 1.2e-3, 1200
L, V = o.shape[0], v.shape[0]
T = np.zeros((L, B)); pl = np.zeros(B); G = np.zeros((V, 3)); data = np.zeros((L, B)); w = np.ones((L, B))
renderer.renderStreamedTransient(o, n, v, f, ns, lb, ub, res, data, pl, 1, 1, ctx=ctx)
renderer.renderStreamedTransient(o, n, v, f, ns, lb, ub, res, T, pl, 10, 1, ctx=ctx)
renderer.renderStreamedGradient(o, n, v, f, ns, lb, ub, res, T, pl, G, data, w, 10, 1, 1, 0, ctx=ctx)
renderer.renderStreamedShadingGradient(o, n, v, f, vn, ns, lb, ub, res, T, pl, G, data, w, 10, 1, 0, 0, ctx=ctx)
renderer.renderStreamedGradientWithAlbedo(o, n, v, f, alb, ns, lb, ub, res, T, pl, G, data, w, 10, 1, 1, 0, ctx=ctx)
renderer.renderStreamedGradientAlbedo(o, n, v, f, alb, ns, lb, ub, res, T, pl, data, w, 10, 1, 1, 0, ctx=ctx)
ggx.renderStreamedGradient(o, n, v, f, 0.3, ns, lb, ub, res, T, pl, G, data, w, 10, 1, 1, ctx=ctx)
ggx.renderStreamedGradientAlpha(o, n, v, f, 0.3, ns, lb, ub, res, T, pl, data, w, 10, 1, ctx=ctx)
I = np.zeros(f.shape[0]); renderer.renderStreamedTriangleIntensity(o, n, v, f, ns, lb, ub, I, ctx=ctx)
Gb = np.zeros((B, 3)); renderer.renderStreamedVertexGradient(o, n, v, f, ns, lb, ub, res, Gb, 5, 10, 1, ctx=ctx)
renderer.renderStreamedNormalSmoothing(v, f, scenes.face_affinity(f), G, ctx=ctx); renderer.renderStreamedCurvatureGradient(v, f, G, ctx=ctx)
ctx.set_option('reuse_visibility', 0)
renderer.renderStreamedGradient(o, n, v, f, ns, lb, ub, res, T, pl, G, data, w, 10, 1, 1, 0, ctx=ctx)
vis, cnt = nb.debug_visibility(o, v, f, ns, ctx=ctx)
ctx.set_option('reuse_visibility', 1)
from nlos_surface_optimization_b200 import jitter, embree_intersector, renderer_sr
jw = np.ascontiguousarray((np.exp(-0.5 * ((np.arange(9) - 3) / 2.0) ** 2) / 5).reshape(-1, 1)); jg = np.ascontiguousarray(np.gradient(jw[:, 0]).reshape(-1, 1))
jitter.renderStreamedTransient(o, n, v, f, ns, lb, ub, res, T, pl, jw, 3, ctx=ctx)
jitter.renderStreamedGradient(o, n, v, f, ns, lb, ub, res, jw, jg, 3, T, pl, G, data, w, 1, ctx=ctx)
renderer_sr.renderStreamedTransient(o, n, v, f, ns, lb, ub, res, T, pl, ctx=ctx)
t1 = np.zeros(B); renderer_sr.renderTransient(o[0], n[0], v, f, ns, lb, ub, res, t1, pl, ctx=ctx)
renderer_sr.renderStreamedGradient(o, n, v, f, ns, lb, ub, res, 3, T, pl, G, data, ctx=ctx)
ro = np.tile(o, (8, 1)).astype(np.float32); rd = np.tile(np.array([[0.02, -0.04, 1.0]], np.float32), (ro.shape[0], 1))
bc = np.zeros((ro.shape[0], 3), np.float32); embree_intersector.embree3_tbb_intersection(ro, rd, v, f, bc)
pw = np.zeros((ro.shape[0], 3), np.float32); embree_intersector.barycoord_to_world(v, f, bc, pw)
# the two grid forward kernels (the auto choice above is the BVH kernel: 9 wall points < SMs): per-point grid with 4, 2, 1 depth slices and
# through its coarsening path; shared grid (wall grouping, binning kernel with shared-memory and with global counters, live lists, batches)
o2, n2 = scenes.wall_grid(5); L2 = o2.shape[0]
T2 = np.zeros((L2, B)); d2 = np.zeros((L2, B)); w2 = np.ones((L2, B))
for algo, opts in ((2, {}), (2, {'grid_slices': 2}), (2, {'grid_slices': 1}), (2, {'grid_res': 40, 'grid_cap': 700}),
                   (3, {}), (3, {'group_side': 2, 'grid_slices': 4}), (3, {'grid_res': 200, 'grid_slices': 16}), (3, {'grid_budget_mb': 1}), (3, {'grid_res': 40, 'grid_cap': 2000})):
    ctx.set_option('forward_algo', algo)
    for k_, v_ in opts.items(): ctx.set_option(k_, v_)
    renderer.renderStreamedGradient(o2, n2, v, f, ns, lb, ub, res, T2, pl, G, d2, w2, 10, 1, 1, 0, ctx=ctx)
    renderer.renderStreamedTransient(o2, n2, v, f, ns, lb, ub, res, T2, pl, 10, 1, ctx=ctx)
    ggx.renderStreamedGradient(o2, n2, v, f, 0.3, ns, lb, ub, res, T2, pl, G, d2, w2, 10, 1, 1, ctx=ctx)
    for k_ in opts: ctx.set_option(k_, 0)
ctx.set_option('forward_algo', 0)
print('sanitize run ok', T.sum(), np.abs(G).sum(), vis.mean(), T2.sum())
