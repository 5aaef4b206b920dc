"""GPU tests of the shared perspective grid (k_group_bin + k_forward_group, DESIGN.md "K1s"): one 3-D grid per GROUP of neighbouring
wall points.  It must decide every sample exactly as the BVH kernel and the oracle do — the grid only selects candidate triangles —
for every group size, resolution and slice count, through the coarsening path, for tilted / unnormalised wall normals, for wall points
that are not on a regular grid or not in one plane, and for wall points that do not see the whole mesh in front of them."""
import numpy as np
import pytest
from helpers import LB, UB, RES, TOL_TRANSIENT, TOL_GRADIENT, rel_l2, make_target

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    import nlos_surface_optimization_b200 as nb
    c = nb.Context(0)
    yield c
    c.close()


def _grad(ctx, o, n, v, f, ns, data, weight, algo, side=0, K=0, gres=0, gcap=0, budget=0, refine=10, sigma=1, vn=None):
    from nlos_surface_optimization_b200 import renderer
    opts = dict(forward_algo=algo, group_side=side, grid_slices=K, grid_res=gres, grid_cap=gcap, grid_budget_mb=budget)
    for k_, v_ in opts.items(): ctx.set_option(k_, v_)
    B = data.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    if vn is None:
        renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, refine, sigma, 1, 0, ctx=ctx)
    else:
        renderer.renderStreamedShadingGradient(o, n, v, f, vn, ns, LB, UB, RES, T, pl, G, data, weight, refine, sigma, 0, 0, ctx=ctx)
    words = ctx.visibility_words()
    algo_used = ctx.work_counters()['forward_algo']
    for k_ in opts: ctx.set_option(k_, 0)
    return T, G, words, algo_used


def test_shared_grid_decides_every_sample_like_the_bvh_kernel(ctx):
    """bunny, 16x16 wall: 1.8e7 samples; visibility words bit-identical, outputs equal to FP64 summation order."""
    from nlos_surface_optimization_b200 import scenes
    o, n = scenes.wall_grid(16); v, f = scenes.bunny(); ns = 20000
    rng = np.random.RandomState(0)
    data = rng.rand(o.shape[0], 1200) * 1e-3; weight = np.ones_like(data)
    T1, G1, w1, a1 = _grad(ctx, o, n, v, f, ns, data, weight, 1)
    T3, G3, w3, a3 = _grad(ctx, o, n, v, f, ns, data, weight, 3)
    assert a1 == 1 and a3 == 3
    assert w1.size == o.shape[0] * ((f.shape[0] + 31) // 32) and w1.any()
    assert np.array_equal(w1, w3), 'visibility words differ in %d of %d' % (int((w1 != w3).sum()), w1.size)
    assert rel_l2(T3, T1) <= 1e-12 and rel_l2(G3, G1) <= 1e-10


@pytest.mark.parametrize('side,K,gres,gcap,budget', [(1, 1, 1, 0, 0), (1, 4, 33, 0, 0), (2, 16, 0, 0, 0), (3, 7, 50, 0, 0), (8, 16, 0, 0, 0), (4, 64, 20, 0, 0),
                                                     (4, 16, 256, 0, 0), (4, 16, 64, 300000, 0), (4, 8, 200, 600000, 0), (4, 16, 0, 0, 16)])
def test_group_size_resolution_slices_coarsening_and_batching_do_not_change_the_answer(side, K, gres, gcap, budget, ctx):
    """Any group size / resolution / slice count, the on-device coarsening when the entry budget is exceeded (grid_cap forces it) and
    the split of the groups into several batches (grid_budget_mb forces it) give the same bits."""
    from nlos_surface_optimization_b200 import scenes
    o, n = scenes.wall_grid(7); v, f = scenes.bunny(); ns = 20000
    data = np.zeros((o.shape[0], 1200)); weight = np.ones_like(data)
    T1, G1, w1, _ = _grad(ctx, o, n, v, f, ns, data, weight, 1)
    T3, G3, w3, a3 = _grad(ctx, o, n, v, f, ns, data, weight, 3, side, K, gres, gcap, budget)
    assert a3 == 3
    assert np.array_equal(w1, w3)
    assert rel_l2(T3, T1) <= 1e-12


def test_shared_grid_with_tilted_and_unnormalised_wall_normals(oracle, ctx):
    from nlos_surface_optimization_b200 import scenes
    o, n = scenes.wall_grid(6)
    rng = np.random.RandomState(5)
    n = n + 0.35 * rng.randn(*n.shape).astype(np.float32); n[:, 2] = np.abs(n[:, 2]) + 0.3
    n = np.ascontiguousarray(n * rng.uniform(0.5, 3.0, size=(n.shape[0], 1)).astype(np.float32))      # not unit length (the reference never normalises)
    v, f = scenes.icosphere(4, 0.1, (0.02, -0.03, 0.45), noise=0.03, seed=3); ns = 2 * f.shape[0]
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=1, loss_flag=0)
    T, G, w3, _ = _grad(ctx, o, n, v, f, ns, data, weight, 3)
    T1, G1, w1, _ = _grad(ctx, o, n, v, f, ns, data, weight, 1)
    assert np.array_equal(w1, w3)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT and rel_l2(G, G_ref) <= TOL_GRADIENT


def test_shared_grid_with_scattered_wall_points_off_the_plane(oracle, ctx):
    """Wall points that are neither on a regular grid nor in one plane: the ray lines have per-ray slopes (delta.n != 0), rays beyond
    the group's slope bound take the BVH query."""
    from nlos_surface_optimization_b200 import scenes
    rng = np.random.RandomState(11)
    L = 150
    o = np.ascontiguousarray(np.stack([rng.uniform(-0.25, 0.25, L), rng.uniform(-0.25, 0.25, L), rng.uniform(-0.02, 0.02, L)], axis=1), dtype=np.float32)
    n = np.tile(np.array([0, 0, 1], dtype=np.float32), (L, 1))
    v, f = scenes.icosphere(4, 0.1, (0.0, 0.02, 0.42), noise=0.03, seed=7); ns = 2 * f.shape[0]
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=1, loss_flag=0)
    for side in (2, 4):
        T, G, w3, _ = _grad(ctx, o, n, v, f, ns, data, weight, 3, side)
        T1, G1, w1, _ = _grad(ctx, o, n, v, f, ns, data, weight, 1)
        assert np.array_equal(w1, w3)
        assert rel_l2(T, T_ref) <= TOL_TRANSIENT and rel_l2(G, G_ref) <= TOL_GRADIENT


def test_shared_grid_falls_back_for_groups_that_do_not_see_the_mesh_in_front(oracle, ctx):
    """A mesh that reaches behind the wall plane, and two "wall points" inside the scene: their groups have no grid (per-ray BVH query)."""
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.merge([scenes.quad(0.40, 0.1), scenes.quad(0.25, 0.05, 0.05, 0.0), scenes.icosphere(2, 0.06, (-0.1, 0.05, 0.02))])   # sphere straddles z = 0
    o, n = scenes.wall_grid(5)
    o = np.ascontiguousarray(np.concatenate([o, np.array([[0.0, 0.0, 0.3], [0.05, 0.0, 0.41]], dtype=np.float32)]))
    n = np.ascontiguousarray(np.concatenate([n, np.array([[0, 0, 1], [0, 0, -1]], dtype=np.float32)]))
    ns = 16 * f.shape[0]
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=1, loss_flag=0)
    T, G, w3, _ = _grad(ctx, o, n, v, f, ns, data, weight, 3)
    T1, G1, w1, _ = _grad(ctx, o, n, v, f, ns, data, weight, 1)
    assert np.array_equal(w1, w3)
    assert T_ref.sum() > 0
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT and rel_l2(G, G_ref) <= TOL_GRADIENT


def test_shared_grid_many_samples_per_triangle_and_few_wall_points(oracle, ctx):
    """The optimisation-loop regime (C-arm: small mesh, spp = 18), and a call with a single wall point (the unit of work is a warp, so
    the kernel does not need many wall points to fill the machine)."""
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.armadillo_init(); ns = 20000
    assert 1 + (ns - 1) // f.shape[0] == 18
    for wall in (6, 1):
        o, n = scenes.wall_grid(wall) if wall > 1 else (np.array([[0.01, -0.02, 0.0]], dtype=np.float32), np.array([[0, 0, 1]], dtype=np.float32))
        data, weight = make_target(oracle, o, n, v, f, ns)
        T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=1, loss_flag=0)
        T, G, w3, a3 = _grad(ctx, o, n, v, f, ns, data, weight, 3)
        T1, G1, w1, _ = _grad(ctx, o, n, v, f, ns, data, weight, 1)
        assert a3 == 3
        assert np.array_equal(w1, w3)
        assert rel_l2(T, T_ref) <= TOL_TRANSIENT and rel_l2(G, G_ref) <= TOL_GRADIENT


def test_shared_grid_with_vertex_normals_keeps_every_triangle(oracle, ctx):
    """Shading normals switch the plane-side cull off: the live list of every group is the whole mesh."""
    from nlos_surface_optimization_b200 import scenes
    o, n = scenes.wall_grid(6)
    v, f = scenes.icosphere(3, 0.1, (0.0, 0.0, 0.45), noise=0.02, seed=1); ns = 4 * f.shape[0]
    vn = scenes.vertex_normals(v, f)
    data, weight = make_target(oracle, o, n, v, f, ns)
    T1, G1, w1, _ = _grad(ctx, o, n, v, f, ns, data, weight, 1, vn=vn)
    T3, G3, w3, a3 = _grad(ctx, o, n, v, f, ns, data, weight, 3, vn=vn)
    assert a3 == 3 and np.array_equal(w1, w3)
    assert rel_l2(T3, T1) <= 1e-12 and rel_l2(G3, G1) <= 1e-10
