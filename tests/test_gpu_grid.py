"""GPU tests of the perspective-grid forward kernel (k_forward_grid, DESIGN.md "K1g"): it must decide every sample exactly as the
BVH kernel and the oracle do — the grid only selects candidate triangles — for every grid resolution, through its coarsening path,
for tilted / unnormalised wall normals, and for wall points that do not see the whole mesh in front of them (in-kernel fallback)."""
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


def _grad(ctx, o, n, v, f, ns, data, weight, algo, gres=0, gcap=0, refine=10, sigma=1):
    from nlos_surface_optimization_b200 import renderer
    ctx.set_option('forward_algo', algo); ctx.set_option('grid_res', gres); ctx.set_option('grid_cap', gcap)
    B = data.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, refine, sigma, 1, 0, ctx=ctx)
    words = ctx.visibility_words()
    ctx.set_option('forward_algo', 0); ctx.set_option('grid_res', 0); ctx.set_option('grid_cap', 0)
    return T, G, words


def test_both_forward_kernels_decide_every_sample_identically(ctx):
    """bunny, 16x16 wall: 1.8e7 samples; visibility words bit-identical, outputs equal to FP64 summation order."""
    from nlos_surface_optimization_b200 import scenes
    o, n = scenes.wall_grid(16); v, f = scenes.bunny(); ns = 20000
    rng = np.random.RandomState(0)
    data = rng.rand(o.shape[0], 1200) * 1e-3; weight = np.ones_like(data)
    T1, G1, w1 = _grad(ctx, o, n, v, f, ns, data, weight, 1)
    T2, G2, w2 = _grad(ctx, o, n, v, f, ns, data, weight, 2)
    assert w1.size == o.shape[0] * ((f.shape[0] + 31) // 32) and w1.any()
    assert np.array_equal(w1, w2), 'visibility words differ in %d of %d' % (int((w1 != w2).sum()), w1.size)
    assert rel_l2(T2, T1) <= 1e-12 and rel_l2(G2, G1) <= 1e-10


@pytest.mark.parametrize('gres,gcap', [(1, 0), (7, 0), (33, 0), (256, 0), (64, 200000), (200, 70000)])
def test_grid_resolution_and_coarsening_do_not_change_the_answer(gres, gcap, ctx):
    """Any resolution, and the on-device coarsening when the entry budget is exceeded (grid_cap forces it), give the same bits."""
    from nlos_surface_optimization_b200 import scenes
    o, n = scenes.wall_grid(6); v, f = scenes.bunny(); ns = 20000
    data = np.zeros((o.shape[0], 1200)); weight = np.ones_like(data)
    T1, G1, w1 = _grad(ctx, o, n, v, f, ns, data, weight, 1)
    T2, G2, w2 = _grad(ctx, o, n, v, f, ns, data, weight, 2, gres, gcap)
    assert np.array_equal(w1, w2)
    assert rel_l2(T2, T1) <= 1e-12


@pytest.mark.parametrize('slices', [1, 2, 3])
def test_fewer_depth_slices_do_not_change_the_answer(slices, ctx):
    """Large meshes trade depth slices for picture resolution (the cell counters must fit shared memory): any slice count gives the same bits."""
    from nlos_surface_optimization_b200 import scenes
    o, n = scenes.wall_grid(6); v, f = scenes.bunny(); ns = 20000
    data = np.zeros((o.shape[0], 1200)); weight = np.ones_like(data)
    T1, G1, w1 = _grad(ctx, o, n, v, f, ns, data, weight, 1)
    ctx.set_option('grid_slices', slices)
    try:
        T2, G2, w2 = _grad(ctx, o, n, v, f, ns, data, weight, 2)
    finally:
        ctx.set_option('grid_slices', 0)
    assert np.array_equal(w1, w2)
    assert rel_l2(T2, T1) <= 1e-12


def test_grid_with_tilted_and_unnormalised_wall_normals(oracle, ctx):
    from nlos_surface_optimization_b200 import scenes, renderer
    o, n = scenes.wall_grid(6)
    rng = np.random.RandomState(5)
    n = n + 0.35 * rng.randn(*n.shape).astype(np.float32); n[:, 2] = np.abs(n[:, 2]) + 0.3
    n = np.ascontiguousarray(n * rng.uniform(0.5, 3.0, size=(n.shape[0], 1)).astype(np.float32))      # not unit length (the reference never normalises)
    v, f = scenes.icosphere(4, 0.1, (0.02, -0.03, 0.45), noise=0.03, seed=3); ns = 2 * f.shape[0]
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=1, loss_flag=0)
    T, G, _ = _grad(ctx, o, n, v, f, ns, data, weight, 2)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT and rel_l2(G, G_ref) <= TOL_GRADIENT
    assert np.array_equal(T > 0, T_ref > 0) or rel_l2(T, T_ref) <= 1e-7


def test_grid_falls_back_for_wall_points_that_do_not_see_the_mesh_in_front(oracle, ctx):
    """A mesh that reaches behind the wall plane of some wall points: those sources take the per-ray BVH query inside the grid kernel."""
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.merge([scenes.quad(0.40, 0.1), scenes.quad(0.25, 0.05, 0.05, 0.0), scenes.icosphere(2, 0.06, (-0.1, 0.05, 0.02))])   # sphere straddles z = 0
    o, n = scenes.wall_grid(5)
    o = np.ascontiguousarray(np.concatenate([o, np.array([[0.0, 0.0, 0.3], [0.05, 0.0, 0.41]], dtype=np.float32)]))    # two "wall points" inside the scene
    n = np.ascontiguousarray(np.concatenate([n, np.array([[0, 0, 1], [0, 0, -1]], dtype=np.float32)]))
    ns = 16 * f.shape[0]
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=1, loss_flag=0)
    T, G, w2 = _grad(ctx, o, n, v, f, ns, data, weight, 2)
    T1, G1, w1 = _grad(ctx, o, n, v, f, ns, data, weight, 1)
    assert np.array_equal(w1, w2)
    assert T_ref.sum() > 0
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT and rel_l2(G, G_ref) <= TOL_GRADIENT


def test_grid_many_samples_per_triangle(oracle, ctx):
    """The optimisation-loop regime (C-arm: small mesh, spp = 18): one grid per wall point serves all its samples."""
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.armadillo_init(); o, n = scenes.wall_grid(6); ns = 20000
    assert 1 + (ns - 1) // f.shape[0] == 18
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=1, loss_flag=0)
    T, G, w2 = _grad(ctx, o, n, v, f, ns, data, weight, 2)
    T1, G1, w1 = _grad(ctx, o, n, v, f, ns, data, weight, 1)
    assert np.array_equal(w1, w2)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT and rel_l2(G, G_ref) <= TOL_GRADIENT


def test_torch_tensors_are_ordered_against_the_current_stream(ctx):
    """device tensors are used in place on the context's own stream; the module orders the call against torch's current
    stream on both sides, so no manual synchronisation is needed around it."""
    import torch
    from nlos_surface_optimization_b200 import scenes, renderer
    dev = torch.device('cuda', 0)
    o, n = scenes.wall_grid(8); v, f = scenes.icosphere(3, 0.1, (0.0, 0.0, 0.45), noise=0.02, seed=1); ns = 4 * f.shape[0]
    B = 1200
    to = lambda a: torch.from_numpy(a).to(dev)
    d_o, d_n, d_f = to(o), to(n), to(f)
    T_h = np.zeros((o.shape[0], B)); pl_h = np.zeros(B)
    v_moved = v.copy(); v_moved[:, 2] += 0.05
    renderer.renderStreamedTransient(o, n, v_moved, f, ns, LB, UB, RES, T_h, pl_h, 1, 1, ctx=ctx)
    d_T = torch.zeros((o.shape[0], B), dtype=torch.float64, device=dev); d_pl = torch.zeros(B, dtype=torch.float64, device=dev)
    side = torch.cuda.Stream(device=dev)
    big = torch.randn(64 << 20, device=dev)
    with torch.cuda.stream(side):
        d_v = to(v)
        for _ in range(20):
            big = big * 1.0001                       # keep the stream busy so that the vertex update below is still pending at call time
        d_v[:, 2] += 0.05                            # written by a torch kernel on `side` right before the call
        renderer.renderStreamedTransient(d_o, d_n, d_v, d_f, ns, LB, UB, RES, d_T, d_pl, 1, 1, ctx=ctx)
        total = d_T.sum()                            # consumed by a torch kernel on `side` right after the call
    got = float(total.item())
    assert abs(got - T_h.sum()) <= 1e-9 * abs(T_h.sum())
