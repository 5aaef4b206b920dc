"""GPU parity: the sm_100a path (through the C ABI / reference-signature modules) against the CPU oracle on
identical fixed-seed samples.  Bars (BASELINE.json north_star): per-sample visibility bit-exact, transient
relative L2 <= 1e-5, gradients relative L2 <= 1e-4."""
import numpy as np
import pytest
from helpers import LB, UB, RES, TOL_TRANSIENT, TOL_GRADIENT, rel_l2, make_target

pytestmark = pytest.mark.gpu


def _scene(name):
    from nlos_surface_optimization_b200 import scenes
    if name == 'fan8':
        v, f = scenes.fan8(); o, n = scenes.wall_grid(4); ns = 8 * 64
    elif name == 'occluders':
        v, f = scenes.merge([scenes.quad(0.40, 0.08), scenes.quad(0.55, 0.2), scenes.icosphere(2, 0.05, (0.1, -0.05, 0.3))])
        o, n = scenes.wall_grid(5); ns = f.shape[0] * 8
    elif name == 'ico':
        v, f = scenes.icosphere(4, 0.1, (0.02, -0.03, 0.45), noise=0.03, seed=3); o, n = scenes.wall_grid(6); ns = 20000
    elif name == 'bunny':
        v, f = scenes.bunny(); o, n = scenes.wall_grid(4); ns = 20000
    elif name == 'heightfield':        # the C-scale mesh family at test size (self-shadowing ridges, open surface)
        v, f = scenes.heightfield(41); o, n = scenes.wall_grid(5); ns = 3 * f.shape[0]
    else:
        raise KeyError(name)
    return o, n, v, f, ns


@pytest.mark.parametrize('name', ['fan8', 'occluders', 'ico', 'bunny', 'heightfield'])
def test_visibility_bit_exact(name, oracle, gpu_ctx):
    import nlos_surface_optimization_b200 as nb
    o, n, v, f, ns = _scene(name)
    ref = oracle.transient(o, n, v, f, ns, LB, UB, RES, want_visibility=True)[2]
    vis, cnt = nb.debug_visibility(o, v, f, ns, ctx=gpu_ctx)
    assert vis.shape == ref.shape
    assert np.array_equal(vis, ref), 'visibility differs in %d of %d samples' % (int((vis != ref).sum()), ref.size)
    assert cnt['rays'] > 0


@pytest.mark.parametrize('name', ['fan8', 'occluders', 'ico', 'bunny'])
@pytest.mark.parametrize('refine,sigma', [(1, 1), (10, 1)])
def test_forward_transient(name, refine, sigma, oracle, gpu_ctx):
    from nlos_surface_optimization_b200 import renderer
    o, n, v, f, ns = _scene(name)
    T_ref, pl_ref = oracle.transient(o, n, v, f, ns, LB, UB, RES, refine, sigma)[:2]
    B = T_ref.shape[1]
    T = np.full((o.shape[0], B), 7.0); pl = np.zeros(B)     # callee must zero the transient (TG.cpp:291)
    renderer.renderStreamedTransient(o, n, v, f, ns, LB, UB, RES, T, pl, refine, sigma, ctx=gpu_ctx)
    assert np.array_equal(pl, pl_ref)
    assert T_ref.sum() > 0
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT
    if refine == 1:
        # unsmoothed histogram: identical set of non-empty bins
        assert np.array_equal(T > 0, T_ref > 0)


@pytest.mark.parametrize('name', ['fan8', 'occluders', 'ico', 'bunny', 'heightfield'])
def test_vertex_gradient(name, oracle, gpu_ctx):
    from nlos_surface_optimization_b200 import renderer
    o, n, v, f, ns = _scene(name)
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, pl_ref = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=1, loss_flag=0)
    B = T_ref.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT
    assert np.linalg.norm(G_ref) > 0
    assert rel_l2(G, G_ref) <= TOL_GRADIENT
    # '+=' semantics (TG.cpp:563): a second call accumulates
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
    assert rel_l2(G, 2 * G_ref) <= TOL_GRADIENT


def test_gradient_without_visibility_reuse(oracle, gpu_ctx):
    """The gradient pass re-tracing its rays gives the same result as consuming the forward pass's bits."""
    from nlos_surface_optimization_b200 import renderer
    o, n, v, f, ns = _scene('ico')
    data, weight = make_target(oracle, o, n, v, f, ns)
    B = data.shape[1]
    out = []
    for reuse in (1, 0):
        gpu_ctx.set_option('reuse_visibility', reuse)
        T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
        renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
        out.append(G)
    gpu_ctx.set_option('reuse_visibility', 1)
    assert rel_l2(out[1], out[0]) <= 1e-12


@pytest.mark.parametrize('loss_flag,sigma', [(1, 1), (0, 5)])
def test_vertex_gradient_variants(loss_flag, sigma, oracle, gpu_ctx):
    from nlos_surface_optimization_b200 import renderer
    o, n, v, f, ns = _scene('ico')
    refine = 4
    data, weight = make_target(oracle, o, n, v, f, ns)
    weight = np.ascontiguousarray(weight * np.linspace(0.5, 1.5, weight.shape[1])[None, :])
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, refine, sigma, testing_flag=1, loss_flag=loss_flag)
    B = T_ref.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, refine, sigma, 1, loss_flag, ctx=gpu_ctx)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT
    assert rel_l2(G, G_ref) <= TOL_GRADIENT


@pytest.mark.parametrize('testing_flag', [0, 1])
def test_shading_gradient(testing_flag, oracle, gpu_ctx):
    from nlos_surface_optimization_b200 import renderer, scenes
    o, n, v, f, ns = _scene('ico')
    vn = scenes.vertex_normals(v, f)
    data, weight = make_target(oracle, o, n, v, f, ns, vertex_normal=vn)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=testing_flag, vertex_normal=vn)
    B = T_ref.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedShadingGradient(o, n, v, f, vn, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, testing_flag, 0, ctx=gpu_ctx)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT
    assert rel_l2(G, G_ref) <= TOL_GRADIENT
    T2 = np.zeros_like(T)
    renderer.renderStreamedTransientShading(o, n, v, vn, f, ns, LB, UB, RES, T2, pl, 1, 1, ctx=gpu_ctx)
    assert rel_l2(T2, T_ref) <= TOL_TRANSIENT


def test_albedo_paths(oracle, gpu_ctx):
    from nlos_surface_optimization_b200 import renderer
    o, n, v, f, ns = _scene('ico')
    rng = np.random.RandomState(0)
    alb = np.ascontiguousarray(0.5 + rng.rand(v.shape[0]), dtype=np.float32)
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, vertex_albedo=alb)
    _, g_ref = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, vertex_albedo=alb, kind=1)
    B = T_ref.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradientWithAlbedo(o, n, v, f, alb, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT
    assert rel_l2(G, G_ref) <= TOL_GRADIENT
    T3 = np.zeros_like(T)
    g = renderer.renderStreamedGradientAlbedo(o, n, v, f, alb, ns, LB, UB, RES, T3, pl, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
    assert rel_l2(T3, T_ref) <= TOL_TRANSIENT
    assert abs(g - g_ref) <= TOL_GRADIENT * abs(g_ref)
    T4 = np.zeros_like(T)
    renderer.renderStreamedTransientwAlbedo(o, n, v, alb, f, ns, LB, UB, RES, T4, pl, 1, 1, ctx=gpu_ctx)
    assert rel_l2(T4, T_ref) <= TOL_TRANSIENT


def test_intensity(oracle, gpu_ctx):
    from nlos_surface_optimization_b200 import renderer, ggx
    o, n, v, f, ns = _scene('ico')
    I_ref = oracle.intensity(o, n, v, f, ns, LB, UB)
    I = np.zeros(f.shape[0])
    renderer.renderStreamedTriangleIntensity(o, n, v, f, ns, LB, UB, I, ctx=gpu_ctx)
    assert rel_l2(I, I_ref) <= TOL_TRANSIENT
    assert np.array_equal(I > 0, I_ref > 0)          # removeTriangle thresholds at 0 (exp_bunny/rendering.py:275-276)
    Ig_ref = oracle.intensity(o, n, v, f, ns, LB, UB, alpha=0.3)
    Ig = np.zeros(f.shape[0])
    ggx.renderStreamedTriangleIntensity(o, n, v, f, 0.3, ns, LB, UB, Ig, ctx=gpu_ctx)
    assert rel_l2(Ig, Ig_ref) <= TOL_TRANSIENT


@pytest.mark.parametrize('alpha', [0.1, 0.5])
def test_ggx(alpha, oracle, gpu_ctx):
    from nlos_surface_optimization_b200 import ggx, scenes
    o, n, v, f, ns = _scene('ico')
    vn = scenes.vertex_normals(v, f)
    data = oracle.transient(o, n, v, f, ns, LB, UB, RES, alpha=alpha * 2)[0]
    weight = np.ones_like(data)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=1, alpha=alpha)
    _, ga_ref = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, alpha=alpha, kind=2)
    B = T_ref.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    ggx.renderStreamedGradient(o, n, v, f, alpha, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, ctx=gpu_ctx)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT
    assert rel_l2(G, G_ref) <= TOL_GRADIENT
    T2 = np.zeros_like(T)
    ga = ggx.renderStreamedGradientAlpha(o, n, v, f, alpha, ns, LB, UB, RES, T2, pl, data, weight, 10, 1, ctx=gpu_ctx)
    assert abs(ga - ga_ref) <= TOL_GRADIENT * abs(ga_ref)
    # shading normals with the normal-variation term (testing_flag = 0)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, testing_flag=0, alpha=alpha, vertex_normal=vn)
    G = np.zeros((v.shape[0], 3))
    ggx.renderStreamedShadingGradient(o, n, v, f, vn, alpha, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 0, ctx=gpu_ctx)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT
    assert rel_l2(G, G_ref) <= TOL_GRADIENT
    T3 = np.zeros_like(T)
    ggx.renderStreamedTransient(o, n, v, f, alpha, ns, LB, UB, RES, T3, pl, 10, 1, ctx=gpu_ctx)
    T3_ref = oracle.transient(o, n, v, f, ns, LB, UB, RES, 10, 1, alpha=alpha)[0]
    assert rel_l2(T3, T3_ref) <= TOL_TRANSIENT


def test_device_tensors_in_place(oracle, gpu_ctx):
    """torch CUDA tensors are used in place (no host staging) and give the same numbers as host arrays."""
    import torch
    from nlos_surface_optimization_b200 import renderer
    o, n, v, f, ns = _scene('ico')
    data, weight = make_target(oracle, o, n, v, f, ns)
    B = data.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
    dev = torch.device('cuda:0')
    to = lambda a: torch.from_numpy(a).to(dev)
    dT = torch.zeros((o.shape[0], B), dtype=torch.float64, device=dev); dpl = torch.zeros(B, dtype=torch.float64, device=dev)
    dG = torch.zeros((v.shape[0], 3), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    renderer.renderStreamedGradient(to(o), to(n), to(v), to(f), ns, LB, UB, RES, dT, dpl, dG, to(data), to(weight), 10, 1, 1, 0, ctx=gpu_ctx)
    gpu_ctx.synchronize()
    assert rel_l2(dT.cpu().numpy(), T) <= 1e-12
    assert rel_l2(dG.cpu().numpy(), G) <= 1e-9


def test_sharded_sources_match_single_call(oracle, gpu_ctx):
    """Two half-slices of the wall with set_source_window reproduce the single-call result (multi-GPU contract)."""
    from nlos_surface_optimization_b200 import renderer
    o, n, v, f, ns = _scene('ico')
    data, weight = make_target(oracle, o, n, v, f, ns)
    L, B = data.shape
    T = np.zeros((L, B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
    Ts = np.zeros((L, B)); Gs = np.zeros((v.shape[0], 3)); h = L // 2
    try:
        for a, b in ((0, h), (h, L)):
            gpu_ctx.set_source_window(a, L)
            Tpart = np.zeros((b - a, B))
            renderer.renderStreamedGradient(np.ascontiguousarray(o[a:b]), np.ascontiguousarray(n[a:b]), v, f, ns, LB, UB, RES, Tpart, pl, Gs,
                                            np.ascontiguousarray(data[a:b]), np.ascontiguousarray(weight[a:b]), 10, 1, 1, 0, ctx=gpu_ctx)
            Ts[a:b] = Tpart
    finally:
        gpu_ctx.set_source_window(0, 0)
    assert rel_l2(Ts, T) <= 1e-12
    assert rel_l2(Gs, G) <= 1e-9


def test_edge_cases(gpu_ctx):
    import nlos_surface_optimization_b200 as nb
    from nlos_surface_optimization_b200 import renderer, scenes
    o, n = scenes.wall_grid(2)
    v, f = scenes.quad(0.4, 0.1)
    B = 1200
    T = np.zeros((4, B)); pl = np.zeros(B)
    # back-facing quad renders all-zero under the product clamp
    renderer.renderStreamedTransient(o, n, v, np.ascontiguousarray(f[:, ::-1]), 64, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    assert T.sum() == 0
    # single triangle (no BVH nodes), and a degenerate triangle in the mesh
    f1 = np.ascontiguousarray(f[:1])
    renderer.renderStreamedTransient(o, n, v, f1, 64, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    assert T.sum() > 0
    fd = np.ascontiguousarray(np.vstack([f, [[0, 0, 1]]]).astype(np.int32))
    renderer.renderStreamedTransient(o, n, v, fd, 64, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    assert np.isfinite(T).all() and T.sum() > 0
    # shape errors are AssertionErrors like the Cython asserts (renderer.pyx:95-110)
    with pytest.raises(AssertionError):
        renderer.renderStreamedTransient(o, n, v, f, 64, LB, UB, RES, np.zeros((4, B - 1)), pl, 1, 1, ctx=gpu_ctx)
    with pytest.raises(ValueError):
        renderer.renderStreamedTransient(o.astype(np.float64), n, v, f, 64, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    # range window that excludes the mesh -> zeros
    from nlos_surface_optimization_b200._arrays import num_bins
    B2 = num_bins(0.0, 0.12, RES)
    T2 = np.zeros((4, B2)); pl2 = np.zeros(B2)
    renderer.renderStreamedTransient(o, n, v, f, 64, 0.0, 0.12, RES, T2, pl2, 1, 1, ctx=gpu_ctx)
    assert T2.sum() == 0


# ------------------------------------------------------------------------------------------------ full-size goldens
def _vis_digest(vis):
    import zlib
    L = vis.shape[0]; flat = vis.reshape(L, -1)
    pop = flat.sum(axis=1).astype(np.int64)
    crc = np.array([zlib.crc32(np.packbits(flat[i]).tobytes()) for i in range(L)], dtype=np.uint32)
    return pop, crc


@pytest.mark.parametrize('fixture', ['c_bunny16', 'c_bunny'])
def test_c_bunny_against_committed_oracle_goldens(fixture, gpu_ctx):
    """BASELINE.json configs[1] at full size (bunny, 64x64 wall, spp=1, r=10, s=1) against oracle output committed under
    tests/golden (tools/make_golden.py): per-sample visibility digests bit-exact, transient <= 1e-5, gradient <= 1e-4."""
    import os
    import nlos_surface_optimization_b200 as nb
    from nlos_surface_optimization_b200 import renderer, scenes
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', fixture + '.npz'))
    wall, ns = int(g['wall']), int(g['num_sample'])
    v, f = scenes.bunny(); o, n = scenes.wall_grid(wall)
    L, B = o.shape[0], 1200
    # target = the displaced mesh rendered by the GPU path; its marginals must match the oracle's
    v2 = v.copy(); v2[:, 2] += 0.01
    data = np.zeros((L, B)); pl = np.zeros(B)
    renderer.renderStreamedTransient(o, n, v2, f, ns, LB, UB, RES, data, pl, 1, 1, ctx=gpu_ctx)
    assert rel_l2(data.sum(1), g['data_row_sum']) <= TOL_TRANSIENT and rel_l2(data.sum(0), g['data_col_sum']) <= TOL_TRANSIENT
    weight = np.ones_like(data)
    T = np.zeros((L, B)); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
    assert rel_l2(T[g['rows']], g['transient_rows']) <= TOL_TRANSIENT
    assert rel_l2(T.sum(1), g['transient_row_sum']) <= TOL_TRANSIENT
    assert rel_l2(T.sum(0), g['transient_col_sum']) <= TOL_TRANSIENT
    assert abs((T * T).sum() - float(g['transient_sq_sum'])) <= 2 * TOL_TRANSIENT * float(g['transient_sq_sum'])
    assert np.array_equal((T > 0).sum(1), g['transient_nnz'])
    assert rel_l2(G, g['gradient']) <= TOL_GRADIENT
    # per-sample visibility, digested per source (popcount + crc32 of the packed bits), in slabs to bound host memory
    slab = 256
    for a in range(0, L, slab):
        vis, _ = nb.debug_visibility(np.ascontiguousarray(o[a:a + slab]), v, f, ns, ctx=_windowed(gpu_ctx, a, L))
        pop, crc = _vis_digest(vis)
        assert np.array_equal(pop, g['vis_pop'][a:a + slab])
        assert np.array_equal(crc, g['vis_crc'][a:a + slab])
    gpu_ctx.set_source_window(0, 0)


@pytest.mark.parametrize('fixture', ['c_ggx16', 'c_ggx'])
def test_c_ggx_against_committed_oracle_goldens(fixture, gpu_ctx):
    """BASELINE.json configs[2] (exp_ggx/test10.py:37,60): GGX render + gradient of the bunny, alpha = 0.1 against data rendered at
    alpha = 0.2, on a 16x16 and on the full 64x64 wall, against oracle output committed under tests/golden (tools/make_golden.py):
    transient <= 1e-5; vertex gradient <= 1e-4 with testing_flag 1 (face normals) and 0 (shading normals); alpha scalar <= 1e-4."""
    import os
    from nlos_surface_optimization_b200 import ggx, scenes
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', fixture + '.npz'))
    wall, ns, alpha, alpha_data = int(g['wall']), int(g['num_sample']), float(g['alpha']), float(g['alpha_data'])
    v, f = scenes.bunny(); o, n = scenes.wall_grid(wall); vn = scenes.vertex_normals(v, f)
    L, B = o.shape[0], 1200
    data = np.zeros((L, B)); pl = np.zeros(B)
    ggx.renderStreamedTransient(o, n, v, f, alpha_data, ns, LB, UB, RES, data, pl, 1, 1, ctx=gpu_ctx)
    assert rel_l2(data.sum(1), g['data_row_sum']) <= TOL_TRANSIENT and rel_l2(data.sum(0), g['data_col_sum']) <= TOL_TRANSIENT
    weight = np.ones_like(data)
    T = np.zeros((L, B)); G = np.zeros((v.shape[0], 3))
    ggx.renderStreamedGradient(o, n, v, f, alpha, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, ctx=gpu_ctx)
    assert rel_l2(T[g['rows']], g['transient_rows']) <= TOL_TRANSIENT
    assert rel_l2(T.sum(1), g['transient_row_sum']) <= TOL_TRANSIENT and rel_l2(T.sum(0), g['transient_col_sum']) <= TOL_TRANSIENT
    assert rel_l2(G, g['gradient_tf1']) <= TOL_GRADIENT
    T0 = np.zeros((L, B)); G0 = np.zeros((v.shape[0], 3))
    ggx.renderStreamedShadingGradient(o, n, v, f, vn, alpha, ns, LB, UB, RES, T0, pl, G0, data, weight, 10, 1, 0, ctx=gpu_ctx)
    assert rel_l2(T0.sum(1), g['shading_transient_row_sum']) <= TOL_TRANSIENT and rel_l2(T0.sum(0), g['shading_transient_col_sum']) <= TOL_TRANSIENT
    assert rel_l2(G0, g['gradient_tf0_shading']) <= TOL_GRADIENT
    Ta = np.zeros((L, B))
    ga = ggx.renderStreamedGradientAlpha(o, n, v, f, alpha, ns, LB, UB, RES, Ta, pl, data, weight, 10, 1, ctx=gpu_ctx)
    assert abs(ga - float(g['alpha_grad'])) <= TOL_GRADIENT * abs(float(g['alpha_grad']))


def test_c_scale_mesh_against_committed_oracle_golden(gpu_ctx):
    """BASELINE.json configs[4] at its full mesh size (height field F = 500 000, B = 2048) on a 16x16 wall: per-sample visibility
    digests bit-exact, transient <= 1e-5, gradient <= 1e-4 (norm, column sums, 1024-vertex block sums and every 64th row)."""
    import os
    import nlos_surface_optimization_b200 as nb
    from nlos_surface_optimization_b200 import renderer, scenes
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'c_scale16.npz'))
    wall, ns, B = int(g['wall']), int(g['num_sample']), int(g['numbins'])
    v, f = scenes.heightfield(501); o, n = scenes.wall_grid(wall)
    assert f.shape[0] == 500000
    L, ub = o.shape[0], B * RES
    v2 = v.copy(); v2[:, 2] += 0.01
    data = np.zeros((L, B)); pl = np.zeros(B)
    renderer.renderStreamedTransient(o, n, v2, f, ns, LB, ub, RES, data, pl, 1, 1, ctx=gpu_ctx)
    assert rel_l2(data.sum(1), g['data_row_sum']) <= TOL_TRANSIENT and rel_l2(data.sum(0), g['data_col_sum']) <= TOL_TRANSIENT
    T = np.zeros((L, B)); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, ub, RES, T, pl, G, data, np.ones_like(data), 10, 1, 1, 0, ctx=gpu_ctx)
    assert rel_l2(T.sum(1), g['transient_row_sum']) <= TOL_TRANSIENT and rel_l2(T.sum(0), g['transient_col_sum']) <= TOL_TRANSIENT
    assert abs((T * T).sum() - float(g['transient_sq_sum'])) <= 2 * TOL_TRANSIENT * float(g['transient_sq_sum'])
    assert rel_l2(G[::64], g['gradient_rows']) <= TOL_GRADIENT
    assert abs(np.linalg.norm(G) - float(g['gradient_l2'])) <= TOL_GRADIENT * float(g['gradient_l2'])
    assert rel_l2(np.add.reduceat(G, np.arange(0, G.shape[0], 1024), axis=0), g['gradient_block_sum']) <= TOL_GRADIENT
    slab = 32
    for a in range(0, L, slab):
        vis, _ = nb.debug_visibility(np.ascontiguousarray(o[a:a + slab]), v, f, ns, ctx=_windowed(gpu_ctx, a, L))
        pop, crc = _vis_digest(vis)
        assert np.array_equal(pop, g['vis_pop'][a:a + slab]) and np.array_equal(crc, g['vis_crc'][a:a + slab])
    gpu_ctx.set_source_window(0, 0)


def _windowed(ctx, offset, total):
    ctx.set_source_window(offset, total)
    return ctx


def test_vertex_gradient_debug_and_regularisers(oracle, gpu_ctx):
    from nlos_surface_optimization_b200 import renderer, scenes
    v, f = scenes.fan8(); o, n = scenes.wall_grid(2)
    ns = 8 * 200; B = 1200
    ref = oracle.vertex_gradient(8, o, n, v, f, ns, LB, UB, RES, 10, 1)
    G = np.zeros((B, 3))
    renderer.renderStreamedVertexGradient(o, n, v, f, ns, LB, UB, RES, G, 8, 10, 1, ctx=gpu_ctx)
    assert np.linalg.norm(ref) > 0 and rel_l2(G, ref) <= TOL_GRADIENT
    for mesh in (scenes.icosphere(3, 0.1, (0, 0, 0.45), noise=0.05, seed=7), scenes.bunny()):
        mv, mf = mesh
        aff = scenes.face_affinity(mf) if mf.shape[0] < 5000 else -np.ones_like(mf)
        if mf.shape[0] >= 5000:      # cheap synthetic adjacency for the big mesh: neighbours by index
            aff = np.ascontiguousarray(np.stack([np.roll(np.arange(mf.shape[0]), 1), np.roll(np.arange(mf.shape[0]), -1), -np.ones(mf.shape[0])], axis=1).astype(np.int32))
        val_ref, G_ref = oracle.normal_smoothing(mv, mf, aff)
        Gn = np.full((mv.shape[0], 3), 5.0)
        val = renderer.renderStreamedNormalSmoothing(mv, mf, aff, Gn, ctx=gpu_ctx)
        assert abs(val - val_ref) <= 1e-9 * max(abs(val_ref), 1e-12)
        assert np.array_equal(Gn, G_ref)
        C_ref = oracle.curvature_grad(mv, mf)
        Gc = np.full((mv.shape[0], 3), 5.0)
        renderer.renderStreamedCurvatureGradient(mv, mf, Gc, ctx=gpu_ctx)
        assert np.array_equal(Gc, C_ref)


def test_embree_intersector_api(oracle, gpu_ctx):
    """batched nearest-hit queries (primID,u,v) on the renderer's LBVH equal the oracle's nearest hit exactly."""
    from nlos_surface_optimization_b200 import embree_intersector as ei, scenes
    rng = np.random.RandomState(5)
    for v, f in (scenes.icosphere(4, 0.1, (0, 0, 0.45), noise=0.04, seed=2), scenes.bunny()):
        N = 20000
        o = np.ascontiguousarray(np.stack([rng.uniform(-.25, .25, N), rng.uniform(-.25, .25, N), np.zeros(N)], 1), dtype=np.float32)
        tgt = v[rng.randint(0, v.shape[0], N)] + rng.normal(0, 0.01, (N, 3))
        d = np.ascontiguousarray((tgt - o) * rng.uniform(0.5, 2.0, (N, 1)), dtype=np.float32)      # not normalised on purpose
        ref3, ref1 = oracle.intersect(o, d, v, f)
        b3 = np.full((N, 3), 7.0, dtype=np.float32); b1 = np.zeros(N, dtype=np.float32)
        ei.embree3_tbb_intersection(o, d, v, f, b3, ctx=gpu_ctx)
        ei.embree3_tbb_short_intersection(o, d, v, f, b1, ctx=gpu_ctx)
        hit = ref1 >= 0
        assert 0.2 < hit.mean() < 1.0
        assert np.array_equal(b1, ref1)
        assert np.array_equal(b3[hit], ref3[hit])
        assert np.array_equal(b3[~hit, 0], ref3[~hit, 0]) and (b3[~hit, 1:] == 7.0).all()      # misses leave (u,v) untouched
        p_ref = oracle.bary_to_world(v, f, ref3)
        p = np.full((N, 3), 9.0, dtype=np.float32)
        ei.PyMesh(v, f, ctx=gpu_ctx).barycoord_to_world(b3, p)
        assert np.array_equal(p[hit], p_ref[hit]) and (p[~hit] == 9.0).all()
        # the hit point lies on the ray
        t = np.linalg.norm(p[hit] - o[hit], axis=1) / np.linalg.norm(d[hit], axis=1)
        assert np.abs(o[hit] + t[:, None] * d[hit] - p[hit]).max() < 1e-4


def test_jitter_temporal_kernel(oracle, gpu_ctx):
    """tabulated SPAD-jitter kernel (asymmetric 40-tap, like jitter/jitter_info.mat) forward + vertex gradient."""
    from nlos_surface_optimization_b200 import jitter
    o, n, v, f, ns = _scene('ico')
    J, off = 40, 12
    x = np.arange(J) - off
    jw = np.exp(-0.5 * (x / 3.0) ** 2) + 0.3 * np.exp(-np.maximum(x, 0) / 8.0) * (x > 0); jw /= jw.sum()
    jg = np.gradient(jw)
    jw2 = np.ascontiguousarray(jw.reshape(J, 1)); jg2 = np.ascontiguousarray(jg.reshape(J, 1))
    T_ref, pl_ref = oracle.jitter_transient(o, n, v, f, ns, LB, UB, RES, jw, off)
    B = T_ref.shape[1]
    T = np.full((o.shape[0], B), 3.0); pl = np.zeros(B)
    jitter.renderStreamedTransient(o, n, v, f, ns, LB, UB, RES, T, pl, jw2, off, ctx=gpu_ctx)
    assert np.array_equal(pl, pl_ref) and rel_l2(T, T_ref) <= TOL_TRANSIENT
    v2 = v.copy(); v2[:, 2] += 0.01
    data = oracle.jitter_transient(o, n, v2, f, ns, LB, UB, RES, jw, off)[0]; weight = np.ones_like(data)
    for tf in (1,):
        T_ref, G_ref, _ = oracle.jitter_gradient(o, n, v, f, ns, LB, UB, RES, jw, jg, off, data, weight, testing_flag=tf)
        T = np.zeros((o.shape[0], B)); G = np.zeros((v.shape[0], 3))
        jitter.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, jw2, jg2, off, T, pl, G, data, weight, tf, ctx=gpu_ctx)
        assert rel_l2(T, T_ref) <= TOL_TRANSIENT
        assert np.linalg.norm(G_ref) > 0 and rel_l2(G, G_ref) <= TOL_GRADIENT


def test_first_generation_api(oracle, gpu_ctx):
    """stratified_transient_raytracer/ module (`renderer_sr`): unclamped forward, single-origin renderTransient,
    gradient with the twice-applied box filter, one tap, normal-variation term always on, '=' into the gradient."""
    from nlos_surface_optimization_b200 import renderer_sr, scenes
    qv, qf = scenes.quad(0.4, 0.05, -0.17, 0.1)
    v, f = scenes.merge([scenes.icosphere(3, 0.1, (0.02, -0.03, 0.45), noise=0.03, seed=3), (qv, np.ascontiguousarray(qf[:, ::-1])),
                         scenes.quad(0.3, 0.04, 0.03, 0.0)])
    o, n = scenes.wall_grid(5); ns = 6 * f.shape[0]
    T_ref, pl_ref = oracle.sr_transient(o, n, v, f, ns, LB, UB, RES)
    T_st = oracle.transient(o, n, v, f, ns, LB, UB, RES)[0]
    assert T_ref.sum() > 1.0005 * T_st.sum()                       # the back-facing quad only shows without the clamp
    B = T_ref.shape[1]
    T = np.full((o.shape[0], B), 3.0); pl = np.zeros(B)
    renderer_sr.renderStreamedTransient(o, n, v, f, ns, LB, UB, RES, T, pl, ctx=gpu_ctx)
    assert np.array_equal(pl, pl_ref) and rel_l2(T, T_ref) <= TOL_TRANSIENT
    vn = scenes.vertex_normals(v, f); va = (0.5 + np.random.RandomState(2).rand(v.shape[0])).astype(np.float32)
    renderer_sr.renderStreamedTransientShading(o, n, v, vn, f, ns, LB, UB, RES, T, pl, ctx=gpu_ctx)
    assert rel_l2(T, oracle.sr_transient(o, n, v, f, ns, LB, UB, RES, vertex_normal=vn)[0]) <= TOL_TRANSIENT
    renderer_sr.renderStreamedTransientwAlbedo(o, n, v, va, f, ns, LB, UB, RES, T, pl, ctx=gpu_ctx)
    assert rel_l2(T, oracle.sr_transient(o, n, v, f, ns, LB, UB, RES, vertex_albedo=va)[0]) <= TOL_TRANSIENT
    # single origin
    t1 = np.zeros(B); renderer_sr.renderTransient(o[7], n[7], v, f, ns, LB, UB, RES, t1, pl, ctx=gpu_ctx)
    assert rel_l2(t1, oracle.sr_transient(o[7:8], n[7:8], v, f, ns, LB, UB, RES)[0][0]) <= TOL_TRANSIENT
    # gradient
    v2 = v.copy(); v2[:, 2] += 0.01
    data = oracle.sr_transient(o, n, v2, f, ns, LB, UB, RES)[0]
    for w in (0, 2, 7):
        T_ref, G_ref, _ = oracle.sr_gradient(o, n, v, f, ns, LB, UB, RES, w, data)
        T = np.zeros((o.shape[0], B)); G = np.full((v.shape[0], 3), 5.0)      # '=' semantics: the 5s must disappear
        renderer_sr.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, w, T, pl, G, data, ctx=gpu_ctx)
        assert rel_l2(T, T_ref) <= TOL_TRANSIENT
        assert np.linalg.norm(G_ref) > 0 and rel_l2(G, G_ref) <= TOL_GRADIENT, (w, rel_l2(G, G_ref))
    # a second-generation call afterwards is unaffected by the first-generation mode
    from nlos_surface_optimization_b200 import renderer
    renderer.renderStreamedTransient(o, n, v, f, ns, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    assert rel_l2(T, T_st) <= TOL_TRANSIENT


def test_zero_area_triangles(oracle, gpu_ctx):
    """Coincident / collinear vertices: such triangles cannot be hit, so they contribute nothing — on both sides, without NaNs."""
    from nlos_surface_optimization_b200 import renderer, scenes
    o, n = scenes.wall_grid(3)
    v, f = scenes.fan8()
    v = np.ascontiguousarray(np.vstack([v, [[0.1, 0.1, 0.3], [0.1, 0.1, 0.3], [0.2, 0.1, 0.3], [0.3, 0.1, 0.3]]]), dtype=np.float32)
    f = np.ascontiguousarray(np.vstack([f, [[9, 10, 11]], [[10, 11, 12]]]), dtype=np.int32)
    ns = 4 * f.shape[0]
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, 1, 0)
    B = T_ref.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
    assert np.isfinite(T).all() and np.isfinite(G).all()
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT and rel_l2(G, G_ref) <= TOL_GRADIENT
    assert np.all(G[9:] == 0)


def test_facade_helpers_on_the_boundary(oracle, gpu_ctx):
    """exp_bunny/rendering.py helpers that call the boundary directly: vertex_gradient (:26-30) and space_carving_projection (:193-206)."""
    from nlos_surface_optimization_b200 import rendering, scenes

    class Obj(object):
        pass
    o, n = scenes.wall_grid(1)
    mesh = Obj(); mesh.v, mesh.f = scenes.icosphere(2, 0.1, (0.0, 0.0, 0.45))
    opt = Obj(); opt.lighting, opt.lighting_normal = o, n
    opt.max_distance_bin, opt.distance_resolution, opt.sample_num, opt.bin_refine_resolution, opt.sigma_bin = 1200, 1.2e-3, 50 * mesh.f.shape[0], 10, 1
    g = rendering.vertex_gradient(mesh, 3, opt)
    g_ref = oracle.vertex_gradient(3, o, n, mesh.v, mesh.f, opt.sample_num, 0.0, 1.44, 1.2e-3, 10, 1)
    assert np.linalg.norm(g_ref) > 0 and rel_l2(g, g_ref) <= TOL_GRADIENT
    # space carving: a flat hull at z = 0.4 over x,y in [-0.1, 0.1]; vertices in front of it are pushed back to 0.4, others stay
    hull = Obj(); hull.v, hull.f = scenes.quad(0.4, 0.1)
    v = np.array([[0.0, 0.0, 0.3], [0.05, -0.05, 0.5], [0.3, 0.3, 0.2]], dtype=np.float32)
    rendering.space_carving_projection(v, hull)
    assert np.allclose(v[:, 2], [0.4, 0.5, 0.2], atol=1e-6)
    nrm, area = rendering.face_normal_and_area(hull.v, hull.f)
    assert np.allclose(area, 0.02, rtol=1e-5) and np.allclose(np.abs(nrm[:, 2]), 1.0, atol=1e-6)


def test_sharded_rendering_nccl_two_gpus(oracle, tmp_path):
    """dist.inverse_rendering_sharded over NCCL on 2 GPUs (skipped on a 1-GPU box): all-reduced gradient and gathered
    transient equal the single-GPU call."""
    import os, subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from nlos_surface_optimization_b200 import renderer, scenes
    import nlos_surface_optimization_b200 as nb
    o, n, v, f, ns = _scene('ico')
    data, weight = make_target(oracle, o, n, v, f, ns)
    L, B = data.shape
    T = np.zeros((L, B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=nb.default_context(0))
    np.savez(str(tmp_path / 'in.npz'), o=o, n=n, v=v, f=f, data=data, weight=weight)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / 'w.py'
    script.write_text('''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from nlos_surface_optimization_b200 import dist as nd
local = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
d = np.load(%r)
T, G, pl = nd.inverse_rendering_sharded(d['o'], d['n'], d['v'], d['f'], %d, %r, %r, %r, d['data'], d['weight'], 10, 1, device=torch.device('cuda', local), gather=True)
if dist.get_rank() == 0: np.savez(%r, T=T, G=G)
dist.barrier(); dist.destroy_process_group()
''' % (root, str(tmp_path / 'in.npz'), ns, LB, UB, RES, str(tmp_path / 'out.npz')))
    rc = subprocess.call([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                          '--master-port', '29533', str(script)], timeout=600)
    assert rc == 0
    out = np.load(str(tmp_path / 'out.npz'))
    assert rel_l2(out['T'], T) <= 1e-12
    assert rel_l2(out['G'], G) <= 1e-9


def test_degenerate_sizes_and_error_paths(gpu_ctx):
    """Empty inputs and bad arguments: no crash, zeros out, errors reported through the status code (the reference printf()s)."""
    import ctypes as C
    import nlos_surface_optimization_b200 as nb
    from nlos_surface_optimization_b200 import renderer, scenes
    v, f = scenes.quad(0.4, 0.1); B = 1200
    # zero sources
    o0 = np.zeros((0, 3), np.float32); T0 = np.zeros((0, B)); pl = np.zeros(B); G = np.ones((v.shape[0], 3))
    renderer.renderStreamedGradient(o0, o0, v, f, 64, LB, UB, RES, T0, pl, G, T0, T0, 10, 1, 1, 0, ctx=gpu_ctx)
    assert (G == 1.0).all() and pl[1] > 0
    # zero faces: transient zeroed, gradient untouched
    o, n = scenes.wall_grid(2)
    f0 = np.zeros((0, 3), np.int32); T = np.full((4, B), 5.0)
    renderer.renderStreamedTransient(o, n, v, f0, 64, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    assert (T == 0).all()
    # bad numBins / resolution are rejected by the C layer with a message
    lib, h = gpu_ctx.lib, gpu_ctx.handle
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float)); dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib.nlos_streamed_render_transient(h, fp(o), 4, fp(n), fp(v), 4, None, None, f.ctypes.data_as(C.POINTER(C.c_int)), 2, 64, 0.0, 1.44, 1.2e-3,
                                            dp(T), dp(pl), 1, 1, 0)
    assert rc == nb._ffi.NLOS_ERR_INVALID and b'numBins' in lib.nlos_last_error(h)
    rc = lib.nlos_streamed_render_transient(h, fp(o), 4, fp(n), fp(v), 4, None, None, f.ctypes.data_as(C.POINTER(C.c_int)), 2, 64, 0.0, 1.44, 0.0,
                                            dp(T), dp(pl), 1, 1, B)
    assert rc == nb._ffi.NLOS_ERR_INVALID
    assert lib.nlos_streamed_render_transient(None, fp(o), 4, fp(n), fp(v), 4, None, None, f.ctypes.data_as(C.POINTER(C.c_int)), 2, 64, 0.0, 1.44, 1.2e-3,
                                              dp(T), dp(pl), 1, 1, B) == nb._ffi.NLOS_ERR_INVALID
    # the context still works afterwards
    renderer.renderStreamedTransient(o, n, v, f, 64, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    assert T.sum() > 0


@pytest.mark.parametrize('spp', [3, 40])
def test_many_samples_per_triangle(spp, oracle, gpu_ctx):
    """spp > 1 (odd and even sample indices share a Philox block), forward + gradient + smoothed forward."""
    from nlos_surface_optimization_b200 import renderer, scenes
    v, f = scenes.icosphere(2, 0.1, (0.0, 0.01, 0.45), noise=0.03, seed=9); o, n = scenes.wall_grid(3)
    ns = spp * f.shape[0]
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1)
    B = T_ref.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT and rel_l2(G, G_ref) <= TOL_GRADIENT
    Ts_ref = oracle.transient(o, n, v, f, ns, LB, UB, RES, 10, 2)[0]
    Ts = np.zeros_like(T)
    renderer.renderStreamedTransient(o, n, v, f, ns, LB, UB, RES, Ts, pl, 10, 2, ctx=gpu_ctx)
    assert rel_l2(Ts, Ts_ref) <= TOL_TRANSIENT


@pytest.mark.parametrize('reuse', [1, 0])
def test_gradient_chunk_sizes_agree(reuse, oracle, gpu_ctx):
    """Sources per block of the gradient kernel (option chunk_gradient; 0 = automatic, small meshes get small chunks): any chunk — also ones
    that leave a partial group of 32 sample slots at the end of a block, with spp > 1 — gives the same gradient, against the oracle and among
    themselves, with the forward pass's visibility words (only the slots with a visible lane are walked) and when re-tracing."""
    from nlos_surface_optimization_b200 import renderer, scenes
    v, f = scenes.icosphere(3, 0.1, (0.01, -0.02, 0.45), noise=0.03, seed=5); o, n = scenes.wall_grid(7)       # 49 wall points
    ns = 3 * f.shape[0]                                                                                       # spp = 3: 147 slots
    data, weight = make_target(oracle, o, n, v, f, ns)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1)
    B = T_ref.shape[1]
    gpu_ctx.set_option('reuse_visibility', reuse)
    out = []
    try:
        for chunk in (0, 1, 4, 13, 37, 128):
            gpu_ctx.set_option('chunk_gradient', chunk)
            T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
            renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, 10, 1, 1, 0, ctx=gpu_ctx)
            assert rel_l2(G, G_ref) <= TOL_GRADIENT
            out.append(G)
    finally:
        gpu_ctx.set_option('chunk_gradient', 0); gpu_ctx.set_option('reuse_visibility', 1)
    for G in out[1:]:
        assert rel_l2(G, out[0]) <= 1e-12
