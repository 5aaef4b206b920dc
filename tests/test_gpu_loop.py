"""GPU test of the optimisation iteration (nlos_surface_optimization_b200/loop.py): the device-resident loop must walk
the same vertex trajectory and produce the same loss sequence as the reference's host-array call sequence
(exp_bunny/test.py:161-216, adam_modified.py:60-107, rendering.py:360-367), and rendering.removeTriangle must act through the facade."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup():
    import nlos_surface_optimization_b200 as nb
    from nlos_surface_optimization_b200 import scenes, loop, rendering
    ctx = nb.Context(0)
    o, n = scenes.wall_grid(12)
    gv, gf = scenes.armadillo(); iv, iF = scenes.armadillo_init()
    opt = loop.RenderOptions(20000, o, n)
    gt_opt = loop.RenderOptions(2 * gf.shape[0], o, n)
    gt_mesh = loop.Mesh(gv, gf)
    from nlos_surface_optimization_b200 import renderer
    gt = np.zeros((o.shape[0], opt.max_distance_bin)); pl = np.zeros(opt.max_distance_bin)
    renderer.renderStreamedTransient(o, n, gt_mesh.v, gt_mesh.f, gt_opt.sample_num, 0, opt.max_distance_bin * opt.distance_resolution, opt.distance_resolution, gt, pl, 1, 1, ctx=ctx)
    weight = rendering.create_weighting_function(gt, opt.gamma)
    return ctx, loop, rendering, opt, gt, weight, iv, iF


def test_device_loop_reproduces_the_host_loop():
    ctx, loop, rendering, opt, gt, weight, iv, iF = _setup()
    lr = 0.0001 / 3                                                   # exp_bunny/test.py:56
    host = loop.HostIteration(loop.Mesh(iv, iF), gt, weight, opt, lr, ctx=ctx)
    dev = loop.DeviceIteration(loop.Mesh(iv, iF), gt, weight, opt, lr, ctx=ctx)
    hl, dl = [], []
    for it in range(10):
        hl.append(host.step()); dl.append(dev.step())
        vh, vd = host.mesh.v, dev.vertices()
        err = np.linalg.norm(vh.astype(np.float64) - vd) / np.linalg.norm(vh)
        assert err <= 1e-6, 'iteration %d: vertex trajectories differ by %.3e' % (it, err)
    hl, dl = np.array(hl), np.array(dl)
    # iteration 0 starts from identical arrays: identical loss to FP64 summation order.  Later iterations start from float32 vertices that
    # may differ in the last bit (NumPy vs torch rounding inside Adam): a last-bit move of a vertex re-bins single Monte-Carlo samples,
    # which shows in the loss at ~1e-5 relative while the trajectories stay within 1e-6 (asserted above)
    assert np.all(np.abs(hl[0] - dl[0]) <= 1e-12 * np.abs(hl[0])), (hl[0], dl[0])
    assert np.all(np.abs(hl - dl) <= 1e-4 * np.abs(hl)), (hl, dl)
    assert hl[-1, 1] < hl[0, 1] and dl[-1, 1] < dl[0, 1]             # and the loop does reduce the data term
    assert np.linalg.norm(host.mesh.v - iv) > 0                       # the loop did move the mesh
    # no host synchronisation variant: device scalars come back, same numbers
    loss_t, l2_t = dev.step(read_loss=False)
    h = host.step()
    assert abs(float(l2_t) - h[1]) <= 1e-4 * abs(h[1])
    ctx.close()


def test_adam_modified_numpy_equals_torch():
    import torch
    from nlos_surface_optimization_b200.loop import AdamModified
    rng = np.random.RandomState(0)
    p0 = rng.randn(50, 3).astype(np.float32)
    a, b = AdamModified(1e-3), AdamModified(1e-3)
    pn = p0.copy(); pt = torch.from_numpy(p0.copy()).cuda()
    for _ in range(5):
        g = rng.randn(50, 3)
        pn = a.step(pn, g.astype(np.float32)); b.step(pt, torch.from_numpy(g).cuda())
    assert np.allclose(pn, pt.cpu().numpy(), rtol=2e-6, atol=1e-7)
    # the denominator is shared by the three components of a vertex (adam_modified.py:99): after ONE step the move is parallel to the gradient
    c = AdamModified(1e-3); g = rng.randn(50, 3).astype(np.float32); p1 = c.step(p0.copy(), g)
    cosang = np.sum((p0 - p1) * g, axis=1) / (np.linalg.norm(p0 - p1, axis=1) * np.linalg.norm(g, axis=1))
    assert np.all(cosang > 0.9999)


def test_remove_triangle_through_the_facade(oracle):
    """exp_bunny/rendering.py:271-278: faces no wall point sees are dropped unless all three neighbours exist."""
    ctx, loop, rendering, opt, gt, weight, iv, iF = _setup()
    from nlos_surface_optimization_b200 import scenes
    # an open sheet facing the wall plus an open sheet facing away from it (never seen, boundary faces go)
    v1, f1 = scenes.quad(0.40, 0.1); v2, f2 = scenes.quad(0.50, 0.1); f2 = f2[:, ::-1].copy()
    v, f = scenes.merge([(v1, f1), (v2, np.ascontiguousarray(f2))])
    mesh = loop.Mesh(v, f)
    opt2 = loop.RenderOptions(64, opt.lighting, opt.lighting_normal)
    inten = oracle.intensity(opt2.lighting, opt2.lighting_normal, mesh.v, mesh.f, opt2.sample_num, 0.0, opt2.max_distance_bin * opt2.distance_resolution)
    keep = np.logical_or(inten > 0, np.sum(mesh.f_affinity < 0, axis=1) == 0)
    rendering.removeTriangle(mesh, opt2, ctx=ctx)
    assert mesh.f.shape[0] == int(keep.sum()) and 0 < mesh.f.shape[0] < f.shape[0]
    assert np.array_equal(mesh.f, f[keep])
    ctx.close()
