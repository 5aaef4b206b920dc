"""GPU tests of the round-2 ABI hardening: face-index validation, long tap tables in the gradient kernel, the visibility
buffer fallback is exercised indirectly (reuse_visibility=0 parity lives in test_gpu_parity.py)."""
import numpy as np
import pytest
from helpers import LB, UB, RES, TOL_TRANSIENT, TOL_GRADIENT, rel_l2, make_target

pytestmark = pytest.mark.gpu


def test_out_of_range_face_indices_are_refused(gpu_ctx):
    import torch
    import nlos_surface_optimization_b200 as nb
    from nlos_surface_optimization_b200 import renderer, scenes
    o, n = scenes.wall_grid(4); v, f = scenes.fan8()
    B = 1200
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B)
    bad = f.copy(); bad[3, 1] = v.shape[0]            # one past the last vertex
    with pytest.raises(nb.NlosError, match='out of range'):
        renderer.renderStreamedTransient(o, n, v, bad, 512, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    bad[3, 1] = -1
    with pytest.raises(nb.NlosError, match='out of range'):
        renderer.renderStreamedTransient(o, n, v, bad, 512, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    # device-resident faces: the build kernels clamp (no out-of-bounds access) and the call fails where it synchronises
    d_bad = torch.from_numpy(bad).cuda()
    with pytest.raises(nb.NlosError, match='out of range'):
        renderer.renderStreamedTransient(o, n, v, d_bad, 512, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    # and the context is usable afterwards
    renderer.renderStreamedTransient(o, n, v, f, 512, LB, UB, RES, T, pl, 1, 1, ctx=gpu_ctx)
    assert T.sum() > 0


def test_gradient_with_a_tap_table_beyond_48kb(oracle, gpu_ctx):
    """refine_scale * sigma_bin = 800 -> K = 3201 taps: 2 x 25.6 KB of prefix tables in the gradient kernel (opt-in shared memory)."""
    from nlos_surface_optimization_b200 import renderer, scenes
    o, n = scenes.wall_grid(3); v, f = scenes.fan8(); ns = 8 * 16
    data, weight = make_target(oracle, o, n, v, f, ns)
    refine, sigma = 160, 5                             # sigma >= 5: the forward pass is smoothed too (SSG.cpp:521-524)
    T_ref, G_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, refine, sigma, testing_flag=1, loss_flag=0)
    B = data.shape[1]
    T = np.zeros((o.shape[0], B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
    renderer.renderStreamedGradient(o, n, v, f, ns, LB, UB, RES, T, pl, G, data, weight, refine, sigma, 1, 0, ctx=gpu_ctx)
    assert rel_l2(T, T_ref) <= TOL_TRANSIENT and rel_l2(G, G_ref) <= TOL_GRADIENT
