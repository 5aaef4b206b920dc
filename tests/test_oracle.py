"""CPU tests of the oracle itself (oracle/nlos_oracle.cpp): known-answer RNG vectors, closed-form renders,
brute-force-vs-BVH agreement, smoothing conservation, finite differences, and the committed golden fixtures.
These tests pin the oracle against mathematics; tests/test_reference_pin.py pins it (statistically) against the
reference's own code compiled with stand-in library headers (oracle/_ref)."""
import os
import numpy as np
import pytest
from helpers import LB, UB, RES, rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _u2f(x):
    return np.uint32((x >> 9) | 0x3f800000).view(np.float32) - np.float32(1.0)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10
    S, T = oracle.philox_st(0, 0, 0, 0)
    assert (S, T) == (float(_u2f(0x6627e8d5)), float(_u2f(0xe169c58d)))
    S, T = oracle.philox_st(0, 0, 0, 1)
    assert (S, T) == (float(_u2f(0xbc57ac4c)), float(_u2f(0x9b00dbd8)))
    S, T = oracle.philox_st(0xffffffffffffffff, -1, -1, -2)
    assert (S, T) == (float(_u2f(0x408f276d)), float(_u2f(0x41c83b0e)))
    S, T = oracle.philox_st(0xffffffffffffffff, -1, -1, -1)
    assert (S, T) == (float(_u2f(0xa20bc7c6)), float(_u2f(0x6d5451fd)))
    # key = (a4093822, 299f31d0), ctr = (243f6a88, 85a308d3, 13198a2e, 03707344): tri=c0, src=c1|c3<<32, k>>1=c2
    seed = 0x299f31d0a4093822; src = (0x03707344 << 32) | 0x85a308d3
    S, T = oracle.philox_st(seed, src, 0x243f6a88, 0x13198a2e * 2)
    assert (S, T) == (float(_u2f(0xd16cfe09)), float(_u2f(0x94fdcceb)))
    S, T = oracle.philox_st(seed, src, 0x243f6a88, 0x13198a2e * 2 + 1)
    assert (S, T) == (float(_u2f(0x5001e420)), float(_u2f(0x24126ea1)))


def test_uniforms_are_in_unit_interval_and_uniform(oracle):
    st = np.array([oracle.philox_st(5489, s, t, k) for s in range(8) for t in range(64) for k in range(4)])
    assert st.min() >= 0.0 and st.max() < 1.0
    assert abs(st.mean() - 0.5) < 0.02 and abs(st.var() - 1 / 12) < 0.01


def test_intersection_against_double_precision(oracle):
    rng = np.random.RandomState(1)
    hits = 0
    for _ in range(400):
        tri = rng.randn(3, 3); tri[:, 2] += 4
        o = rng.randn(3) * 0.2
        bary = rng.rand() < 0.7
        w = rng.dirichlet([1, 1, 1]) if bary else rng.randn(3)
        target = w @ tri
        d = target - o; d /= np.linalg.norm(d)
        hit, tuv = oracle.isect(tri, o, d)
        inside = bary and (w >= 1e-4).all()
        if inside:
            assert hit
        if hit:
            hits += 1
            t, u, v = [float(x) for x in tuv]
            t32 = tri.astype(np.float32).astype(np.float64)
            p_bary = (1 - u - v) * t32[0] + u * t32[1] + v * t32[2]
            p_ray = o.astype(np.float32).astype(np.float64) + t * d.astype(np.float32).astype(np.float64)
            assert np.linalg.norm(p_bary - p_ray) < 2e-5 * max(1.0, t)
            assert -1e-6 <= u <= 1 + 1e-6 and -1e-6 <= v <= 1 + 1e-6 and u + v <= 1 + 1e-6
    assert hits > 200


def test_plane_closed_form_integral(oracle):
    """8-triangle fan at z=.38 (exp_bunny/weight_test.py:84-85): sum_b T[s,b] -> integral of z^4/r^8 dA."""
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.fan8(); o, n = scenes.wall_grid(3)
    T = oracle.transient(o, n, v, f, 8 * 20000, LB, UB, RES)[0]
    z = 0.38; xs = (np.arange(2000) + 0.5) / 2000 * 0.5 - 0.25
    X, Y = np.meshgrid(xs, xs)
    for s in range(o.shape[0]):
        r2 = (X - o[s, 0]) ** 2 + (Y - o[s, 1]) ** 2 + z * z
        exact = (z ** 4 / r2 ** 4).mean() * 0.25
        assert abs(T[s].sum() - exact) / exact < 0.01
    # first photon cannot arrive before 2*z
    first = np.argmax(T > 0, axis=1)
    assert (first >= int(2 * z / RES) - 1).all()


def _numpy_forward_no_occlusion(oracle, o, n_o, v, f, num_sample, seed=5489):
    """Independent float32 NumPy restatement of the forward task (TG.cpp:184-232) for scenes with no occluders."""
    F = f.shape[0]; spp = 1 + (num_sample - 1) // F
    B = oracle.num_bins(LB, UB, RES)
    T = np.zeros((o.shape[0], B))
    f32 = np.float32
    for s in range(o.shape[0]):
        for t in range(F):
            v1, v2, v3 = v[f[t, 0]], v[f[t, 1]], v[f[t, 2]]
            N = np.cross((v2 - v1).astype(np.float64), (v3 - v1).astype(np.float64)); A = np.linalg.norm(N) / 2; nf = N / (2 * A)
            for k in range(spp):
                S, Tt = oracle.philox_st(seed, s, t, k)
                sq = np.sqrt(f32(Tt)); u = f32(1) - sq; vv = (f32(1) - f32(S)) * sq; w = f32(S) * sq
                p = u.astype(np.float64) * v1 + vv.astype(np.float64) * v2 + w.astype(np.float64) * v3
                q = p - o[s]; r = np.linalg.norm(q); d = q / r
                if not (LB / 2 <= r <= UB / 2):
                    continue
                ff = max(0.0, -(nf @ d) * (n_o[s] @ d) / r / r)
                b = int(np.floor((2 * r - LB) / RES))
                T[s, b] += A * ff * ff / spp
    return T


def test_forward_matches_independent_numpy_restatement(oracle):
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.fan8(); o, n = scenes.wall_grid(2)
    T = oracle.transient(o, n, v, f, 8 * 16, LB, UB, RES)[0]
    T_np = _numpy_forward_no_occlusion(oracle, o, n, v, f, 8 * 16)
    # float32-vs-float64 rounding may move a handful of samples across a bin edge; compare cumulative curves
    assert abs(T.sum() - T_np.sum()) / T_np.sum() < 1e-5
    c1, c2 = np.cumsum(T, axis=1), np.cumsum(T_np, axis=1)
    assert np.abs(c1 - c2).max() / c2.max() < 5e-3
    assert rel_l2(T, T_np) < 0.05


def test_occlusion_front_quad_hides_back_quad(oracle):
    from nlos_surface_optimization_b200 import scenes
    front = scenes.quad(0.30, 0.05); back = scenes.quad(0.50, 0.20)
    v, f = scenes.merge([front, back]); o, n = scenes.wall_grid(3)
    spp = 128
    T, pl, vis = oracle.transient(o, n, v, f, f.shape[0] * spp, LB, UB, RES, want_visibility=True)
    assert vis[:, :2, :].all()                          # nothing in front of the front quad
    # a back-quad sample is hidden iff the segment origin->sample crosses the front quad (checked in float64)
    checked = 0
    for s in range(o.shape[0]):
        for t in (2, 3):
            v1, v2, v3 = [v[i].astype(np.float64) for i in f[t]]
            for k in range(spp):
                S, Tt = oracle.philox_st(5489, s, t, k)
                sq = np.sqrt(np.float32(Tt)); u = 1 - sq; vv = (1 - np.float32(S)) * sq; w = np.float32(S) * sq
                p = float(u) * v1 + float(vv) * v2 + float(w) * v3
                lam = 0.30 / p[2]
                x = o[s, 0] + lam * (p[0] - o[s, 0]); y = o[s, 1] + lam * (p[1] - o[s, 1])
                m = max(abs(x), abs(y))
                if abs(m - 0.05) < 1e-4:
                    continue                             # too close to the occluder's edge to call in float64
                assert bool(vis[s, t, k]) == (m > 0.05)
                checked += 1
    assert checked > 1000
    # and hidden samples contribute nothing: re-render without the front quad and compare far-range energy
    T_back = oracle.transient(o, n, back[0], back[1], 2 * spp, LB, UB, RES)[0]
    assert T[:, 700:].sum() < T_back[:, 700:].sum()


@pytest.mark.parametrize('mesh', ['ico', 'quads'])
def test_bvh_equals_brute_force(mesh, oracle):
    from nlos_surface_optimization_b200 import scenes
    if mesh == 'ico':
        v, f = scenes.icosphere(3, 0.1, (0.01, 0.02, 0.45), noise=0.05, seed=2)
    else:
        v, f = scenes.merge([scenes.quad(0.4, 0.1), scenes.quad(0.5, 0.2), scenes.quad(0.45, 0.03, 0.1, 0.1)])
    o, n = scenes.wall_grid(4)
    a = oracle.transient(o, n, v, f, 5000, LB, UB, RES, want_visibility=True)
    b = oracle.transient(o, n, v, f, 5000, LB, UB, RES, want_visibility=True, brute=True)
    assert np.array_equal(a[2], b[2])
    assert np.array_equal(a[0], b[0])


def test_smoothing_conserves_energy_and_matches_kernel(oracle):
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.quad(0.4, 0.1); o, n = scenes.wall_grid(2)      # well inside the time range: no taps fall off the ends
    T1 = oracle.transient(o, n, v, f, 512, LB, UB, RES, 1, 1)[0]
    T10 = oracle.transient(o, n, v, f, 512, LB, UB, RES, 10, 1)[0]
    w, sigma2 = oracle.taps(RES, 10, 1)
    assert len(w) == 41 and abs(w.sum() - 0.9999987) < 1e-6
    assert abs(T10.sum() / T1.sum() - w.sum()) < 1e-9
    assert abs(np.sqrt(sigma2) - RES / 2.355) < 1e-9
    # smoothing spreads each coarse bin over at most 2 neighbours on each side
    nz1 = np.flatnonzero(T1[0] > 0); nz10 = np.flatnonzero(T10[0] > 0)
    assert nz10.min() >= nz1.min() - 3 and nz10.max() <= nz1.max() + 3


def test_pathlengths_and_numbins(oracle):
    assert oracle.num_bins(0, 1200 * 1.2e-3, 1.2e-3) == 1200
    assert oracle.num_bins(0, 2048 * 1.2e-3, 1.2e-3) == 2048
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.fan8(); o, n = scenes.wall_grid(1)
    pl = oracle.transient(o, n, v, f, 8, 0.1, 1.3, 5e-3)[1]
    expect = (np.float32(0.1) + np.arange(len(pl)).astype(np.float32) * np.float32(5e-3)).astype(np.float64)
    assert np.array_equal(pl, expect)


def _surrogate_energy(oracle, o, n, v, f, ns, D, r, s, vn):
    """E(v) = -(2/L) sum_{s,b} D[s,b] * T~[s,b], T~ = the (r,s)-smoothed forward: the functional whose vertex derivative the
    reference's gradient computes when diff == D (TG.cpp:972-980 differentiates the Gaussian splat)."""
    T = oracle.transient(o, n, v, f, ns, LB, UB, RES, r, s, vertex_normal=vn)[0]
    return -2.0 * (D * T).sum() / o.shape[0]


@pytest.mark.parametrize('with_gn', [False, True])
def test_vertex_gradient_finite_difference(with_gn, oracle):
    """Whole-gradient FD on the consistent surrogate: flat fan, shading normals == face normals,
    diff pinned to a smooth D by choosing data = T + D.  with_gn=False: normals held fixed (testing_flag=1 drops gn);
    with_gn=True: face normals follow the vertices and the gradient includes the normal-variation term gn."""
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.fan8(); o, n = scenes.wall_grid(3)
    vn = np.ascontiguousarray(np.tile(np.array([0, 0, -1], dtype=np.float32), (v.shape[0], 1)))
    ns = 8 * 4000; r, s = 10, 1
    B = oracle.num_bins(LB, UB, RES)
    b = np.arange(B)
    D = np.stack([np.exp(-0.5 * ((b - (640 + 7 * k)) / 25.0) ** 2) * (1 + 0.1 * k) for k in range(o.shape[0])])
    T0 = oracle.transient(o, n, v, f, ns, LB, UB, RES, vertex_normal=vn)[0]
    data = T0 + D; weight = np.ones_like(D)
    _, G, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, r, s, testing_flag=0 if with_gn else 1, vertex_normal=vn)
    h = 2e-3
    checked = 0
    for vi, ax in ((8, 2), (8, 0), (4, 2), (5, 1)):
        vp = v.copy(); vp[vi, ax] += h
        vm = v.copy(); vm[vi, ax] -= h
        vn_fd = None if with_gn else vn       # with_gn: the forward recomputes face normals from the moved vertices
        fd = (_surrogate_energy(oracle, o, n, vp, f, ns, D, r, s, vn_fd) - _surrogate_energy(oracle, o, n, vm, f, ns, D, r, s, vn_fd)) / (2 * h)
        scale = np.abs(G).max()
        assert abs(G[vi, ax] - fd) < 0.05 * scale + 0.08 * abs(fd), (vi, ax, G[vi, ax], fd)
        checked += 1
    assert checked == 4 and np.abs(G).max() > 0


def test_gradient_is_linear_in_residual_and_accumulates(oracle):
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.icosphere(2, 0.1, (0, 0, 0.45)); o, n = scenes.wall_grid(2)
    ns = 2000
    T0 = oracle.transient(o, n, v, f, ns, LB, UB, RES)[0]
    rng = np.random.RandomState(0); D = rng.rand(*T0.shape); w = np.ones_like(D)
    _, G1, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, T0 + D, w, 10, 1)
    _, G3, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, T0 + 3 * D, w, 10, 1)
    assert rel_l2(G3, 3 * G1) < 1e-6
    _, Gw, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, T0 + D, 2 * w, 10, 1)
    assert rel_l2(Gw, 2 * G1) < 1e-6
    Gacc = G1.copy()
    oracle.gradient(o, n, v, f, ns, LB, UB, RES, T0 + D, w, 10, 1, gradient_inout=Gacc)
    assert rel_l2(Gacc, 2 * G1) < 1e-12
    # loss_flag = 1: diff = 2 d^3 w  (SSG.cpp:546-549)
    _, Gq, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, T0 + D, w, 10, 1, loss_flag=1)
    _, Gq_ref, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, T0 + 2 * D ** 3, w, 10, 1)
    assert rel_l2(Gq, Gq_ref) < 1e-6


def test_albedo_scalar_gradient_is_derivative_of_uniform_albedo(oracle):
    """d/d(albedo) of -2/L sum D T~ with T~ linear in a uniform albedo equals the albedo-scalar kernel at albedo=1."""
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.icosphere(2, 0.1, (0, 0, 0.45)); o, n = scenes.wall_grid(2)
    ns = 4000
    alb = np.ones(v.shape[0], dtype=np.float32)
    T0 = oracle.transient(o, n, v, f, ns, LB, UB, RES, vertex_albedo=alb)[0]
    rng = np.random.RandomState(1); D = rng.rand(*T0.shape)
    _, g = oracle.gradient(o, n, v, f, ns, LB, UB, RES, T0 + D, np.ones_like(D), 10, 1, vertex_albedo=alb, kind=1)
    Ts = oracle.transient(o, n, v, f, ns, LB, UB, RES, 10, 1, vertex_albedo=alb)[0]
    assert abs(g - (-2.0 * (D * Ts).sum() / o.shape[0])) < 1e-5 * abs(g)


@pytest.mark.parametrize('alpha', [0.1, 0.35, 0.8])
def test_ggx_derivatives_finite_difference(alpha, oracle):
    rng = np.random.RandomState(3)
    for _ in range(20):
        n = rng.randn(3); n /= np.linalg.norm(n)
        w = n + 0.6 * rng.randn(3); w /= np.linalg.norm(w)
        if n @ w < 0.15 or n @ w > 0.98:
            continue
        f0 = oracle.ggx(0, alpha, n, w)
        assert f0 > 0
        h = 1e-3
        fd_a = (oracle.ggx(0, alpha + h, n, w) - oracle.ggx(0, alpha - h, n, w)) / (2 * h)
        an = oracle.ggx(1, alpha, n, w)
        assert abs(an - fd_a) < 2e-2 * max(abs(fd_a), abs(f0))
        dn, dw = oracle.ggx(2, alpha, n, w)
        for ax in range(3):
            e = np.zeros(3); e[ax] = 2e-3
            fd_n = (oracle.ggx(0, alpha, n + e, w) - oracle.ggx(0, alpha, n - e, w)) / 4e-3
            fd_w = (oracle.ggx(0, alpha, n, w + e) - oracle.ggx(0, alpha, n, w - e)) / 4e-3
            tol = 3e-2 * max(np.abs(dn).max(), np.abs(dw).max(), 1e-3)
            assert abs(dn[ax] - fd_n) < tol and abs(dw[ax] - fd_w) < tol


def test_ggx_forward_is_lambert_times_brdf_on_a_plane(oracle):
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.fan8(); o, n = scenes.wall_grid(2)
    Tl = oracle.transient(o, n, v, f, 800, LB, UB, RES)[0]
    Tg = oracle.transient(o, n, v, f, 800, LB, UB, RES, alpha=0.5)[0]
    nz = Tl > 0
    assert np.array_equal(nz, Tg > 0)
    ratio = Tg[nz] / Tl[nz]
    assert ratio.min() > 0 and np.isfinite(ratio).all()


def test_intensity_and_regularisers(oracle):
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.icosphere(2, 0.1, (0, 0, 0.45)); o, n = scenes.wall_grid(3)
    I = oracle.intensity(o, n, v, f, 2000, LB, UB)
    T = oracle.transient(o, n, v, f, 2000, LB, UB, RES)[0]
    assert abs(I.sum() - T.sum()) < 1e-9 * T.sum()          # same samples, unbinned
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    nrm = np.cross(b - a, c - a); cen = (a + b + c) / 3
    facing = np.stack([((cen - o[s]) * nrm).sum(1) < 0 for s in range(o.shape[0])])     # [L,F] front-facing w.r.t. source s
    assert (I[~facing.any(0)] == 0).all()               # faces that look away from every wall point get nothing
    assert (I[facing.all(0)] > 0).mean() > 0.9
    aff = scenes.face_affinity(f)
    val, G = oracle.normal_smoothing(v, f, aff)
    assert val > 0 and np.isfinite(G).all()
    val_flat, G_flat = oracle.normal_smoothing(*scenes.fan8(), scenes.face_affinity(scenes.fan8()[1]))
    assert abs(val_flat) < 1e-6 and np.abs(G_flat).max() < 1e-6
    C = oracle.curvature_grad(v, f)
    assert np.isfinite(C).all() and np.abs(C).max() > 0


def test_vertex_gradient_per_bin_sums_to_unit_residual_gradient(oracle):
    """renderStreamedVertexGradient (TG.cpp:697-840) = gradient per time bin of one vertex; summing it against diff == -1/2
    reproduces the gn-including vertex gradient for a single source."""
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.fan8(); o, n = scenes.wall_grid(1)
    vn = np.ascontiguousarray(np.tile(np.array([0, 0, -1], dtype=np.float32), (v.shape[0], 1)))
    ns = 8 * 500
    Gb = oracle.vertex_gradient(8, o, n, v, f, ns, LB, UB, RES, 10, 1)
    T0 = oracle.transient(o, n, v, f, ns, LB, UB, RES, vertex_normal=vn)[0]
    D = -0.5 * np.ones_like(T0)
    _, G, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, T0 + D, np.ones_like(D), 10, 1, testing_flag=0, vertex_normal=vn)
    assert rel_l2(Gb.sum(0), G[8]) < 1e-5


def test_golden_fixtures_pin_the_oracle(oracle):
    from nlos_surface_optimization_b200 import scenes
    g = np.load(os.path.join(GOLD, 'c_tiny.npz'))
    v, f = scenes.fan8(); o, n = scenes.wall_grid(4); ns = 8 * 64
    T, pl, vis = oracle.transient(o, n, v, f, ns, LB, UB, RES, want_visibility=True)
    assert np.array_equal(T, g['transient']) and np.array_equal(pl, g['pathlengths'])
    assert np.array_equal(np.packbits(vis), g['vis'])
    assert np.array_equal(oracle.transient(o, n, v, f, ns, LB, UB, RES, 10, 1)[0], g['transient_smoothed'])
    _, G, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, g['data'], np.ones_like(T), 10, 1, 1, 0)
    assert rel_l2(G, g['gradient']) < 1e-12               # OpenMP reduction order may differ between runs
    # two rows of the bunny fixture, rendered with the global source index as RNG key
    gb = np.load(os.path.join(GOLD, 'c_bunny16.npz'))
    bv, bf = scenes.bunny(); bo, bn = scenes.wall_grid(int(gb['wall']))
    for row, ref in zip(gb['rows'][:2], gb['transient_rows'][:2]):
        Tr = oracle.transient(bo[row:row + 1], bn[row:row + 1], bv, bf, int(gb['num_sample']), LB, UB, RES, src_offset=int(row))[0]
        assert np.array_equal(Tr[0], ref)


def test_ray_query_oracle_bvh_equals_brute_force(oracle):
    """embree_intersector restatement: nearest hit (primID,u,v) per ray; BVH and brute force agree, hit points lie on the rays."""
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.icosphere(3, 0.1, (0, 0, 0.45), noise=0.05, seed=4)
    rng = np.random.RandomState(2); N = 3000
    o = np.stack([rng.uniform(-.25, .25, N), rng.uniform(-.25, .25, N), np.zeros(N)], 1).astype(np.float32)
    d = ((v[rng.randint(0, v.shape[0], N)] + rng.normal(0, 0.02, (N, 3)) - o) * rng.uniform(0.5, 2, (N, 1))).astype(np.float32)
    b3, b1 = oracle.intersect(o, d, v, f)
    c3, c1 = oracle.intersect(o, d, v, f, brute=True)
    assert np.array_equal(b3, c3) and np.array_equal(b1, c1)
    hit = b1 >= 0
    assert 0.3 < hit.mean() < 1
    p = oracle.bary_to_world(v, f, b3)
    t = np.linalg.norm(p[hit] - o[hit], axis=1) / np.linalg.norm(d[hit], axis=1)
    assert np.abs(o[hit] + t[:, None] * d[hit] - p[hit]).max() < 1e-5


def test_jitter_kernel_restatement(oracle):
    """jitter/: a delta kernel reproduces the raw histogram, a shifted delta shifts it, energy is conserved, and the vertex
    gradient with (jitter_weight, jitter_grad) = a sampled Gaussian matches the Gaussian-tap gradient of smoothed_transient/."""
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.icosphere(2, 0.1, (0, 0, 0.45)); o, n = scenes.wall_grid(2); ns = 4000
    H = oracle.transient(o, n, v, f, ns, LB, UB, RES)[0]
    delta = np.zeros(9); delta[4] = 1.0
    assert np.array_equal(oracle.jitter_transient(o, n, v, f, ns, LB, UB, RES, delta, 4)[0], H)
    Ts = oracle.jitter_transient(o, n, v, f, ns, LB, UB, RES, delta, 1)[0]          # offset 1: T[b] = H[b + 1 - 4]
    assert np.array_equal(Ts[:, 3:], H[:, :-3])
    J, off = 41, 20
    x = np.arange(J) - off
    w = np.exp(-0.5 * (x / 2.0) ** 2); w /= w.sum()
    T = oracle.jitter_transient(o, n, v, f, ns, LB, UB, RES, w, off)[0]
    assert abs(T.sum() - H.sum()) < 1e-9 * H.sum()
    # whole-bin Gaussian taps: jitter gradient == Gaussian gradient built from the same taps (refine_scale=1 => taps on whole bins)
    s_bin = 10
    wg, sigma2 = oracle.taps(RES, 1, s_bin)                                           # K = 4*1*10+1 = 41 taps, delta_i = (i-20)*res
    delta_i = (np.arange(41) - 20) * np.float32(RES)
    jg = wg * delta_i / sigma2 * 2 * (-np.float32(RES) / 2.0)                         # so that jg*(-2)/res == w*delta/sigma^2*2
    D = np.random.RandomState(0).rand(*H.shape)
    # same residual for both: data = T_forward + D with each path's own forward
    Tj = oracle.jitter_transient(o, n, v, f, ns, LB, UB, RES, wg, 20)[0]
    _, Gj, _ = oracle.jitter_gradient(o, n, v, f, ns, LB, UB, RES, wg, jg, 20, Tj + D, np.ones_like(D))
    _, Gg, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, H + D, np.ones_like(D), 1, s_bin)       # sigma>=5 & r=1: forward raw
    assert rel_l2(Gj, Gg) < 1e-5


def test_canonical_traversal_counter(oracle):
    """the accounting traversal (Karras LBVH, one triangle per leaf, near child first, nearest hit) answers visibility exactly as
    the oracle's own query does, and its counts follow the closed form on a mesh where that is known."""
    from nlos_surface_optimization_b200 import scenes
    o, n = scenes.wall_grid(4)
    v, f = scenes.fan8()
    c = oracle.canonical_counts(o, v, f, 8 * 64)
    # 8 coplanar triangles, every ray hits exactly one: root-to-leaf descent of a 3-level tree tests 3 x 2 boxes and, with the boxes of
    # neighbouring triangles touching, at most a few extra; every sample is visible
    assert c['rays'] == o.shape[0] * 8 * 64 and c['visible_frac'] == 1.0
    assert 6.0 <= c['box_per_ray'] <= 12.0 and 1.0 <= c['tri_per_ray'] <= 3.0
    v, f = scenes.icosphere(3, 0.1, (0.02, -0.03, 0.45), noise=0.03, seed=3)
    ns = 2 * f.shape[0]
    vis = oracle.transient(o, n, v, f, ns, LB, UB, RES, want_visibility=True)[2]
    c = oracle.canonical_counts(o, v, f, ns)
    assert c['rays'] == vis.size and abs(c['visible_frac'] * c['rays'] - vis.sum()) < 0.5
    import math
    assert c['box_per_ray'] >= 2 * math.floor(math.log2(f.shape[0])) * 0.5 and c['tri_per_ray'] >= 1.0
