"""Generate tests/golden/*.npz with the CPU oracle (run by hand in the build container; takes minutes).

No golden vectors exist in the reference, so these pin the ORACLE's own output for the
headline configuration, letting the GPU path be checked at BASELINE.json's full size without re-running minutes
of CPU work on the GPU box.  Everything is derived from seeded inputs that ship with the repo.

    python tools/make_golden.py [c_bunny|c_bunny16|c_tiny|c_ggx16|c_ggx|c_scale|all]
"""
import os, sys, time, zlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from oracle import oracle
from nlos_surface_optimization_b200 import scenes

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')
LB, UB, RES = 0.0, 1.44, 1.2e-3
ROWS = np.array([0, 63, 777, 2016, 2080, 3000, 4032, 4095])


def vis_digest(vis):
    """per-source popcount and crc32 of the packed visibility bits."""
    L = vis.shape[0]
    flat = vis.reshape(L, -1)
    pop = flat.sum(axis=1).astype(np.int64)
    crc = np.array([zlib.crc32(np.packbits(flat[i]).tobytes()) for i in range(L)], dtype=np.uint32)
    return pop, crc


def c_bunny(wall=64, name='c_bunny'):
    v, f = scenes.bunny(); o, n = scenes.wall_grid(wall)
    ns = 20000
    t0 = time.time()
    v2 = v.copy(); v2[:, 2] += 0.01
    data = oracle.transient(o, n, v2, f, ns, LB, UB, RES)[0]
    print('data', time.time() - t0, flush=True)
    T, pl, vis = oracle.transient(o, n, v, f, ns, LB, UB, RES, want_visibility=True)
    print('fwd', time.time() - t0, flush=True)
    pop, crc = vis_digest(vis)
    weight = np.ones_like(data)
    T2, G, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, 1, 0)
    print('grad', time.time() - t0, flush=True)
    assert np.array_equal(T, T2)
    rows = ROWS[ROWS < o.shape[0]]
    np.savez_compressed(os.path.join(OUT, name + '.npz'), wall=wall, num_sample=ns, seed=oracle.DEFAULT_SEED,
                        transient_row_sum=T.sum(1), transient_col_sum=T.sum(0), transient_sq_sum=(T * T).sum(), rows=rows, transient_rows=T[rows],
                        data_row_sum=data.sum(1), data_col_sum=data.sum(0), vis_pop=pop, vis_crc=crc, gradient=G.astype(np.float64),
                        transient_nnz=(T > 0).sum(1))


def c_tiny():
    v, f = scenes.fan8(); o, n = scenes.wall_grid(4)
    ns = 8 * 64
    v2 = v.copy(); v2[:, 2] += 0.01
    data = oracle.transient(o, n, v2, f, ns, LB, UB, RES)[0]; weight = np.ones_like(data)
    T, pl, vis = oracle.transient(o, n, v, f, ns, LB, UB, RES, want_visibility=True)
    Ts = oracle.transient(o, n, v, f, ns, LB, UB, RES, 10, 1)[0]
    _, G, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, 1, 0)
    np.savez_compressed(os.path.join(OUT, 'c_tiny.npz'), transient=T, transient_smoothed=Ts, pathlengths=pl, vis=np.packbits(vis), gradient=G, data=data)


def c_ggx(wall=64, name='c_ggx'):
    """C-ggx (SURVEY 8d / BASELINE configs[2], exp_ggx/test10.py:37,60): the bunny, GGX alpha = 0.1, data rendered at alpha = 0.2;
    vertex gradient with testing_flag 1 (face normals) and 0 (shading normals: the normal-variation term is on), alpha scalar."""
    v, f = scenes.bunny(); o, n = scenes.wall_grid(wall)
    vn = scenes.vertex_normals(v, f)
    ns = 20000
    t0 = time.time()
    data = oracle.transient(o, n, v, f, ns, LB, UB, RES, alpha=0.2)[0]; weight = np.ones_like(data)
    print('data', time.time() - t0, flush=True)
    T, G1, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, 1, 0, alpha=0.1)
    print('grad tf=1', time.time() - t0, flush=True)
    Tn, G0, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, 0, 0, alpha=0.1, vertex_normal=vn)
    print('grad tf=0', time.time() - t0, flush=True)
    _, ga = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, alpha=0.1, kind=2)
    print('alpha', time.time() - t0, flush=True)
    rows = ROWS[ROWS < o.shape[0]]
    np.savez_compressed(os.path.join(OUT, name + '.npz'), wall=wall, num_sample=ns, alpha=0.1, alpha_data=0.2,
                        transient_row_sum=T.sum(1), transient_col_sum=T.sum(0), rows=rows, transient_rows=T[rows],
                        shading_transient_row_sum=Tn.sum(1), shading_transient_col_sum=Tn.sum(0),
                        gradient_tf1=G1, gradient_tf0_shading=G0, alpha_grad=ga, data_row_sum=data.sum(1), data_col_sum=data.sum(0))


def c_scale(wall=16, name='c_scale16'):
    """C-scale (BASELINE configs[4]) at its full MESH size: height field F = 500 000, B = 2048, on a 16x16 wall (256 of the 65 536 wall
    points; the oracle needs minutes for these).  Visibility digests per source, transient marginals, and the gradient as norms + every
    64th vertex row (the full [V,3] array would be 6 MB)."""
    v, f = scenes.heightfield(501); o, n = scenes.wall_grid(wall)
    ns = f.shape[0]; ub = 2048 * RES
    t0 = time.time()
    v2 = v.copy(); v2[:, 2] += 0.01
    data = oracle.transient(o, n, v2, f, ns, LB, ub, RES)[0]
    print('data', time.time() - t0, flush=True)
    T, pl, vis = oracle.transient(o, n, v, f, ns, LB, ub, RES, want_visibility=True)
    pop, crc = vis_digest(vis)
    print('fwd', time.time() - t0, flush=True)
    T2, G, _ = oracle.gradient(o, n, v, f, ns, LB, ub, RES, data, np.ones_like(data), 10, 1, 1, 0)
    print('grad', time.time() - t0, flush=True)
    assert np.array_equal(T, T2)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), wall=wall, num_sample=ns, numbins=2048, transient_row_sum=T.sum(1), transient_col_sum=T.sum(0),
                        transient_sq_sum=(T * T).sum(), data_row_sum=data.sum(1), data_col_sum=data.sum(0), vis_pop=pop, vis_crc=crc,
                        gradient_rows=G[::64].astype(np.float64), gradient_l2=np.linalg.norm(G), gradient_col_sum=G.sum(0), gradient_abs_sum=np.abs(G).sum(),
                        gradient_block_sum=np.add.reduceat(G, np.arange(0, G.shape[0], 1024), axis=0))


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    os.makedirs(OUT, exist_ok=True)
    if which in ('c_tiny', 'all'): c_tiny()
    if which in ('c_ggx16', 'all'): c_ggx(16, 'c_ggx16')
    if which in ('c_ggx', 'all'): c_ggx()
    if which in ('c_scale', 'all'): c_scale()
    if which in ('c_bunny16', 'all'): c_bunny(16, 'c_bunny16')
    if which in ('c_bunny', 'all'): c_bunny()
