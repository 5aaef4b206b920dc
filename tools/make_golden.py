"""Generate tests/golden/*.npz with the CPU oracle (run by hand in the build container; takes minutes).

No golden vectors exist in the reference, so these pin the ORACLE's own output for the
headline configuration, letting the GPU path be checked at BASELINE.json's full size without re-running minutes
of CPU work on the GPU box.  Everything is derived from seeded inputs that ship with the repo.

    python tools/make_golden.py [c_bunny|c_bunny16|c_tiny|c_ggx|all]
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


def c_ggx():
    v, f = scenes.icosphere(4, 0.1, (0.02, -0.03, 0.45), noise=0.03, seed=3); o, n = scenes.wall_grid(8)
    ns = 20000
    data = oracle.transient(o, n, v, f, ns, LB, UB, RES, alpha=0.2)[0]; weight = np.ones_like(data)
    T, G, _ = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, 1, 0, alpha=0.1)
    _, ga = oracle.gradient(o, n, v, f, ns, LB, UB, RES, data, weight, 10, 1, alpha=0.1, kind=2)
    np.savez_compressed(os.path.join(OUT, 'c_ggx.npz'), transient_row_sum=T.sum(1), transient_col_sum=T.sum(0), gradient=G, alpha_grad=ga,
                        data_col_sum=data.sum(0))


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    os.makedirs(OUT, exist_ok=True)
    if which in ('c_tiny', 'all'): c_tiny()
    if which in ('c_ggx', 'all'): c_ggx()
    if which in ('c_bunny16', 'all'): c_bunny(16, 'c_bunny16')
    if which in ('c_bunny', 'all'): c_bunny()
