import sys, os, zlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import scenes
from oracle import oracle
g = np.load('tests/golden/c_bunny.npz')
v, f = scenes.bunny(); o, n = scenes.wall_grid(64)
ctx = nb.default_context(0)
L = o.shape[0]; bad = []
for a in range(0, L, 256):
    ctx.set_source_window(a, L)
    vis, _ = nb.debug_visibility(np.ascontiguousarray(o[a:a+256]), v, f, 20000, ctx=ctx)
    pop = vis.reshape(256, -1).sum(1)
    d = np.flatnonzero(pop != g['vis_pop'][a:a+256])
    for i in d:
        bad.append((a + i, vis[i, :, 0].copy()))
print('sources with popcount mismatch:', [b[0] for b in bad])
for s, gv in bad[:6]:
    ov = oracle.transient(o[s:s+1], n[s:s+1], v, f, 20000, 0, 1.44, 1.2e-3, want_visibility=True, src_offset=int(s))[2][0, :, 0]
    bv = oracle.transient(o[s:s+1], n[s:s+1], v, f, 20000, 0, 1.44, 1.2e-3, want_visibility=True, src_offset=int(s), brute=True)[2][0, :, 0]
    print('source', s, 'gpu!=oracle_bvh at tris', np.flatnonzero(gv != ov), 'gpu!=brute', np.flatnonzero(gv != bv), 'oracle_bvh!=brute', np.flatnonzero(ov != bv))
