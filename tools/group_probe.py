"""GPU probe: the shared-grid forward kernel (forward_algo 3) against the BVH kernel (1) and the per-point grid (2):
visibility words compared bit for bit, timings per phase.  python tools/group_probe.py WALL MESH NS "side,K,G;side,K,G;..." """
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import renderer, scenes

wall = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mesh = sys.argv[2] if len(sys.argv) > 2 else 'bunny'
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
variants = [tuple(int(x) for x in v.split(',')) for v in (sys.argv[4] if len(sys.argv) > 4 else '4,16,0').split(';')]
ctx = nb.Context(0)
dev = torch.device('cuda', 0)
o, n = scenes.wall_grid(wall); v, f = getattr(scenes, mesh)()
L = o.shape[0]; B = 1200
to = lambda a: torch.from_numpy(a).to(dev)
d_o, d_n, d_v, d_f = to(o), to(n), to(v), to(f)
v2 = v.copy(); v2[:, 2] += 0.01
d_data = torch.zeros((L, B), dtype=torch.float64, device=dev); d_pl = torch.zeros(B, dtype=torch.float64, device=dev)
ctx.set_option('forward_algo', 1)
renderer.renderStreamedTransient(d_o, d_n, to(v2), d_f, ns, 0.0, 1.44, 1.2e-3, d_data, d_pl, 1, 1, ctx=ctx)
d_w = torch.ones((L, B), dtype=torch.float64, device=dev)
ctx.set_option('timing', 1)

def run(algo, side=0, K=0, G=0, steps=3):
    ctx.set_option('forward_algo', algo); ctx.set_option('grid_res', G); ctx.set_option('group_side', side); ctx.set_option('grid_slices', K)
    T = torch.zeros((L, B), dtype=torch.float64, device=dev); Gr = torch.zeros((v.shape[0], 3), dtype=torch.float64, device=dev)
    best = None
    for i in range(steps):
        T.zero_(); Gr.zero_()
        renderer.renderStreamedGradient(d_o, d_n, d_v, d_f, ns, 0.0, 1.44, 1.2e-3, T, d_pl, Gr, d_data, d_w, 10, 1, 1, 0, ctx=ctx)
        ctx.synchronize()
        t = ctx.timing()
        if best is None or t['forward_ms'] < best['forward_ms']: best = t
    return T.cpu().numpy(), Gr.cpu().numpy(), ctx.visibility_words().copy(), best

T0, G0, w0, t0 = run(1)
print('bvh        forward %.3f ms gradient %.3f total %.3f' % (t0['forward_ms'], t0['gradient_ms'], t0['total_ms']), flush=True)
for a in sys.argv:
    if a.startswith('--g2='):
        for g in [int(x) for x in a[5:].split(',')]:
            T1, G1, w1, t1 = run(2, 0, 0, g)
            print('grid G=%3d  forward %.3f ms gradient %.3f total %.3f | words equal %s' % (g, t1['forward_ms'], t1['gradient_ms'], t1['total_ms'], np.array_equal(w0, w1)), flush=True)
if '--no2' not in sys.argv:
    T1, G1, w1, t1 = run(2)
    print('grid       forward %.3f ms gradient %.3f total %.3f | words equal %s' % (t1['forward_ms'], t1['gradient_ms'], t1['total_ms'], np.array_equal(w0, w1)), flush=True)
for side, K, G in variants:
    T1, G1, w1, t1 = run(3, side, K, G)
    et = np.linalg.norm(T1 - T0) / np.linalg.norm(T0); eg = np.linalg.norm(G1 - G0) / np.linalg.norm(G0)
    print('group side=%d K=%2d G=%3d forward %.3f ms gradient %.3f total %.3f | words equal %s (%d differ) transient rel %.2e gradient rel %.2e' % (
        side, K, G, t1['forward_ms'], t1['gradient_ms'], t1['total_ms'], np.array_equal(w0, w1), int((w0 != w1).sum()), et, eg), flush=True)
