"""GPU probe: gradient-kernel time against the sources-per-block chunk (option chunk_gradient; 0 = the library's automatic choice).
   python tools/sweep_chunk.py [bunny|armadillo_init|...] [chunk list]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import renderer, scenes
mesh = sys.argv[1] if len(sys.argv) > 1 else 'bunny'
chunks = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 32, 64, 128, 256]
ctx = nb.Context(0); dev = torch.device('cuda', 0)
o, n = scenes.wall_grid(64); v, f = getattr(scenes, mesh)(); L = o.shape[0]; B = 1200
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_o, d_n, d_v, d_f = to(o), to(n), to(v), to(f)
v2 = v.copy(); v2[:, 2] += 0.01
d_data = torch.zeros((L, B), dtype=torch.float64, device=dev); d_pl = torch.zeros(B, dtype=torch.float64, device=dev)
renderer.renderStreamedTransient(d_o, d_n, to(v2), d_f, 20000, 0.0, 1.44, 1.2e-3, d_data, d_pl, 1, 1, ctx=ctx)
d_w = torch.ones((L, B), dtype=torch.float64, device=dev); d_T = torch.zeros((L, B), dtype=torch.float64, device=dev); d_G = torch.zeros((v.shape[0], 3), dtype=torch.float64, device=dev)
torch.cuda.synchronize(); ctx.set_option('timing', 1)
ref = None
for cg in chunks:
    ctx.set_option('chunk_gradient', cg)
    best = 1e9
    for i in range(5):
        d_T.zero_(); d_G.zero_()
        renderer.renderStreamedGradient(d_o, d_n, d_v, d_f, 20000, 0.0, 1.44, 1.2e-3, d_T, d_pl, d_G, d_data, d_w, 10, 1, 1, 0, ctx=ctx); ctx.synchronize()
        best = min(best, ctx.timing()['gradient_ms'])
    g = d_G.cpu().numpy()
    if ref is None: ref = g
    print('%s F=%d chunk_gradient %4d  gradient %.3f ms  forward %.3f ms | gradient rel diff vs first %.2e' % (mesh, f.shape[0], cg, best, ctx.timing()['forward_ms'], np.linalg.norm(g - ref) / max(np.linalg.norm(ref), 1e-300)), flush=True)
