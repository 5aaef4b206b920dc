"""One renderStreamedGradient step on C-bunny with a chosen forward algorithm (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import renderer, scenes
algo = int(sys.argv[1]) if len(sys.argv) > 1 else 2
gres = int(sys.argv[2]) if len(sys.argv) > 2 else 0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mesh = sys.argv[4] if len(sys.argv) > 4 else 'bunny'
ns = int(sys.argv[5]) if len(sys.argv) > 5 else 20000
ctx = nb.Context(0); dev = torch.device('cuda', 0)
o, n = scenes.wall_grid(64); v, f = getattr(scenes, mesh)()
L = o.shape[0]; B = 1200
to = lambda a: torch.from_numpy(a).to(dev)
d_o, d_n, d_v, d_f = to(o), to(n), to(v), to(f)
d_data = torch.zeros((L, B), dtype=torch.float64, device=dev); d_pl = torch.zeros(B, dtype=torch.float64, device=dev)
d_w = torch.ones((L, B), dtype=torch.float64, device=dev)
T = torch.zeros((L, B), dtype=torch.float64, device=dev); G = torch.zeros((v.shape[0], 3), dtype=torch.float64, device=dev)
ctx.set_option('forward_algo', algo); ctx.set_option('grid_res', gres); ctx.set_option('timing', 1)
for i in range(steps):
    renderer.renderStreamedGradient(d_o, d_n, d_v, d_f, ns, 0.0, 1.44, 1.2e-3, T, d_pl, G, d_data, d_w, 10, 1, 1, 0, ctx=ctx)
    ctx.synchronize(); print(ctx.timing())
