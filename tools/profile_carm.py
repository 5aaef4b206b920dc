import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import renderer, scenes
ctx = nb.Context(0); dev = torch.device('cuda', 0)
o, n = scenes.wall_grid(64); v, f = scenes.armadillo_init()
L = o.shape[0]; B = 1200
to = lambda a: torch.from_numpy(a).to(dev)
d_o, d_n, d_v, d_f = to(o), to(n), to(v), to(f)
d_data = torch.zeros((L, B), dtype=torch.float64, device=dev); d_pl = torch.zeros(B, dtype=torch.float64, device=dev)
d_w = torch.ones((L, B), dtype=torch.float64, device=dev)
d_T = torch.zeros((L, B), dtype=torch.float64, device=dev); d_G = torch.zeros((v.shape[0], 3), dtype=torch.float64, device=dev)
ctx.set_option('timing', 1)
for i in range(3):
    renderer.renderStreamedGradient(d_o, d_n, d_v, d_f, 20000, 0.0, 1.44, 1.2e-3, d_T, d_pl, d_G, d_data, d_w, 10, 1, 1, 0, ctx=ctx)
    ctx.synchronize(); print(i, ctx.timing())
