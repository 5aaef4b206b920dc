"""One-GPU driver for ncu captures: N device-resident renderStreamedGradient steps of the C-bunny workload."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import renderer, scenes

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wall = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = nb.Context(0)
dev = torch.device('cuda', 0)
o, n = scenes.wall_grid(wall); v, f = scenes.bunny()
order = sys.argv[3] if len(sys.argv) > 3 else 'row'
if order != 'row':        # experiment: feed the wall points in a spatially compact order (tiles / Morton) instead of row-major
    iy, ix = np.divmod(np.arange(wall * wall), wall)
    if order == 'random':
        key = np.random.RandomState(0).permutation(wall * wall)
    elif order == 'stride':
        key = (np.arange(wall * wall) % 64) * 64 + np.arange(wall * wall) // 64      # chunk = one column = 64 points spread over y
    elif order == 'scatter':
        key = (ix % 8) * 512 + (iy % 8) * 64 + (iy // 8) * 8 + ix // 8                    # chunk = an 8x8 lattice spread over the whole wall
    elif order == 'tile8':
        key = ((iy // 8) * (wall // 8) + ix // 8) * 64 + (iy % 8) * 8 + ix % 8
    else:
        def spread(a):
            r = np.zeros_like(a)
            for b in range(8):
                r |= ((a >> b) & 1) << (2 * b)
            return r
        key = spread(ix) | (spread(iy) << 1)
    perm = np.argsort(key, kind='stable'); o = np.ascontiguousarray(o[perm]); n = np.ascontiguousarray(n[perm])
L = o.shape[0]; B = 1200
to = lambda a: torch.from_numpy(a).to(dev)
d_o, d_n, d_v, d_f = to(o), to(n), to(v), to(f)
v2 = v.copy(); v2[:, 2] += 0.01
d_data = torch.zeros((L, B), dtype=torch.float64, device=dev); d_pl = torch.zeros(B, dtype=torch.float64, device=dev)
renderer.renderStreamedTransient(d_o, d_n, to(v2), d_f, 20000, 0.0, 1.44, 1.2e-3, d_data, d_pl, 1, 1, ctx=ctx)
d_w = torch.ones((L, B), dtype=torch.float64, device=dev)
d_T = torch.zeros((L, B), dtype=torch.float64, device=dev); d_G = torch.zeros((v.shape[0], 3), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
ctx.set_option('timing', 1)
for i in range(steps):
    renderer.renderStreamedGradient(d_o, d_n, d_v, d_f, 20000, 0.0, 1.44, 1.2e-3, d_T, d_pl, d_G, d_data, d_w, 10, 1, 1, 0, ctx=ctx)
    ctx.synchronize()
    print(i, ctx.timing())
