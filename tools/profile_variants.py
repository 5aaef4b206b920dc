"""Forward/gradient time of the non-default kernel instantiations on the C-bunny workload (shading normals, albedo, GGX, smoothed forward)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import renderer, ggx, scenes
ctx = nb.Context(0); dev = torch.device('cuda', 0)
o, n = scenes.wall_grid(64); v, f = scenes.bunny(); L = o.shape[0]; B = 1200
vn = scenes.vertex_normals(v, f); va = np.ones(v.shape[0], np.float32)
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_o, d_n, d_v, d_f, d_vn, d_va = to(o), to(n), to(v), to(f), to(vn), to(va)
d_data = torch.zeros((L, B), dtype=torch.float64, device=dev); d_pl = torch.zeros(B, dtype=torch.float64, device=dev)
d_w = torch.ones((L, B), dtype=torch.float64, device=dev); d_T = torch.zeros((L, B), dtype=torch.float64, device=dev); d_G = torch.zeros((v.shape[0], 3), dtype=torch.float64, device=dev)
torch.cuda.synchronize(); ctx.set_option('timing', 1)
A = (20000, 0.0, 1.44, 1.2e-3)
cases = {
    'plain gradient': lambda: renderer.renderStreamedGradient(d_o, d_n, d_v, d_f, *A, d_T, d_pl, d_G, d_data, d_w, 10, 1, 1, 0, ctx=ctx),
    'shading gradient (tf=0)': lambda: renderer.renderStreamedShadingGradient(d_o, d_n, d_v, d_f, d_vn, *A, d_T, d_pl, d_G, d_data, d_w, 10, 1, 0, 0, ctx=ctx),
    'albedo gradient': lambda: renderer.renderStreamedGradientWithAlbedo(d_o, d_n, d_v, d_f, d_va, *A, d_T, d_pl, d_G, d_data, d_w, 10, 1, 1, 0, ctx=ctx),
    'ggx gradient': lambda: ggx.renderStreamedGradient(d_o, d_n, d_v, d_f, 0.5, *A, d_T, d_pl, d_G, d_data, d_w, 10, 1, 1, ctx=ctx),
    'ggx shading gradient': lambda: ggx.renderStreamedShadingGradient(d_o, d_n, d_v, d_f, d_vn, 0.5, *A, d_T, d_pl, d_G, d_data, d_w, 10, 1, 0, ctx=ctx),
    'smoothed forward (r=10)': lambda: renderer.renderStreamedTransient(d_o, d_n, d_v, d_f, *A, d_T, d_pl, 10, 1, ctx=ctx),
}
for name, fn in cases.items():
    for _ in range(3):
        fn(); ctx.synchronize()
    t = ctx.timing(); print('%-26s forward %.2f ms  gradient %.2f ms' % (name, t['forward_ms'], t['gradient_ms']))
