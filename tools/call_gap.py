"""Host-array call overhead: wall time of one reference-signature call (pageable NumPy arrays, as the reference's drivers pass them)
against the kernels' own time inside it (library CUDA events).  python tools/call_gap.py [arm|bunny]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import renderer, scenes

which = sys.argv[1] if len(sys.argv) > 1 else 'arm'
ctx = nb.default_context(0)
o, n = scenes.wall_grid(64)
v, f = scenes.armadillo_init() if which == 'arm' else scenes.bunny()
L, B = o.shape[0], 1200
data = np.random.RandomState(0).rand(L, B) * 1e-4; weight = np.ones((L, B))
T = np.zeros((L, B)); pl = np.zeros(B); G = np.zeros((v.shape[0], 3))
ctx.set_option('timing', 1)
walls, devs = [], []
for i in range(8):
    G[:] = 0
    t0 = time.perf_counter()
    renderer.renderStreamedGradient(o, n, v, f, 20000, 0.0, 1.44, 1.2e-3, T, pl, G, data, weight, 10, 1, 1, 0, ctx=ctx)
    walls.append(1e3 * (time.perf_counter() - t0)); devs.append(ctx.timing())
w = float(np.median(walls[2:])); d = devs[-1]
k = d['build_ms'] + d['forward_ms'] + d['residual_ms'] + d['gradient_ms']
print('%s: call wall %.2f ms, kernels %.2f ms (build %.2f fwd %.2f res %.2f grad %.2f), device total %.2f ms, gap wall - kernels %.2f ms; H2D %.1f MB, D2H %.1f MB'
      % (which, w, k, d['build_ms'], d['forward_ms'], d['residual_ms'], d['gradient_ms'], d['total_ms'], w - k,
         (data.nbytes + weight.nbytes + o.nbytes + n.nbytes + v.nbytes + f.nbytes + G.nbytes) / 1e6, (T.nbytes + G.nbytes + pl.nbytes) / 1e6))
