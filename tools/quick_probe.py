"""GPU A/B of kernel variants on C-bunny: for every library given (default first) one subprocess runs renderStreamedGradient 6 times and prints the best
forward / gradient / total ms and SHA-1 digests of the visibility-dependent outputs; the digests of transient row sums are compared loosely (FP64 atomics
reorder), the visibility words exactly.   python tools/quick_probe.py [name ...]   (names of build/variants/libnlos_q_<name>.so)"""
import sys, os, subprocess, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
CHILD = r'''
import sys, os, json, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import renderer, scenes
cfg = sys.argv[1]
ctx = nb.Context(0); dev = torch.device('cuda', 0)
if cfg == 'arm':
    v, f = scenes.armadillo_init(); ns = 20000
else:
    v, f = scenes.bunny(); ns = 20000
o, n = scenes.wall_grid(64); L = o.shape[0]; B = 1200
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_o, d_n, d_v, d_f = to(o), to(n), to(v), to(f)
v2 = v.copy(); v2[:, 2] += 0.01
d_data = torch.zeros((L, B), dtype=torch.float64, device=dev); d_pl = torch.zeros(B, dtype=torch.float64, device=dev)
ctx.set_option('forward_algo', 2)
renderer.renderStreamedTransient(d_o, d_n, to(v2), d_f, ns, 0.0, 1.44, 1.2e-3, d_data, d_pl, 1, 1, ctx=ctx)
d_w = torch.ones((L, B), dtype=torch.float64, device=dev)
ctx.set_option('timing', 1)
T = torch.zeros((L, B), dtype=torch.float64, device=dev); G = torch.zeros((v.shape[0], 3), dtype=torch.float64, device=dev)
best = None; ts = []
for i in range(6):
    T.zero_(); G.zero_()
    renderer.renderStreamedGradient(d_o, d_n, d_v, d_f, ns, 0.0, 1.44, 1.2e-3, T, d_pl, G, d_data, d_w, 10, 1, 1, 0, ctx=ctx)
    ctx.synchronize(); t = ctx.timing(); ts.append(t['forward_ms'])
    if best is None or t['forward_ms'] < best['forward_ms']: best = t
Tn = T.cpu().numpy(); Gn = G.cpu().numpy()
nz = hashlib.sha1((Tn != 0).tobytes()).hexdigest()[:12]
print(json.dumps({'forward_ms': best['forward_ms'], 'gradient_ms': best['gradient_ms'], 'total_ms': best['total_ms'], 'all_fwd': ts,
                  'T_sum': float(Tn.sum()), 'T_nz': nz, 'G_abs': float(np.abs(Gn).sum()), 'T_rows': Tn.sum(1)[:8].tolist()}))
''' % ROOT
cfg = 'bunny'
names = sys.argv[1:]
if names and names[0] in ('bunny', 'arm'):
    cfg, names = names[0], names[1:]
libs = [('default', os.path.join(ROOT, 'nlos_surface_optimization_b200', 'libnlos_b200.so'))] + [(n, os.path.join(ROOT, 'build', 'variants', 'libnlos_q_%s.so' % n)) for n in names]
ref = None
for name, path in libs:
    env = dict(os.environ, NLOS_B200_LIB=path)
    r = subprocess.run([sys.executable, '-c', CHILD, cfg], env=env, capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        print('%-16s FAILED: %s' % (name, r.stderr[-400:]), flush=True); continue
    d = json.loads(r.stdout.strip().splitlines()[-1])
    if ref is None: ref = d
    same = (d['T_nz'] == ref['T_nz'] and abs(d['T_sum'] - ref['T_sum']) <= 1e-12 * abs(ref['T_sum']) and abs(d['G_abs'] - ref['G_abs']) <= 1e-10 * abs(ref['G_abs']))
    print('%-16s forward %.3f ms gradient %.3f total %.3f | %s | all %s' % (name, d['forward_ms'], d['gradient_ms'], d['total_ms'], 'outputs identical' if same else 'OUTPUTS DIFFER %r vs %r' % ((d['T_sum'], d['T_nz'], d['G_abs']), (ref['T_sum'], ref['T_nz'], ref['G_abs'])), ' '.join('%.2f' % x for x in d['all_fwd'])), flush=True)
