"""Host probe of the group perspective grid (tests/emul/ggrid_emul.cpp): statistics + mismatch count against the BVH query.
python tools/ggrid_probe.py MESH WALL G K GX GY STRIDE"""
import sys, os, ctypes as C, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
from nlos_surface_optimization_b200 import scenes
mesh, wall, G, K, gx, gy, stride = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'emul', 'libggrid.so'))
o, n = scenes.wall_grid(wall); v, f = getattr(scenes, mesh)()
out = np.zeros(16)
fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
t = time.time()
rc = lib.ggrid_emul(fp(o), fp(n), o.shape[0], fp(v), v.shape[0], f.ctypes.data_as(C.POINTER(C.c_int)), f.shape[0], G, K, wall, gx, gy, stride, out.ctypes.data_as(C.POINTER(C.c_double)))
names = ['rays', 'entries/tri', 'lookups/ray', 'nonempty/ray', 'scanned/ray', 'rect/ray', 'edge/ray', 'exact/ray', 'MISMATCH', 'occ frac', 'fallback rays', 'groups', 'nogrid', 'never frac', 'max e units']
print('rc', rc, '%.1fs' % (time.time() - t), ' '.join('%s=%.4g' % (k, x) for k, x in zip(names, out)))
