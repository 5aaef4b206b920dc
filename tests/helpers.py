import numpy as np

LB, UB, RES = 0.0, 1.44, 1.2e-3       # exp_bunny/test.py:33-34 (1200 bins x 1.2 mm)
TOL_TRANSIENT = 1e-5                   # BASELINE.json north_star: relative L2
TOL_GRADIENT = 1e-4


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / den if den > 0 else np.linalg.norm(a)


def make_target(oracle, origin, normal, v, f, num_sample, dz=0.01, **kw):
    """`data` = oracle render of the same mesh displaced by dz, `weight` = ones (gamma=0)."""
    v2 = v.copy(); v2[:, 2] += dz
    data = oracle.transient(origin, normal, v2, f, num_sample, LB, UB, RES, **kw)[0]
    weight = np.ones_like(data)
    return np.ascontiguousarray(data), np.ascontiguousarray(weight)
