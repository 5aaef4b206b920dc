"""Synthetic workloads of SURVEY.md section 8(d): meshes and relay-wall grids (numpy only, no GPU).

Shapes follow the reference's experiment drivers:
  wall grid            exp_bunny/test.py:17-32   (meshgrid over [-.25,.25]^2 at z=0, normal (0,0,1))
  8-triangle fan       exp_bunny/weight_test.py:84-85
  2-triangle quad      smoothed_transient/test.py:17-19 (winding flipped so it faces the wall)
  bunny                mesh/bunny_centered.obj (vendored as assets/bunny.npz by tools/make_bunny_npz.py)
"""
import os
import numpy as np

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'assets')


def wall_grid(resolution, half_extent=0.25):
    """Confocal relay-wall measurement points: (origin[L,3] f32, normal[L,3] f32), L = resolution^2."""
    lin = np.linspace(-half_extent, half_extent, resolution)
    gx, gy = np.meshgrid(lin, lin)
    origin = np.stack([np.concatenate(gx), np.concatenate(gy), np.zeros(resolution * resolution)], axis=1)
    normal = np.tile(np.array([0, 0, 1], dtype=np.float32), (resolution * resolution, 1))
    return np.ascontiguousarray(origin, dtype=np.float32), np.ascontiguousarray(normal, dtype=np.float32)


def fan8(z=0.38):
    v = np.array([[-.25, -.25, z], [.25, -.25, z], [.25, .25, z], [-.25, .25, z], [0, -.25, z], [.25, 0, z], [0, .25, z],
                  [-.25, 0, z], [0, 0, z]], dtype=np.float32)
    f = np.array([[0, 8, 4], [0, 7, 8], [4, 8, 1], [1, 8, 5], [8, 2, 5], [8, 6, 2], [7, 3, 8], [3, 6, 8]], dtype=np.int32)
    return np.ascontiguousarray(v), np.ascontiguousarray(f)


def quad(z=0.4, half=0.1, cx=0.0, cy=0.0):
    """Two triangles facing the wall (n_z < 0)."""
    v = np.array([[cx - half, cy - half, z], [cx + half, cy - half, z], [cx + half, cy + half, z], [cx - half, cy + half, z]], dtype=np.float32)
    f = np.array([[0, 2, 1], [0, 3, 2]], dtype=np.int32)
    return np.ascontiguousarray(v), np.ascontiguousarray(f)


def merge(meshes):
    vs, fs, off = [], [], 0
    for v, f in meshes:
        vs.append(v); fs.append(f + off); off += v.shape[0]
    return np.ascontiguousarray(np.concatenate(vs), dtype=np.float32), np.ascontiguousarray(np.concatenate(fs), dtype=np.int32)


def _asset(name):
    d = np.load(os.path.join(_ASSETS, name + '.npz'))
    return np.ascontiguousarray(d['v'], dtype=np.float32), np.ascontiguousarray(d['f'], dtype=np.int32)


def bunny():
    return _asset('bunny')


def armadillo():
    """GT armadillo (exp_armadillo/setup/armadillo.obj, V=43 243 F=86 482)."""
    return _asset('armadillo')


def armadillo_init():
    """CNLOS-thresholded initial mesh the optimisation loop starts from (exp_armadillo/setup/cnlos_armadillo_threshold.obj)."""
    return _asset('armadillo_init')


def _value_noise(x, y, seed, octaves=3):
    rng = np.random.RandomState(seed)
    out = np.zeros_like(x)
    amp, freq = 1.0, 4.0
    for _ in range(octaves):
        n = int(freq) + 2
        lattice = rng.rand(n, n)
        fx = (x + 0.25) / 0.5 * freq; fy = (y + 0.25) / 0.5 * freq
        ix = np.clip(np.floor(fx).astype(int), 0, n - 2); iy = np.clip(np.floor(fy).astype(int), 0, n - 2)
        tx = fx - ix; ty = fy - iy
        tx = tx * tx * (3 - 2 * tx); ty = ty * ty * (3 - 2 * ty)
        v = (lattice[iy, ix] * (1 - tx) + lattice[iy, ix + 1] * tx) * (1 - ty) + (lattice[iy + 1, ix] * (1 - tx) + lattice[iy + 1, ix + 1] * tx) * ty
        out += amp * (v - 0.5)
        amp *= 0.5; freq *= 2
    return out


def heightfield(n=501, seed=1, half_extent=0.25):
    """C-scale mesh: (n x n) vertices => F = 2 (n-1)^2 (n=501 -> 500 000), z = .45 + .04 sin(6 pi x) cos(4 pi y) + .01 noise,
    wound so that n_z < 0 (faces the wall)."""
    lin = np.linspace(-half_extent, half_extent, n)
    x, y = np.meshgrid(lin, lin)
    z = 0.45 + 0.04 * np.sin(6 * np.pi * x) * np.cos(4 * np.pi * y) + 0.01 * _value_noise(x, y, seed)
    v = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.float32)
    idx = np.arange(n * n).reshape(n, n)
    a = idx[:-1, :-1].ravel(); b = idx[:-1, 1:].ravel(); c = idx[1:, 1:].ravel(); d = idx[1:, :-1].ravel()
    f = np.concatenate([np.stack([a, c, b], axis=1), np.stack([a, d, c], axis=1)]).astype(np.int32)
    return np.ascontiguousarray(v), np.ascontiguousarray(f)


def icosphere(subdiv=3, radius=0.1, center=(0, 0, 0.45), noise=0.0, seed=0):
    """Closed, outward-wound test mesh (front half faces the wall, back half is self-occluded)."""
    t = (1 + 5 ** 0.5) / 2
    v = [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]]
    f = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
         [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []
        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]; v.append(m / np.linalg.norm(m)); cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f = nf
    v = np.array(v)
    if noise > 0:
        rng = np.random.RandomState(seed)
        v = v * (1 + noise * rng.randn(v.shape[0], 1))
    v = v * radius + np.asarray(center)
    return np.ascontiguousarray(v, dtype=np.float32), np.ascontiguousarray(np.array(f), dtype=np.int32)


def vertex_normals(v, f):
    """Area-weighted per-vertex normals (host stand-in for cgal_api.per_vertex_normal, exp_bunny/rendering.py:220-221)."""
    a, b, c = v[f[:, 0]].astype(np.float64), v[f[:, 1]].astype(np.float64), v[f[:, 2]].astype(np.float64)
    n = np.cross(b - a, c - a)
    vn = np.zeros((v.shape[0], 3))
    for k in range(3):
        np.add.at(vn, f[:, k], n)
    vn /= np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-30)
    return np.ascontiguousarray(vn, dtype=np.float32)


def face_affinity(f):
    """f_affinity[F,3]: the face across edge (k, k+1) or -1 (host stand-in for cgal_api.face_affinity)."""
    edges = {}
    for fi in range(f.shape[0]):
        for k in range(3):
            a, b = int(f[fi, k]), int(f[fi, (k + 1) % 3])
            edges.setdefault((min(a, b), max(a, b)), []).append((fi, k))
    aff = -np.ones((f.shape[0], 3), dtype=np.int32)
    for lst in edges.values():
        if len(lst) == 2:
            (f0, k0), (f1, k1) = lst
            aff[f0, k0] = f1; aff[f1, k1] = f0
    return aff
