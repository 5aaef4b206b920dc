"""Vendor the reference's ground-truth bunny mesh as a compact .npz (data asset, not source).

Source: <reference>/transient_rendering_cython/mesh/bunny_centered.obj  (V=34817, F=69630; the
Stanford bunny re-centred to z in [0.356, 0.549]).  The reference
path is only read when this script is run by hand in the build container; the resulting
assets/bunny.npz travels with the repo so nothing on the GPU box needs /root/reference.

    python tools/make_bunny_npz.py [/root/reference]
"""
import sys, os
import numpy as np

def read_obj(path):
    v, f = [], []
    with open(path) as fh:
        for line in fh:
            if line.startswith('v '):
                v.append([float(x) for x in line.split()[1:4]])
            elif line.startswith('f '):
                f.append([int(tok.split('/')[0]) - 1 for tok in line.split()[1:4]])
    return np.asarray(v, dtype=np.float64), np.asarray(f, dtype=np.int32)

if __name__ == '__main__':
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    base = os.path.join(ref, 'transient_rendering_cython')
    for name, rel in (('bunny', 'mesh/bunny_centered.obj'),
                      ('armadillo', 'exp_armadillo/setup/armadillo.obj'),                     # GT mesh of config C-arm
                      ('armadillo_init', 'exp_armadillo/setup/cnlos_armadillo_threshold.obj')):  # CNLOS initialisation (exp_bunny/test.py:89-93)
        v, f = read_obj(os.path.join(base, rel))
        # the reference loads with igl.readOBJ (double) then casts to float32 (exp_bunny/main_create_gt.py:66-67)
        v32 = v.astype(np.float32)
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'assets', name + '.npz')
        np.savez_compressed(out, v=v32, f=f)
        print(out, v32.shape, f.shape, v32.min(0), v32.max(0), os.path.getsize(out))
