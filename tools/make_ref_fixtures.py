"""Generate tests/golden/ref_pin.npz: outputs of the REFERENCE's own renderer code on the cases of tests/ref_cases.py.

Runs only where /root/reference exists (this container): `make -C oracle ref` compiles the reference's unmodified translation units
against the stand-in headers of oracle/ref_shim (Embree/TBB/MKL/Boost are absent) into oracle/_ref/, and this script calls them through
oracle/reference.py.  The reference's seed is fixed inside its sources, so independent draws are obtained by rendering R relabelled
copies of each scene (random permutations of the source order and of the face order; the estimated quantity is unchanged) and the
fixture stores the per-element mean and standard deviation over the R draws.  Deterministic entry points are stored as they are.

    python tools/make_ref_fixtures.py            # everything, ~8 minutes (the reference's gradient tap loop is slow)
    python tools/make_ref_fixtures.py NAME...    # regenerate the named cases only (also: `same`, `deterministic`) and merge into the existing file
"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
from oracle import oracle, reference
import ref_cases as rc
from nlos_surface_optimization_b200 import scenes

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'ref_pin.npz')
R = 8


def main():
    assert reference.build(), 'reference modules not built (needs /root/reference)'
    only = sys.argv[1:]
    out = {k: v for k, v in np.load(OUT).items()} if only and os.path.exists(OUT) else {}
    out['R'] = R
    o, n = rc.wall()
    for name, c in rc.CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        reference.set_threads(1 if c['kind'] == 'intensity' else os.cpu_count())
        v, f = rc.SCENES[c['scene']]()
        rng = np.random.RandomState(sum(map(ord, name)))
        draws = []
        for r in range(R):
            relabel = (rng.permutation(o.shape[0]), rng.permutation(f.shape[0]))
            if c['kind'] == 'vertex_gradient':
                relabel = (np.arange(o.shape[0]), relabel[1])
            draws.append(rc.run_case(reference, oracle, c, relabel=relabel))
        for key in draws[0]:
            d = np.stack([x[key] for x in draws])
            assert np.isfinite(d).all(), (name, key)
            out['%s/%s/mean' % (name, key)] = d.mean(0)
            out['%s/%s/std' % (name, key)] = d.std(0, ddof=1)
        print('%-24s %.1fs' % (name, time.time() - t0), flush=True)
        np.savez_compressed(OUT, **out)
    if not only or 'same' in only:
        # same-sample cases: ONE reference worker, no relabelling -> the sample stream is worker 0's, stored once for all cases
        reference.set_threads(1)
        cases = rc.same_sample_cases()
        out['same/stream'] = reference.sampler_stream(max(rc.stream_length(c) for c in cases.values()))
        for name, c in cases.items():
            t0 = time.time()
            res = rc.run_case(reference, oracle, c)
            for key, val in res.items():
                assert np.isfinite(val).all(), (name, key)
                out['same/%s/%s' % (name, key)] = val
            print('same/%-24s %.1fs' % (name, time.time() - t0), flush=True)
        np.savez_compressed(OUT, **out)
    if only and 'deterministic' not in only:
        return
    # deterministic entry points (one thread: serial triangle order, see tests/ref_cases.py)
    reference.set_threads(1)
    v, f = rc.ico2(); aff = scenes.face_affinity(f)
    val, g = reference.normal_smoothing(v, f, aff)
    out['normal_smoothing/value'] = np.array([val]); out['normal_smoothing/grad'] = g
    out['curvature_grad/grad'] = reference.curvature_grad(v, f)
    ro, rd = ray_bundle()
    reference.set_threads(os.cpu_count())
    v, f = rc.occluder()
    out['intersect/bary'] = reference.intersect(ro, rd, v, f)
    out['intersect/short'] = reference.intersect(ro, rd, v, f, short=True)
    out['intersect/world'] = reference.bary_to_world(v, f, out['intersect/bary'])
    np.savez_compressed(OUT, **out)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


def ray_bundle(N=4096):
    o, _ = rc.wall()
    rng = np.random.RandomState(1)
    ro = np.tile(o, (N // o.shape[0], 1)).astype(np.float32)
    tgt = np.stack([rng.uniform(-.2, .2, N), rng.uniform(-.2, .2, N), np.full(N, 0.45)], 1)
    rd = tgt - ro
    return ro, np.ascontiguousarray(rd / np.linalg.norm(rd, axis=1, keepdims=True), dtype=np.float32)


if __name__ == '__main__':
    main()
