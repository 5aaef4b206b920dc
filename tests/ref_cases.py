"""Case table shared by tools/make_ref_fixtures.py (runs the REFERENCE's own code, oracle/_ref) and tests/test_reference_pin.py
(runs the oracle and, on a GPU, the CUDA path on the same inputs).  Every case is a small seeded scene; stochastic outputs are
compared statistically (the reference's Mersenne-Twister sample stream is unrelated to the counter-based one used here), deterministic
ones (regularisers, ray queries) directly.

Scenes avoid the places where the reference itself is undefined (found while pinning):
  * render_intensity adds into intensity[triangle] from all workers unsynchronised  -> reference run with one thread;
  * streamed_render_normal_smoothing / _curvature_grad assign (not add) per vertex and then sum per-thread buffers, so the result depends
    on how triangles fall on threads                                               -> reference run with one thread (= serial order);
  * render_smoothed_vertex_gradients allocates 3*numBins per thread but clears 3*numVertices (transient_and_gradient.cpp:401,406)
                                                                                     -> case built with numVertices == numBins;
  * ggx evalNWDiff leaves its outputs unset when n.w <= 0 (ggx_confocal.cpp:138-166), NaN with shading normals on silhouettes
                                                                                     -> GGX + shading gradient pinned on a height field;
  * the gradient tap loop reads difference[source*numBins + bin] for bin >= numBins when a sample sits within 2*sigma_bin bins of the
    upper bound (transient_and_gradient.cpp:685-690, no range check)                 -> every scene keeps 2r + 2 bins below the upper bound.
"""
import numpy as np
from nlos_surface_optimization_b200 import scenes

LB, UB = 0.0, 1.44
RES = 4.8e-3                      # 300 bins
RES_VG = 0.0089                   # 162 bins == V of icosphere(2)


def wall():
    return scenes.wall_grid(4)


def ico():
    return scenes.icosphere(1, 0.12, (0.03, -0.02, 0.45), noise=0.05, seed=3)       # V=42 F=80, back half self-occluded


def ico2():
    return scenes.icosphere(2, 0.12, (0.03, -0.02, 0.45), noise=0.03, seed=3)       # V=162 F=320


def field():
    v, f = scenes.heightfield(6)                                                      # V=36 F=50
    v[:, 2] = 0.45 + 0.25 * (v[:, 2] - 0.45)                                          # gentle slopes: n.w > 0 for every wall point
    v[:, :2] *= 0.5                                                                   # 2r + 2 bins < upper bound (see below)
    return np.ascontiguousarray(v), f


def occluder():
    """bumpy sphere partly hidden behind a quad: visibility against OTHER geometry, not only self-occlusion."""
    return scenes.merge([ico(), scenes.quad(0.3, 0.05, 0.02, 0.0)])


def backface():
    """bumpy sphere plus a quad wound AWAY from the wall (visible from behind): separates the unclamped first-generation form factor
    (contributes ff^2) from the clamped second-generation one (contributes 0)."""
    qv, qf = scenes.quad(0.4, 0.05, -0.17, 0.1)
    return scenes.merge([ico(), (qv, np.ascontiguousarray(qf[:, ::-1]))])


def albedo(v):
    return (0.5 + np.random.RandomState(0).rand(v.shape[0])).astype(np.float32)


def jitter_kernel():
    x = np.arange(21) - 8.0
    jw = np.exp(-0.5 * (x / 3.0) ** 2) * (1 + 0.3 * np.sin(x)); jw /= jw.sum()
    return jw, np.gradient(jw)


def target(oracle, v, f, num_sample, res=RES, nsrc=None, **kw):
    """ground-truth transient = the oracle's render of the mesh pushed back by 1 cm (deterministic: fixed seed)."""
    o, n = wall()
    if nsrc:
        o, n = o[:nsrc], n[:nsrc]
    v2 = v.copy(); v2[:, 2] += 0.01
    return oracle.transient(o, n, v2, f, num_sample, LB, UB, res, 10, 1, seed=5, **kw)[0]


# name -> dict(kind, scene, samples, kwargs).  kinds: transient, intensity, gradient (T and G), scalar (albedo / alpha derivative),
# vertex_gradient, jitter_transient, jitter_gradient
S = 400000
CASES = {
    'transient_r10_s1':     dict(kind='transient', scene='ico', S=S, rs=10, sb=1),
    'transient_r1_s1':      dict(kind='transient', scene='ico', S=S, rs=1, sb=1),
    'transient_r4_s2':      dict(kind='transient', scene='ico', S=S, rs=4, sb=2),
    'transient_occluder':   dict(kind='transient', scene='occluder', S=S, rs=10, sb=1),
    'transient_shading':    dict(kind='transient', scene='ico', S=S, rs=10, sb=1, shading=True),
    'transient_albedo':     dict(kind='transient', scene='ico', S=S, rs=10, sb=1, albedo=True),
    'intensity':            dict(kind='intensity', scene='ico', S=S),
    'gradient_t1_l0':       dict(kind='gradient', scene='ico', S=S, rs=10, sb=1, tf=1, lf=0),
    'gradient_t1_l1':       dict(kind='gradient', scene='ico', S=S, rs=10, sb=1, tf=1, lf=1),
    'gradient_t0_l0':       dict(kind='gradient', scene='ico', S=S, rs=4, sb=2, tf=0, lf=0),
    'gradient_occluder':    dict(kind='gradient', scene='occluder', S=S, rs=10, sb=1, tf=1, lf=0),
    'gradient_shading':     dict(kind='gradient', scene='ico', S=S, rs=10, sb=1, tf=1, lf=0, shading=True),
    'gradient_w_albedo':    dict(kind='gradient', scene='ico', S=S, rs=10, sb=1, tf=1, lf=0, albedo=True),
    'gradient_albedo':      dict(kind='scalar', scene='ico', S=S, rs=10, sb=1, albedo=True),
    'ggx_transient_a03':    dict(kind='transient', scene='ico', S=S, rs=10, sb=1, alpha=0.3),
    'ggx_transient_a08_sh': dict(kind='transient', scene='ico', S=S, rs=10, sb=1, alpha=0.8, shading=True),
    'ggx_intensity':        dict(kind='intensity', scene='ico', S=S, alpha=0.5),
    'ggx_gradient_a03':     dict(kind='gradient', scene='ico', S=S, rs=10, sb=1, tf=1, lf=0, alpha=0.3),
    'ggx_gradient_field_sh': dict(kind='gradient', scene='field', S=S, rs=10, sb=1, tf=1, lf=0, alpha=0.5, shading=True),
    'ggx_gradient_alpha':   dict(kind='scalar', scene='ico', S=S, rs=10, sb=1, alpha=0.5),
    'vertex_gradient_v20':  dict(kind='vertex_gradient', scene='ico2', S=S, rs=10, sb=1, vertex=20),
    'vertex_gradient_v100': dict(kind='vertex_gradient', scene='ico2', S=S, rs=10, sb=1, vertex=100),
    'jitter_transient':     dict(kind='jitter_transient', scene='ico', S=S, offset=8),
    'jitter_transient_sh':  dict(kind='jitter_transient', scene='ico', S=S, offset=8, shading=True),
    'jitter_gradient_t1':   dict(kind='jitter_gradient', scene='ico', S=S, offset=8, tf=1),
    'jitter_gradient_t0':   dict(kind='jitter_gradient', scene='occluder', S=S, offset=8, tf=0),
    'jitter_gradient_off3': dict(kind='jitter_gradient', scene='ico', S=S, offset=3, tf=1),
    # the headline mesh (69 630 triangles, real self-occlusion), 2 samples per triangle and source
    'bunny_transient':      dict(kind='transient', scene='bunny', S=2 * 69630, rs=10, sb=1),
    'bunny_gradient':       dict(kind='gradient', scene='bunny', S=2 * 69630, rs=10, sb=1, tf=1, lf=0),
    # first-generation renderer (stratified_transient_raytracer/)
    'sr_transient_backface': dict(kind='sr_transient', scene='backface', S=S),
    'sr_transient_albedo':  dict(kind='sr_transient', scene='occluder', S=S, albedo=True),
    'sr_single_origin':     dict(kind='sr_single', scene='backface', S=4 * S, source=5),
    'sr_gradient_w0':       dict(kind='sr_gradient', scene='occluder', S=S, w=0),
    'sr_gradient_w3':       dict(kind='sr_gradient', scene='backface', S=S, w=3),
}


def same_sample_cases():
    """The cases above at a small sample count, to be run by the oracle on the REFERENCE'S OWN sample stream (one reference worker):
    outputs must then agree to float rounding, not just statistically.  spp = 20 (bunny: 2 sources, spp = 1)."""
    out = {}
    for name, c in CASES.items():
        d = dict(c); F = SCENES[c['scene']]()[1].shape[0]
        if c['scene'] == 'bunny':
            d['S'] = F; d['nsrc'] = 2
        elif c['kind'] == 'vertex_gradient':
            d['S'] = 50 * F
        else:
            d['S'] = 20 * F
        out[name] = d
    return out


def stream_length(c):
    """floats of the reference's sample stream one call of case c consumes: 2 per sample, tasks in (source, triangle) order."""
    F = SCENES[c['scene']]()[1].shape[0]
    L = 1 if c['kind'] in ('vertex_gradient', 'sr_single') else (c.get('nsrc') or wall()[0].shape[0])
    spp = 1 + (c['S'] - 1) // F
    passes = 2 if c['kind'] == 'sr_gradient' else 1      # the first-generation gradient call draws its two passes from one stream
    return 2 * L * F * spp * passes


SCENES = {'ico': ico, 'ico2': ico2, 'field': field, 'occluder': occluder, 'backface': backface, 'bunny': scenes.bunny}


def run_case(impl, oracle, c, seed=None, relabel=None, ext_stream=None):
    """Run one case on `impl` (oracle / reference / gpu adapter: same function names as oracle.oracle).  `seed` is passed where the
    implementation takes one; `relabel` = (source permutation, face permutation) renders a relabelled copy of the same scene (how
    independent draws are obtained from the reference, whose seed is fixed) and un-permutes the outputs.  `ext_stream` (oracle only)
    installs the reference's own (S,T) sample stream around the call under test — the "same-sample" cases.  Returns a dict of arrays."""
    o, n = wall(); v, f = SCENES[c['scene']]()
    if c.get('nsrc'):
        o, n = np.ascontiguousarray(o[:c['nsrc']]), np.ascontiguousarray(n[:c['nsrc']])

    def call(fn, *a, **k):
        if ext_stream is not None:
            oracle.set_external_samples(ext_stream)
        try:
            return fn(*a, **k)
        finally:
            if ext_stream is not None:
                oracle.set_external_samples(None)
    kw = {}
    if seed is not None:
        kw['seed'] = seed
    vn = scenes.vertex_normals(v, f) if c.get('shading') else None
    va = albedo(v) if c.get('albedo') else None
    sp = np.arange(o.shape[0]); fp = np.arange(f.shape[0])
    if relabel is not None:
        sp, fp = relabel
    o_, n_, f_ = o[sp], n[sp], f[fp]
    inv_s = np.argsort(sp); inv_f = np.argsort(fp)
    alpha = c.get('alpha')
    akw = {} if alpha is None else {'alpha': alpha}
    k = c['kind']
    if k == 'transient':
        T = call(impl.transient, o_, n_, v, f_, c['S'], LB, UB, RES, c['rs'], c['sb'], vertex_normal=vn, vertex_albedo=va, **akw, **kw)[0]
        return {'T': T[inv_s]}
    if k == 'intensity':
        I = call(impl.intensity, o_, n_, v, f_, c['S'], LB, UB, vertex_normal=vn, **akw, **kw)
        return {'I': I[inv_f]}
    if k in ('gradient', 'scalar'):
        tkw = dict(akw)
        if va is not None:
            tkw['vertex_albedo'] = va
        if vn is not None:
            tkw['vertex_normal'] = vn
        data = target(oracle, v, f, c['S'], nsrc=c.get('nsrc'), **tkw); w = np.ones_like(data)
        if k == 'gradient':
            T, G, _ = call(impl.gradient, o_, n_, v, f_, c['S'], LB, UB, RES, data[sp], w, c['rs'], c['sb'], c['tf'], c['lf'], vertex_normal=vn, vertex_albedo=va, **akw, **kw)
            return {'T': T[inv_s], 'G': G}
        if alpha is None:
            T, g = call(impl.gradient_albedo, o_, n_, v, f_, c['S'], LB, UB, RES, data[sp], w, c['rs'], c['sb'], va, **kw)
        else:
            T, g = call(impl.gradient_alpha, o_, n_, v, f_, c['S'], LB, UB, RES, data[sp], w, c['rs'], c['sb'], alpha, **kw)
        return {'g': np.array([g])}
    if k == 'vertex_gradient':
        G = call(impl.vertex_gradient, c['vertex'], o[:1], n[:1], v, f_, c['S'], LB, UB, RES_VG, c['rs'], c['sb'], **kw)
        return {'VG': np.asarray(G).reshape(-1, 3)}
    if k == 'sr_transient':
        T = call(impl.sr_transient, o_, n_, v, f_, c['S'], LB, UB, RES, vertex_normal=vn, vertex_albedo=va, **kw)[0]
        return {'T': T[inv_s]}
    if k == 'sr_single':
        i = c['source']
        return {'T': np.asarray(call(impl.sr_render_transient, o[i], n[i], v, f_, c['S'], LB, UB, RES, **kw)[0]).reshape(1, -1)}
    if k == 'sr_gradient':
        v2 = v.copy(); v2[:, 2] += 0.01
        data = oracle.sr_transient(o, n, v2, f, c['S'], LB, UB, RES, seed=5)[0]
        T, G, _ = call(impl.sr_gradient, o_, n_, v, f_, c['S'], LB, UB, RES, c['w'], data[sp], **kw)
        return {'T': T[inv_s], 'G': G}
    jw, jg = jitter_kernel()
    if k == 'jitter_transient':
        T = call(impl.jitter_transient, o_, n_, v, f_, c['S'], LB, UB, RES, jw, c['offset'], vertex_normal=vn, **kw)[0]
        return {'T': T[inv_s]}
    if k == 'jitter_gradient':
        v2 = v.copy(); v2[:, 2] += 0.01
        data = oracle.jitter_transient(o, n, v2, f, c['S'], LB, UB, RES, jw, c['offset'], seed=5)[0]; w = np.ones_like(data)
        T, G, _ = call(impl.jitter_gradient, o_, n_, v, f_, c['S'], LB, UB, RES, jw, jg, c['offset'], data[sp], w, c['tf'], **kw)
        return {'T': T[inv_s], 'G': G}
    raise KeyError(k)


class OracleAdapter(object):
    """oracle.oracle with the albedo / alpha scalar derivatives under the reference's names."""
    def __init__(self, oracle):
        self.o = oracle
        for name in ('transient', 'intensity', 'gradient', 'vertex_gradient', 'jitter_transient', 'jitter_gradient'):
            setattr(self, name, getattr(oracle, name))

    # first generation: the reference's gradient is pinned with its output-index slips reproduced (typos=True; see render_gradients_sr)
    def sr_transient(self, *a, **kw):
        return self.o.sr_transient(*a, **kw)

    def sr_render_transient(self, o, n, v, f, S, lb, ub, res, seed=None):
        T, pl = self.o.sr_transient(np.reshape(o, (1, 3)), np.reshape(n, (1, 3)), v, f, S, lb, ub, res, **({} if seed is None else {'seed': seed}))
        return T[0], pl

    def sr_gradient(self, o, n, v, f, S, lb, ub, res, w, data, seed=None):
        return self.o.sr_gradient(o, n, v, f, S, lb, ub, res, w, data, typos=True, **({} if seed is None else {'seed': seed}))

    def gradient_albedo(self, o, n, v, f, S, lb, ub, res, data, w, rs, sb, va, seed=None):
        return self.o.gradient(o, n, v, f, S, lb, ub, res, data, w, rs, sb, 1, 0, vertex_albedo=va, kind=1, **({} if seed is None else {'seed': seed}))

    def gradient_alpha(self, o, n, v, f, S, lb, ub, res, data, w, rs, sb, alpha, seed=None):
        return self.o.gradient(o, n, v, f, S, lb, ub, res, data, w, rs, sb, 1, 0, alpha=alpha, kind=2, **({} if seed is None else {'seed': seed}))


def two_sample_z(mean_a, std_a, n_a, mean_b, std_b, n_b):
    """Two-sample z per element with the pooled standard deviation (both sides are the same estimator with the same sample count, so
    their per-element variances are equal; pooling keeps z well behaved when one side has few draws).  Elements without spread on
    either side are skipped.  Returns (z, relative L2 distance of the means, the same distance expected from noise alone)."""
    pooled = np.sqrt(((n_a - 1) * std_a ** 2 + (n_b - 1) * std_b ** 2) / (n_a + n_b - 2))
    se = pooled * np.sqrt(1.0 / n_a + 1.0 / n_b)
    ok = se > 0
    z = (mean_a[ok] - mean_b[ok]) / se[ok]
    rel = np.linalg.norm(mean_a - mean_b) / max(np.linalg.norm(mean_b), 1e-300)
    noise_rel = np.sqrt((se ** 2).sum()) / max(np.linalg.norm(mean_b), 1e-300)      # what pure Monte-Carlo noise makes of `rel`
    return z, rel, noise_rel


def residual_events(got, want, thr_rel):
    """Structure of a same-sample residual D = got - want.  A sample whose fine bin (last-bit difference in r) or visibility (grazing
    ray) differs between the two sides changes a transient row in a short run of neighbouring bins, an intensity in one entry and a
    gradient in the 3 vertex rows of its triangle — everything else must agree to float rounding.  Returns (events, residual outside
    the events relative to ||want||): events = runs of adjacent entries (along the last axis) above thr_rel * max|want|, each run
    widened by one entry (a partner bin just below the threshold)."""
    D = np.atleast_2d(np.asarray(got, dtype=np.float64) - want); W = np.atleast_2d(want)
    mask = np.abs(D) > thr_rel * np.abs(W).max()
    dil = mask.copy(); dil[:, 1:] |= mask[:, :-1]; dil[:, :-1] |= mask[:, 1:]
    events = int(sum(int(r[0]) + int(np.sum((~r[:-1]) & r[1:])) for r in dil))
    return events, float(np.linalg.norm(D[~dil]) / max(np.linalg.norm(W), 1e-300))


# Bars of the event analysis (measured: tests/test_reference_pin.py docstring of the same-sample tests): outside at most MAX_EVENTS
# localized events the two sides agree to RESIDUAL_OUTSIDE.
EVENT_THR = {'T': 1e-6, 'I': 1e-6, 'G': 1e-5, 'VG': 1e-5}
RESIDUAL_OUTSIDE = {'T': 1e-6, 'I': 1e-6, 'G': 2e-5, 'VG': 2e-5}


def max_events(case, key, shape):
    """transients / intensities: at most 8 events per case (bunny: 2 x 69 630 samples); gradients: the 3 vertex rows of at most 2
    triangles, for the bunny at most 1 % of the vertex rows."""
    if key in ('T', 'I'):
        return 8
    return max(6, shape[0] // 100) if case.get('scene') == 'bunny' else 6


def assert_residual_is_a_few_flipped_samples(case, name, key, got, want):
    if np.ndim(got) == 0 or np.size(got) < 4 or key not in EVENT_THR:
        return
    ev, res = residual_events(got, want, EVENT_THR[key])
    print('same/%s/%s events %d residual outside %.2e' % (name, key, ev, res))
    assert ev <= max_events(case, key, np.shape(got)), (name, key, ev)
    assert res <= RESIDUAL_OUTSIDE[key], (name, key, res)
