#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 transient renderer (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C-bunny"): bunny (F=69 630), 64x64 confocal wall, B=1200 bins of 1.2 mm,
sample_num=20 000 (spp=1), refine_scale=10, sigma_bin=1, testing_flag=1, loss_flag=0.  One STEP = one
renderStreamedGradient call = forward pass + residual + vertex-gradient pass (+ scene build, + NCCL all-reduce
of the gradient when N>1) = 2*L*F*spp path samples.  N>1: each rank renders its own 64x64 slice of a (64N)x64
wall (weak scaling), the gradient is all-reduced once per step.

Prints ONE JSON line (rank 0).  `value` = path samples/s with all inputs resident in HBM; `e2e` = the same through
the C ABI with pinned HOST buffers (H2D/D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LB, UB, RES = 0.0, 1.44, 1.2e-3
SAMPLE_NUM, REFINE, SIGMA = 20000, 10, 1
WALL = 64
METRIC = 'transient path samples/sec (fwd+vertex grad)'
UNIT = 'path samples/s'


def workload(rank, world):
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.bunny()
    # (64*world) x 64 wall on [-.25,.25]^2; rank r owns the interleaved rows r, r+world, ... (every rank sees the whole
    # wall extent => equal work per rank); global source index = rank*L + local index (RNG key, nlos_ctx_set_source_window)
    lin_x = np.linspace(-.25, .25, WALL)
    lin_y = np.linspace(-.25, .25, WALL * world)[rank::world]
    gx, gy = np.meshgrid(lin_x, lin_y)
    o = np.ascontiguousarray(np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1).astype(np.float32))
    L = o.shape[0]
    L_global = L * world
    n = np.ascontiguousarray(np.tile(np.array([0, 0, 1], dtype=np.float32), (L, 1)))
    return o, n, v, f, L, L_global


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index; self.rows = []; self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        busy = [x for x in sm if x > 0]
        return {'sm_mhz': float(np.median(busy)) if busy else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_baseline(v, f, n_sources, want_stats=True, repeats=1):
    """The oracle port (reference-restated CPU path, NOT the Embree build) on a bounded sample of the same workload."""
    from oracle import oracle
    from nlos_surface_optimization_b200 import scenes
    o, n = scenes.wall_grid(WALL)
    idx = np.linspace(0, o.shape[0] - 1, n_sources).astype(int)       # spread over the wall, not one corner
    o = np.ascontiguousarray(o[idx]); n = np.ascontiguousarray(n[idx])
    B = oracle.num_bins(LB, UB, RES)
    data = np.zeros((n_sources, B)); weight = np.ones((n_sources, B))
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        oracle.gradient(o, n, v, f, SAMPLE_NUM, LB, UB, RES, data, weight, REFINE, SIGMA, 1, 0)
        times.append(time.perf_counter() - t0)
    spp = 1 + (SAMPLE_NUM - 1) // f.shape[0]
    samples = 2 * n_sources * f.shape[0] * spp
    stats = None
    if want_stats:
        st = oracle.transient(o[:8], n[:8], v, f, SAMPLE_NUM, LB, UB, RES, want_stats=True)[2]
        stats = {'box_per_ray': st['box_tests'] / st['rays'], 'tri_per_ray': st['tri_tests'] / st['rays']}
    return samples / min(times), oracle.threads(), times, stats


def reference_build_rate(v, f, n_sources=32):
    """Throughput of the reference's OWN translation units (oracle/_ref, built against the stand-in Embree/TBB/MKL/Boost headers) on a
    small sample, for the record: its ray query is the stand-in's scalar BVH, so it is SLOWER than the port and not used as the arm."""
    try:
        from oracle import oracle, reference
        if not reference.available():
            return None
        from nlos_surface_optimization_b200 import scenes
        o, n = scenes.wall_grid(WALL)
        idx = np.linspace(0, o.shape[0] - 1, n_sources).astype(int)
        o = np.ascontiguousarray(o[idx]); n = np.ascontiguousarray(n[idx])
        B = oracle.num_bins(LB, UB, RES)
        data = np.zeros((n_sources, B)); weight = np.ones((n_sources, B))
        reference.set_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        reference.gradient(o, n, v, f, SAMPLE_NUM, LB, UB, RES, data, weight, REFINE, SIGMA, 1, 0)
        dt = time.perf_counter() - t0
        spp = 1 + (SAMPLE_NUM - 1) // f.shape[0]
        return {'value': 2 * n_sources * f.shape[0] * spp / dt, 'unit': UNIT, 'sample': '%d wall points, %.1f s' % (n_sources, dt),
                'note': "the reference's unmodified sources on stand-in library headers (oracle/ref_shim): scalar double-precision ray query instead of "
                        "Embree, so slower than the port above; reported for transparency, not used for the ratio"}
    except Exception as e:     # the arm must not fail because the optional build is absent or broken
        return {'unavailable': str(e)[:200]}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on the host cores, each step a bounded sample of the workload.
    The arm is the oracle port: of the two CPU implementations available (the port, and the reference's own sources compiled against
    stand-in library headers, oracle/_ref) it is the FASTER one — the conservative choice for a GPU/CPU ratio; the other is reported
    under `reference_build`."""
    if rank != 0:
        return
    from nlos_surface_optimization_b200 import scenes
    v, f = scenes.bunny()
    n_src = 256
    for _ in range(max(args.warmup, 0) and 1):
        cpu_baseline(v, f, 8, want_stats=False)
    vals, all_t = [], []
    for _ in range(args.steps):
        val, threads, times, _ = cpu_baseline(v, f, n_src, want_stats=False)
        vals.append(val); all_t += times
    value = float(np.mean(vals))
    spp = 1 + (SAMPLE_NUM - 1) // f.shape[0]
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * float(np.mean(all_t)), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 math / f64 accumulation',
            'data': 'synthetic', 'config': {'workload': 'C-bunny (bunny F=69630, 64x64 wall, B=1200, spp=1, r=10, s=1); each step = %d of 4096 wall points' % n_src},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                             'sample': '%d of 4096 wall points x all 69630 triangles, forward+gradient, OpenMP oracle (reference-restated CPU path, not the Embree build)' % n_src},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'ms_per_iteration_extrapolated': 1e3 * (2 * 4096 * f.shape[0] * spp) / value}
    rb = reference_build_rate(v, f)
    if rb is not None:
        line['reference_build'] = rb
    print(json.dumps(line))


def fp32_peak_tflops(torch, dev, cx):
    """Measured FP32 FMA throughput (the denominator of the FP32 roofline): a torch elementwise FMA chain is not a
    pure-pipe benchmark, so use the library's own micro-kernel when present; else the nominal 148 SM x 128 lanes x 2 x clock."""
    try:
        v = float(cx.lib.nlos_microbench_fp32(cx.handle))
        if v > 0:
            return v, 'measured here (in-repo FFMA chain, csrc/microbench.cu)'
    except Exception:
        pass
    props = torch.cuda.get_device_properties(dev)
    return props.multi_processor_count * 128 * 2 * 1.965e9 / 1e12, 'nominal (SMs x 128 x 2 x 1965 MHz)'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--cpu-sources', type=int, default=1024, help='wall points of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import nlos_surface_optimization_b200 as nb
    from nlos_surface_optimization_b200 import renderer
    if args.warmup < 3:
        args.warmup = 3
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    ctx = nb.Context(local)                       # fails loudly without the CUDA library / a B200
    try:
        line = run(args, ctx, torch, dist, nb, renderer, dev, rank, world, local)
    finally:
        # ordered teardown: every tensor that lives on the context's stream must be gone before the stream is destroyed
        import gc
        gc.collect(); torch.cuda.synchronize(); torch.cuda.empty_cache()
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        ctx.close()
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)


def run(args, ctx, torch, dist, nb, renderer, dev, rank, world, local):
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)

    o, n, v, f, L, L_global = workload(rank, world)
    V, F = v.shape[0], f.shape[0]
    B = nb._arrays.num_bins(LB, UB, RES)
    spp = 1 + (SAMPLE_NUM - 1) // F
    ctx.set_source_window(rank * L, L_global)
    samples_per_step_rank = 2 * L * F * spp

    # target: the same mesh displaced by +1 cm, rendered by this library (bench needs a plausible residual, not parity)
    v2 = v.copy(); v2[:, 2] += 0.01
    data_h = np.zeros((L, B)); pl_h = np.zeros(B)
    renderer.renderStreamedTransient(o, n, v2, f, SAMPLE_NUM, LB, UB, RES, data_h, pl_h, 1, 1, ctx=ctx)
    weight_h = np.ones((L, B))

    with torch.cuda.stream(ext):
        to = lambda a: torch.from_numpy(a).to(dev)
        d_o, d_n, d_v, d_f, d_data, d_w = to(o), to(n), to(v), to(f), to(data_h), to(weight_h)
        d_T = torch.zeros((L, B), dtype=torch.float64, device=dev); d_pl = torch.zeros(B, dtype=torch.float64, device=dev)
        d_G = torch.zeros((V, 3), dtype=torch.float64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

        def step_device():
            d_G.zero_()
            renderer.renderStreamedGradient(d_o, d_n, d_v, d_f, SAMPLE_NUM, LB, UB, RES, d_T, d_pl, d_G, d_data, d_w, REFINE, SIGMA, 1, 0, ctx=ctx)
            if world > 1:
                dist.all_reduce(d_G)              # per-rank gradients are already normalised by the GLOBAL source count

        for _ in range(args.warmup):
            flush.zero_(); step_device()
        ext.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local); sampler.start()
        launches0 = ctx.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        ctx.set_option('timing', 0)
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()                          # L2 flush between timed iterations (outside the per-step events)
            ev[i][0].record(ext); step_device(); ev[i][1].record(ext)
        ext.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = ctx.launch_count() - launches0
        step_ms = [a.elapsed_time(b) for a, b in ev]
        # per-kernel breakdown of one extra step (CUDA events inside the library, on the launching stream)
        ctx.set_option('timing', 1); flush.zero_(); step_device(); ext.synchronize(); phase = ctx.timing(); ctx.set_option('timing', 0)
        clocks = sampler.stop()

    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = samples_per_step_rank * world / (ms_per_step * 1e-3)

    # ---- e2e: the reference-facing call with pinned HOST buffers, copies inside the timed region
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    h_o, h_n, h_v, h_f, h_data, h_w = pin(o), pin(n), pin(v), pin(f), pin(data_h), pin(weight_h)
    h_T = torch.zeros((L, B), dtype=torch.float64).pin_memory().numpy(); h_pl = torch.zeros(B, dtype=torch.float64).pin_memory().numpy()
    h_G = torch.zeros((V, 3), dtype=torch.float64).pin_memory().numpy()
    h2d = h_o.nbytes + h_n.nbytes + h_v.nbytes + h_f.nbytes + h_data.nbytes + h_w.nbytes + (h_G.nbytes if world == 1 else 0)
    d2h = h_T.nbytes + h_pl.nbytes + h_G.nbytes

    def step_e2e():
        if world == 1:
            h_G[:] = 0
            renderer.renderStreamedGradient(h_o, h_n, h_v, h_f, SAMPLE_NUM, LB, UB, RES, h_T, h_pl, h_G, h_data, h_w, REFINE, SIGMA, 1, 0, ctx=ctx)
        else:
            with torch.cuda.stream(ext):
                d_G.zero_()
                renderer.renderStreamedGradient(h_o, h_n, h_v, h_f, SAMPLE_NUM, LB, UB, RES, h_T, h_pl, d_G, h_data, h_w, REFINE, SIGMA, 1, 0, ctx=ctx)
                dist.all_reduce(d_G)
                torch.from_numpy(h_G).copy_(d_G, non_blocking=True)
                ext.synchronize()

    for _ in range(2):
        step_e2e()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = samples_per_step_rank * world * args.steps / e2e_s

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        cpu = None; stats = None
        if world == 1 and not args.no_cpu_baseline:
            cv, threads, times, stats = cpu_baseline(v, f, args.cpu_sources)
            cpu = {'value': cv, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                   'sample': '%d of 4096 wall points x all 69630 triangles, forward+gradient, %.1f s; OpenMP oracle = reference-restated CPU path (not the Embree build)' % (args.cpu_sources, times[0]),
                   'ms_per_iteration_extrapolated': 1e3 * samples_per_step_rank / cv}
        # FP32 roofline of the dominant kernel (forward sample kernel): algorithmic flops per path sample from the
        # canonical nearest-hit traversal counted by the oracle, divided by the kernel's event time.
        # canonical per-ray counts of the nearest-hit traversal (oracle BVH: object-median split, <=4 triangles per leaf,
        # boxes padded by scale/65536), measured once on 8 wall points of this workload and frozen here so that the
        # accounting does not move with the checker's (deliberately very conservative) culling slack
        box, tri = 58.7, 11.9
        rho = 0.41
        flops_fwd_sample = 32 + 23 * box + 50 * tri + rho * 48
        fwd_ms = phase['forward_ms']
        achieved = L * F * spp * flops_fwd_sample / (fwd_ms * 1e-3) / 1e12
        peak, peak_how = fp32_peak_tflops(torch, dev, ctx)
        try:
            red_peak = float(ctx.lib.nlos_microbench_red_f64(ctx.handle, L * B))       # FP64 RED.ADD over the transient's address range
        except Exception:
            red_peak = None
        algo_bytes = (V * 12 + F * 12 + F * 128 + (F - 1) * 64) + L * B * 8 * 4 + 3 * V * 8 + L * spp * ((F + 31) // 32) * 4 * 2
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 math / f64 accumulation', 'data': 'synthetic',
            'config': {'workload': 'C-bunny: bunny V=34817 F=69630, %dx%d confocal wall (%d points per GPU), B=1200 x 1.2 mm, sample_num=20000 (spp=1), refine_scale=10, sigma_bin=1; step = renderStreamedGradient (scene build + forward + residual + vertex gradient%s)' % (WALL * world, WALL, L, ' + NCCL all-reduce' if world > 1 else ''),
                       'l2': 'flushed between timed iterations (256 MiB write)', 'ms_per_iteration': ms_per_step, 'wall_ms_per_step_incl_flush': 1e3 * t_wall / args.steps,
                       'phase_ms': phase, 'visibility_reuse': True},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h), 'ms_per_step': 1e3 * e2e_s / args.steps},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'fp32', 'kernel': 'k_forward', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                         'traffic': 30.4e6,     # bytes: dram__bytes_read.sum + dram__bytes_write.sum of one k_forward launch, ncu --set full (profiles/r1_ncu_full_summary.txt)
                         'traffic_unit': 'bytes of DRAM per k_forward launch (ncu); the unit of achieved/peak is TFLOP/s',
                         'peak_source': peak_how, 'flops_per_path_sample': flops_fwd_sample, 'canonical_box_tests_per_ray': box, 'canonical_tri_tests_per_ray': tri,
                         'kernel_ms': fwd_ms,
                         'note': 'achieved = CANONICAL flops (oracle nearest-hit traversal counts, SURVEY 8d) / kernel time: an effective rate; the kernel executes fewer (any-hit query, zero-contribution samples never traced)',
                         'atomic': {'fp64_red_peak_Gops': red_peak, 'fp64_red_achieved_Gops': rho * L * F * spp / (fwd_ms * 1e-3) / 1e9,
                                    'frac': (rho * L * F * spp / (fwd_ms * 1e-3) / 1e9) / red_peak if red_peak else None},
                         'hbm': {'algorithmic_bytes_per_step': int(algo_bytes), 'achieved_GBps': algo_bytes / (ms_per_step * 1e-3) / 1e9, 'peak_GBps': peaks.get('hbm_gbs'),
                                 'frac': (algo_bytes / (ms_per_step * 1e-3) / 1e9) / peaks['hbm_gbs'] if peaks.get('hbm_gbs') else None}},
        }
        if cpu:
            line['cpu_baseline'] = cpu
        return line
    return None


if __name__ == '__main__':
    main()
