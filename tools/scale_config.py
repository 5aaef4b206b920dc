"""C-scale (BASELINE.json configs[4]): synthetic 500k-triangle height field, 256x256 confocal wall, 2048 bins, sources
sharded over the ranks (strong scaling), one NCCL all-reduce of the gradient per iteration.

    python tools/scale_config.py [--wall 256] [--n 501] [--steps 2]                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/scale_config.py ...
Prints one JSON line on rank 0 (device-resident inputs, CUDA-event timing, max over ranks)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np
import torch
import torch.distributed as dist
import nlos_surface_optimization_b200 as nb
from nlos_surface_optimization_b200 import renderer, scenes
from nlos_surface_optimization_b200.dist import shard_range

ap = argparse.ArgumentParser(); ap.add_argument('--wall', type=int, default=256); ap.add_argument('--n', type=int, default=501)
ap.add_argument('--steps', type=int, default=2); ap.add_argument('--bins', type=int, default=2048)
args = ap.parse_args()
rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1)); local = int(os.environ.get('LOCAL_RANK', 0))
dev = torch.device('cuda', local); torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
ctx = nb.Context(local)
try:
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    v, f = scenes.heightfield(args.n)
    o, n = scenes.wall_grid(args.wall)
    Lg = o.shape[0]; a, b = shard_range(Lg, rank, world); L = b - a
    res = 1.2e-3; B = args.bins; ub = float(np.float32(B) * np.float32(res))
    assert nb._arrays.num_bins(0.0, ub, res) == B
    ns = f.shape[0]                       # spp = 1
    ctx.set_source_window(a, Lg)
    with torch.cuda.stream(ext):
        to = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        d_o, d_n, d_v, d_f = to(o[a:b]), to(n[a:b]), to(v), to(f)
        d_T = torch.zeros((L, B), dtype=torch.float64, device=dev); d_pl = torch.zeros(B, dtype=torch.float64, device=dev)
        d_data = torch.zeros((L, B), dtype=torch.float64, device=dev); d_w = torch.ones((L, B), dtype=torch.float64, device=dev)
        d_G = torch.zeros((v.shape[0], 3), dtype=torch.float64, device=dev)
        v2 = v.copy(); v2[:, 2] += 0.01
        renderer.renderStreamedTransient(d_o, d_n, to(v2), d_f, ns, 0.0, ub, res, d_data, d_pl, 1, 1, ctx=ctx)
        ext.synchronize()
        times = []; phases = None
        for i in range(args.steps + 1):
            ctx.set_option('timing', 1 if i == args.steps else 0)
            d_G.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            renderer.renderStreamedGradient(d_o, d_n, d_v, d_f, ns, 0.0, ub, res, d_T, d_pl, d_G, d_data, d_w, 10, 1, 1, 0, ctx=ctx)
            if world > 1:
                dist.all_reduce(d_G)
            e1.record(ext); ext.synchronize()
            if i > 0:
                times.append(e0.elapsed_time(e1))
            if i == args.steps:
                phases = ctx.timing()
        ms = float(np.mean(times)); t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        gsum = float(d_G.abs().sum().item()); tsum = float(d_T.sum().item())
    if rank == 0:
        print(json.dumps({'config': 'C-scale', 'F': int(f.shape[0]), 'V': int(v.shape[0]), 'L': int(Lg), 'bins': B, 'n_gpus': world, 'sources_per_gpu': int(L),
                          'ms_per_iteration': ms, 'path_samples_per_s': 2.0 * Lg * f.shape[0] / (ms * 1e-3), 'phase_ms_rank0': phases,
                          'grad_abs_sum': gsum, 'transient_sum_rank0': tsum, 'scaling': 'strong'}), flush=True)
    del d_o, d_n, d_v, d_f, d_T, d_pl, d_data, d_w, d_G
finally:
    import gc; gc.collect(); torch.cuda.synchronize(); torch.cuda.empty_cache()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    ctx.close()
