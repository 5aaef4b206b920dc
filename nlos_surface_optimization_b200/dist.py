"""Multi-GPU sharding of the renderer: one process per GPU, each holds the full mesh and BVH and
renders a contiguous slice of the relay-wall sources; transient rows are disjoint (no reduction), the vertex /
scalar gradient is summed with ONE all-reduce per iteration (NCCL over NVLink for CUDA tensors; gloo in the
CPU tests).  The reference has no counterpart (single process, TBB threads only).

Samples are keyed by the GLOBAL source index (nlos_ctx_set_source_window), and every rank normalises by the
GLOBAL source count, so the sum over ranks equals the single-GPU result to FP64 round-off.
"""
import numpy as np


def shard_range(num_sources, rank, world_size):
    """Contiguous slice [start, stop) of the sources owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(num_sources), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _cuda_render(origin, normal, vertices, faces, num_sample, lower, upper, resolution, data, weight, refine_scale, sigma_bin, testing_flag,
                 loss_flag, src_offset, num_sources_global, ctx=None):
    """Default per-shard render: the sm_100a path; the gradient comes back normalised by the GLOBAL source count."""
    from . import _ffi, renderer
    cx = ctx or _ffi.default_context()
    L, B = data.shape
    transient = np.zeros((L, B)); pathlengths = np.zeros(B); gradient = np.zeros((vertices.shape[0], 3))
    cx.set_source_window(src_offset, num_sources_global)
    try:
        renderer.renderStreamedGradient(origin, normal, vertices, faces, num_sample, lower, upper, resolution, transient, pathlengths, gradient,
                                        data, weight, refine_scale, sigma_bin, testing_flag, loss_flag, ctx=cx)
    finally:
        cx.set_source_window(0, 0)
    return transient, gradient, pathlengths


def all_reduce_sum(array, group=None, device=None):
    """Sum a NumPy array over the process group (in place) through torch.distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return array
    t = torch.from_numpy(array)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    if device is not None:
        array[...] = t.cpu().numpy()
    return array


def inverse_rendering_sharded(origin, normal, vertices, faces, num_sample, lower, upper, resolution, data, weight, refine_scale, sigma_bin,
                              testing_flag=1, loss_flag=0, rank=None, world_size=None, group=None, device=None, render_fn=None, gather=False):
    """Forward + vertex gradient of ALL sources, computed cooperatively.

    Every rank passes the full `origin/normal/data/weight` (or at least its own slice's rows are read); returns
    (transient, gradient, pathlengths) where `gradient` [V,3] is the all-reduced global-mean gradient (identical on
    all ranks) and `transient` is this rank's rows [start:stop] (or the full [L,B] array when gather=True).
    """
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_initialized() else 1
    L = origin.shape[0]
    a, b = shard_range(L, rank, world_size)
    fn = render_fn or _cuda_render
    sl = slice(a, b)
    T, G, pl = fn(np.ascontiguousarray(origin[sl]), np.ascontiguousarray(normal[sl]), vertices, faces, num_sample, lower, upper, resolution,
                  np.ascontiguousarray(data[sl]), np.ascontiguousarray(weight[sl]), refine_scale, sigma_bin, testing_flag, loss_flag, a, L)
    all_reduce_sum(G, group=group, device=device)
    if gather and world_size > 1:
        import torch
        parts = [None] * world_size
        dist.all_gather_object(parts, (a, b, T), group=group)
        full = np.zeros((L, T.shape[1]))
        for pa, pb, pt in parts:
            full[pa:pb] = pt
        T = full
    return T, G, pl
