"""Ray-parallel multi-GPU helpers (one process per GPU, torch.distributed; NCCL on GPUs).

The path shards naturally: every op of render_rays is per ray, so the forward render
needs no data-path collective -- each rank renders a contiguous stripe of the rays with replicated
weights (4.77 MB).  Collectives appear only at the edges:
  * gather_pixels : all-gather of the per-rank output stripes ([rays, C]) -- full-image render;
  * allreduce_grads: ONE all-reduce of a flat fp32 gradient buffer per training step (data parallel);
  * broadcast_params: weights from rank 0 at start / after a checkpoint load.
The same code runs on the `gloo` backend with CPU tensors (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced stripe [lo, hi) of n rays for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays: torch.Tensor, rank: int | None = None, world: int | None = None, dim: int = -2):
    """Slice a ray tensor ([2,N,3] or [N,C]) along the ray axis for this rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(rays.shape[dim], rank, world)
    return rays.narrow(dim, lo, hi - lo)


def gather_pixels(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather ragged per-rank stripes [n_r, ...] into [n_total, ...] in ray order."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((maxn,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


def allreduce_grads(params, average: bool = True, group=None):
    """One all-reduce over a flat fp32 buffer holding every .grad (1 191 688 elements = 4.77 MB for
    coarse + fine).  With each rank's loss being the mean over its own N/world rays, averaging the
    gradients reproduces the gradient of the global-batch mean (run_nerf_helpers.py:9)."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    world = dist.get_world_size(group)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def broadcast_params(params, src: int = 0, group=None):
    params = list(params)
    if dist.get_world_size(group) == 1 or not params:
        return
    with torch.no_grad():
        flat = torch.cat([p.reshape(-1) for p in params])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))        # in-place on the Parameter: bumps its version -> NeRF.packed() re-packs
            off += n


def render_sharded(render_fn, rays: torch.Tensor, gather: bool = True, **kw):
    """Ray-parallel render of rays [2,N,3]: each rank renders its stripe with `render_fn`
    (nerf_b200.render signature) and, if `gather`, every rank receives the full rgb/disp/acc maps."""
    n = rays.shape[1]
    local = shard_rays(rays, dim=1)
    rgb, disp, acc, extras = render_fn(rays=local, **kw)
    if not gather:
        return rgb, disp, acc, extras
    packed = torch.cat([rgb, disp[:, None], acc[:, None]], -1)
    full = gather_pixels(packed, n)
    return full[:, :3], full[:, 3], full[:, 4], extras
