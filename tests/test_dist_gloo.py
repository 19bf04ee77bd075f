"""Host-side multi-process logic (ray sharding, pixel gather, flat gradient all-reduce) on the gloo
backend, world_size 2, CPU tensors.  The render itself is stubbed by a per-ray function so that the
N-rank result must equal the 1-rank result exactly."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(rays, **kw):
    o, d = rays
    rgb = torch.sin(o * 3.0 + d)                     # any per-ray function
    return rgb, (o * d).sum(-1), d.norm(dim=-1), {}


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerf_pytorch_b200 import dist as nd
    g = torch.Generator().manual_seed(0)
    rays = torch.randn(2, n, 3, generator=g)
    rgb, disp, acc, _ = nd.render_sharded(_fake_render, rays)
    ref = _fake_render(rays)
    ok = torch.equal(rgb, ref[0]) and torch.equal(disp, ref[1]) and torch.equal(acc, ref[2])
    # data-parallel gradient: mean over the global batch == average of per-rank means
    w = torch.nn.Parameter(torch.arange(6.0).reshape(2, 3))
    b = torch.nn.Parameter(torch.ones(3))
    lo, hi = nd.shard_bounds(n - n % world, rank, world)           # equal shards for an exact identity
    x = rays[0, lo:hi]
    loss = (((x * b).sum(-1)) ** 2).mean() + (w ** 2).sum() * x.mean()
    loss.backward()
    nd.allreduce_grads([w, b])
    xs = rays[0, : n - n % world]
    w2 = torch.nn.Parameter(torch.arange(6.0).reshape(2, 3)); b2 = torch.nn.Parameter(torch.ones(3))
    # reference: average of the per-shard losses
    tot = 0
    for r in range(world):
        l2, h2 = nd.shard_bounds(n - n % world, r, world)
        xr = xs[l2:h2]
        tot = tot + (((xr * b2).sum(-1)) ** 2).mean() + (w2 ** 2).sum() * xr.mean()
    (tot / world).backward()
    ok = ok and torch.allclose(w.grad, w2.grad, atol=1e-6) and torch.allclose(b.grad, b2.grad, atol=1e-6)
    p = torch.nn.Parameter(torch.full((4,), float(rank)))
    v0 = p._version
    nd.broadcast_params([p], src=0)
    ok = ok and bool((p == 0).all())
    # the broadcast writes through the Parameter itself (not .data): the version counter that NeRF.packed() keys its
    # fp16 weight cache on must move, otherwise non-zero ranks would keep rendering with stale tensor-core weights
    ok = ok and p._version > v0
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7])
def test_two_rank_shard_gather_allreduce(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from nerf_pytorch_b200 import dist as nd
    for n in (0, 1, 5, 4096, 640000):
        for world in (1, 2, 3, 8):
            spans = [nd.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
