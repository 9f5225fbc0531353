"""The N>1 path on CPU: world_size-2 gloo processes exercise the sharding and the bucketed gradient all-reduce that
bench.py uses under torchrun (NCCL there, gloo here).  No GPU, no kernels."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparenet_b200.dist import allreduce_gradients, max_over_ranks, shard_batch
    torch.manual_seed(0)                       # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3), torch.nn.Linear(3, 1, bias=False))
    net[3].weight.requires_grad_(True)
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0       # the global batch, same on every rank
    lo, hi = shard_batch(8, rank, world)
    loss = net(x[lo:hi]).pow(2).sum() / 8.0    # per-rank share of the global mean loss
    loss.backward()
    calls = allreduce_gradients(list(net.parameters()), world, bucket_bytes=256)   # tiny buckets -> several collectives
    for p in net.parameters():
        p.grad.mul_(world)                     # undo the average: sum over shards == full-batch gradient
    ref = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3), torch.nn.Linear(3, 1, bias=False))
    ref.load_state_dict(net.state_dict())
    (ref(x).pow(2).sum() / 8.0).backward()
    ok = all(torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-6) for a, b in zip(net.parameters(), ref.parameters()))
    mx = max_over_ranks(float(rank + 1), torch.device("cpu"))
    out[rank] = (ok, calls, mx, (lo, hi))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_full_batch():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        ok, calls, mx, shard = out[rank]
        assert ok and calls >= 2 and mx == 2.0
    assert out[0][3] == (0, 4) and out[1][3] == (4, 8)


def test_shard_batch_covers_everything():
    from sparenet_b200.dist import shard_batch
    for gb in (32, 33, 255, 256):
        for world in (1, 2, 3, 8):
            cuts = [shard_batch(gb, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == gb
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1


def _arena_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparenet_b200.dist import GradArena, allreduce_gradients, shard_batch
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 17), torch.nn.ReLU(), torch.nn.Linear(17, 3), torch.nn.Linear(3, 1, bias=False))
    arena = GradArena(net.parameters())
    views = [p.grad for p in net.parameters()]
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0
    lo, hi = shard_batch(8, rank, world)
    ok = True
    for step in range(2):                      # the views must survive a second step (zero() instead of zero_grad(set_to_none))
        arena.zero()
        (net(x[lo:hi]).pow(2).sum() / 8.0).backward()
        ok &= all(p.grad is v for p, v in zip(net.parameters(), views))                       # autograd accumulated IN the arena
        calls = arena.allreduce(world, chunk_bytes=64)                                         # several in-place collectives
        ref = torch.nn.Sequential(torch.nn.Linear(6, 17), torch.nn.ReLU(), torch.nn.Linear(17, 3), torch.nn.Linear(3, 1, bias=False))
        ref.load_state_dict(net.state_dict())
        (ref(x).pow(2).sum() / 8.0).backward()
        ok &= all(torch.allclose(a.grad * world, b.grad, rtol=1e-5, atol=1e-6) for a, b in zip(net.parameters(), ref.parameters()))
    ok &= all(v.data_ptr() % 16 == 0 for v in views)
    # the flat-bucket path: nothing to reduce must be a no-op, mixed dtypes must not be promoted
    empty = [torch.nn.Parameter(torch.zeros(3))]
    ok &= allreduce_gradients(empty, world) == 0
    a, b = torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(4, dtype=torch.float64))
    a.grad, b.grad = torch.full((4,), float(rank)), torch.full((4,), float(rank) + 0.25, dtype=torch.float64)
    allreduce_gradients([a, b], world)
    ok &= a.grad.dtype == torch.float32 and b.grad.dtype == torch.float64
    ok &= torch.allclose(a.grad, torch.full((4,), 0.5)) and torch.allclose(b.grad, torch.full((4,), 0.75, dtype=torch.float64))
    out[rank] = (bool(ok), calls)
    dist.destroy_process_group()


def test_two_rank_grad_arena_allreduce_in_place():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_arena_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for rank in range(world):
        ok, calls = out[rank]
        assert ok and calls >= 2


def test_grad_arena_pack_mode_cpu():
    """GradArena(own_grads=False): autograd keeps assigning fresh gradient tensors; pack() gathers them into the flat buffer (zeros
    for a parameter without gradient), views stay 16-byte aligned, and a second pack after new gradients overwrites the first."""
    from sparenet_b200.dist import GradArena
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(*s)) for s in ((5, 3), (7,), (2, 2, 2), (1,))]
    arena = GradArena(ps, own_grads=False)
    assert all(p.grad is None for p in ps)
    for round_ in range(2):
        for i, p in enumerate(ps):
            p.grad = None if i == 1 else torch.full_like(p, float(10 * round_ + i + 1))
        arena.pack()
        flat = arena.flats[0]
        for i, p in enumerate(ps):
            v = arena.views[id(p)]
            assert (v.data_ptr() - flat.data_ptr()) % 16 == 0
            want = 0.0 if i == 1 else float(10 * round_ + i + 1)
            assert torch.equal(v, torch.full_like(p, want))
    assert flat.numel() == 16 + 8 + 8 + 4      # every view padded to a multiple of 4 elements
    # a parameter that HAD a gradient and then has none: its slot is cleared (once); one that never had any costs nothing
    ps[0].grad = None
    arena.pack()
    assert torch.equal(arena.views[id(ps[0])], torch.zeros_like(ps[0])) and torch.equal(arena.views[id(ps[1])], torch.zeros_like(ps[1]))
    assert id(ps[0]) not in arena._written and id(ps[1]) not in arena._written and id(ps[2]) in arena._written
