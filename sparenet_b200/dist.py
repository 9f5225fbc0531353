"""Data-parallel plumbing for the hot path: one process per GPU, batches sharded by sample, ONE collective per optimizer
step -- the gradient all-reduce (the reference wraps everything in nn.DataParallel,
runners/base_runner.py:100-104).  Every op on the path is independent per sample, and the two cross-sample couplings
(train-mode BatchNorm statistics, ComputeDepthMaps' depth min/max) are per-replica in the reference too, so rank-local
statistics with local B=32 reproduce its single-GPU numerics on every rank.

torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is the transport; there is no compute here.
"""
import torch
import torch.distributed as dist


def shard_batch(global_batch: int, rank: int, world: int):
    """Contiguous sample shard [lo, hi) of rank `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradArena:
    """All gradients of a parameter list in ONE flat buffer per (dtype, device): every p.grad is a 16-byte-aligned view into it, so
    the gradient all-reduce runs in place on a few large chunks -- no concatenation before and no copy back after the collective
    (round 1's flat-cat buckets moved ~1.3 GB per step for 330 MB of gradients).  Gradients ACCUMULATE into the views: call zero()
    at the start of every step instead of zero_grad(set_to_none=True) (which would drop the views)."""

    def __init__(self, params, own_grads=True):
        """own_grads=True: every p.grad becomes a view into the arena (gradients ACCUMULATE into it: one add per parameter and
        backward).  own_grads=False: autograd keeps assigning fresh gradient tensors (no accumulation launches) and pack() gathers
        them into the arena with a few multi-tensor copies after the backward."""
        self.params = [p for p in params if p.requires_grad]
        self.own_grads = own_grads
        groups = {}
        for p in self.params:
            groups.setdefault((p.dtype, p.device), []).append(p)
        self.flats, self.views = [], {}
        for (dt, dev), ps in groups.items():
            offs, n = [], 0
            for p in ps:
                offs.append(n)
                n += (p.numel() + 3) // 4 * 4                     # keep every view 16-byte aligned for vectorised optimizers
            flat = torch.zeros(max(n, 4), dtype=dt, device=dev)
            for p, o in zip(ps, offs):
                self.views[id(p)] = flat[o:o + p.numel()].view_as(p)
                if own_grads:
                    p.grad = self.views[id(p)]
            self.flats.append(flat)
        self._written = set()      # own_grads=False: ids of the parameters whose arena slot pack() has ever copied a gradient into

    def pack(self):
        """own_grads=False: copy the parameters' current .grad tensors into the arena (multi-tensor copies; a parameter without a
        gradient contributes zeros).  Graph-capturable.  The arena starts as zeros and only pack() writes into it (the all-reduce
        averages zeros to zeros), so the slot of a parameter that has never had a gradient is still zero: it is cleared only after
        an earlier pack() copied something there -- the generator has ~100 such parameters (biases that cancel under the instance
        norm, the unused top-level conv1), which used to cost one fill launch each per step."""
        dst, src, clear = [], [], []
        for p in self.params:
            v = self.views[id(p)]
            if p.grad is None:
                if id(p) in self._written:
                    clear.append(v)
                    self._written.discard(id(p))
            elif p.grad.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(p.grad)
                self._written.add(id(p))
        if clear:
            torch._foreach_zero_(clear)
        if not dst:
            return
        if all(s.is_cuda and s.dtype == torch.float32 and s.is_contiguous() and v.dtype == torch.float32 for v, s in zip(dst, src)):
            self._pack_one_launch(dst, src)
        else:
            torch._foreach_copy_(dst, src)

    def _pack_one_launch(self, dst, src):
        """One launch per 1024 tensors over a pointer table (csrc/optim.cu: snb_multi_copy) instead of ~11 multi-tensor launches.  The
        table is a kernel parameter: nothing is uploaded, and under CUDA-graph capture the call is an ordinary kernel node (the
        captured tensors' addresses are static)."""
        import ctypes
        from . import _lib
        n = len(dst)
        srcs = (ctypes.c_void_p * n)(*[s.data_ptr() for s in src])
        dsts = (ctypes.c_void_p * n)(*[v.data_ptr() for v in dst])
        ns = (ctypes.c_longlong * n)(*[s.numel() for s in src])
        with torch.cuda.device(dst[0].device):
            _lib.check(_lib.load().snb_multi_copy(srcs, dsts, ns, n, _lib.stream_ptr()), "multi_copy")

    def zero(self):
        for f in self.flats:
            f.zero_()

    def allreduce(self, world: int = None, chunk_bytes: int = 128 << 20):
        """Average over ranks, in place, chunk by chunk (NCCL averages inside the collective; gloo sums and divides)."""
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
        if world == 1:
            return 0
        avg = dist.get_backend() == "nccl"
        n_calls = 0
        for f in self.flats:
            step = max(1, chunk_bytes // f.element_size())
            for c in f.split(step):
                if avg:
                    dist.all_reduce(c, op=dist.ReduceOp.AVG)
                else:
                    dist.all_reduce(c)
                    c.div_(world)
                n_calls += 1
        return n_calls


def allreduce_gradients(params, world: int = None, bucket_bytes: int = 256 << 20):
    """Average .grad over ranks in flat buckets (few large NCCL calls: 330 MB of generator gradients per step).
    Parameters whose .grad is None on this rank are skipped consistently (same set on every rank by construction).
    Per bucket: ONE concatenation kernel, one all-reduce (NCCL averages in the collective; gloo sums and divides), and one
    multi-tensor copy back -- not one launch per parameter (the generator has ~800 of them)."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return 0
    avg = dist.get_backend() == "nccl"
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    if len({g.dtype for g in grads}) > 1:                         # one pass per dtype: torch.cat would type-promote a mixed bucket
        return sum(allreduce_gradients([p for p in params if p.grad is not None and p.grad.dtype == dt], world, bucket_bytes)
                   for dt in sorted({g.dtype for g in grads}, key=str))
    n_calls, bucket, size = 0, [], 0
    for g in grads + [None]:
        if g is not None and (size + g.numel() * g.element_size() <= bucket_bytes or not bucket):
            bucket.append(g)
            size += g.numel() * g.element_size()
            continue
        flat = torch.cat([t.reshape(-1) for t in bucket])
        if avg:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(flat)
            flat.div_(world)
        torch._foreach_copy_(bucket, [v.view_as(t) for v, t in zip(flat.split([t.numel() for t in bucket]), bucket)])
        n_calls += 1
        bucket, size = ([g], g.numel() * g.element_size()) if g is not None else ([], 0)
    return n_calls


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
