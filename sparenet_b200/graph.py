"""CUDA-graph capture of one training step's forward + backward (host plumbing; no arithmetic here).

The generator step issues ~2000 launches, most of them microsecond-sized closed-form normalisation math between the big
kernels; replaying them as one graph removes the per-launch gaps.  The optimizer step (and, under data parallelism, the
gradient all-reduce) stays outside the graph so the same capture serves 1 and N GPUs.

    step = GraphedForwardBackward(loss_fn, params, example_inputs)   # warms up on a side stream, then captures
    loss = step(partial, gt)        # copies into the static inputs (host or device sources), replays, returns the static loss
    optimizer.step()                # gradients live in static .grad tensors: never call zero_grad(set_to_none=True) afterwards

Everything the sm_100a kernels need during capture is capture-safe: they launch on the current stream, take their scratch
from the caching allocator (graph-private pool) and never synchronise.
"""
import torch


class GraphedForwardBackward:
    def __init__(self, loss_fn, params, example_inputs, warmup=3, zero_fn=None, post_fn=None):
        """zero_fn: when the gradients live in a persistent buffer (sparenet_b200.dist.GradArena) they are zeroed by this callable --
        captured at the head of the graph -- instead of being dropped (p.grad = None) before the capture.
        post_fn: captured right after the backward (e.g. GradArena.pack: gather the gradients into the flat arena)."""
        self.params = [p for p in params if p.requires_grad]
        self.static_inputs = [torch.empty_like(t, device=self.params[0].device) for t in example_inputs]
        for s, t in zip(self.static_inputs, example_inputs):
            s.copy_(t)
        dev = self.params[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                if zero_fn is not None:
                    zero_fn()
                else:
                    for p in self.params:
                        p.grad = None
                loss_fn(*self.static_inputs).backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        if zero_fn is None:
            for p in self.params:
                p.grad = None
        self.graph = torch.cuda.CUDAGraph()
        # Captured on a HIGH-priority stream: kernel nodes inherit the priority of the stream they were captured from, so work that
        # loss_fn forks onto ordinary (lowest-priority) side streams -- bench.py's coarse / middle Chamfer searches beside the
        # sampler -- never holds back the main chain's blocks: the block scheduler serves pending high-priority blocks first.
        # (Measured without it: the 5 us expansion_mean launch waited ~0.7 ms behind the side stream's 4096-block Chamfer query.)
        cap = torch.cuda.Stream(device=dev, priority=-1)
        with torch.cuda.graph(self.graph, stream=cap):
            if zero_fn is not None:
                zero_fn()
            self.static_loss = loss_fn(*self.static_inputs)
            self.static_loss.backward()
            if post_fn is not None:
                post_fn()

    def __call__(self, *inputs):
        for s, t in zip(self.static_inputs, inputs):
            if t is not s:
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_loss
