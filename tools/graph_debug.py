"""Find what breaks CUDA-graph capture of the bench step (development tool)."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class A:
    batch = int(os.environ.get("PB", "32"))


dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
step, hp, hg = bench.build_gpu(A, dev, 0)
p, g = hp.to(dev), hg.to(dev)
for _ in range(2):
    step(p, g)
torch.cuda.synchronize()
import gc  # noqa: E402

live = [o for o in gc.get_objects() if isinstance(o, torch.Tensor) and o.grad_fn is not None]
print("tensors still holding a graph after the eager steps:", len(live))
for o in live[:20]:
    print("   ", tuple(o.shape), type(o.grad_fn).__name__, "referrers:", [type(r).__name__ for r in gc.get_referrers(o)][:6])
del live
gc.collect()
live = [o for o in gc.get_objects() if isinstance(o, torch.Tensor) and o.grad_fn is not None]
print("after gc.collect():", len(live))
del live
mode = os.environ.get("MODE", "fwd")
for prm in step.params:
    prm.grad = None
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    loss = step.loss_fn(p, g)
    if mode != "fwd":
        loss.backward()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
for prm in step.params:
    prm.grad = None
gr = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(gr, capture_error_mode=os.environ.get("CEM", "global")):
        loss = step.loss_fn(p, g)
        if mode != "fwd":
            loss.backward()
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    print("capture OK, mode", mode, "loss", float(loss))
except Exception:
    traceback.print_exc()
