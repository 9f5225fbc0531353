"""Side-stream overlap of the coarse/middle Chamfer losses (bench.py) must not change the step: first-step loss and gradients of
overlap vs no-overlap vs a repeat of no-overlap (the float-atomic noise floor).  Development tool."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda:0")


def first_step(no_overlap, graph=False):
    args = types.SimpleNamespace(gpus=1, steps=1, warmup=1, no_graph=True, no_cpu_baseline=True, batch=8, no_overlap=no_overlap)
    step, hp, hg = bench.build_gpu(args, dev, 0)
    p, g = hp.to(dev), hg.to(dev)
    if graph:
        from sparenet_b200.graph import GraphedForwardBackward
        gfb = GraphedForwardBackward(step.loss_fn, step.params, (p, g), warmup=1)
        loss = gfb(p, g)
    else:
        loss = step.loss_fn(p, g)
        loss.backward()
    torch.cuda.synchronize()
    return loss.item(), [q.grad.clone() for q in step.params if q.grad is not None]


def dev_(a, b):
    gmax = max(x.abs().max().item() for x in a)
    return max(((x - y).abs().max().item()) / (x.abs().max().item() + 1e-4 * gmax) for x, y in zip(a, b))


l0, g0 = first_step(True)
l1, g1 = first_step(True)
l2, g2 = first_step(False)
l3, g3 = first_step(False, graph=True)
print(f"loss no-overlap {l0:.9f} repeat {l1:.9f} overlap {l2:.9f} overlap+graph {l3:.9f}")
print(f"worst relative gradient deviation: repeat {dev_(g0, g1):.3e}  overlap {dev_(g0, g2):.3e}  overlap+graph {dev_(g0, g3):.3e}")
