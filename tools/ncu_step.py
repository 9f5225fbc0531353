"""Two steps of the bench workload for ncu captures (development tool; numbers printed under ncu are never bench values)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class A:
    batch = 32


dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
step, hp, hg = bench.build_gpu(A, dev, 0)
p, g = hp.to(dev), hg.to(dev)
for _ in range(int(os.environ.get("NSTEPS", "2"))):
    step(p, g)
torch.cuda.synchronize()
