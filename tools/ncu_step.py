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
n = int(os.environ.get("NSTEPS", "2"))
rng = os.environ.get("NCU_RANGE") == "1"   # with `ncu --profile-from-start off`: profile only the last step
for i in range(n):
    if rng and i == n - 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    step(p, g)
torch.cuda.synchronize()
if rng:
    torch.cuda.cudart().cudaProfilerStop()
