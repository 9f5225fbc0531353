"""Every tcgen05 GEMM call of one eager bench step with its shape, time and TFLOP/s (development tool).  python tools/gemm_shapes.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sparenet_b200 import gemm  # noqa: E402


class A:
    batch = 32
    torch_adam = False


dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
step, hp, hg = bench.build_gpu(A, dev, 0)
p, g = hp.to(dev), hg.to(dev)
for _ in range(3):
    step(p, g)
torch.cuda.synchronize()
gemm.SHAPE_LOG = []
step(p, g)
torch.cuda.synchronize()
log, gemm.SHAPE_LOG = gemm.SHAPE_LOG, None
tot = 0.0
rows = []
for label, a, b, fl in log:
    ms = a.elapsed_time(b)
    tot += ms
    rows.append((ms, label, fl))
print(f"{len(rows)} GEMM calls, {tot:.3f} ms (eager, events around each call)")
for i, (ms, label, fl) in enumerate(rows):
    print(f"{i:3d} {ms * 1e3:8.1f} us {fl / (ms * 1e-3) / 1e12:7.1f} TFLOP/s  {label}")
