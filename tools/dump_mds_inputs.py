"""Dump the (xyz, m, mean_mst_length) of the MDS calls of one bench step (development tool) -> gpurun_out/mds_inputs.pt"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sparenet_b200 import functional as F_  # noqa: E402

dev = torch.device("cuda:0")
calls = []
orig = F_.mds_sample


def spy(xyz, npoint, mml):
    calls.append((xyz.detach().cpu().clone(), int(npoint), mml.detach().cpu().clone()))
    return orig(xyz, npoint, mml)


F_.mds_sample = spy
args = types.SimpleNamespace(gpus=1, steps=1, warmup=1, no_graph=True, no_cpu_baseline=True, batch=bench.LOCAL_B)
step, h_partial, h_gt = bench.build_gpu(args, dev, 0)
nsteps = int(os.environ.get("NSTEPS", "1"))
for _ in range(nsteps):
    calls.clear()
    step(h_partial.to(dev), h_gt.to(dev))
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
torch.save(calls, os.path.join(ROOT, "gpurun_out", "mds_inputs.pt"))
for x, m, mml in calls:
    print("mds call: xyz", tuple(x.shape), "m", m, "mml mean %.5f min %.5f max %.5f" % (mml.mean(), mml.min(), mml.max()),
          "extent", (x.amax(1) - x.amin(1)).mean(0).tolist())
