"""How much of the MDS update is exactly zero?  (development tool)
expf(-d/t) == 0 for d > ~104 t, so a round only changes the densities of the points within sqrt(105 t) of the chosen point.
Captures the MDS inputs of one bench step and reports the fraction of (chosen point, point) pairs inside that radius."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sparenet_b200 import functional as F_  # noqa: E402


class A:
    batch = 32


dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
calls = []
orig = F_.mds_sample


def spy(xyz, npoint, mml):
    idx = orig(xyz, npoint, mml)
    calls.append((xyz.detach().clone(), mml.detach().clone(), idx.clone()))
    return idx


F_.mds_sample = spy
import sparenet_b200.dropin.cuda.MDS.MDS_module as M  # noqa: E402
M.F_.mds_sample = spy
step, hp, hg = bench.build_gpu(A, dev, 0)
p, g = hp.to(dev), hg.to(dev)
for it in range(3):
    calls.clear()
    step(p, g)
torch.cuda.synchronize()
for ci, (xyz, mml, idx) in enumerate(calls):
    B, n, _ = xyz.shape
    t = 5.0 * mml.double() ** 2
    ext = (xyz.amax(1) - xyz.amin(1)).mean(0).tolist()
    sel = idx[:, torch.linspace(0, idx.shape[1] - 1, 256, device=dev).long()].long()
    w = torch.gather(xyz, 1, sel[..., None].expand(-1, -1, 3))
    d = torch.cdist(w.double(), xyz.double()) ** 2
    inside = (d <= 105.0 * t[:, None, None]).double().mean().item()
    print(f"call {ci}: n={n} m={idx.shape[1]} mml mean {mml.mean().item():.5f}  cutoff radius {torch.sqrt(105 * t).mean().item():.4f} "
          f"extent {['%.3f' % e for e in ext]}  pairs inside the cutoff: {100 * inside:.2f} %")
