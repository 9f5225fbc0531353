"""Generates tests/golden/chamfer_cpu_ref.npz by running the REFERENCE's own C++ CPU Chamfer
(cuda/chamfer_distance/chamfer_distance.cpp:57-180, built unmodified by oracle/build_ref.py into
oracle/_ref/cd.so) in the build container.  Committed together with its output; the GPU box never
runs this (no /root/reference there)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.build_ref import load_ref  # noqa: E402

cd = load_ref("cd")


def run(x, y, g1, g2):
    B, N, _ = x.shape
    M = y.shape[1]
    d1, d2 = torch.zeros(B, N), torch.zeros(B, M)
    i1, i2 = torch.zeros(B, N, dtype=torch.int32), torch.zeros(B, M, dtype=torch.int32)
    cd.forward(x, y, d1, d2, i1, i2)
    gx, gy = torch.zeros_like(x), torch.zeros_like(y)
    cd.backward(x, y, gx, gy, g1, g2, i1, i2)
    return d1, d2, i1, i2, gx, gy


out = {}
# case A: BASELINE config 1: B=2, N=M=1024, seed 0
torch.manual_seed(0)
x, y = torch.rand(2, 1024, 3), torch.rand(2, 1024, 3)
g1, g2 = torch.rand(2, 1024), torch.rand(2, 1024)
# case B: ragged N != M
torch.manual_seed(1)
xb, yb = torch.rand(3, 300, 3) - 0.5, torch.rand(3, 517, 3) - 0.5
g1b, g2b = torch.rand(3, 300), torch.rand(3, 517)
# case C: exact ties (duplicated reference points + zero-padded tail, as the real loader pads)
torch.manual_seed(2)
xc, yc = torch.rand(1, 256, 3), torch.rand(1, 128, 3)
yc = torch.cat([yc, yc], 1)            # every reference point appears twice
xc[:, 200:] = 0.0                      # zero padding
g1c, g2c = torch.rand(1, 256), torch.rand(1, 256)
for tag, args in (("a", (x, y, g1, g2)), ("b", (xb, yb, g1b, g2b)), ("c", (xc, yc, g1c, g2c))):
    res = run(*args)
    for name, t in zip(("x", "y", "g1", "g2"), args):
        out[f"{tag}_{name}"] = t.numpy()
    for name, t in zip(("d1", "d2", "i1", "i2", "gx", "gy"), res):
        out[f"{tag}_{name}"] = t.numpy()
np.savez_compressed(os.path.join(os.path.dirname(__file__), "chamfer_cpu_ref.npz"), **out)
print("wrote chamfer_cpu_ref.npz", {k: v.shape for k, v in out.items() if k.startswith("a_")})
