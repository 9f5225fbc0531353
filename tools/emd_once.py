"""One EMD launch (development tool, e.g. under ncu): python tools/emd_once.py [N] [iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparenet_b200 import functional as F_  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
it = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda:0")
torch.manual_seed(4)
x = torch.rand(32, N, 3, device=dev)
torch.manual_seed(5)
y = torch.rand(32, N, 3, device=dev)
d, a = F_.emd_forward(x, y, 0.005, it)
torch.cuda.synchronize()
print("emd", N, it, float(d.sqrt().mean()))
