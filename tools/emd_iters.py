"""EMD time as a function of the iteration count (development tool): the increments are the cost of each group of rounds."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparenet_b200 import functional as F_  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
B = 32
dev = torch.device("cuda:0")
torch.manual_seed(4)
x = torch.rand(B, N, 3, device=dev)
torch.manual_seed(5)
y = torch.rand(B, N, 3, device=dev)
prev, pi = 0.0, 0
for it in (1, 2, 3, 5, 10, 20, 30, 40, 50):
    for _ in range(2):
        d, a = F_.emd_forward(x, y, 0.005, it)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d, a = F_.emd_forward(x, y, 0.005, it)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[1]
    una = int((a < 0).sum())
    print(f"N={N} iters={it:3d}: {t:8.3f} ms  (+{(t - prev) / (it - pi):7.3f} ms/iter over {pi}..{it})  unassigned after: {una}", flush=True)
    prev, pi = t, it
