"""MDS time as a function of the number of rounds m at fixed n (development tool): the differences are the cost of the rounds
in each compaction level.  `python tools/mds_levels.py [n] [B]`"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparenet_b200 import functional as F_  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 18432
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.rand(B, n, 3, device=dev) - 0.5
mml = torch.full((B,), 0.01, device=dev)
prev = 0.0
pm = 0
for m in [1024, 2048, 4096, 6144, 8192, 10240, 12288, 14336, 16384]:
    if m > n:
        break
    for _ in range(2):
        F_.mds_sample(x, m, mml)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        F_.mds_sample(x, m, mml)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[1]
    print(f"n={n} m={m:6d}  {t:8.3f} ms   rounds {pm}..{m}: {(t - prev) * 1e6 / (m - pm):7.1f} ns/round", flush=True)
    prev, pm = t, m
