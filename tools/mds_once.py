"""One MDS launch on the dumped bench inputs through the product library (development tool, e.g. under ncu).
`python tools/mds_once.py [call index] [repeats]`"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparenet_b200 import functional as F_  # noqa: E402

dev = torch.device("cuda:0")
i = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
calls = torch.load(os.path.join(ROOT, "tools", "_data", "mds_inputs.pt"))
x, m, mml = calls[i]
x, mml = x.to(dev).contiguous(), mml.to(dev).contiguous()
ts = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    idx = F_.mds_sample(x, m, mml)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"mds call {i}: n={x.shape[1]} m={m} B={x.shape[0]} layout={os.environ.get('SNB_MDS_LAYOUT', 'default')}: " + " ".join(f"{t:.3f}" for t in ts) + " ms; checksum", int(idx.long().sum()))
