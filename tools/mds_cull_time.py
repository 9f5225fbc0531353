"""Times snb_mds_sample with and without the culling variant (SNB_MDS_CULL=1) on uniform clouds at the two kernel-width regimes of a
training step (development measurement, GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparenet_b200 import functional as F_  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(21)
x = torch.rand(32, 18432, 3, device=dev) * 1.2 - 0.6


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for mml in (0.0227, 0.048):
    mm = torch.full((32,), mml, device=dev)
    os.environ.pop("SNB_MDS_CULL", None)
    ref = F_.mds_sample(x, 16384, mm)
    t0 = timed(lambda: F_.mds_sample(x, 16384, mm))
    os.environ["SNB_MDS_CULL"] = "1"
    got = F_.mds_sample(x, 16384, mm)
    t1 = timed(lambda: F_.mds_sample(x, 16384, mm))
    same = torch.equal(ref, got)
    first = int((ref != got).any(0).nonzero()[0]) if not same else -1
    print(f"mml {mml}: default {t0:.2f} ms, cull {t1:.2f} ms, identical {same}, first differing pick {first}", flush=True)
