"""EMD forward timings, box-pruned Bid vs exhaustive Bid vs the reference extension (development tool):
python tools/emd_time.py  ->  one line per (cloud kind, N)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sparenet_b200 import functional as F_  # noqa: E402

dev = torch.device("cuda:0")
try:
    from conftest import ref_ext
    import refcalls
    ext = ref_ext("emd")
except Exception as e:  # noqa: BLE001
    print("reference extension unavailable:", e)
    ext = None


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


B = 32
for N in (8192, 16384):
    torch.manual_seed(4)
    y = torch.rand(B, N, 3, device=dev) - 0.5
    clouds = {"iid": torch.rand(B, N, 3, device=dev) - 0.5,
              "near": torch.stack([y[b, torch.randperm(N, device=dev)] for b in range(B)]) + 0.01 * torch.randn(B, N, 3, device=dev),
              "blob": 0.05 * torch.randn(B, N, 3, device=dev)}
    for kind, x in clouds.items():
        t_tree = timeit(lambda: F_.emd_forward(x, y, 0.005, 50))
        t_scan = timeit(lambda: F_.emd_forward(x, y, 0.005, 50, exhaustive=True), reps=2)
        t_ref = timeit(lambda: refcalls.emd_fwd(ext, x, y, 0.005, 50), reps=2) if ext is not None else float("nan")
        same = torch.equal(F_.emd_forward(x, y, 0.005, 50)[1], F_.emd_forward(x, y, 0.005, 50, exhaustive=True)[1])
        print(f"emd_fwd {kind:5s} B{B} N{N} iters 50: pruned {t_tree:8.3f} ms  exhaustive {t_scan:8.3f} ms  reference ext {t_ref:8.3f} ms  "
              f"identical={same}", flush=True)
