"""Per-op timing on the GPU box: ours (C ABI) next to the reference's own extensions (oracle/_ref/*.so).
CUDA events, warm-up 3, median of `reps`.  Writes gpurun_out/opbench.json.  Development tool, not bench.py."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402
from sparenet_b200 import functional as F_  # noqa: E402
from tests import refcalls  # noqa: E402

dev = torch.device("cuda:0")
only = set(sys.argv[1:])


def timeit(fn, reps=7, warm=3):
    for _ in range(warm):
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
    ts.sort()
    return ts[len(ts) // 2]


def ext(name):
    try:
        return build_ref.load_ref(name) if build_ref.available(name) else None
    except Exception as e:
        print("ref ext", name, "unavailable:", e)
        return None


res = {}


def rec(name, ours, ref=None, **kw):
    res[name] = dict(ours_ms=ours, ref_ms=ref, speedup=(ref / ours if ref else None), **kw)
    print(f"{name:34s} ours {ours:9.3f} ms   ref {ref if ref is None else round(ref, 3)} ms   x{(ref / ours) if ref else float('nan'):.2f}", flush=True)


def want(k):
    return not only or k in only


B = 32
if want("chamfer"):
    torch.manual_seed(2)
    x, y = torch.rand(B, 16384, 3, device=dev) - 0.5, torch.rand(B, 16384, 3, device=dev) - 0.5
    e = ext("chamfer")
    d1, d2, i1, i2 = F_.chamfer_forward(x, y)
    g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
    rec("chamfer_fwd B32 N16384", timeit(lambda: F_.chamfer_forward(x, y)), timeit(lambda: refcalls.chamfer_fwd(e, x, y)) if e else None)
    rec("chamfer_bwd B32 N16384", timeit(lambda: F_.chamfer_backward(x, y, i1, i2, g1, g2)),
        timeit(lambda: refcalls.chamfer_bwd(e, x, y, i1, i2, g1, g2)) if e else None)
if want("emd"):
    e = ext("emd")
    for N in (8192, 16384):
        torch.manual_seed(4)
        x, y = torch.rand(B, N, 3, device=dev), torch.rand(B, N, 3, device=dev)
        rec(f"emd_fwd iid B32 N{N}", timeit(lambda: F_.emd_forward(x, y, 0.005, 50), reps=3, warm=1),
            timeit(lambda: refcalls.emd_fwd(e, x, y, 0.005, 50), reps=3, warm=1) if e else None)
        torch.manual_seed(6)
        xn = torch.stack([y[b, torch.randperm(N, device=dev)] for b in range(B)]) + 0.02 * torch.randn(B, N, 3, device=dev)
        rec(f"emd_fwd near B32 N{N}", timeit(lambda: F_.emd_forward(xn, y, 0.005, 50), reps=3, warm=1),
            timeit(lambda: refcalls.emd_fwd(e, xn, y, 0.005, 50), reps=3, warm=1) if e else None)
if want("expansion"):
    e = ext("expansion_penalty")
    for N, p in ((16384, 512), (8192, 256)):
        torch.manual_seed(3)
        x = torch.rand(B, N, 3, device=dev)
        rec(f"expansion_fwd B32 N{N} p{p}", timeit(lambda: F_.expansion_forward(x, p, 1.5)),
            timeit(lambda: refcalls.expansion_fwd(e, x, p, 1.5), reps=3, warm=1) if e else None)
if want("mds"):
    e = ext("MDS")
    for n, m in ((18432, 16384), (10240, 8192)):
        torch.manual_seed(5)
        x = torch.rand(B, n, 3, device=dev)
        mml = torch.full((B,), 0.6 / n ** (1 / 3), device=dev)
        rec(f"mds B32 n{n} m{m}", timeit(lambda: F_.mds_sample(x, m, mml), reps=3, warm=1),
            timeit(lambda: refcalls.mds(e, x, m, mml), reps=3, warm=1) if e else None)
if want("p2i"):
    e = ext("ext")
    torch.manual_seed(3)
    n = 16384
    binds = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(n)
    bg = torch.zeros(B, 1, 256, 256, device=dev)
    for kind in ("uniform", "shell"):
        if kind == "uniform":
            pts = torch.rand(B * n, 2, device=dev) * 255
        else:
            v = torch.randn(B * n, 3, device=dev)
            v = v / v.norm(dim=1, keepdim=True) * 0.5
            pts = (v[:, :2] * 0.75 + 1) / 2 * 255
        feat = torch.rand(B * n, 1, device=dev)
        for R in (5.0, 10.0):
            out, ids = F_.p2i_max_forward(pts, feat, binds, bg, 0, R)
            go = torch.rand_like(out)
            rec(f"p2i_max_fwd {kind} R{R}", timeit(lambda: F_.p2i_max_forward(pts, feat, binds, bg, 0, R)),
                timeit(lambda: e.p2i_max_forward_gpu(pts, feat, binds, bg, 0, R), reps=3, warm=1) if e else None)
            rec(f"p2i_max_bwd {kind} R{R}", timeit(lambda: F_.p2i_max_backward(go, ids, pts, feat, 0, R)),
                timeit(lambda: e.p2i_max_backward_gpu(go, ids, pts, feat, 0, R), reps=3, warm=1) if e else None)
if want("knn"):
    for C in (3, 256, 512):
        torch.manual_seed(7)
        x = torch.rand(B, C, 2048, device=dev)

        def torch_knn():
            inner = -2 * torch.matmul(x.transpose(2, 1), x)
            xx = torch.sum(x ** 2, dim=1, keepdim=True)
            return (-xx - inner - xx.transpose(2, 1)).topk(k=8, dim=-1)[1]
        rec(f"knn B32 C{C} N2048 k8 (ref=torch fallback :872-875)", timeit(lambda: F_.knn_indices(x, 8)), timeit(torch_knn))

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "opbench.json"), "w") as f:
    json.dump(res, f, indent=1)
