"""GPU diagnostics #3: edge_reduce backward mismatch at (B=2,C=8,N=4096,k=16)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparenet_b200 import fused
from tests import fused_ref as R
dev = torch.device("cuda:0")
for (B, C, N, k) in ((2, 8, 4096, 16), (2, 8, 4096, 8), (2, 8, 2048, 16), (1, 1, 4096, 16), (1, 2, 3000, 9)):
    torch.manual_seed(B * 1000 + N)
    a = torch.randn(B, C, N, device=dev); c = torch.randn(B, C, N, device=dev)
    idx = torch.stack([torch.stack([torch.randperm(N, device=dev)[:k] for _ in range(N)]) for _ in range(B)]).int()
    for which in ("max", "min", "s1", "s2"):
        a1, c1 = a.clone().requires_grad_(), c.clone().requires_grad_()
        a2, c2 = a.double().requires_grad_(), c.double().requires_grad_()
        o1 = fused.edge_reduce(a1, c1, idx); o2 = R.edge_reduce(a2, c2, idx)
        j = {"max": 0, "min": 1, "s1": 2, "s2": 3}[which]
        w = torch.randn_like(o1[j])
        (o1[j] * w).sum().backward(); (o2[j] * w.double()).sum().backward()
        da = (a1.grad.double() - a2.grad).abs()
        nbad = int((da > 1e-4 * a2.grad.abs().max()).sum())
        print(f"[diag3] B={B} C={C} N={N} k={k} term={which}: bad={nbad} maxerr={da.max().item():.3e} scale={a2.grad.abs().max().item():.3e}")
        if nbad and which in ("max", "min"):
            loc = (da > 1e-4 * a2.grad.abs().max()).nonzero()[:3]
            for l in loc.tolist():
                print("    at", l, "ours", a1.grad[tuple(l)].item(), "ref", a2.grad[tuple(l)].item())
