"""One-off GPU diagnostics: where do we differ from the reference extensions? (development tool)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sparenet_b200", "dropin"))
from oracle import build_ref  # noqa: E402
from sparenet_b200 import functional as F_  # noqa: E402
from tests import refcalls  # noqa: E402

dev = torch.device("cuda:0")

# ---------------------------------------------------------------- MDS
MDS = build_ref.load_ref("MDS")
for (n, m, mmlv) in ((64, 32, 0.2), (2048, 512, 0.05), (2048, 512, 0.01), (9216, 64, 0.02)):
    torch.manual_seed(15)
    x = torch.rand(2, n, 3, device=dev)
    mml = torch.full((2,), mmlv, device=dev)
    a = F_.mds_sample(x, m, mml)
    r = refcalls.mds(MDS, x, m, mml)
    same = (a == r).float().mean().item()
    first = (a != r).nonzero()
    print(f"[mds] n={n} m={m} mml={mmlv} same={same:.4f} first diff at {first[0].tolist() if len(first) else None}")
    print("   ours", a[0, :12].tolist())
    print("   ref ", r[0, :12].tolist())

# ---------------------------------------------------------------- EMD: who wins the GetMax race in round 0?
EMD = build_ref.load_ref("emd")
torch.manual_seed(4)
B, n = 4, 8192
x, y = torch.rand(B, n, 3, device=dev), torch.rand(B, n, 3, device=dev)


def z(*s, dt=torch.float32):
    return torch.zeros(*s, device=dev, dtype=dt)


for iters in (1,):
    dist = z(B, n)
    assignment = z(B, n, dt=torch.int32) - 1
    assignment_inv = z(B, n, dt=torch.int32) - 1
    price, bid, bid_inc, max_inc = z(B, n), z(B, n, dt=torch.int32), z(B, n), z(B, n)
    unass_idx, max_idx = z(B * n, dt=torch.int32), z(B * n, dt=torch.int32)
    c1, c2, c3 = z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32)
    EMD.forward(x, y, dist, assignment, price, assignment_inv, bid, bid_inc, max_inc, unass_idx, c1, c2, c3, max_idx, 0.005, iters)
    torch.cuda.synchronize()
    max_idx = max_idx.view(B, n)
    bidl, inc = bid.long(), bid_inc.double()
    nconf = 0
    rules = {"largest": 0, "smallest": 0, "other": 0}
    for b in range(B):
        mx = torch.full((n,), -1e30, dtype=torch.float64, device=dev).scatter_reduce(0, bidl[b], inc[b], reduce="amax")
        qual = (inc[b] - 1e-6 <= mx[bidl[b]]) & (mx[bidl[b]] <= inc[b] + 1e-6)
        cnt = torch.zeros(n, dtype=torch.long, device=dev).scatter_add(0, bidl[b][qual], torch.ones_like(bidl[b][qual]))
        for o in (cnt >= 2).nonzero().flatten().tolist():
            js = ((bidl[b] == o) & qual).nonzero().flatten().tolist()
            w = int(max_idx[b, o])
            nconf += 1
            key = "largest" if w == max(js) else ("smallest" if w == min(js) else "other")
            rules[key] += 1
            if nconf <= 12:
                print(f"[emd] b={b} obj={o} qualifying={js} incs={[float(inc[b, j]) for j in js]} winner={w}")
    print(f"[emd] round-0 conflicts={nconf} rule stats={rules}")

# our result vs reference after k rounds: first round where they diverge
for iters in (1, 2, 3, 5, 10, 50):
    rd, ra = refcalls.emd_fwd(EMD, x, y, 0.005, iters)
    d, a = F_.emd_forward(x, y, 0.005, iters)
    print(f"[emd] iters={iters} identical={(a == ra).float().mean().item():.6f}")

# ---------------------------------------------------------------- p2i
EXT = build_ref.load_ref("ext")
torch.manual_seed(23)
B, n, H, R = 4, 4096, 128, 5.0
pts = (torch.rand(B * n, 2, device=dev) * 1.3 - 0.15) * (H - 1)
feat = torch.rand(B * n, 1, device=dev)
binds = torch.arange(B, dtype=torch.int32, device=dev).repeat_interleave(n)
bg = torch.rand(B, 1, H, H, device=dev) * 0.3
out, ids = F_.p2i_max_forward(pts, feat, binds, bg, 0, R)
rout, rids = EXT.p2i_max_forward_gpu(pts, feat, binds, bg, 0, R)
neq = out != rout
print(f"[p2i] value mismatches={int(neq.sum())} of {out.numel()} max abs diff={float((out - rout).abs().max()):.3e} id mismatches={int((ids != rids).sum())}")
if neq.any():
    i = neq.nonzero()[0].tolist()
    print("   at", i, float(out[tuple(i)]), float(rout[tuple(i)]), int(ids[tuple(i)]), int(rids[tuple(i)]), float(bg[tuple(i)]))
    less = (out < rout).sum().item()
    print("   ours<ref:", less, " ours>ref:", (out > rout).sum().item())

# fp64 gradcheck failure: which entries?
from cuda.p2i_op import p2i  # noqa: E402
torch.manual_seed(24)
for reduce in ("sum", "max"):
    for trial in range(3):
        p = (torch.rand(2, 2, dtype=torch.float64, device=dev) * 1.2 - 0.6).requires_grad_()
        f = torch.rand(2, 2, dtype=torch.float64, device=dev).requires_grad_()
        bi = torch.zeros(2, dtype=torch.int32, device=dev)
        b0 = torch.zeros(1, 2, 8, 8, dtype=torch.float64, device=dev).requires_grad_()
        o = p2i(p, f, bi, b0, 3.0, "cos", reduce)
        go = torch.rand_like(o)
        gb, = torch.autograd.grad(o, b0, go)
        # numeric: perturb bg everywhere by +-eps
        eps = 1e-6
        op = p2i(p.detach(), f.detach(), bi, b0.detach() + eps, 3.0, "cos", reduce)
        om = p2i(p.detach(), f.detach(), bi, b0.detach() - eps, 3.0, "cos", reduce)
        num = (op - om) / (2 * eps)
        ana = gb / go
        bad = (num - ana).abs() > 1e-4
        print(f"[p2i-grad] {reduce} trial {trial}: bad entries {int(bad.sum())}; out==0 count {(o == 0).sum().item()}")
        if bad.any():
            i = tuple(bad.nonzero()[0].tolist())
            print("    e.g.", i, "num", float(num[i]), "ana", float(ana[i]), "out", float(o[i]))
