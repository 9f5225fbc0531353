"""GPU parity tests (run with -m gpu on the B200 box).  Every op is called through the C ABI (via the drop-in
wrappers / sparenet_b200.functional) and compared with (1) the CPU oracle on seeded inputs, (2) the
reference's own extension rebuilt for sm_100a when oracle/_ref/ holds it, (3) size-independent properties at
BASELINE.json's full sizes.  Bar: bit-exact for indices, and for values wherever the arithmetic is
replicated; tolerances are written next to each comparison otherwise."""
import sys

import os

import pytest
import torch

import oracle
from tests.conftest import ref_ext
from tests import refcalls

pytestmark = pytest.mark.gpu


def _dropin():
    import sparenet_b200
    p = sparenet_b200.dropin_path()
    if p not in sys.path:
        sys.path.insert(0, p)


# ================================================================ Chamfer
@pytest.mark.parametrize("B,N,M,seed", [(2, 1024, 1024, 0), (3, 300, 517, 1), (1, 1, 5, 2), (2, 4099, 2048, 3), (1, 7, 7, 4)])
def test_chamfer_vs_oracle(cuda, B, N, M, seed):
    from sparenet_b200 import functional as F_
    torch.manual_seed(seed)
    x, y = torch.rand(B, N, 3), torch.rand(B, M, 3)
    d1, d2, i1, i2 = F_.chamfer_forward(x.to(cuda), y.to(cuda))
    od1, od2, oi1, oi2 = oracle.chamfer_fwd(x, y)
    assert torch.equal(i1.cpu(), oi1) and torch.equal(i2.cpu(), oi2)       # bit-exact indices
    assert torch.equal(d1.cpu(), od1) and torch.equal(d2.cpu(), od2)       # bit-exact distances
    g1, g2 = torch.rand(B, N), torch.rand(B, M)
    gx, gy = F_.chamfer_backward(x.to(cuda), y.to(cuda), i1, i2, g1.to(cuda), g2.to(cuda))
    ogx, ogy = oracle.chamfer_bwd(x, y, oi1, oi2, g1, g2)
    # float atomics: summation order differs -> tolerance (<= 1e-5 rel as BASELINE.json asks)
    assert torch.allclose(gx.cpu(), ogx, rtol=1e-5, atol=1e-6) and torch.allclose(gy.cpu(), ogy, rtol=1e-5, atol=1e-6)


def test_chamfer_ties_and_zero_padding(cuda):
    from sparenet_b200 import functional as F_
    torch.manual_seed(5)
    y = torch.rand(2, 256, 3)
    y = torch.cat([y, y], 1)                 # every reference point twice -> exact ties, lowest index must win
    x = torch.rand(2, 777, 3)
    x[:, 600:] = 0                           # zero-padded tail like the real loader (data_transforms.py:170-173)
    d1, d2, i1, i2 = F_.chamfer_forward(x.to(cuda), y.to(cuda))
    od1, od2, oi1, oi2 = oracle.chamfer_fwd(x, y)
    assert torch.equal(i1.cpu(), oi1) and torch.equal(i2.cpu(), oi2)
    assert (i1 < 256).all()
    assert torch.equal(d1.cpu(), od1) and torch.equal(d2.cpu(), od2)


def test_chamfer_vs_reference_extension(cuda):
    ext = ref_ext("chamfer")
    if ext is None:
        pytest.skip("oracle/_ref/chamfer.so not present")
    from sparenet_b200 import functional as F_
    torch.manual_seed(11)
    x, y = torch.rand(4, 4096, 3, device=cuda), torch.rand(4, 2048, 3, device=cuda)
    d1, d2, i1, i2 = F_.chamfer_forward(x, y)
    r1, r2, j1, j2 = refcalls.chamfer_fwd(ext, x, y)
    assert torch.equal(i1, j1) and torch.equal(i2, j2) and torch.equal(d1, r1) and torch.equal(d2, r2)
    g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
    gx, gy = F_.chamfer_backward(x, y, i1, i2, g1, g2)
    rx, ry = refcalls.chamfer_bwd(ext, x, y, j1, j2, g1, g2)
    assert torch.allclose(gx, rx, rtol=1e-5, atol=1e-6) and torch.allclose(gy, ry, rtol=1e-5, atol=1e-6)


def test_chamfer_full_size_properties(cuda):
    """B=32, N=M=16384 (BASELINE config): spot-check 64 queries per sample against fp64 brute force."""
    from sparenet_b200 import functional as F_
    torch.manual_seed(2)
    x = torch.rand(32, 16384, 3, device=cuda) - 0.5
    y = torch.rand(32, 16384, 3, device=cuda) - 0.5
    d1, d2, i1, i2 = F_.chamfer_forward(x, y)
    assert i1.min() >= 0 and i1.max() < 16384 and i2.min() >= 0 and i2.max() < 16384
    # dist is the distance to the reported index
    yy = torch.gather(y, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3))
    assert torch.allclose(d1, (yy - x).pow(2).sum(-1), rtol=1e-6, atol=0)
    q = torch.randint(0, 16384, (64,), device=cuda)
    D = torch.cdist(x[:, q].double(), y.double()).pow(2)
    m, am = D.min(-1)
    assert torch.allclose(d1[:, q].double(), m, rtol=1e-6, atol=1e-12)
    # symmetric call swaps the outputs exactly
    e2, e1, k2, k1 = F_.chamfer_forward(y, x)
    assert torch.equal(e1, d1) and torch.equal(k1, i1) and torch.equal(e2, d2) and torch.equal(k2, i2)


@pytest.mark.parametrize("kind", ["uniform", "blob_vs_uniform", "duplicates", "two_clusters", "ragged", "line"])
def test_chamfer_pruned_search_equals_brute_force_bit_for_bit(cuda, kind):
    """The spatially pruned search (Morton-sorted clusters + box lower bounds) must return the very same bits as the
    brute-force kernel on any distribution, including exact ties (lowest index wins) and degenerate boxes."""
    from sparenet_b200 import functional as F_
    torch.manual_seed(hash(kind) % 1000)
    B, N, M = 3, 4096, 4096
    x, y = torch.rand(B, N, 3, device=cuda), torch.rand(B, M, 3, device=cuda)
    if kind == "blob_vs_uniform":                 # the bench's situation: predictions in a tiny blob, targets spread out
        x = 0.5 + 0.01 * torch.randn(B, N, 3, device=cuda)
    elif kind == "duplicates":
        y[:, M // 2:] = y[:, :M // 2]
        x[:, -500:] = 0
        y[:, -100:] = 0
    elif kind == "two_clusters":
        x[:, : N // 2] = 0.02 * torch.rand(B, N // 2, 3, device=cuda)
        y[:, M // 3:] = 0.97 + 0.03 * torch.rand(B, M - M // 3, 3, device=cuda)
    elif kind == "ragged":
        x, y = torch.rand(B, 1000, 3, device=cuda), torch.rand(B, 5003, 3, device=cuda)
    elif kind == "line":                          # degenerate extent on two axes
        x[..., 1:] = 0.25
        y[..., 1:] = 0.25
    outs = []
    for pruned in (True, False):
        F_.CHAMFER_PRUNED = pruned
        F_._CHAMFER_MEMO.clear()
        outs.append(F_.chamfer_forward(x, y))
    F_.CHAMFER_PRUNED = True
    F_._CHAMFER_MEMO.clear()
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_chamfer_dropin_modules_autograd(cuda):
    _dropin()
    from cuda.chamfer_dist import ChamferDistance as CDa, ChamferFunction
    from cuda.chamfer_distance import ChamferDistance as CDb, ChamferDistanceMean
    from sparenet_b200 import functional as F_
    # the de-duplication of identical searches is opt-in and scoped: identical values, one launch pair less, invalidated by edits
    xr, yr = torch.rand(2, 512, 3, device=cuda), torch.rand(2, 600, 3, device=cuda)
    n0 = F_.LAUNCHES["count"]
    a = F_.chamfer_forward(xr, yr)
    k1 = F_.LAUNCHES["count"] - n0
    b = F_.chamfer_forward(xr, yr)
    assert k1 > 0 and F_.LAUNCHES["count"] - n0 == 2 * k1 and a[0].data_ptr() != b[0].data_ptr()      # no scope: two searches
    with F_.chamfer_reuse():
        n1 = F_.LAUNCHES["count"]
        c = F_.chamfer_forward(xr, yr)
        per = F_.LAUNCHES["count"] - n1
        d = F_.chamfer_forward(xr, yr)
        assert F_.LAUNCHES["count"] - n1 == per and d[0].data_ptr() == c[0].data_ptr() and all(torch.equal(p, q) for p, q in zip(c, d))
        c[0].mul_(2.0)                                                        # an in-place edit of a returned tensor drops the entry
        e = F_.chamfer_forward(xr, yr)
        assert F_.LAUNCHES["count"] - n1 == 2 * per and torch.equal(e[0], a[0])
    assert not F_._CHAMFER_MEMO
    torch.manual_seed(0)
    x = torch.rand(2, 512, 3, device=cuda, requires_grad=True)
    y = torch.rand(2, 640, 3, device=cuda, requires_grad=True)
    loss = ChamferDistanceMean()(x, y)
    loss.backward()
    xd, yd = x.detach().double().requires_grad_(), y.detach().double().requires_grad_()
    Dd = (xd[:, :, None] - yd[:, None]).pow(2).sum(-1)
    ref = Dd.min(2)[0].mean() + Dd.min(1)[0].mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-6 * ref.item() + 1e-9
    assert torch.allclose(x.grad.double(), xd.grad, rtol=1e-4, atol=1e-8) and torch.allclose(y.grad.double(), yd.grad, rtol=1e-4, atol=1e-8)
    assert abs(CDa()(x, y).item() - loss.item()) < 1e-7
    d1, d2 = CDb()(x, y)
    e1, e2 = ChamferFunction.apply(x, y)
    assert torch.equal(d1, e1) and torch.equal(d2, e2)
    # ignore_zeros only acts for batch size 1 (chamfer_dist/__init__.py:28-32)
    xz = torch.rand(1, 64, 3, device=cuda)
    xz[:, 50:] = 0
    yz = torch.rand(1, 80, 3, device=cuda)
    assert abs(CDa(ignore_zeros=True)(xz, yz).item() - CDa()(xz[:, :50], yz).item()) < 1e-7


# ================================================================ EMD
@pytest.mark.parametrize("B,N,eps,iters,seed,kind", [(2, 1024, 0.005, 50, 4, "iid"), (3, 2048, 0.005, 50, 6, "near"),
                                                     (1, 3072, 0.002, 120, 7, "iid"), (2, 1024, 0.005, 1, 8, "iid"),
                                                     (1, 2048, 0.005, 50, 9, "dups")])
@pytest.mark.parametrize("exhaustive", [False, True])
def test_emd_vs_oracle_bit_exact(cuda, B, N, eps, iters, seed, kind, exhaustive):
    from sparenet_b200 import functional as F_
    torch.manual_seed(seed)
    y = torch.rand(B, N, 3)
    if kind == "near":
        x = torch.stack([y[b, torch.randperm(N)] for b in range(B)]) + 0.02 * torch.randn(B, N, 3)
    elif kind == "dups":                      # exact ties: duplicated targets and a zero-padded prediction tail
        x = torch.rand(B, N, 3)
        y[:, N // 2:] = y[:, :N // 2]
        x[:, -256:] = 0
    else:
        x = torch.rand(B, N, 3)
    dist, ass = F_.emd_forward(x.to(cuda), y.to(cuda), eps, iters, exhaustive=exhaustive)
    odist, oass = oracle.emd_fwd(x, y, eps, iters)
    assert torch.equal(ass.cpu(), oass)
    assert torch.equal(dist.cpu(), odist)
    g = torch.rand(B, N)
    gx = F_.emd_backward(x.to(cuda), y.to(cuda), g.to(cuda), ass)
    assert torch.equal(gx.cpu(), oracle.emd_bwd(x, y, g, oass))


def test_emd_vs_reference_extension(cuda):
    """The reference is racy (GetMax last-writer-wins, emd_cuda.cu:188-191): compare reference-vs-reference first and
    hold our result to the same band."""
    ext = ref_ext("emd")
    if ext is None:
        pytest.skip("oracle/_ref/emd.so not present")
    from sparenet_b200 import functional as F_
    torch.manual_seed(4)
    x, y = torch.rand(4, 8192, 3, device=cuda), torch.rand(4, 8192, 3, device=cuda)
    rd1, ra1 = refcalls.emd_fwd(ext, x, y, 0.005, 50)
    rd2, ra2 = refcalls.emd_fwd(ext, x, y, 0.005, 50)
    torch.cuda.synchronize()
    dist, ass = F_.emd_forward(x, y, 0.005, 50)
    ref_self = (ra1 == ra2).float().mean().item()
    ours = (ass == ra1).float().mean().item()
    e_ref1, e_ref2, e_ours = rd1.sqrt().mean().item(), rd2.sqrt().mean().item(), dist.sqrt().mean().item()
    print(f"[emd] ref-vs-ref identical={ref_self:.6f} ours-vs-ref identical={ours:.6f} emd ref={e_ref1:.7f}/{e_ref2:.7f} ours={e_ours:.7f}")
    # tests/perf/diag2.py (crafted duplicate bidders) shows the reference's GetMax winner is scheduling dependent: lower lane
    # inside a warp, otherwise whichever warp/block stores last -- no index rule reproduces it.  Our rule (largest index)
    # therefore departs from it on a few bids per round; the loss value agrees to ~1e-5 and >99% of matches coincide.
    band = max(abs(e_ref1 - e_ref2), 1e-5 * e_ref1)
    assert abs(e_ours - e_ref1) <= 10 * band + 5e-5 * e_ref1
    assert ours >= 0.99
    # rounds before the first in-window conflict are bit-identical to the reference
    for it in (1, 2):
        rd, ra = refcalls.emd_fwd(ext, x, y, 0.005, it)
        d, a = F_.emd_forward(x, y, 0.005, it)
        assert torch.equal(a, ra) and torch.equal(d, rd)


def test_emd_full_size_properties(cuda):
    from sparenet_b200 import functional as F_
    torch.manual_seed(5)
    x, y = torch.rand(32, 8192, 3, device=cuda), torch.rand(32, 8192, 3, device=cuda)
    dist, ass = F_.emd_forward(x, y, 0.005, 50)
    assert ass.min() >= 0 and ass.max() < 8192
    yy = torch.gather(y, 1, ass.long().unsqueeze(-1).expand(-1, -1, 3))
    assert torch.allclose(dist, (x - yy).pow(2).sum(-1), rtol=1e-6, atol=0)
    # the auction spreads the matches: most targets are used exactly once
    uniq = torch.tensor([ass[b].unique().numel() for b in range(32)], dtype=torch.float32).mean().item()
    assert uniq > 0.9 * 8192
    # determinism: same input, same bits
    d2, a2 = F_.emd_forward(x, y, 0.005, 50)
    assert torch.equal(a2, ass) and torch.equal(d2, dist)
    # identical clouds -> identity assignment, zero distance
    d0, a0 = F_.emd_forward(x[:2], x[:2].clone(), 0.005, 5)
    assert (a0 == torch.arange(8192, device=cuda, dtype=torch.int32)).all() and (d0 == 0).all()


@pytest.mark.parametrize("kind", ["iid", "blob_vs_shell", "near", "planes", "dups"])
@pytest.mark.parametrize("N,iters", [(16384, 50), (8192, 50), (2048, 300)])
def test_emd_pruned_bid_identical_to_exhaustive(cuda, kind, N, iters):
    """The box-pruned Bid (snb_emd_fwd) against the exhaustive one (snb_emd_fwd_scan, the kernel the oracle pins bit for bit) at
    sizes the oracle cannot reach: assignments and distances must be the same bits, also when prices pile up (a collapsed
    prediction against a spread target), on coplanar clouds (degenerate boxes) and with exact ties."""
    from sparenet_b200 import functional as F_
    torch.manual_seed(N + iters)
    B = 8
    y = torch.rand(B, N, 3, device=cuda) - 0.5
    if kind == "iid":
        x = torch.rand(B, N, 3, device=cuda) - 0.5
    elif kind == "blob_vs_shell":
        y = torch.nn.functional.normalize(torch.randn(B, N, 3, device=cuda), dim=-1) * 0.5
        x = 0.02 * torch.randn(B, N, 3, device=cuda)
    elif kind == "near":
        x = torch.stack([y[b, torch.randperm(N, device=cuda)] for b in range(B)]) + 0.01 * torch.randn(B, N, 3, device=cuda)
    elif kind == "planes":
        x = torch.rand(B, N, 3, device=cuda) - 0.5
        x[..., 2] = 0.25
        y[..., 0] = -0.1
    else:
        x = torch.rand(B, N, 3, device=cuda) - 0.5
        y[:, N // 2:] = y[:, :N // 2]
        x[:, -512:] = 0
    d_t, a_t = F_.emd_forward(x, y, 0.005, iters)
    d_s, a_s = F_.emd_forward(x, y, 0.005, iters, exhaustive=True)
    same = (a_t == a_s).float().mean().item()
    assert torch.equal(a_t, a_s), f"{kind} N={N}: {same * 100:.4f} % of the assignments agree"
    assert torch.equal(d_t, d_s)


def test_emd_limits_and_dropin(cuda):
    _dropin()
    from cuda.emd.emd_module import emdModule
    from sparenet_b200 import functional as F_
    from sparenet_b200._lib import SnbError
    with pytest.raises(SnbError):
        F_.emd_forward(torch.rand(1, 1000, 3, device=cuda), torch.rand(1, 1000, 3, device=cuda), 0.005, 10)
    with pytest.raises(AssertionError):
        emdModule()(torch.rand(1, 1000, 3, device=cuda), torch.rand(1, 1000, 3, device=cuda), 0.005, 10)
    x = torch.rand(2, 1024, 3, device=cuda, requires_grad=True)
    y = torch.rand(2, 1024, 3, device=cuda, requires_grad=True)
    dist, ass = emdModule()(x, y, 0.005, 50)
    torch.sqrt(dist).mean(1).mean().backward()
    yy = torch.gather(y.detach(), 1, ass.long().unsqueeze(-1).expand(-1, -1, 3))
    diff = x.detach() - yy
    ref = diff / diff.norm(dim=-1, keepdim=True) / (2 * 1024)
    assert torch.allclose(x.grad, ref, rtol=1e-4, atol=1e-8)
    assert (y.grad == 0).all()             # xyz2 receives no gradient (emd_module.py:84-87)


# ================================================================ expansion penalty
@pytest.mark.parametrize("B,N,p,seed", [(2, 1024, 256, 1), (2, 2048, 512, 2), (3, 512, 64, 3), (1, 256, 32, 4), (2, 64, 8, 5), (1, 4, 2, 6)])
def test_expansion_vs_oracle(cuda, B, N, p, seed):
    from sparenet_b200 import functional as F_
    torch.manual_seed(seed)
    x = torch.rand(B, N, 3)
    dist, ass, mml = F_.expansion_forward(x.to(cuda), p, 1.5)
    od, oa, om = oracle.expansion_fwd(x, p, 1.5)
    assert torch.equal(ass.cpu(), oa) and torch.equal(dist.cpu(), od) and torch.equal(mml.cpu(), om)
    g = torch.rand(B, N)
    gx = F_.expansion_backward(x.to(cuda), g.to(cuda), ass)
    assert torch.equal(gx.cpu(), oracle.expansion_bwd(x, g, oa))


def test_expansion_duplicate_points(cuda):
    from sparenet_b200 import functional as F_
    torch.manual_seed(8)
    x = torch.rand(2, 512, 3)
    x[:, 100:140] = x[:, 60:100]       # zero-length edges and exact ties in the Prim arg-min
    x[:, 300:] = 0.25
    dist, ass, mml = F_.expansion_forward(x.to(cuda), 256, 1.5)
    od, oa, om = oracle.expansion_fwd(x, 256, 1.5)
    assert torch.equal(ass.cpu(), oa) and torch.equal(dist.cpu(), od) and torch.equal(mml.cpu(), om)


def test_expansion_vs_reference_extension(cuda):
    ext = ref_ext("expansion_penalty")
    if ext is None:
        pytest.skip("oracle/_ref/expansion_penalty.so not present")
    from sparenet_b200 import functional as F_
    torch.manual_seed(3)
    x = torch.rand(8, 8192, 3, device=cuda)
    dist, ass, mml = F_.expansion_forward(x, 256, 1.5)
    rd, ra, rm = refcalls.expansion_fwd(ext, x, 256, 1.5)
    same = (ass == ra).float().mean().item()
    print(f"[expansion] per-index identical to reference: {same:.6f}")
    assert torch.allclose(mml, rm, rtol=1e-6, atol=0)                       # atomicAdd order only
    assert abs(dist.sum().item() - rd.sum().item()) <= 1e-5 * rd.sum().item()   # sum of tagged edges is race-invariant
    assert same > 0.995                                                      # leaf-attribution race (:128-139)
    g = torch.rand_like(dist)
    assert torch.allclose(F_.expansion_backward(x, g, ra), refcalls.expansion_bwd(ext, x, g, ra), rtol=1e-6, atol=1e-7)


def test_expansion_limits(cuda):
    from sparenet_b200 import functional as F_
    from sparenet_b200._lib import SnbError
    x = torch.rand(1, 768, 3, device=cuda)
    for p in (384, 1024, 1):
        with pytest.raises(SnbError):
            F_.expansion_forward(x, p, 1.5)


# ================================================================ MDS + gather
@pytest.mark.parametrize("B,n,m,seed", [(2, 640, 320, 9), (3, 2304, 2048, 10), (1, 100, 60, 11), (2, 9216, 1024, 12), (33, 1100, 64, 13),
                                        (150, 700, 64, 17), (5, 4099, 1500, 18)])
def test_mds_vs_oracle(cuda, B, n, m, seed):
    """Host expf is not bit-identical to CUDA's, so the sequence is replayed step by step: every choice must be an
    arg-min of the oracle's own densities within 1e-5 rel; in practice the sequences coincide."""
    from sparenet_b200 import functional as F_
    torch.manual_seed(seed)
    x = torch.rand(B, n, 3)
    mml = 0.6 / (n ** (1 / 3)) * (0.8 + 0.4 * torch.rand(B))
    idx = F_.mds_sample(x.to(cuda), m, mml.to(cuda)).cpu()
    assert (idx[:, 0] == 0).all() and idx.min() >= 0 and idx.max() < n
    tot_mism = 0
    for b in range(min(B, 3)):
        bad, mism = oracle.mds_check(x[b], mml[b], idx[b], 1e-5)
        assert bad == 0
        tot_mism += mism
        assert idx[b].unique().numel() == m
    assert tot_mism <= 0.01 * m * min(B, 3)


def test_mds_ties_zero_padding_and_oversampling(cuda):
    from sparenet_b200 import functional as F_
    torch.manual_seed(14)
    x = torch.rand(2, 9000, 3)
    x[:, 8500:] = 0                       # identical padded points: exact density ties, key = (k % 1024, k)
    mml = torch.tensor([0.03, 0.05])
    idx = F_.mds_sample(x.to(cuda), 3000, mml.to(cuda)).cpu()
    for b in range(2):
        bad, mism = oracle.mds_check(x[b], mml[b], idx[b], 1e-5)
        assert bad == 0 and mism <= 30
    # m > n: once everything is parked index 0 repeats (MDS_cuda.cu:121-133)
    idx2 = F_.mds_sample(x[:, :64].contiguous().to(cuda), 80, mml.to(cuda)).cpu()
    assert (idx2[:, 64:] == 0).all() and idx2[0, :64].unique().numel() == 64


def test_mds_vs_reference_extension(cuda):
    ext = ref_ext("MDS")
    if ext is None:
        pytest.skip("oracle/_ref/MDS.so not present")
    from sparenet_b200 import functional as F_
    torch.manual_seed(15)
    x = torch.rand(4, 9216, 3, device=cuda)
    mml = torch.tensor([0.02, 0.025, 0.03, 0.035], device=cuda)
    idx = F_.mds_sample(x, 4096, mml)
    ridx = refcalls.mds(ext, x, 4096, mml)
    same = (idx == ridx).float().mean().item()
    print(f"[mds] identical to reference: {same:.6f}")
    assert same == 1.0                      # same expf, same tie key -> bit-exact (reference race aside)
    f = torch.rand(4, 4, 9216, device=cuda)
    assert torch.equal(F_.gather_forward(f, idx), ext.gather_points(f, ridx))
    g = torch.rand(4, 4, 4096, device=cuda)
    assert torch.equal(F_.gather_backward(g, idx, 9216), ext.gather_points_grad(g, ridx, 9216))


@pytest.mark.parametrize("mml", [0.0227, 0.048])
def test_mds_full_size_vs_reference_extension(cuda, mml):
    """BASELINE size (B=32, n=18432 = 16384 generated + 2048 partial points, m=16384) in the two regimes the bench step meets:
    a neighbourhood-sized kernel width and one where every pick moves every density.  Bit-exact sequence."""
    ext = ref_ext("MDS")
    if ext is None:
        pytest.skip("oracle/_ref/MDS.so not present")
    from sparenet_b200 import functional as F_
    torch.manual_seed(19)
    x = torch.rand(32, 18432, 3, device=cuda) * 1.2 - 0.6
    mm = torch.full((32,), mml, device=cuda) * (0.9 + 0.2 * torch.rand(32, device=cuda))
    idx = F_.mds_sample(x, 16384, mm)
    ridx = refcalls.mds(ext, x, 16384, mm)
    assert torch.equal(idx, ridx)
    assert all(idx[b].unique().numel() == 16384 for b in (0, 31))


def test_gather_vs_oracle_and_autograd(cuda):
    _dropin()
    from cuda.MDS.MDS_module import gather_operation, minimum_density_sample
    torch.manual_seed(16)
    f = torch.rand(3, 4, 500)
    idx = torch.stack([torch.randperm(500)[:200] for _ in range(3)]).int()
    fc = f.to(cuda).requires_grad_()
    out = gather_operation(fc, idx.to(cuda))
    assert torch.equal(out.detach().cpu(), oracle.gather_fwd(f, idx))
    g = torch.rand(3, 4, 200)
    out.backward(g.to(cuda))
    assert torch.equal(fc.grad.cpu(), oracle.gather_bwd(g, idx, 500))
    s = minimum_density_sample(torch.rand(2, 300, 3, device=cuda), 100, torch.tensor([0.1, 0.1], device=cuda))
    assert s.dtype == torch.int32 and s.shape == (2, 100) and not s.requires_grad


# ================================================================ p2i
def _p2i_inputs(B, n, C, H, dtype, seed, radius):
    torch.manual_seed(seed)
    pts = (torch.rand(B * n, 2, dtype=dtype) * 1.3 - 0.15) * (H - 1)     # some footprints cross the border
    feat = torch.rand(B * n, C, dtype=dtype)
    binds = torch.arange(B, dtype=torch.int32).repeat_interleave(n)
    binds[::37] = B + 3                                                    # out-of-range batch ids are skipped
    bg = torch.rand(B, C, H, H, dtype=dtype) * 0.3
    return pts, feat, binds, bg


@pytest.mark.parametrize("dtype,C,radius", [(torch.float32, 1, 5.0), (torch.float32, 3, 2.5), (torch.float64, 2, 3.3), (torch.float32, 1, 10.0)])
def test_p2i_max_vs_oracle(cuda, dtype, C, radius):
    from sparenet_b200 import functional as F_
    B, n, H = 2, 400, 48
    pts, feat, binds, bg = _p2i_inputs(B, n, C, H, dtype, 21, radius)
    out, ids = F_.p2i_max_forward(pts.to(cuda), feat.to(cuda), binds.to(cuda), bg.to(cuda), 0, radius)
    oo, oi = oracle.p2i_max_fwd(pts, feat, binds, bg, radius)
    # CUDA's fp64 cos and glibc's can differ by 1 ulp(double); after rounding to T that is <= 1 ulp(T), rarely
    tol = 2e-7 if dtype == torch.float32 else 1e-14
    assert torch.allclose(out.cpu(), oo, rtol=tol, atol=0)
    neq = (ids.cpu() != oi)
    assert neq.float().mean().item() < 2e-3          # only 1-ulp near-ties between two candidates may flip the winner
    gout = torch.rand_like(bg)
    gp, gf, gb = F_.p2i_max_backward(gout.to(cuda), ids, pts.to(cuda), feat.to(cuda), 0, radius)
    ogp, ogf, ogb = oracle.p2i_max_bwd(gout, ids.cpu(), pts, feat, radius)
    rt = 1e-5 if dtype == torch.float32 else 1e-12
    assert torch.equal(gb.cpu(), ogb)
    assert torch.allclose(gf.cpu(), ogf, rtol=rt, atol=rt) and torch.allclose(gp.cpu(), ogp, rtol=rt, atol=rt * 10)


@pytest.mark.parametrize("dtype,C,radius", [(torch.float32, 1, 5.0), (torch.float64, 2, 3.3)])
def test_p2i_sum_vs_oracle(cuda, dtype, C, radius):
    from sparenet_b200 import functional as F_
    B, n, H = 2, 300, 40
    pts, feat, binds, bg = _p2i_inputs(B, n, C, H, dtype, 22, radius)
    out = F_.p2i_sum_forward(pts.to(cuda), feat.to(cuda), binds.to(cuda), bg.to(cuda), 0, radius)
    oo = oracle.p2i_sum_fwd(pts, feat, binds, bg, radius)
    rt = 1e-5 if dtype == torch.float32 else 1e-12
    assert torch.allclose(out.cpu(), oo, rtol=rt, atol=rt)
    gout = torch.rand_like(bg)
    gp, gf = F_.p2i_sum_backward(gout.to(cuda), pts.to(cuda), feat.to(cuda), binds.to(cuda), 0, radius)
    ogp, ogf = oracle.p2i_sum_bwd(gout, pts, feat, binds, radius)
    assert torch.allclose(gf.cpu(), ogf, rtol=rt, atol=rt) and torch.allclose(gp.cpu(), ogp, rtol=rt * 10, atol=rt * 10)


def test_p2i_vs_reference_extension(cuda):
    ext = ref_ext("ext")
    if ext is None:
        pytest.skip("oracle/_ref/ext.so not present")
    from sparenet_b200 import functional as F_
    B, n, H, R = 4, 4096, 128, 5.0
    pts, feat, binds, bg = [t.to(cuda) for t in _p2i_inputs(B, n, 1, H, torch.float32, 23, R)]
    out, ids = F_.p2i_max_forward(pts, feat, binds, bg, 0, R)
    rout, rids = ext.p2i_max_forward_gpu(pts, feat, binds, bg, 0, R)
    assert torch.equal(out, rout)                         # same fp64 cosine on the same GPU -> bit-exact values
    assert (ids != rids).float().mean().item() < 1e-4     # winner id differs only on exact value ties (reference: first locker)
    gout = torch.rand_like(bg)
    gp, gf, gb = F_.p2i_max_backward(gout, rids, pts, feat, 0, R)
    rgp, rgf, rgb = ext.p2i_max_backward_gpu(gout, rids, pts, feat, 0, R)
    assert torch.equal(gb, rgb) and torch.allclose(gf, rgf, rtol=1e-5, atol=1e-6) and torch.allclose(gp, rgp, rtol=1e-5, atol=1e-5)
    s = F_.p2i_sum_forward(pts, feat, binds, bg, 0, R)
    rs = ext.p2i_sum_forward_gpu(pts, feat, binds, bg, 0, R)
    assert torch.allclose(s, rs, rtol=1e-5, atol=1e-5)
    sgp, sgf = F_.p2i_sum_backward(gout, pts, feat, binds, 0, R)
    rsgp, rsgf = ext.p2i_sum_backward_gpu(gout, pts, feat, binds, 0, R)
    assert torch.allclose(sgf, rsgf, rtol=1e-5, atol=1e-5) and torch.allclose(sgp, rsgp, rtol=1e-4, atol=1e-4)


def test_p2i_dropin_gradcheck_fp64(cuda):
    """The reference's own test (cuda/p2i_op/p2i_test.py:24-35): gradcheck of p2i sum and max in float64."""
    _dropin()
    from cuda.p2i_op import p2i
    torch.manual_seed(24)
    for reduce in ("sum", "max"):
        for _ in range(3):
            pts = (torch.rand(2, 2, dtype=torch.float64, device=cuda) * 1.2 - 0.6).requires_grad_()
            feat = torch.rand(2, 2, dtype=torch.float64, device=cuda).requires_grad_()
            binds = torch.zeros(2, dtype=torch.int32, device=cuda)
            # background well below every splatted value: max() has no kink within the finite-difference step
            bg = torch.full((1, 2, 8, 8), -0.5, dtype=torch.float64, device=cuda).requires_grad_()
            assert torch.autograd.gradcheck(lambda p, f, b: p2i(p, f, binds, b, 3.0, "cos", reduce), (pts, feat, bg), eps=1e-6, atol=1e-5)
    with pytest.raises(RuntimeError):
        p2i(pts, feat, binds, bg, 3.0, "cos", "mean")


# ================================================================ kNN
@pytest.mark.parametrize("B,C,N,k,seed", [(2, 3, 512, 8, 31), (2, 64, 300, 8, 32), (1, 256, 1024, 8, 33), (2, 5, 130, 16, 34)])
def test_knn_vs_oracle_sets(cuda, B, C, N, k, seed):
    from sparenet_b200 import functional as F_
    torch.manual_seed(seed)
    x = torch.rand(B, C, N)
    idx = F_.knn_indices(x.to(cuda), k).cpu()
    oidx, od = oracle.knn(x, k, return_dist=True)
    assert (idx[:, :, 0] == torch.arange(N)).all()                           # self first (distance exactly 0)
    same = (idx.sort(-1)[0] == oidx.sort(-1)[0]).all(-1)
    # a set may differ from the fp64 oracle only where the k-th and (k+1)-th fp32 distances are within rounding
    if not same.all():
        xt = x.transpose(1, 2).double()
        for b, i in (~same).nonzero().tolist():
            dm = (xt[b, idx[b, i].long()] - xt[b, i]).pow(2).sum(-1).max()
            do = (xt[b, oidx[b, i].long()] - xt[b, i]).pow(2).sum(-1).max()
            assert abs(dm - do) <= 1e-5 * do
    assert same.float().mean().item() > 0.999


@pytest.mark.parametrize("B,C,N,k,offset", [(2, 64, 512, 8, 0.0), (3, 256, 2048, 8, 3.0), (2, 512, 1024, 16, 1.0), (1, 128, 320, 8, 0.0),
                                              (1, 64, 4096, 8, 0.5), (2, 256, 2048, 32, 0.0), (2, 64, 96, 32, 0.0)])
def test_knn_pruned_identical_to_brute_force(cuda, monkeypatch, B, C, N, k, offset):
    """tensor-core Gram matrix as a pruning filter + exact re-evaluation == the brute-force kernel, index for index; `offset` adds
    a common mean to the features (large norms, small distances: the cancellation case the bound has to survive); the last case
    has duplicated and all-zero points."""
    from sparenet_b200 import functional as F_
    torch.manual_seed(B * 100 + C)
    x = torch.randn(B, C, N, device=cuda) + offset
    if C == 128:
        x[:, :, 200:260] = x[:, :, 100:160]
        x[:, :, 280:] = 0
    a = F_.knn_indices_pruned(x, k)                      # Gram matrix from the tcgen05 TF32 GEMM (positions % 32 == 0)
    monkeypatch.setenv("SNB_KNN_PRUNE_CACHE", "0")       # the variant without the register cache of the approximate distances
    a2 = F_.knn_indices_pruned(x, k)
    monkeypatch.setenv("SNB_KNN_PRUNE", "0")             # the reference side is always the brute-force kernels
    ref = F_.knn_indices(x, k)
    assert torch.equal(a, ref) and torch.equal(a2, ref)


@pytest.mark.parametrize("B,C,N,k", [(32, 3, 2048, 8), (2, 3, 130, 16), (3, 4, 1000, 8), (2, 3, 40, 32), (2, 2, 2048, 20)])
def test_knn_small_identical_to_two_kernel_path(cuda, monkeypatch, B, C, N, k):
    """Fused distance + selection for narrow features (the encoder's xyz layer) == distance matrix + top-k kernels, index for index,
    including duplicated points (ties by index, list overflow) and a query whose neighbours are all identical."""
    from sparenet_b200 import functional as F_
    torch.manual_seed(N + k)
    x = torch.rand(B, C, N, device=cuda)
    if N >= 1000:
        x[:, :, 100:400] = x[:, :, 500:501]               # 300 copies of one point: more ties than the list holds
        x[:, :, 900:910] = x[:, :, 10:20]
    a = F_._knn_brute(x, k)
    monkeypatch.setenv("SNB_KNN_SMALL", "0")
    assert torch.equal(a, F_._knn_brute(x, k))


def test_transpose_cn(cuda):
    from sparenet_b200 import _lib
    from sparenet_b200._lib import check, ptr, stream_ptr
    torch.manual_seed(3)
    for B, C, N in ((2, 256, 2048), (3, 37, 100), (1, 4, 31)):
        x = torch.randn(B, C, N, device=cuda)
        xT = torch.empty(B, N, C, device=cuda)
        check(_lib.load().snb_transpose_cn(ptr(x), B, C, N, ptr(xT), stream_ptr()), "transpose_cn")
        assert torch.equal(xT, x.transpose(1, 2).contiguous())


def test_knn_cuda_shim(cuda):
    _dropin()
    from knn_cuda import KNN
    torch.manual_seed(35)
    ref = torch.rand(2, 200, 16, device=cuda)
    dist, idx = KNN(k=8, transpose_mode=True)(ref, ref)
    assert idx.dtype == torch.int64 and idx.shape == (2, 200, 8) and dist.shape == (2, 200, 8)
    D = torch.cdist(ref, ref)
    assert torch.equal(idx.sort(-1)[0], D.topk(8, largest=False)[1].sort(-1)[0])


def test_integration_snippet_runs(cuda):
    """docs/integration_snippet_chamfer.py (quoted in INTEGRATION.md section B) executed as written."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("integration_snippet_chamfer", os.path.join(root, "docs", "integration_snippet_chamfer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(1)
    x, y = torch.rand(2, 1024, 3), torch.rand(2, 1024, 3)
    d1, d2, i1, i2 = mod.chamfer_forward(x.to(cuda), y.to(cuda))
    od1, od2, oi1, oi2 = oracle.chamfer_fwd(x, y)
    assert torch.equal(d1.cpu(), od1) and torch.equal(i1.cpu(), oi1) and torch.equal(d2.cpu(), od2) and torch.equal(i2.cpu(), oi2)
