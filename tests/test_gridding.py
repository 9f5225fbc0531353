"""Gridding / GriddingReverse (GRNet ops named by BASELINE.json's drop-in list): oracle sanity on CPU, CUDA parity on the GPU
against the oracle and the reference's own extension (oracle/_ref/gridding.so)."""
import sys

import pytest
import torch

import oracle
from tests.conftest import ref_ext

B6 = (-4.0, 3.0, -4.0, 3.0, -4.0, 3.0)          # scale 8 -> half-scale 4: bounds [-4, 3], 8^3 vertices


def _cloud(B, n, seed, half=4):
    torch.manual_seed(seed)
    p = (torch.rand(B, n, 3) * 2 - 1) * (half - 1.01)
    p[:, ::7] = torch.round(p[:, ::7] * 0.6)      # points exactly on grid planes (floor == ceil branch), upper corner still inside
    return p


def test_oracle_gridding_partition_of_unity_and_grad():
    p = _cloud(2, 50, 1)
    grid, w, ix = oracle.gridding_fwd(p, B6)
    assert torch.allclose(grid.sum(1), torch.full((2,), 50.0), rtol=1e-5)          # trilinear weights of a point sum to 1
    assert (ix >= 0).all() and ix.max() < 512
    gg = torch.rand(2, 512)
    g = oracle.gridding_bwd(w, ix, gg)
    pd = p.double().requires_grad_()
    lo = torch.floor(pd.detach())
    fr = pd - lo
    tot = 0
    for t in range(8):
        u = torch.tensor([(t >> 2) & 1, (t >> 1) & 1, t & 1], dtype=torch.float64)
        wt = torch.where(u.bool(), fr, 1 - fr).prod(-1)
        idx = (((lo[..., 0] + u[0] + 4) * 8 + (lo[..., 1] + u[1] + 4)) * 8 + (lo[..., 2] + u[2] + 4)).long()
        tot = tot + (wt * torch.gather(gg.double(), 1, idx)).sum()
    tot.backward()
    on_plane = (p == torch.round(p)).any(-1)                                       # kinks of the trilinear kernel
    assert torch.allclose(g[~on_plane].double(), pd.grad[~on_plane], rtol=1e-4, atol=1e-5)


def test_oracle_gridding_reverse_roundtrip_and_grad():
    torch.manual_seed(2)
    S = 6
    grid = torch.rand(2, S, S, S)
    grid[0, :2] = 0                                                                # empty cells are skipped (sum < 1e-6)
    pts = oracle.gridding_rev_fwd(grid, S)
    assert pts.shape == (2, S ** 3, 3)
    gd = grid.double().requires_grad_()
    # reference formula in torch: weighted centroid of the 8-corner cell below each vertex
    acc = torch.zeros(2, S, S, S, 3, dtype=torch.float64)
    wsum = torch.zeros(2, S, S, S, dtype=torch.float64)
    rng = torch.arange(S, dtype=torch.float64) - S // 2
    for t in range(8):
        dx, dy, dz = (t >> 2) & 1, (t >> 1) & 1, t & 1
        wv = torch.zeros(2, S, S, S, dtype=torch.float64)
        wv[:, 1:, 1:, 1:] = gd[:, dx:S - 1 + dx, dy:S - 1 + dy, dz:S - 1 + dz]
        cx = (rng - 1 + dx).view(1, S, 1, 1).expand(2, S, S, S)
        cy = (rng - 1 + dy).view(1, 1, S, 1).expand(2, S, S, S)
        cz = (rng - 1 + dz).view(1, 1, 1, S).expand(2, S, S, S)
        acc = acc + wv.unsqueeze(-1) * torch.stack((cx, cy, cz), -1)
        wsum = wsum + wv
    ok = wsum >= 1e-6
    ref = torch.where(ok.unsqueeze(-1), acc / wsum.clamp_min(1e-30).unsqueeze(-1), torch.zeros_like(acc))
    assert torch.allclose(pts.double(), ref.view(2, -1, 3), rtol=1e-5, atol=1e-5)
    gp = torch.rand(2, S ** 3, 3)
    (ref.view(2, -1, 3) * gp.double()).sum().backward()
    g = oracle.gridding_rev_bwd(pts, grid, gp, S)
    assert torch.allclose(g.double(), gd.grad.view(2, -1), rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_gridding_cuda_vs_oracle_and_reference(cuda):
    from sparenet_b200 import functional as F_
    p = _cloud(3, 700, 3)
    grid, w, ix = F_.gridding_forward(p.to(cuda), B6)
    og, ow, oi = oracle.gridding_fwd(p, B6)
    assert torch.equal(ix.cpu(), oi) and torch.equal(w.cpu(), ow)
    assert torch.allclose(grid.cpu(), og, rtol=1e-5, atol=1e-6)                    # float atomics: order only
    gg = torch.rand(3, 512)
    g = F_.gridding_backward(w, ix, gg.to(cuda))
    assert torch.equal(g.cpu(), oracle.gridding_bwd(ow, oi, gg))
    S = 8
    torch.manual_seed(4)
    gr = torch.rand(2, S, S, S)
    gr[1, 5:] = 0
    pts = F_.gridding_reverse_forward(gr.view(2, -1).to(cuda), S)
    opts = oracle.gridding_rev_fwd(gr, S)
    assert torch.allclose(pts.cpu(), opts, rtol=1e-6, atol=1e-6)
    gp = torch.rand(2, S ** 3, 3)
    gb = F_.gridding_reverse_backward(pts, gr.view(2, -1).to(cuda), gp.to(cuda), S)
    assert torch.allclose(gb.cpu(), oracle.gridding_rev_bwd(opts, gr, gp, S), rtol=1e-5, atol=1e-6)
    ext = ref_ext("gridding")
    if ext is not None:
        rg, rw, ri = ext.forward(*B6, p.to(cuda))
        assert torch.equal(ri, ix) and torch.equal(rw, w) and torch.allclose(rg, grid, rtol=1e-5, atol=1e-6)
        assert torch.allclose(ext.backward(rw, ri, gg.to(cuda)), g, rtol=1e-6, atol=1e-7)
        rp = ext.rev_forward(S, gr.to(cuda))
        assert torch.allclose(rp, pts, rtol=1e-6, atol=1e-6)
        assert torch.allclose(ext.rev_backward(rp, gr.to(cuda), gp.to(cuda)).view(2, -1), gb, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_gridding_dropin_modules(cuda):
    import sparenet_b200
    if sparenet_b200.dropin_path() not in sys.path:
        sys.path.insert(0, sparenet_b200.dropin_path())
    from cuda.gridding import Gridding, GriddingReverse
    torch.manual_seed(5)
    pc = (torch.rand(2, 300, 3, device=cuda) - 0.5) * 1.6
    pc[0, 250:] = 0                                                                # zero-padded rows are dropped per sample
    pc.requires_grad_()
    grid = Gridding(scale=16)(pc)
    assert grid.shape == (2, 16 ** 3)
    assert abs(grid[0].sum().item() - 250) < 1e-2 and abs(grid[1].sum().item() - 300) < 1e-2
    out = GriddingReverse(scale=16)(grid.view(2, 16, 16, 16))
    assert out.shape == (2, 16 ** 3, 3)
    out.square().sum().backward()
    assert torch.isfinite(pc.grad).all() and pc.grad[0, 250:].abs().sum() == 0 and pc.grad.abs().sum() > 0
