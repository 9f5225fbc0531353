"""CPU tests of the oracle itself: pinned against the reference's golden vectors (tests/golden/, generated
by the reference's own C++ CPU Chamfer) and checked for the invariants the other ops must satisfy."""
import os

import numpy as np
import pytest
import torch

import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_chamfer_oracle_vs_reference_cpu_golden(tag):
    g = np.load(os.path.join(GOLD, "chamfer_cpu_ref.npz"))
    x, y = torch.from_numpy(g[f"{tag}_x"]), torch.from_numpy(g[f"{tag}_y"])
    d1, d2, i1, i2 = oracle.chamfer_fwd(x, y)
    # the reference CPU path does not contract to FMA and keeps `best` in double: <= 1 ulp on values
    assert torch.allclose(d1, torch.from_numpy(g[f"{tag}_d1"]), rtol=3e-7, atol=0)
    assert torch.allclose(d2, torch.from_numpy(g[f"{tag}_d2"]), rtol=3e-7, atol=0)
    for mine, ref, q, r in ((i1, g[f"{tag}_i1"], x, y), (i2, g[f"{tag}_i2"], y, x)):
        ref = torch.from_numpy(ref)
        bad = (mine != ref).nonzero()
        for b, j in bad.tolist():  # any disagreement must be a 1-ulp near-tie
            da = (r[b, mine[b, j]] - q[b, j]).double().pow(2).sum()
            db = (r[b, ref[b, j]] - q[b, j]).double().pow(2).sum()
            assert abs(da - db) <= 1e-6 * max(da, db)
    gx, gy = oracle.chamfer_bwd(x, y, torch.from_numpy(g[f"{tag}_i1"]), torch.from_numpy(g[f"{tag}_i2"]),
                                torch.from_numpy(g[f"{tag}_g1"]), torch.from_numpy(g[f"{tag}_g2"]))
    assert torch.allclose(gx, torch.from_numpy(g[f"{tag}_gx"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(gy, torch.from_numpy(g[f"{tag}_gy"]), rtol=1e-6, atol=1e-7)


def test_chamfer_tie_rule_lowest_index():
    x = torch.zeros(1, 4, 3)
    y = torch.tensor([[[1.0, 0, 0], [0, 1.0, 0], [1.0, 0, 0], [0, 0, 1.0]]])
    _, _, i1, _ = oracle.chamfer_fwd(x, y)
    assert i1.tolist() == [[0, 0, 0, 0]]


def test_emd_properties():
    torch.manual_seed(4)
    x, y = torch.rand(2, 1024, 3), torch.rand(2, 1024, 3)
    dist, ass, ev = oracle.emd_fwd(x, y, 0.005, 50, return_evals=True)
    assert ass.min() >= 0 and ass.max() < 1024
    gathered = torch.gather(y, 1, ass.long().unsqueeze(-1).expand(-1, -1, 3))
    assert torch.allclose(dist, (x - gathered).pow(2).sum(-1), rtol=1e-5, atol=1e-9)
    assert (ev >= 1024 * 1024).all()  # round 0 alone evaluates n*n pairs
    # identical clouds: every point must win itself in round 0 and the distance is exactly 0
    d0, a0 = oracle.emd_fwd(x, x.clone(), 0.005, 5)
    assert (a0 == torch.arange(1024, dtype=torch.int32)).all() and (d0 == 0).all()
    with pytest.raises(ValueError):
        oracle.emd_fwd(torch.rand(1, 1000, 3), torch.rand(1, 1000, 3), 0.005, 5)


def test_expansion_properties():
    torch.manual_seed(7)
    x = torch.rand(2, 1024, 3)
    dist, idx, mml = oracle.expansion_fwd(x, 256, 1.5)
    assert ((dist > 0) == (idx >= 0)).all()
    # tagged edges stay inside their primitive and are longer than alpha * mean of that primitive
    prim = torch.arange(1024) // 256
    sel = idx >= 0
    assert (prim.expand(2, -1)[sel] == (idx[sel] // 256)).all()
    # brute-force Prim total per primitive (double) == mean * (p-1)
    for b in range(2):
        tot = 0.0
        for y in range(4):
            p = x[b, y * 256:(y + 1) * 256].double()
            D = (p[:, None] - p[None]).pow(2).sum(-1).sqrt()
            in_tree = torch.zeros(256, dtype=torch.bool); in_tree[0] = True
            best = D[0].clone(); w = 0.0
            for _ in range(255):
                best_m = best.masked_fill(in_tree, float("inf"))
                v = int(best_m.argmin()); w += float(best_m[v]); in_tree[v] = True
                best = torch.minimum(best, D[v])
            tot += w / 255
        assert abs(tot / 4 - float(mml[b])) < 1e-5
    with pytest.raises(ValueError):
        oracle.expansion_fwd(x, 384, 1.5)


def test_mds_properties():
    torch.manual_seed(9)
    x = torch.rand(2, 640, 3)
    mml = torch.tensor([0.05, 0.08])
    idx = oracle.mds(x, 320, mml)
    assert (idx[:, 0] == 0).all()
    for b in range(2):
        assert idx[b].unique().numel() == 320  # no repeats while unsampled points remain
        bad, mism = oracle.mds_check(x[b], mml[b], idx[b])
        assert bad == 0 and mism == 0
    # m > n: after all points are parked the sampler keeps returning index 0 (MDS_cuda.cu:121-133)
    idx2 = oracle.mds(x[:, :16], 20, mml)
    assert (idx2[:, 16:] == 0).all() and idx2[0, :16].unique().numel() == 16


def test_gather_roundtrip():
    torch.manual_seed(3)
    f = torch.rand(2, 4, 50)
    idx = torch.stack([torch.randperm(50)[:30], torch.randperm(50)[:30]]).int()
    out = oracle.gather_fwd(f, idx)
    assert torch.equal(out, torch.gather(f, 2, idx.long().unsqueeze(1).expand(-1, 4, -1)))
    g = oracle.gather_bwd(out, idx, 50)
    assert torch.equal(torch.gather(g, 2, idx.long().unsqueeze(1).expand(-1, 4, -1)), out)


@pytest.mark.parametrize("reduce", ["max", "sum"])
def test_p2i_oracle_gradients_fp64(reduce):
    """finite differences in float64 -- the reference's own p2i_test.py:24-35 strategy."""
    torch.manual_seed(11)
    H = W = 12
    pts = (torch.rand(5, 2, dtype=torch.float64) * 0.6 + 0.2) * (H - 1)
    feat = torch.rand(5, 2, dtype=torch.float64) + 0.5
    binds = torch.tensor([0, 0, 1, 1, 0], dtype=torch.int32)
    bg = torch.zeros(2, 2, H, W, dtype=torch.float64)
    gout = torch.rand(2, 2, H, W, dtype=torch.float64)
    R = 3.3

    def fwd(p, f):
        return oracle.p2i_max_fwd(p, f, binds, bg, R)[0] if reduce == "max" else oracle.p2i_sum_fwd(p, f, binds, bg, R)

    if reduce == "max":
        out, ids = oracle.p2i_max_fwd(pts, feat, binds, bg, R)
        gp, gf, gb = oracle.p2i_max_bwd(gout, ids, pts, feat, R)
        assert torch.equal(gb, torch.where(ids < 0, gout, torch.zeros_like(gout)))
    else:
        gp, gf = oracle.p2i_sum_bwd(gout, pts, feat, binds, R)
    eps = 1e-6
    for t, g in ((pts, gp), (feat, gf)):
        num = torch.zeros_like(t)
        for i in range(t.numel()):
            tp, tm = t.clone().view(-1), t.clone().view(-1)
            tp[i] += eps; tm[i] -= eps
            a = fwd(tp.view_as(t), feat) if t is pts else fwd(pts, tp.view_as(t))
            b = fwd(tm.view_as(t), feat) if t is pts else fwd(pts, tm.view_as(t))
            num.view(-1)[i] = ((a - b) * gout).sum() / (2 * eps)
        assert torch.allclose(num, g, rtol=1e-4, atol=1e-6), (num - g).abs().max()


def test_p2i_max_semantics():
    pts = torch.tensor([[4.0, 4.0], [4.0, 4.0], [20.0, 20.0]])
    feat = torch.tensor([[0.5], [0.5], [1.0]])
    binds = torch.tensor([0, 0, 5], dtype=torch.int32)  # third point: batch index out of range -> skipped
    bg = torch.full((1, 1, 9, 9), 0.25)
    out, ids = oracle.p2i_max_fwd(pts, feat, binds, bg, 2.0)
    assert out[0, 0, 4, 4] == 0.5 and ids[0, 0, 4, 4] == 0      # exact tie -> lowest point id
    assert out[0, 0, 0, 0] == 0.25 and ids[0, 0, 0, 0] == -1    # background untouched
    assert (out >= 0.25).all()


def test_knn_oracle():
    torch.manual_seed(5)
    x = torch.rand(2, 6, 40)
    idx, d = oracle.knn(x, 5, return_dist=True)
    D = (x.transpose(1, 2)[:, :, None] - x.transpose(1, 2)[:, None]).pow(2).sum(-1)
    ref = D.topk(5, dim=-1, largest=False)
    assert torch.equal(idx.long().sort(-1)[0], ref.indices.sort(-1)[0])
    assert (idx[:, :, 0] == torch.arange(40)).all()  # self first
