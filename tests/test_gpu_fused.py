"""GPU tests of the fused generator kernels (snb_edge_reduce_*, snb_row_*) through their autograd Functions, against the
plain PyTorch definitions in tests/fused_ref.py evaluated in float64.  Floating point: <= 1e-5 relative."""
import pytest
import torch

from tests import fused_ref as R

pytestmark = pytest.mark.gpu


def _close(a, b, rtol=1e-5, atol=1e-6):
    b = b.to(a.dtype)
    return torch.allclose(a, b, rtol=rtol, atol=atol * (b.abs().max().item() + 1e-30))


@pytest.mark.parametrize("B,C,N,k", [(2, 5, 300, 8), (3, 64, 2048, 8), (1, 16, 97, 3), (2, 8, 4096, 16)])
def test_edge_reduce_forward_backward(cuda, B, C, N, k):
    from sparenet_b200 import fused
    torch.manual_seed(B * 1000 + N)
    # distinct values: on an exact tie inside a neighbourhood torch.amax splits/moves the gradient differently (diag3)
    a = (torch.randperm(B * C * N, device=cuda).float() / (B * C * N) * 6 - 3).view(B, C, N)
    c = torch.randn(B, C, N, device=cuda)
    idx = torch.stack([torch.stack([torch.randperm(N, device=cuda)[:k] for _ in range(N)]) for _ in range(B)]).int()
    a1, c1 = a.clone().requires_grad_(), c.clone().requires_grad_()
    a2, c2 = a.double().requires_grad_(), c.double().requires_grad_()
    o1 = fused.edge_reduce(a1, c1, idx)
    o2 = R.edge_reduce(a2, c2, idx)
    assert torch.equal(o1[0], o2[0].float()) and torch.equal(o1[1], o2[1].float())      # max/min of fp32 sums: exact
    assert _close(o1[2], o2[2], 1e-6) and _close(o1[3], o2[3], 1e-6)
    w = [torch.randn_like(t) for t in o1]
    sum((x * y).sum() for x, y in zip(o1, w)).backward()
    sum((x * y.double()).sum() for x, y in zip(o2, w)).backward()
    ea = (a1.grad.double() - a2.grad).abs().max().item() / a2.grad.abs().max().item()
    ec = (c1.grad.double() - c2.grad).abs().max().item() / c2.grad.abs().max().item()
    print(f"[edge_reduce] B={B} C={C} N={N} k={k}: max grad err / scale  a={ea:.2e} c={ec:.2e}")
    assert ea < 1e-5 and ec < 1e-5      # fp32 accumulation of ~k terms per point against float64


@pytest.mark.parametrize("B,C,N,k", [(2, 5, 300, 8), (3, 64, 2048, 8), (1, 16, 97, 3)])
def test_edge_reduce_sel_forward_backward(cuda, B, C, N, k):
    """the per-channel selected extremum (no torch.where, one slot tensor) against the two-output reference"""
    from sparenet_b200 import fused
    torch.manual_seed(B * 1000 + N + 1)
    a = (torch.randperm(B * C * N, device=cuda).float() / (B * C * N) * 6 - 3).view(B, C, N)
    c = torch.randn(B, C, N, device=cuda)
    idx = torch.stack([torch.stack([torch.randperm(N, device=cuda)[:k] for _ in range(N)]) for _ in range(B)]).int()
    sel = torch.rand(C, device=cuda) > 0.4
    a1, c1 = a.clone().requires_grad_(), c.clone().requires_grad_()
    a2, c2 = a.double().requires_grad_(), c.double().requires_grad_()
    o1 = fused.edge_reduce_sel(a1, c1, idx, sel)
    o2 = R.edge_reduce_sel(a2, c2, idx, sel)
    assert torch.equal(o1[0], o2[0].float())
    assert _close(o1[1], o2[1], 1e-6) and _close(o1[2], o2[2], 1e-6)
    w = [torch.randn_like(t) for t in o1]
    sum((x * y).sum() for x, y in zip(o1, w)).backward()
    sum((x * y.double()).sum() for x, y in zip(o2, w)).backward()
    ea = (a1.grad.double() - a2.grad).abs().max().item() / a2.grad.abs().max().item()
    ec = (c1.grad.double() - c2.grad).abs().max().item() / c2.grad.abs().max().item()
    assert ea < 1e-5 and ec < 1e-5


@pytest.mark.parametrize("shape", [(4, 7, 512), (3, 5, 2048), (2, 3, 333), (1, 2, 16384), (6, 1)])
def test_row_stats_forward_backward(cuda, shape):
    from sparenet_b200 import fused
    torch.manual_seed(sum(shape))
    h = torch.randn(*shape, device=cuda) * 0.3 + 5.0            # large mean / small spread: the cancellation-prone case
    h1, h2 = h.clone().requires_grad_(), h.double().requires_grad_()
    m1, v1 = fused.row_stats(h1)
    m2, v2 = R.row_stats(h2)
    assert _close(m1, m2, 1e-6) and _close(v1, v2, 1e-5)
    wm, wv = torch.randn_like(m1), torch.randn_like(v1)
    ((m1 * wm).sum() + (v1 * wv).sum()).backward()
    ((m2 * wm.double()).sum() + (v2 * wv.double()).sum()).backward()
    assert _close(h1.grad, h2.grad, 1e-5, 1e-6)


@pytest.mark.parametrize("rows,L,in_div,slope", [((4, 6), 512, 1, 0.0), ((3, 5), 2048, 1, 0.2), ((2, 9), 333, 1, 0.0),
                                                ((3, 4), 512, 5, 0.0), ((2, 3), 100, 4, 0.2), ((2, 7), 128, 3, 0.2), ((3, 3), 256, 32, 0.0),
                                                ((2, 2), 1024, 6, 0.2), ((1, 5), 640, 4, 0.0)])
def test_row_affine_act_forward_backward(cuda, rows, L, in_div, slope):
    from sparenet_b200 import fused
    torch.manual_seed(L + in_div)
    h = torch.randn(*rows, L, device=cuda)
    R_out = rows[0] * rows[1] * in_div
    sc = torch.randn(R_out, device=cuda)
    sh = torch.randn(R_out, device=cuda) * 0.5
    out_shape = (*rows, in_div, L) if in_div > 1 else (*rows, L)
    t1 = [t.clone().requires_grad_() for t in (h, sc, sh)]
    t2 = [t.double().requires_grad_() for t in (h, sc, sh)]
    y1 = fused.row_affine_act(t1[0], t1[1], t1[2], slope=slope, in_div=in_div, out_shape=out_shape)
    y2 = R.row_affine_act(t2[0], t2[1], t2[2], slope=slope, in_div=in_div, out_shape=out_shape)
    assert y1.shape == y2.shape and _close(y1, y2, 1e-6)
    w = torch.randn_like(y1)
    (y1 * w).sum().backward()
    (y2 * w.double()).sum().backward()
    for a, b in zip(t1, t2):
        assert _close(a.grad, b.grad, 2e-5, 2e-6)


def test_row_minmax(cuda):
    from sparenet_b200 import fused
    torch.manual_seed(3)
    h = torch.randn(5, 33, 1000, device=cuda)
    h[0, 0, 10] = h[0, 0, 500] = 9.0                            # tie: the first position takes the gradient
    h1 = h.clone().requires_grad_()
    vmax, vmin = fused.row_minmax(h1)
    assert torch.equal(vmax, h.amax(-1)) and torch.equal(vmin, h.amin(-1))
    (vmax.sum() + 2 * vmin.sum()).backward()
    assert h1.grad[0, 0, 10] == 1 and h1.grad[0, 0, 500] == 0
    assert h1.grad.sum().item() == pytest.approx(3 * 5 * 33)


@pytest.mark.parametrize("B,C,L,slope", [(4, 16, 512, 0.0), (3, 8, 2048, 0.2), (2, 5, 333, 0.0)])
def test_row_norm_act_forward_backward(cuda, B, C, L, slope):
    """BN o SE o (leaky)ReLU as one node: two-phase backward (snb_row_act_bwd_reduce / snb_row_norm_act_bwd) vs autograd in fp64."""
    from sparenet_b200 import fused
    torch.manual_seed(B * 100 + L)
    h = torch.randn(B, C, L, device=cuda) * 0.7 + 0.3
    g, beta, rb = torch.rand(C, device=cuda) + 0.5, torch.randn(C, device=cuda) * 0.1, torch.randn(B, C, device=cuda) * 0.2
    w1, w2 = torch.randn(max(C // 4, 1), C, device=cuda) * 0.5, torch.randn(C, max(C // 4, 1), device=cuda) * 0.5

    def tail(m_bc, v_bc, rb, g, beta, w1, w2):                  # the refiner's closed form (PointNetRes._bn_se_relu)
        m = m_bc + rb
        mean = m.mean(0)
        var = v_bc.mean(0) + ((m - mean) ** 2).mean(0)
        inv = torch.rsqrt(var + 1e-5)
        scale, shift = g * inv, beta - g * inv * mean
        gate = torch.sigmoid(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(m * scale + shift, w1)), w2))
        return gate * scale, gate * shift + rb * gate * scale

    leaves32 = [t.clone().requires_grad_() for t in (h, rb, g, beta, w1, w2)]
    leaves64 = [t.double().requires_grad_() for t in (h, rb, g, beta, w1, w2)]
    y1 = fused.row_norm_act(leaves32[0], tail, tuple(leaves32[1:]), slope)
    y2 = R.row_norm_act(leaves64[0], tail, tuple(leaves64[1:]), slope)
    assert _close(y1, y2, 1e-5, 1e-6)
    w = torch.randn_like(y1)
    (y1 * w).sum().backward()
    (y2 * w.double()).sum().backward()
    for name, a, b in zip(("h", "row_bias", "gamma", "beta", "w1", "w2"), leaves32, leaves64):
        err = (a.grad.double() - b.grad).abs().max().item() / (b.grad.abs().max().item() + 1e-30)
        print(f"[row_norm_act] B={B} C={C} L={L} grad {name}: err/scale {err:.2e}")
        assert err < 2e-5, name


@pytest.mark.parametrize("B,Ci,Co,N", [(3, 16, 40, 700), (2, 128, 1024, 4096)])
def test_conv_row_reduce_forward_backward(cuda, B, Ci, Co, N):
    """h = W x reduced to row mean/var/max/min in one pass; Gram-matrix backward vs autograd through the explicit h (fp64)."""
    from sparenet_b200 import fused
    torch.manual_seed(Co + N)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False      # fp32 GEMMs: test the algebra, not TF32
    try:
        x, W = torch.randn(B, Ci, N, device=cuda), torch.randn(Co, Ci, 1, device=cuda) / Ci ** 0.5
        x1, W1 = x.clone().requires_grad_(), W.clone().requires_grad_()
        x2, W2 = x.double().requires_grad_(), W.double().requires_grad_()
        o1 = fused.conv_row_reduce(x1, W1)
        o2 = R.conv_row_reduce(x2, W2)
        for a, b in zip(o1, o2):                                  # h has unit scale: the row means nearly cancel, so the floor is absolute
            assert torch.allclose(a.double(), b, rtol=1e-5, atol=2e-6)
        ws = [torch.randn_like(t) for t in o1]
        sum((a * w).sum() for a, w in zip(o1, ws)).backward()
        sum((a * w.double()).sum() for a, w in zip(o2, ws)).backward()
        ex = (x1.grad.double() - x2.grad).abs().max().item() / x2.grad.abs().max().item()
        ew = (W1.grad.double() - W2.grad).abs().max().item() / W2.grad.abs().max().item()
        print(f"[conv_row_reduce] B={B} {Ci}->{Co} N={N}: grad err/scale x={ex:.2e} W={ew:.2e}")
        # the two Gram-matrix products of the backward run in TF32 by design (they stand for cuDNN's TF32 data / weight gradients):
        # TF32 bound here; the fp32 algebra itself is covered by the CPU float64 tests
        assert ex < 2e-3 and ew < 2e-3
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("B,C,L,slope", [(4, 16, 512, 0.2), (3, 8, 2048, 0.0), (2, 5, 333, 0.2)])
def test_row_norm_act_pool_forward_backward(cuda, B, C, L, slope):
    """[max | mean over the row] of BN o LeakyReLU without the activated tensor (snb_row_act_pool_*) vs autograd in fp64."""
    from sparenet_b200 import fused
    torch.manual_seed(B * 10 + L)
    h = torch.randn(B, C, L, device=cuda) * 0.7 + 0.3
    g, beta = torch.rand(C, device=cuda) + 0.5, torch.randn(C, device=cuda) * 0.1

    def tail(m_bc, v_bc, g, beta):                              # the encoder's BN5 as one per-channel scale/shift
        mean = m_bc.mean(0)
        var = v_bc.mean(0) + ((m_bc - mean) ** 2).mean(0)
        scale = g * torch.rsqrt(var + 1e-5)
        return scale.expand(m_bc.size(0), -1), (beta - scale * mean).expand(m_bc.size(0), -1)

    l32 = [t.clone().requires_grad_() for t in (h, g, beta)]
    l64 = [t.double().requires_grad_() for t in (h, g, beta)]
    pmax, pmean = fused.row_norm_act_pool(l32[0], tail, tuple(l32[1:]), slope)
    qmax, qmean = R.row_norm_act_pool(l64[0], tail, tuple(l64[1:]), slope)
    assert _close(pmax, qmax, 1e-5, 1e-6) and _close(pmean, qmean, 1e-5, 1e-6)
    w1, w2 = torch.randn_like(pmax), torch.randn_like(pmean)
    ((pmax * w1).sum() + (pmean * w2).sum()).backward()
    ((qmax * w1.double()).sum() + (qmean * w2.double()).sum()).backward()
    for name, a, b in zip(("h", "gamma", "beta"), l32, l64):
        err = (a.grad.double() - b.grad).abs().max().item() / (b.grad.abs().max().item() + 1e-30)
        print(f"[row_norm_act_pool] B={B} C={C} L={L} grad {name}: err/scale {err:.2e}")
        assert err < 2e-5, name


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,per_sample,training", [(32, 64, False, True), (32, 512, True, True), (5, 128, False, True),
                                                     (32, 256, False, False), (3, 48, True, False)])
def test_bn_se_tail_matches_pytorch_closed_form(cuda, B, C, per_sample, training):
    """snb_bn_se_tail_fwd/bwd (one launch per direction) against the same closed form written as PyTorch ops
    (PointNetRes._bn_se_tail's tail_torch): scale/shift, the BatchNorm running statistics and every gradient, fp32 both sides,
    <= 2e-5 of each tensor's scale (different summation order only)."""
    import torch.nn as nn
    from sparenet_b200 import fused
    from sparenet_b200.dropin.models import sparenet_generator as G
    torch.manual_seed(C + B)
    L, H = 4096, max(C // 16, 1)
    bn_a, bn_b = nn.BatchNorm1d(C).to(cuda), nn.BatchNorm1d(C).to(cuda)
    for bn in (bn_a, bn_b):
        bn.train(training)
        with torch.no_grad():
            bn.running_mean.copy_(torch.linspace(-0.3, 0.4, C))
            bn.running_var.copy_(torch.linspace(0.5, 1.5, C))
    se = G.SELayer1D(C).to(cuda)
    m0, v0 = torch.randn(B, C, device=cuda), torch.rand(B, C, device=cuda) + 0.1
    rb0 = torch.randn(B, C, device=cuda) if per_sample else torch.randn(C, device=cuda)
    g0, b0 = torch.randn(C, device=cuda), torch.randn(C, device=cuda)
    gS, gT = torch.randn(B, C, device=cuda), torch.randn(B, C, device=cuda)

    def run(fn, bn):
        leaves = [t.clone().requires_grad_() for t in (m0, v0, rb0, g0, b0, se.fc[0].weight.detach(), se.fc[2].weight.detach())]
        S, T = fn(bn, *leaves)
        grads = torch.autograd.grad((S, T), leaves, (gS, gT), allow_unused=True)
        return S, T, grads

    def ref(bn, m, v, rb, g, beta, w1, w2):
        mm = m + rb
        mean, var = G._bn_from_rows(bn, mm, v, L)
        inv = torch.rsqrt(var + bn.eps)
        scale, shift = g * inv, beta - g * inv * mean
        gate = torch.sigmoid(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(mm * scale + shift, w1)), w2))
        gs = gate * scale
        return gs, gate * shift + rb * gs

    S1, T1, G1 = run(ref, bn_a)
    S2, T2, G2 = run(lambda bn, *a: fused.bn_se_tail(*a, bn, L), bn_b)
    names = ["scale", "shift", "g_mean", "g_var", "g_bias", "g_gamma", "g_beta", "g_w1", "g_w2"]
    for name, a, b in zip(names, [S1, T1, *G1], [S2, T2, *G2]):
        if a is None:                       # eval mode: the row variance does not reach the outputs
            assert b is None or float(b.abs().max()) == 0.0, name
            continue
        scale = max(float(a.detach().abs().max()), 1e-6)
        if name == "g_bias" and not per_sample:
            # a bias shared by the batch cancels EXACTLY in train mode (it shifts every row mean and the batch mean alike): the
            # true gradient is 0 and both sides return rounding noise -- judged against the size of the terms that cancel
            scale = max(scale, float(G1[0].detach().abs().sum(0).max()))
        err = float((a - b).detach().abs().max()) / scale
        print(f"[bn_se_tail B={B} C={C} per_sample={per_sample} train={training}] {name}: err/scale {err:.2e}")
        assert err < 2e-5, name
    assert torch.allclose(bn_a.running_mean, bn_b.running_mean, rtol=1e-6, atol=1e-7)
    assert torch.allclose(bn_a.running_var, bn_b.running_var, rtol=1e-6, atol=1e-7)
    assert int(bn_a.num_batches_tracked) == int(bn_b.num_batches_tracked)


@pytest.mark.gpu
@pytest.mark.parametrize("own_grads", [False, True])
def test_flat_adam_matches_torch_adam(cuda, own_grads):
    """sparenet_b200.optim.FlatAdam (one launch over the flat parameter arena) against torch.optim.Adam on the same parameters and
    gradients, 6 steps, odd tensor sizes (padding inside the arena), weight decay on: parameters agree to 2e-6 relative (fp32, the
    same update rule; only fma contraction differs).  own_grads=False exercises GradArena.pack (fresh gradient tensors gathered by
    multi-tensor copies), own_grads=True the accumulate-into-views mode."""
    import torch.nn as nn
    from sparenet_b200.dist import GradArena
    from sparenet_b200.optim import FlatAdam
    torch.manual_seed(3)
    shapes = [(7, 5), (33,), (64, 31, 1), (1,), (128, 128)]
    pa = [nn.Parameter(torch.randn(*s, device=cuda)) for s in shapes]
    pb = [nn.Parameter(p.detach().clone()) for p in pa]
    ref = torch.optim.Adam(pa, lr=3e-3, betas=(0.5, 0.9), eps=1e-8, weight_decay=1e-2)
    arena = GradArena(pb, own_grads=own_grads)
    opt = FlatAdam(pb, lr=3e-3, betas=(0.5, 0.9), eps=1e-8, weight_decay=1e-2, arena=arena)
    for step in range(6):
        grads = [torch.randn(*s, device=cuda) * (1 + step) for s in shapes]
        ref.zero_grad(set_to_none=True)
        opt.zero_grad()
        for p, q, g in zip(pa, pb, grads):
            p.grad = g.clone()
            if own_grads:
                q.grad.add_(g)
            else:
                q.grad = g.clone()
        ref.step()
        opt.step()
    for p, q in zip(pa, pb):
        err = (p - q).abs().max().item() / max(p.abs().max().item(), 1e-6)
        assert err < 2e-6, (tuple(p.shape), err)
    # the parameters live in the arena now: views of one flat buffer, 16-byte aligned
    base = opt.flat_params[0].data_ptr()
    assert all(base <= q.data_ptr() < base + opt.flat_params[0].numel() * 4 and q.data_ptr() % 16 == 0 for q in pb)


@pytest.mark.gpu
@pytest.mark.parametrize("P,C,Cp,B", [(4, 48, 64, 8), (32, 513, 544, 32), (3, 256, 256, 5), (2, 16, 32, 32), (4, 1026, 1056, 32),
                                      (2, 1056, 1056, 4)])
def test_adain_tail_matches_pytorch_closed_form(cuda, P, C, Cp, B):
    """snb_adain_tail_fwd/bwd (one launch per direction for all primitives) against the same closed form as PyTorch ops -- the
    decoders' instance-norm . AdaIN . BatchNorm . SE tail of SpareNetDecode (tail() / _bn_se): scale/shift, the BatchNorm batch
    statistics and all eight gradients, fp32 both sides, <= 3e-5 of each tensor's scale (summation order only)."""
    from sparenet_b200 import fused
    torch.manual_seed(P * 100 + C)
    H, eps = max(C // 16, 1), 1e-5
    mean0 = torch.randn(P, Cp, B, device=cuda)
    var0 = torch.rand(P, Cp, B, device=cuda) * 2 + 0.05
    ws0, bs0 = torch.randn(B, C, device=cuda) * 0.5 + 1.0, torch.randn(B, C, device=cuda) * 0.5
    gam0, bet0 = torch.randn(P, C, 1, device=cuda) * 0.3 + 1.0, torch.randn(P, C, 1, device=cuda) * 0.3
    w10, w20 = torch.randn(P, H, C, device=cuda) / C ** 0.5, torch.randn(P, C, H, device=cuda) / H ** 0.5
    gsc, gsh = torch.randn(P, Cp, B, device=cuda), torch.randn(P, Cp, B, device=cuda)

    def ref(mean, var, wsty, bsty, gam, bet, w1, w2):
        rstd = torch.rsqrt(var + eps)
        v = (var / (var + eps))[:, :C]
        wt, bt = wsty.t().unsqueeze(0), bsty.t().unsqueeze(0)
        mu = bt.mean(-1, keepdim=True).expand(P, -1, -1)
        q = (wt * wt * v + bt * bt).mean(-1, keepdim=True) - mu * mu
        inv = torch.rsqrt(q + eps)
        squeeze = gam * (bt - mu) * inv + bet
        gate = torch.sigmoid(torch.bmm(w2, torch.relu(torch.bmm(w1, squeeze))))
        A = gate * gam * inv * wt
        D = gate * (gam * inv * (bt - mu) + bet)
        pad = (0, 0, 0, Cp - C)
        sc = torch.nn.functional.pad(A, pad) * rstd
        return sc, torch.nn.functional.pad(D, pad) - sc * mean, mu[:, :, 0], q[:, :, 0]

    outs = []
    for fn in (ref, lambda *a: fused.adain_tail(*a, eps)):
        leaves = [t.clone().requires_grad_() for t in (mean0, var0, ws0, bs0, gam0, bet0, w10, w20)]
        sc, sh, mu, q = fn(*leaves)
        grads = torch.autograd.grad((sc, sh), leaves, (gsc, gsh))
        outs.append((sc, sh, mu, q, *grads))
    names = ["scale", "shift", "bn_mean", "bn_var", "g_mean", "g_var", "g_wsty", "g_bsty", "g_gamma", "g_beta", "g_w1", "g_w2"]
    for name, a, b in zip(names, *outs):
        scale = max(float(a.detach().abs().max()), 1e-6)
        err = float((a - b).detach().abs().max()) / scale
        print(f"[adain_tail P={P} C={C} Cp={Cp} B={B}] {name}: err/scale {err:.2e}")
        assert err < 3e-5, name


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,training", [(32, 1024, True), (5, 70, True), (4, 128, False)])
def test_bn_max_tail_matches_pytorch_closed_form(cuda, B, C, training):
    """snb_bn_max_tail_fwd/bwd (PointNetRes conv3 -> bn3 -> max over the points from row statistics and extrema, one launch per
    direction) against the same closed form as PyTorch ops: the global feature, the BatchNorm running statistics and every gradient,
    fp32 both sides (summation order only); the conv bias has no gradient in train mode (it shifts h* and the batch mean alike)."""
    from sparenet_b200 import fused
    torch.manual_seed(B * 1000 + C)
    L = 4096
    m0, v0 = torch.randn(B, C, device=cuda), torch.rand(B, C, device=cuda) + 0.1
    hmax0 = m0 + 3 * v0.sqrt() * (1 + torch.rand(B, C, device=cuda))
    hmin0 = m0 - 3 * v0.sqrt() * (1 + torch.rand(B, C, device=cuda))
    bias0, g0, beta0 = torch.randn(C, device=cuda), torch.randn(C, device=cuda), torch.randn(C, device=cuda)
    gglob = torch.randn(B, C, device=cuda)

    def make_bn():
        bn = torch.nn.BatchNorm1d(C).to(cuda)
        with torch.no_grad():
            bn.running_mean.copy_(torch.linspace(-1, 1, C))
            bn.running_var.copy_(torch.linspace(0.5, 2, C))
        return bn.train(training)

    def ref(bn, m_bc, v_bc, hmax, hmin, bias, g, beta):
        m = m_bc + bias
        if bn.training:
            mean = m.mean(0)
            var = v_bc.mean(0) + ((m - mean) ** 2).mean(0)
            n = B * L
            with torch.no_grad():
                bn.running_mean.mul_(0.9).add_(mean.detach(), alpha=0.1)
                bn.running_var.mul_(0.9).add_(var.detach() * (n / (n - 1)), alpha=0.1)
                bn.num_batches_tracked.add_(1)
        else:
            mean, var = bn.running_mean, bn.running_var
        hstar = torch.where((g > 0).view(1, -1), hmax, hmin) + bias
        return (hstar - mean) * (g * torch.rsqrt(var + bn.eps)) + beta

    outs, bns = [], []
    for fn in (ref, lambda bn, *a: fused.bn_max_tail(*a, bn, L)):
        bn = make_bn()
        leaves = [t.clone().requires_grad_() for t in (m0, v0, hmax0, hmin0, bias0, g0, beta0)]
        glob = fn(bn, *leaves)
        grads = torch.autograd.grad(glob, leaves, gglob, allow_unused=True)
        grads = [torch.zeros_like(t) if gr is None else gr for gr, t in zip(grads, leaves)]
        outs.append((glob, bn.running_mean.clone(), bn.running_var.clone(), *grads))
        bns.append(bn)
    assert int(bns[0].num_batches_tracked) == int(bns[1].num_batches_tracked)
    names = ["glob", "running_mean", "running_var", "g_m", "g_v", "g_hmax", "g_hmin", "g_bias", "g_gamma", "g_beta"]
    for name, a, b in zip(names, *outs):
        scale = max(float(a.detach().abs().max()), 1e-6)
        err = float((a - b).detach().abs().max()) / scale
        print(f"[bn_max_tail B={B} C={C} train={training}] {name}: err/scale {err:.2e}")
        if name == "g_bias" and training:      # exactly 0 in the kernel; autograd's two cancelling sums leave rounding noise
            assert float(b.abs().max()) == 0.0 and float(a.abs().max()) < 1e-3 * float(outs[0][8].abs().max() + 1)
        else:
            assert err < 3e-5, name


@pytest.mark.gpu
def test_conv_row_reduce_backward_extrema_kernel(cuda):
    """conv_row_reduce_backward on fp32 CUDA tensors (the extrema adjoint through snb_conv_extrema_bwd, one launch for max and min)
    against the same function in float64 (pure PyTorch path): gx, gW and the row term."""
    from sparenet_b200 import fused
    torch.manual_seed(11)
    B, Ci, Co, N = 3, 64, 200, 512
    x, W = torch.randn(B, Ci, N, device=cuda), torch.randn(Co, Ci, device=cuda) / 8
    h = torch.matmul(W.double(), x.double())
    mean, imax, imin = h.mean(-1), h.argmax(-1).int(), h.argmin(-1).int()
    gmean, gvar, gmax, gmin = (torch.randn(B, Co, device=cuda) for _ in range(4))
    up = torch.rand(Co, device=cuda) > 0.5
    gmax, gmin = gmax * up, gmin * ~up                          # as the BatchNorm sign rule leaves them: one of the two is zero
    old = fused.tf32_matmul.enabled
    fused.tf32_matmul.enabled = False
    try:
        got = fused.conv_row_reduce_backward(x, W, mean.float(), imax, imin, gmean, gvar, gmax, gmin, split_row_term=True)
        ref = fused.conv_row_reduce_backward(x.double(), W.double(), mean, imax, imin, gmean.double(), gvar.double(), gmax.double(),
                                             gmin.double(), split_row_term=True)
        only_min = fused.conv_row_reduce_backward(x, W, mean.float(), imax, imin, gmean, None, None, gmin)
        only_min_ref = fused.conv_row_reduce_backward(x.double(), W.double(), mean, imax, imin, gmean.double(), None, None, gmin.double())
    finally:
        fused.tf32_matmul.enabled = old
    for name, a, b in list(zip(("gx", "gW", "row_term"), got, ref)) + list(zip(("gx(min only)", "gW(min only)"), only_min, only_min_ref)):
        err = float((a.double() - b).abs().max()) / max(float(b.abs().max()), 1e-9)
        print(f"[conv_extrema_bwd] {name}: err/scale {err:.2e}")
        assert err < 2e-5, name


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,N,k", [(3, 64, 512, 8), (2, 6, 100, 4)])
def test_edge_reduce_sel_stacked_equals_pair(cuda, B, C, N, k):
    """a and c as the two channel halves of ONE [B,2C,N] tensor (snb_edge_reduce_sel_*_stacked) == the two-tensor calls: outputs,
    statistics and gc bit for bit, ga (accumulated with shared-memory atomics in both) to rounding; the gradients land in the two
    halves of one gradient tensor."""
    from sparenet_b200 import fused
    torch.manual_seed(C + N)
    ac = torch.randn(B, 2 * C, N, device=cuda)
    idx = torch.randint(0, N, (B, N, k), device=cuda, dtype=torch.int32)
    sel = torch.rand(C, device=cuda) > 0.5
    gu, g1, g2 = torch.randn(B, C, N, device=cuda), torch.randn(B, C, device=cuda).double(), torch.randn(B, C, device=cuda).double()
    x1 = ac.clone().requires_grad_()
    o1 = fused.edge_reduce_sel_stacked(x1, idx, sel)
    torch.autograd.backward(o1, (gu, g1, g2))
    a, c = ac[:, :C].clone().requires_grad_(), ac[:, C:].clone().requires_grad_()
    o2 = fused.edge_reduce_sel(a, c, idx, sel)
    torch.autograd.backward(o2, (gu, g1, g2))
    for u, v in zip(o1, o2):
        assert torch.equal(u, v)
    assert torch.equal(x1.grad[:, C:], c.grad)                   # gc: plain stores
    assert torch.allclose(x1.grad[:, :C], a.grad, rtol=1e-5, atol=1e-5)   # ga: shared-memory atomics, summation order varies run to run


@pytest.mark.gpu
@pytest.mark.parametrize("B,K,O,bias", [(32, 4096, 4096, True), (32, 4096, 3590, True), (5, 64, 10, True), (1, 8, 3, False), (17, 520, 300, False)])
def test_small_batch_linear_matches_float64(cuda, B, K, O, bias):
    """fused.linear (csrc/linear.cu: exact-fp32 weight-streaming kernels for <= 32 rows) against the float64 product: output and the
    three gradients within fp32 summation error (and no further from it than torch's own fp32 F.linear)."""
    from sparenet_b200 import fused
    torch.manual_seed(B + K + O)
    x0, W0 = torch.randn(B, K, device=cuda), torch.randn(O, K, device=cuda) / K ** 0.5
    b0 = torch.randn(O, device=cuda) if bias else None
    gy = torch.randn(B, O, device=cuda)

    def run(fn, dt):
        leaves = [t.detach().to(dt).requires_grad_() for t in ([x0, W0] + ([b0] if bias else []))]
        y = fn(*leaves) if bias else fn(leaves[0], leaves[1], None)
        return (y,) + torch.autograd.grad(y, leaves, gy.to(dt))
    ours = run(fused.linear, torch.float32)
    ref32 = run(torch.nn.functional.linear, torch.float32)
    ref64 = run(torch.nn.functional.linear, torch.float64)
    for name, a, r32, r64 in zip(("y", "gx", "gW", "gbias"), ours, ref32, ref64):
        scale = float(r64.abs().max())
        err, err32 = float((a.double() - r64).abs().max()) / scale, float((r32.double() - r64).abs().max()) / scale
        print(f"[linear B={B} K={K} O={O}] {name}: err/scale {err:.2e} (torch fp32: {err32:.2e})")
        assert err < max(3e-6, 4 * err32), name


@pytest.mark.gpu
@pytest.mark.parametrize("G,Ci,Co,N,wbatched,xbcast", [(32, 3, 512, 2048, False, False), (8, 2, 1056, 512, True, True), (4, 256, 3, 4096, True, False),
                                                      (3, 4, 64, 1024, False, False), (3, 128, 3, 1024, False, False), (2, 7, 5, 36, False, False)])
def test_thin_conv_kernels_match_float64(cuda, G, Ci, Co, N, wbatched, xbcast):
    """fused.thin_conv on csrc/thinconv.cu (expand / reduce / wgrad streaming kernels, exact fp32) against the float64 products: output,
    data gradient and weight gradient, shared and per-batch weights, batch-broadcast input (the decoders' lattice)."""
    from sparenet_b200 import fused
    torch.manual_seed(G * 7 + Ci + Co)
    x0 = torch.randn(1 if xbcast else G, Ci, N, device=cuda)
    W0 = torch.randn(*((G, Co, Ci) if wbatched else (Co, Ci)), device=cuda)
    gy = torch.randn(G, Co, N, device=cuda)

    def ref(x, W):
        return torch.matmul(W if W.dim() == 3 else W.unsqueeze(0), x)

    outs = []
    for fn, dt in ((fused.thin_conv, torch.float32), (ref, torch.float64)):
        x, W = x0.to(dt).requires_grad_(), W0.to(dt).requires_grad_()
        y = fn(x, W)
        outs.append((y,) + torch.autograd.grad(y, (x, W), gy.to(dt)))
    for name, a, b in zip(("y", "gx", "gW"), *outs):
        assert a.shape == b.shape, (name, a.shape, b.shape)
        err = float((a.double() - b).abs().max()) / max(float(b.abs().max()), 1e-9)
        print(f"[thin_conv G={G} Ci={Ci} Co={Co} N={N}] {name}: err/scale {err:.2e}")
        assert err < 1e-5, name
