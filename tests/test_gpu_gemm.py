"""GPU tests of the tcgen05 TF32 GEMM (csrc/gemm_tc.cu, C ABI snb_gemm_tf32) through sparenet_b200.gemm / fused.

Reference = plain PyTorch in float64 on the same inputs (models/sparenet_generator.py's Conv1d/Conv2d(kernel_size=1) are exactly
these contractions).  Tolerances, written out:
  * products: TF32 operands (10-bit mantissa, truncated by the tensor core) with fp32 accumulation -> |err| <= 2^-10 * sum|a||b| * 2
    per dot product; asserted as max|err| <= 3e-3 * sqrt(K) * rms scale -- what the reference's own cuDNN TF32 convolutions give;
    against the SAME operands truncated to TF32 on the host the error must be <= 2e-5 relative (fp32 accumulation order only).
  * row statistics / extrema from the epilogue: computed from the fp32 accumulators -> <= 1e-5 relative against statistics of the
    stored output; extrema and their positions exact.
  * the prologue T(x) = leaky_relu(scale*x + shift) is fp32 fma + max: bit-exact operand, so the same bounds apply.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def tf32(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def act(x, sc, sh, slope):
    t = torch.addcmul(sh, x, sc)   # fma
    return torch.where(t > 0, t, t * slope)


@pytest.mark.parametrize("G,Cout,Cin,pos,batched", [(2, 128, 64, (256,), False), (3, 200, 72, (2, 512), True), (2, 544, 1056, (4, 512), True),
                                                    (4, 64, 128, (2048,), False), (1, 1024, 128, (4096,), False)])
def test_conv_fwd_plain(cuda, G, Cout, Cin, pos, batched):
    from sparenet_b200 import gemm
    torch.manual_seed(Cout + Cin)
    x = torch.randn(G, Cin, *pos, device=cuda)
    W = torch.randn(*((G,) if batched else ()), Cout, Cin, device=cuda) / Cin ** 0.5
    y, _ = gemm.conv_fwd(x, W)
    x3 = x.view(G, Cin, -1)
    ref = torch.matmul(W.double(), x3.double()).view_as(y)
    ref_t = torch.matmul(tf32(W).double(), tf32(x3).double()).view_as(y)
    scale = ref.abs().max().item()
    e, et = (y.double() - ref).abs().max().item(), (y.double() - ref_t).abs().max().item()
    print(f"[conv_fwd] G={G} {Cin}->{Cout} pos={pos}: err/scale vs fp64 {e / scale:.2e}, vs tf32-truncated operands {et / scale:.2e}")
    assert e <= 3e-3 * scale and et <= 2e-5 * scale


@pytest.mark.parametrize("G,Cout,Cin,pos,seg,slope", [(2, 128, 96, (512,), 512, 0.0), (2, 256, 104, (2, 512), 512, 0.2), (3, 128, 64, (1024,), 1024, 0.0)])
def test_conv_fwd_prologue_stats_minmax(cuda, G, Cout, Cin, pos, seg, slope):
    from sparenet_b200 import gemm
    torch.manual_seed(Cout * 3 + Cin)
    N = 1
    for p in pos:
        N *= p
    S = N // seg
    x = torch.randn(G, Cin, *pos, device=cuda)
    W = torch.randn(Cout, Cin, device=cuda) / Cin ** 0.5
    sc, sh = torch.rand(G, Cin, S, device=cuda) + 0.5, torch.randn(G, Cin, S, device=cuda) * 0.3
    y, st = gemm.conv_fwd(x, W, scale=sc, shift=sh, slope=slope, seg=seg, stats_seg=seg, minmax=True)
    xa = act(x.view(G, Cin, S, seg), sc.unsqueeze(-1), sh.unsqueeze(-1), slope).view(G, Cin, N)
    ref_t = torch.matmul(tf32(W).double(), tf32(xa).double())
    scale = ref_t.abs().max().item()
    # (a rare last-bit difference between the host and device fma can flip one TF32 truncation: 1e-4 instead of 2e-5)
    assert (y.view(G, Cout, N).double() - ref_t).abs().max().item() <= 1e-4 * scale
    ys = y.view(G, Cout, S, seg).double()
    assert torch.allclose(st["mean"].double(), ys.mean(-1), rtol=1e-5, atol=1e-6 * scale)
    assert torch.allclose(st["var"].double(), ys.var(-1, unbiased=False), rtol=2e-5, atol=1e-7 * scale * scale)
    yf = y.view(G, Cout, N)
    assert torch.equal(st["max"], yf.amax(-1)) and torch.equal(st["min"], yf.amin(-1))
    assert torch.equal(yf.gather(2, st["imax"].long().unsqueeze(-1)).squeeze(-1), st["max"])
    assert torch.equal(yf.gather(2, st["imin"].long().unsqueeze(-1)).squeeze(-1), st["min"])
    # the same statistics without storing the product
    y2, st2 = gemm.conv_fwd(x, W, scale=sc, shift=sh, slope=slope, seg=seg, stats_seg=seg, minmax=True, store=False)
    assert y2 is None
    for k in ("mean", "var", "max", "min", "imax", "imin"):
        assert torch.equal(st[k], st2[k]), k


@pytest.mark.parametrize("G,Cout,Cin,pos,batched", [(2, 128, 64, (256,), False), (3, 72, 160, (2, 512), True), (2, 544, 1056, (2, 512), True)])
def test_conv_dgrad_wgrad(cuda, G, Cout, Cin, pos, batched):
    from sparenet_b200 import gemm
    torch.manual_seed(Cout + 7 * Cin)
    x = torch.randn(G, Cin, *pos, device=cuda)
    gy = torch.randn(G, Cout, *pos, device=cuda)
    W = torch.randn(*((G,) if batched else ()), Cout, Cin, device=cuda) / Cin ** 0.5
    gx = gemm.conv_dgrad(gy, W)
    x3, g3 = x.view(G, Cin, -1), gy.view(G, Cout, -1)
    ref = torch.matmul(tf32(W).double().transpose(-1, -2), tf32(g3).double()).view_as(gx)
    assert (gx.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    gW = gemm.conv_wgrad(gy, x, batched=batched)
    refw = torch.matmul(tf32(g3).double(), tf32(x3).double().transpose(1, 2))
    if not batched:
        refw = refw.sum(0)
    assert gW.shape == refw.shape
    assert (gW.double() - refw).abs().max().item() <= 3e-5 * refw.abs().max().item()     # + split-K / batch reduction order


def test_conv_wgrad_prologue(cuda):
    from sparenet_b200 import gemm
    torch.manual_seed(5)
    G, Cout, Cin, B, L = 3, 136, 96, 2, 512
    x, gy = torch.randn(G, Cin, B, L, device=cuda), torch.randn(G, Cout, B, L, device=cuda)
    sc, sh = torch.rand(G, Cin, B, device=cuda) + 0.5, torch.randn(G, Cin, B, device=cuda) * 0.3
    gW = gemm.conv_wgrad(gy, x, batched=True, scale=sc, shift=sh, slope=0.0, seg=L)
    xa = act(x, sc.unsqueeze(-1), sh.unsqueeze(-1), 0.0).view(G, Cin, -1)
    refw = torch.matmul(tf32(gy.view(G, Cout, -1)).double(), tf32(xa).double().transpose(1, 2))
    assert (gW.double() - refw).abs().max().item() <= 1e-4 * refw.abs().max().item()


def test_strided_weight_views(cuda):
    """EdgeConv's W_a = W[:, :C] and PointNetRes' W4[:, 1024:] are strided views: no copy, the TMA descriptor carries the row stride."""
    from sparenet_b200 import fused
    torch.manual_seed(11)
    x = torch.randn(2, 64, 1024, device=cuda, requires_grad=True)
    Wfull = torch.randn(128, 192, device=cuda, requires_grad=True)
    y = fused.conv1x1(x, Wfull[:, 128:])
    ref = torch.matmul(tf32(Wfull[:, 128:].detach()).double(), tf32(x.detach()).double())
    assert (y.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    y.sum().backward()
    assert Wfull.grad[:, :128].abs().max().item() == 0 and Wfull.grad[:, 128:].abs().max().item() > 0


def _tail(m_bc, v_bc, rb, g, beta, w1, w2):                    # the refiner's closed form (PointNetRes._bn_se_tail)
    m = m_bc + rb
    mean = m.mean(0)
    var = v_bc.mean(0) + ((m - mean) ** 2).mean(0)
    inv = torch.rsqrt(var + 1e-5)
    scale, shift = g * inv, beta - g * inv * mean
    gate = torch.sigmoid(torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(m * scale + shift, w1)), w2))
    return gate * scale, gate * shift + rb * gate * scale


def tf32_ste(x64):
    """float64 tensor carrying the VALUE the tensor core sees (fp32 rounding of x, truncated to TF32) with a straight-through
    gradient: the float64 reference then has the same pre-activations as the device path up to fp32 accumulation order, so ReLU /
    max switches agree and what remains is the smooth TF32 error of the backward GEMMs."""
    return x64 + (tf32(x64.detach().float()).double() - x64).detach()


def _grad_close(name, g32, g64, l2=5e-3, frac=1e-4):
    d = (g32.double() - g64).abs()
    scale = g64.abs().max().item() + 1e-30
    rel_l2 = (d.norm() / (g64.norm() + 1e-30)).item()
    off = (d > 1e-2 * scale).double().mean().item()
    print(f"[grad {name}] relative L2 {rel_l2:.2e}, elements off by > 1e-2 of scale: {off:.2e}")
    assert rel_l2 < l2 and off <= frac, name


def test_act_conv_chain_matches_unfused(cuda):
    """conv -> [BN o SE o ReLU folded] -> conv with the tail applied inside the next GEMM (Prologue/ActConv) against the explicit
    composition normalise + relu + matmul in float64 autograd on TF32-truncated operands (tf32_ste).  Outputs <= 1e-4 of scale;
    gradients of the input, both weights and every tail parameter <= 5e-3 relative L2 (TF32 truncation of the backward operands),
    with at most 1e-4 of the elements off by more than 1e-2 of scale (a ReLU deciding differently on a ~1e-7 pre-activation)."""
    from sparenet_b200 import fused
    torch.manual_seed(21)
    B, C0, C1, C2, N = 4, 64, 128, 96, 1024
    h0 = torch.randn(B, C0, N, device=cuda) * 0.7 + 0.2
    W1, W2 = torch.randn(C1, C0, device=cuda) / C0 ** 0.5, torch.randn(C2, C1, device=cuda) / C1 ** 0.5

    def params(C):
        return [torch.randn(B, C, device=cuda) * 0.2, torch.rand(C, device=cuda) + 0.5, torch.randn(C, device=cuda) * 0.1,
                torch.randn(max(C // 4, 1), C, device=cuda) * 0.5, torch.randn(C, max(C // 4, 1), device=cuda) * 0.5]
    p0, p1 = params(C0), params(C1)
    l32 = [t.clone().requires_grad_() for t in [h0, W1, W2] + p0 + p1]
    l64 = [t.double().requires_grad_() for t in [h0, W1, W2] + p0 + p1]

    a, w1, w2, q0, q1 = l32[0], l32[1], l32[2], l32[3:8], l32[8:13]
    m0, v0 = fused.row_stats_nograd(a)
    pro0 = fused.Prologue(a, m0, v0, _tail, tuple(q0))
    h1, m1, v1 = fused.act_conv(w1, pro0, stats_seg=N, h=a)
    pro1 = fused.Prologue(h1, m1, v1, _tail, tuple(q1))
    h2 = fused.act_conv(w2, pro1, stats_seg=None, h=h1)

    A, Wd1, Wd2, Q0, Q1 = l64[0], l64[1], l64[2], l64[3:8], l64[8:13]

    def rna(h, prm):
        var, mean = torch.var_mean(h, dim=-1, unbiased=False)
        sc, sh = _tail(mean, var, *prm)
        return torch.relu(h * sc.unsqueeze(-1) + sh.unsqueeze(-1))
    H1 = torch.matmul(tf32_ste(Wd1), tf32_ste(rna(A, Q0)))
    H2 = torch.matmul(tf32_ste(Wd2), tf32_ste(rna(H1, Q1)))
    s = H2.abs().max().item()
    e1, e2 = (h1.double() - H1).abs().max().item() / H1.abs().max().item(), (h2.double() - H2).abs().max().item() / s
    print(f"[act_conv chain] forward err/scale: layer 1 {e1:.2e}, layer 2 {e2:.2e}")
    assert e1 <= 5e-4 and e2 <= 1e-3        # a last-bit difference of the fp32 vs fp64 activation flips a few TF32 truncations (2^-11 each)
    assert torch.allclose(m1.double(), H1.mean(-1), atol=1e-4 * H1.abs().max().item())
    assert torch.allclose(v1.double(), H1.var(-1, unbiased=False), rtol=1e-3, atol=1e-6)
    w = torch.randn_like(h2)
    (h2 * w).sum().backward()
    (H2 * w.double()).sum().backward()
    names = ["h0", "W1", "W2"] + [f"p0.{i}" for i in range(5)] + [f"p1.{i}" for i in range(5)]
    for n, x32, x64 in zip(names, l32, l64):
        _grad_close(n, x32.grad, x64.grad)


def test_act_conv_row_reduce(cuda):
    """(mean, var, max, min) of W . relu(scale*h + shift) from the GEMM epilogue, product never stored; Gram-matrix backward.  The
    float64 reference takes its extrema at OUR positions (gather), so the gradient routing of near-ties is compared like for like."""
    from sparenet_b200 import fused
    torch.manual_seed(31)
    B, Ci, Co, N = 2, 128, 1024, 2048
    h = torch.randn(B, Ci, N, device=cuda) * 0.8
    W = torch.randn(Co, Ci, 1, device=cuda) / Ci ** 0.5
    prm = [torch.randn(B, Ci, device=cuda) * 0.2, torch.rand(Ci, device=cuda) + 0.5, torch.randn(Ci, device=cuda) * 0.1,
           torch.randn(Ci // 4, Ci, device=cuda) * 0.5, torch.randn(Ci, Ci // 4, device=cuda) * 0.5]
    h32, W32, p32 = h.clone().requires_grad_(), W.clone().requires_grad_(), [t.clone().requires_grad_() for t in prm]
    h64, W64, p64 = h.double().requires_grad_(), W.double().requires_grad_(), [t.double().requires_grad_() for t in prm]
    m, v = fused.row_stats_nograd(h32)
    pro = fused.Prologue(h32, m, v, _tail, tuple(p32))
    o1 = fused.act_conv_row_reduce(W32, pro, h=h32)
    var, mean = torch.var_mean(h64, dim=-1, unbiased=False)
    sc, sh = _tail(mean, var, *p64)
    H = torch.matmul(tf32_ste(W64.view(Co, Ci)), tf32_ste(torch.relu(h64 * sc.unsqueeze(-1) + sh.unsqueeze(-1))))
    hv, hm = torch.var_mean(H, dim=-1, unbiased=False)
    # the positions our epilogue reported must hold the extrema of the float64 product up to the fp32 accumulation error
    node = o1[0].grad_fn
    imax, imin = node.saved_tensors[2].long(), node.saved_tensors[3].long()
    gx, gn = H.gather(2, imax.unsqueeze(-1)).squeeze(-1), H.gather(2, imin.unsqueeze(-1)).squeeze(-1)
    s = H.abs().max().item()
    assert (H.amax(-1) - gx).abs().max().item() <= 1e-5 * s and (gn - H.amin(-1)).abs().max().item() <= 1e-5 * s
    o2 = (hm, hv, gx, gn)
    for a, b, name in zip(o1, o2, ("mean", "var", "max", "min")):
        d = (a.double() - b).abs()
        wi = int(d.argmax())
        print(f"[act_conv_row_reduce] {name}: max |diff| {d.max().item():.3e} at flat {wi} (ours {a.reshape(-1)[wi].item():.6f}, ref {b.reshape(-1)[wi].item():.6f}), "
              f"scale {s:.3f}, finite {bool(torch.isfinite(a).all())}")
    for a, b, name in zip(o1, o2, ("mean", "var", "max", "min")):
        assert torch.allclose(a.double(), b, rtol=1e-4, atol=2e-5 * s), name
    ws = [torch.randn_like(t) for t in o1]
    sum((a * w).sum() for a, w in zip(o1, ws)).backward()
    sum((a * w.double()).sum() for a, w in zip(o2, ws)).backward()
    for n, a, b in [("h", h32, h64), ("W", W32, W64)] + [(f"p{i}", x, y) for i, (x, y) in enumerate(zip(p32, p64))]:
        _grad_close(n, a.grad, b.grad)


def test_gemm_rejects_unserved_shapes(cuda):
    from sparenet_b200 import gemm
    from sparenet_b200.functional import SnbValueError
    with pytest.raises(SnbValueError):
        gemm.conv_fwd(torch.randn(2, 64, 100, device=cuda), torch.randn(32, 64, device=cuda))      # positions % 32
    with pytest.raises(SnbValueError):
        gemm.conv_dgrad(torch.randn(2, 64, 128, device=cuda), torch.randn(64, 40, device=cuda))     # Cin % 32
    with pytest.raises(SnbValueError):
        gemm.conv_fwd(torch.randn(2, 64, 128), torch.randn(32, 64))                                 # CPU tensors


def test_bcast_act_conv_matches_materialised(cuda):
    """The decoders' first layer: x_hat [P,C,L] shared by all samples, per-sample scale/shift applied in the GEMM prologue (tiled
    operand, snb_gemm_desc.b_pos_mod) against the explicit [P,C,B,L] activation + matmul in float64 on TF32-truncated operands."""
    from sparenet_b200 import fused
    torch.manual_seed(41)
    P, C, Co, B, L = 3, 96, 136, 4, 512
    xhat = torch.randn(P, C, L, device=cuda)
    A, D = torch.rand(P, C, B, device=cuda) + 0.5, torch.randn(P, C, B, device=cuda) * 0.3
    W = torch.randn(P, Co, C, device=cuda) / C ** 0.5
    l32 = [t.clone().requires_grad_() for t in (xhat, A, D, W)]
    l64 = [t.double().requires_grad_() for t in (xhat, A, D, W)]
    y, m, v = fused.bcast_act_conv(*l32)
    X, A6, D6, W6 = l64
    xa = torch.relu(X.unsqueeze(2) * A6.unsqueeze(-1) + D6.unsqueeze(-1))               # [P,C,B,L]
    Y = torch.matmul(tf32_ste(W6), tf32_ste(xa).reshape(P, C, B * L)).view(P, Co, B, L)
    s = Y.abs().max().item()
    assert y.shape == (P, Co, B, L) and (y.double() - Y).abs().max().item() <= 1e-4 * s
    assert torch.allclose(m.double(), Y.mean(-1), atol=1e-5 * s) and torch.allclose(v.double(), Y.var(-1, unbiased=False), rtol=1e-4, atol=1e-7)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    (Y * w.double()).sum().backward()
    for n, a, b in zip(("xhat", "A", "D", "W"), l32, l64):
        _grad_close(n, a.grad, b.grad)


@pytest.mark.parametrize("G,Cout,Cin,N", [(2, 128, 64, 512), (32, 512, 256, 2048), (3, 256, 256, 96)])
def test_conv1x1_add_into_matches_separate_add(cuda, G, Cout, Cin, N):
    """acc += W x through the epilogue's TMA reduce-add (fused.conv1x1_add_into, the EdgeConv residual branch) against
    acc + conv1x1(x, W): the SAME tensor-core product, so values agree to the fp32 addition (<= 1e-6 of scale); the gradient passes
    through to acc unchanged and reaches x and W through the usual data / weight gradients."""
    from sparenet_b200 import fused
    torch.manual_seed(Cout + N)
    base = torch.randn(G, Cout, N, device=cuda)
    x0 = torch.randn(G, Cin, N, device=cuda)
    W0 = torch.randn(Cout, Cin, device=cuda) / Cin ** 0.5
    gy = torch.randn(G, Cout, N, device=cuda)
    outs = []
    for fusedp in (False, True):
        b, x, W = base.clone().requires_grad_(), x0.clone().requires_grad_(), W0.clone().requires_grad_()
        acc = b * 1.0                                   # a non-leaf, like the row_affine_act output it is used on
        y = fused.conv1x1_add_into(acc, x, W) if fusedp else acc + fused.conv1x1(x, W)
        y.backward(gy)
        outs.append((y.detach(), b.grad, x.grad, W.grad))
    for name, a, c in zip(("y", "g_acc", "g_x", "g_W"), outs[0], outs[1]):
        err = (a - c).abs().max().item() / max(a.abs().max().item(), 1e-6)
        print(f"[conv1x1_add_into G={G} {Cin}->{Cout} N={N}] {name}: err/scale {err:.2e}")
        assert err < 2e-6, name
