"""Autograd Functions over the fused generator kernels (snb_edge_reduce_*, snb_row_*): host plumbing only.

Used by sparenet_b200/dropin/models/sparenet_generator.py.  CUDA float32 only; no fallback.
"""
import torch

from . import _lib
from . import gemm
from ._lib import check, ptr, stream_ptr
from .functional import _op


class tf32_matmul:
    """TF32 for the library GEMMs inside the block (host-side dispatch flag, restored on exit).  Used only where the reference's own
    arithmetic is a TF32 convolution (the Gram-matrix form of conv3's backward) or where the product merely PRUNES candidates that
    are re-evaluated exactly (kNN); nn.Linear layers keep fp32 like the reference's."""
    enabled = True      # tests of the fp32 algebra switch this off (tf32_matmul.enabled = False): the blocks then run in fp32

    def __enter__(self):
        self.old = torch.backends.cuda.matmul.allow_tf32
        if tf32_matmul.enabled:
            torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *a):
        torch.backends.cuda.matmul.allow_tf32 = self.old
        return False


class ThinConv(torch.autograd.Function):
    """y[g] = W x[g] for the THIN 1x1 convolutions (3-8 channels on one side: xyz inputs, xyz outputs, the 2-d lattice), whose rows
    are shorter than a TMA box of the tensor-core GEMM.  W [Co, Ci] or [G, Co, Ci], x [G, Ci, N] (or [1, Ci, N] against a batched W).
    CUDA fp32 with N % 4 == 0: the exact-fp32 streaming kernels of csrc/thinconv.cu (one pass over the wide tensor per product, forward
    and both gradients).  Anything else: torch.bmm on an EXPANDED (stride-0) operand, TF32 allowed as in the reference's cuDNN."""
    USE_KERNELS = True     # measurement switch: False routes everything through the library GEMM

    @staticmethod
    def _own(x, W):
        return (ThinConv.USE_KERNELS and x.is_cuda and x.dtype == torch.float32 and W.dtype == torch.float32 and x.dim() == 3
                and x.shape[2] % 4 == 0 and x.shape[2] > 0 and min(W.shape[-2], W.shape[-1]) <= 8 and W.shape[-1] == x.shape[1])

    @staticmethod
    def forward(ctx, x, W):
        G = max(x.shape[0], W.shape[0] if W.dim() == 3 else 1)
        ctx.own = ThinConv._own(x, W)
        if ctx.own:
            x, W = x.contiguous(), W.contiguous()
            ctx.save_for_backward(x, W)
            Co, Ci = W.shape[-2], W.shape[-1]
            N = x.shape[2]
            x_bs = Ci * N if x.shape[0] == G else 0
            w_bs = Co * Ci if W.dim() == 3 else 0
            y = torch.empty(G, Co, N, device=x.device, dtype=torch.float32)
            lib = _lib.load()
            with torch.cuda.device(x.device), _op("thin_conv", 1, 4 * (G * Ci * N + y.numel())):
                if Ci <= 8:
                    check(lib.snb_thin_expand(ptr(x), x_bs, ptr(W), w_bs, Ci, 1, G, Ci, Co, N, ptr(y), stream_ptr()), "thin_expand")
                else:
                    check(lib.snb_thin_reduce(ptr(x), x_bs, ptr(W), w_bs, Ci, 1, G, Co, Ci, N, ptr(y), Co * N, stream_ptr()), "thin_reduce")
            return y
        ctx.save_for_backward(x, W)
        Wb = W.unsqueeze(0).expand(G, -1, -1) if W.dim() == 2 else W
        xb = x.expand(G, -1, -1)
        with tf32_matmul():
            return torch.bmm(Wb, xb)

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        gy = gy.contiguous()     # a transposed view (the loss hands back [B,N,3]^T) sends cuBLAS to a 4x slower kernel: 6 MB copy instead
        G = gy.shape[0]
        if ctx.own:
            Co, Ci = W.shape[-2], W.shape[-1]
            N = x.shape[2]
            x_bs = Ci * N if x.shape[0] == G else 0
            w_bs = Co * Ci if W.dim() == 3 else 0
            lib = _lib.load()
            gx = gW = None
            with torch.cuda.device(x.device):
                if ctx.needs_input_grad[0]:
                    gx = torch.empty(G, Ci, N, device=x.device, dtype=torch.float32)
                    with _op("thin_conv_bwd", 1, 4 * (gy.numel() + gx.numel())):
                        if Ci <= 8:     # thin input: gx = W^T gy reduces over the Co wide channels
                            check(lib.snb_thin_reduce(ptr(gy), Co * N, ptr(W), w_bs, 1, Ci, G, Ci, Co, N, ptr(gx), Ci * N, stream_ptr()), "thin_reduce")
                        else:           # thin output: gx = W^T gy expands the Co thin channels
                            check(lib.snb_thin_expand(ptr(gy), Co * N, ptr(W), w_bs, 1, Ci, G, Co, Ci, N, ptr(gx), stream_ptr()), "thin_expand")
                    if x.shape[0] != G:
                        gx = gx.sum(0, keepdim=True)
                if ctx.needs_input_grad[1]:
                    thin_in = Ci <= 8
                    L, S = (Co, Ci) if thin_in else (Ci, Co)
                    out = torch.empty(G, L, 8, device=x.device, dtype=torch.float32)
                    with _op("thin_conv_bwd", 1, 4 * (gy.numel() + G * Ci * N)):
                        if thin_in:
                            check(lib.snb_thin_wgrad(ptr(gy), Co * N, ptr(x), x_bs, G, S, L, N, ptr(out), stream_ptr()), "thin_wgrad")
                        else:
                            check(lib.snb_thin_wgrad(ptr(x), x_bs, ptr(gy), Co * N, G, S, L, N, ptr(out), stream_ptr()), "thin_wgrad")
                    gW = out[:, :, :S] if thin_in else out[:, :, :S].transpose(1, 2)      # [G,Co,Ci]
                    gW = gW.sum(0) if W.dim() == 2 else gW.contiguous()
            return gx, gW
        Wb = W.unsqueeze(0).expand(G, -1, -1) if W.dim() == 2 else W
        xb = x.expand(G, -1, -1)
        gx = gW = None
        with tf32_matmul():
            if ctx.needs_input_grad[0]:
                gx = torch.bmm(Wb.transpose(1, 2), gy)
                if x.shape[0] != G:
                    gx = gx.sum(0, keepdim=True)
            if ctx.needs_input_grad[1]:
                gW = torch.bmm(gy, xb.transpose(1, 2))
                if W.dim() == 2:
                    gW = gW.sum(0)
        return gx, gW


def thin_conv(x, W):
    return ThinConv.apply(x, W)


class EdgeReduce(torch.autograd.Function):
    """(a, c [B,C,N], idx [B,N,k] int32) -> umax, umin [B,C,N], S1, S2 [B,C] (float64) of u = a[idx] + c."""
    @staticmethod
    def forward(ctx, a, c, idx):
        a, c, idx = a.contiguous(), c.contiguous(), idx.contiguous()
        B, C, N = a.shape
        k = idx.shape[2]
        dev = a.device
        umax, umin = torch.empty_like(a), torch.empty_like(a)
        smax = torch.empty(B, C, N, dtype=torch.uint8, device=dev)
        smin = torch.empty(B, C, N, dtype=torch.uint8, device=dev)
        S1 = torch.empty(B, C, dtype=torch.float64, device=dev)
        S2 = torch.empty(B, C, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev), _op("edge_reduce_fwd", 1):
            check(_lib.load().snb_edge_reduce_fwd(ptr(a), ptr(c), ptr(idx), B, C, N, k, ptr(umax), ptr(umin), ptr(smax), ptr(smin), ptr(S1), ptr(S2),
                                                  stream_ptr()), "edge_reduce_fwd")
        ctx.save_for_backward(a, c, idx, smax, smin)
        return umax, umin, S1, S2

    @staticmethod
    def backward(ctx, gmax, gmin, gS1, gS2):
        a, c, idx, smax, smin = ctx.saved_tensors
        B, C, N = a.shape
        k = idx.shape[2]
        ga, gc = torch.empty_like(a), torch.empty_like(c)
        gmax = (gmax if gmax is not None else torch.zeros_like(a)).contiguous()
        gmin = (gmin if gmin is not None else torch.zeros_like(a)).contiguous()
        gS1 = (gS1 if gS1 is not None else torch.zeros(B, C, dtype=torch.float64, device=a.device)).contiguous()
        gS2 = (gS2 if gS2 is not None else torch.zeros(B, C, dtype=torch.float64, device=a.device)).contiguous()
        with torch.cuda.device(a.device), _op("edge_reduce_bwd", 1):
            check(_lib.load().snb_edge_reduce_bwd(ptr(a), ptr(c), ptr(idx), ptr(smax), ptr(smin), ptr(gmax), ptr(gmin), ptr(gS1), ptr(gS2), B, C, N, k,
                                                  ptr(ga), ptr(gc), stream_ptr()), "edge_reduce_bwd")
        return ga, gc, None


class EdgeReduceSel(torch.autograd.Function):
    """(a, c [B,C,N], idx [B,N,k] int32, sel_max [C] bool) -> ustar [B,C,N] = max_m u where sel_max else min_m u, and
    S1, S2 [B,C] (float64) of u = a[idx] + c: one extremum tensor and one slot tensor instead of two (no torch.where)."""
    @staticmethod
    def forward(ctx, a, c, idx, sel_max):
        a, c, idx = a.contiguous(), c.contiguous(), idx.contiguous()
        sel = sel_max.to(torch.uint8).contiguous()
        B, C, N = a.shape
        k = idx.shape[2]
        dev = a.device
        ustar = torch.empty_like(a)
        slot = torch.empty(B, C, N, dtype=torch.uint8, device=dev)
        S1 = torch.empty(B, C, dtype=torch.float64, device=dev)
        S2 = torch.empty(B, C, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev), _op("edge_reduce_fwd", 1):
            check(_lib.load().snb_edge_reduce_sel_fwd(ptr(a), ptr(c), ptr(idx), ptr(sel), B, C, N, k, ptr(ustar), ptr(slot), ptr(S1), ptr(S2),
                                                      stream_ptr()), "edge_reduce_sel_fwd")
        ctx.save_for_backward(a, c, idx, slot)
        return ustar, S1, S2

    @staticmethod
    def backward(ctx, gu, gS1, gS2):
        a, c, idx, slot = ctx.saved_tensors
        B, C, N = a.shape
        k = idx.shape[2]
        ga, gc = torch.empty_like(a), torch.empty_like(c)
        gu = (gu if gu is not None else torch.zeros_like(a)).contiguous()
        gS1 = (gS1 if gS1 is not None else torch.zeros(B, C, dtype=torch.float64, device=a.device)).contiguous()
        gS2 = (gS2 if gS2 is not None else torch.zeros(B, C, dtype=torch.float64, device=a.device)).contiguous()
        with torch.cuda.device(a.device), _op("edge_reduce_bwd", 1):
            check(_lib.load().snb_edge_reduce_sel_bwd(ptr(a), ptr(c), ptr(idx), ptr(slot), ptr(gu), ptr(gS1), ptr(gS2), B, C, N, k, ptr(ga), ptr(gc),
                                                      stream_ptr()), "edge_reduce_sel_bwd")
        return ga, gc, None, None


class EdgeReduceSelStacked(torch.autograd.Function):
    """EdgeReduceSel for a and c stored as the two channel halves of ONE tensor ac [B,2C,N] -- the output of a single GEMM with the
    stacked weight [W_a ; W_b - W_a].  The backward fills the halves of one [B,2C,N] gradient, so the pair's data and weight gradients
    are single GEMMs as well and the two data gradients are never added elementwise."""
    @staticmethod
    def forward(ctx, ac, idx, sel_max):
        ac, idx = ac.contiguous(), idx.contiguous()
        sel = sel_max.to(torch.uint8).contiguous()
        B, C2, N = ac.shape
        C = C2 // 2
        k = idx.shape[2]
        dev = ac.device
        ustar = torch.empty(B, C, N, device=dev, dtype=torch.float32)
        slot = torch.empty(B, C, N, dtype=torch.uint8, device=dev)
        S1 = torch.empty(B, C, dtype=torch.float64, device=dev)
        S2 = torch.empty(B, C, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev), _op("edge_reduce_fwd", 1):
            check(_lib.load().snb_edge_reduce_sel_fwd_stacked(ptr(ac), ptr(idx), ptr(sel), B, C, N, k, ptr(ustar), ptr(slot), ptr(S1), ptr(S2),
                                                              stream_ptr()), "edge_reduce_sel_fwd_stacked")
        ctx.save_for_backward(ac, idx, slot)
        return ustar, S1, S2

    @staticmethod
    def backward(ctx, gu, gS1, gS2):
        ac, idx, slot = ctx.saved_tensors
        B, C2, N = ac.shape
        C = C2 // 2
        k = idx.shape[2]
        gac = torch.empty_like(ac)
        gu = (gu if gu is not None else torch.zeros(B, C, N, device=ac.device)).contiguous()
        gS1 = (gS1 if gS1 is not None else torch.zeros(B, C, dtype=torch.float64, device=ac.device)).contiguous()
        gS2 = (gS2 if gS2 is not None else torch.zeros(B, C, dtype=torch.float64, device=ac.device)).contiguous()
        with torch.cuda.device(ac.device), _op("edge_reduce_bwd", 1):
            check(_lib.load().snb_edge_reduce_sel_bwd_stacked(ptr(ac), ptr(idx), ptr(slot), ptr(gu), ptr(gS1), ptr(gS2), B, C, N, k, ptr(gac),
                                                              stream_ptr()), "edge_reduce_sel_bwd_stacked")
        return gac, None, None


class RowStats(torch.autograd.Function):
    """h [..., L] -> (mean, biased var) over the last dim, shape h.shape[:-1]."""
    @staticmethod
    def forward(ctx, h):
        h = h.contiguous()
        L = h.shape[-1]
        R = h.numel() // L
        mean = torch.empty(h.shape[:-1], device=h.device, dtype=torch.float32)
        var = torch.empty_like(mean)
        with torch.cuda.device(h.device), _op("row_stats", 1, 4 * h.numel()):
            check(_lib.load().snb_row_stats(ptr(h), R, L, ptr(mean), ptr(var), stream_ptr()), "row_stats")
        ctx.save_for_backward(h, mean)
        return mean, var

    @staticmethod
    def backward(ctx, gmean, gvar):
        h, mean = ctx.saved_tensors
        L = h.shape[-1]
        R = h.numel() // L
        gmean = (gmean if gmean is not None else torch.zeros_like(mean)).contiguous()
        gvar = (gvar if gvar is not None else torch.zeros_like(mean)).contiguous()
        gh = torch.empty_like(h)
        with torch.cuda.device(h.device), _op("row_stats_bwd", 1, 8 * h.numel()):
            check(_lib.load().snb_row_stats_bwd(ptr(h), ptr(mean), ptr(gmean), ptr(gvar), R, L, ptr(gh), stream_ptr()), "row_stats_bwd")
        return gh


class RowAffineAct(torch.autograd.Function):
    """y[r,:] = leaky_relu(h[r // in_div, :] * scale[r] + shift[r], slope).  h [Rin, L] (any leading shape), scale/shift
    with R = Rin * in_div elements; returns y of shape out_shape (R rows of L)."""
    @staticmethod
    def forward(ctx, h, scale, shift, in_div, slope, out_shape):
        h, scale, shift = h.contiguous(), scale.contiguous().float(), shift.contiguous().float()
        L = h.shape[-1]
        R = scale.numel()
        assert h.numel() // L * in_div == R and shift.numel() == R
        y = torch.empty(out_shape, device=h.device, dtype=torch.float32)
        assert y.numel() == R * L
        with torch.cuda.device(h.device), _op("row_affine_act_fwd", 1):
            check(_lib.load().snb_row_affine_act_fwd(ptr(h), ptr(scale), ptr(shift), R, L, int(in_div), float(slope), ptr(y), stream_ptr()),
                  "row_affine_act_fwd")
        ctx.save_for_backward(h, scale, shift)
        ctx.meta = (int(in_div), float(slope))
        return y

    @staticmethod
    def backward(ctx, gy):
        h, scale, shift = ctx.saved_tensors
        in_div, slope = ctx.meta
        L = h.shape[-1]
        R = scale.numel()
        gy = gy.contiguous()
        gh = torch.empty_like(h)
        gsc, gsh = torch.empty_like(scale), torch.empty_like(shift)
        with torch.cuda.device(h.device), _op("row_affine_act_bwd", 1):
            check(_lib.load().snb_row_affine_act_bwd(ptr(gy), ptr(h), ptr(scale), ptr(shift), R, L, in_div, slope, ptr(gh), ptr(gsc), ptr(gsh),
                                                     stream_ptr()), "row_affine_act_bwd")
        return gh, gsc, gsh, None, None, None


class RowMinMax(torch.autograd.Function):
    """h [..., L] -> (max, min) over the last dim; the gradient goes to the first position attaining each."""
    @staticmethod
    def forward(ctx, h):
        h = h.contiguous()
        L = h.shape[-1]
        R = h.numel() // L
        vmax = torch.empty(h.shape[:-1], device=h.device, dtype=torch.float32)
        vmin = torch.empty_like(vmax)
        imax = torch.empty(h.shape[:-1], device=h.device, dtype=torch.int32)
        imin = torch.empty_like(imax)
        with torch.cuda.device(h.device), _op("row_minmax", 1):
            check(_lib.load().snb_row_minmax(ptr(h), R, L, ptr(vmax), ptr(vmin), ptr(imax), ptr(imin), stream_ptr()), "row_minmax")
        ctx.save_for_backward(imax, imin)
        ctx.hshape = h.shape
        return vmax, vmin

    @staticmethod
    def backward(ctx, gmax, gmin):
        imax, imin = ctx.saved_tensors
        gh = torch.zeros(ctx.hshape, device=imax.device, dtype=torch.float32)
        flat = gh.view(-1, ctx.hshape[-1])
        if gmax is not None:
            flat.scatter_add_(1, imax.view(-1, 1).long(), gmax.reshape(-1, 1))
        if gmin is not None:
            flat.scatter_add_(1, imin.view(-1, 1).long(), gmin.reshape(-1, 1))
        return gh


class RowNormAct(torch.autograd.Function):
    """y = leaky_relu(scale * h + shift, slope) with (scale, shift) = fn(row_mean(h), row_var(h), *tensors): the whole
    BatchNorm o SE o ReLU tail of a dense layer as ONE autograd node.  Forward: row_stats kernel, the small closed form `fn`
    (plain torch on [rows]-sized tensors, recorded on detached leaves), affine+act kernel.  Backward: phase A reduces
    (gscale, gshift) without touching gh, autograd pulls them through the recorded small graph to (gmean, gvar) and the
    gradients of `tensors`, phase B writes gh once -- the two gradient paths into h are summed inside the kernel instead of
    by a third full-size pass (see snb_row_act_bwd_reduce / snb_row_norm_act_bwd)."""
    @staticmethod
    def forward(ctx, h, fn, slope, stats, *tensors):
        h = h.contiguous()
        L = h.shape[-1]
        R = h.numel() // L
        lib = _lib.load()
        if stats is not None:                                   # row statistics already produced by the GEMM epilogue
            mean, var = (t.detach().reshape(h.shape[:-1]).contiguous().float() for t in stats)
        else:
            mean = torch.empty(h.shape[:-1], device=h.device, dtype=torch.float32)
            var = torch.empty_like(mean)
            with torch.cuda.device(h.device), _op("row_stats", 1, 4 * h.numel()):
                check(lib.snb_row_stats(ptr(h), R, L, ptr(mean), ptr(var), stream_ptr()), "row_stats")
        with torch.enable_grad():
            m_, v_ = mean.requires_grad_(True), var.requires_grad_(True)
            ts = [t.detach().requires_grad_(t.requires_grad) for t in tensors]
            scale, shift = fn(m_, v_, *ts)
        assert scale.shape == mean.shape and shift.shape == mean.shape, "fn must return per-row scale/shift"
        sc, sh = scale.detach().contiguous().float(), shift.detach().contiguous().float()
        y = torch.empty_like(h)
        with torch.cuda.device(h.device), _op("row_affine_act_fwd", 1, 8 * h.numel()):
            check(lib.snb_row_affine_act_fwd(ptr(h), ptr(sc), ptr(sh), R, L, 1, float(slope), ptr(y), stream_ptr()), "row_affine_act_fwd")
        ctx.save_for_backward(h, sc, sh)
        ctx.graph = (m_, v_, ts, scale, shift)
        ctx.slope = float(slope)
        return y

    @staticmethod
    def backward(ctx, gy):
        h, sc, sh = ctx.saved_tensors
        m_, v_, ts, scale, shift = ctx.graph
        L = h.shape[-1]
        R = h.numel() // L
        lib = _lib.load()
        gy = gy.contiguous()
        gsc, gsh = torch.empty_like(sc), torch.empty_like(sh)
        with torch.cuda.device(h.device), _op("row_act_bwd_reduce", 1, 8 * h.numel()):
            check(lib.snb_row_act_bwd_reduce(ptr(gy), ptr(h), ptr(sc), ptr(sh), R, L, ctx.slope, ptr(gsc), ptr(gsh), None, stream_ptr()),
                  "row_act_bwd_reduce")
        wanted = [m_, v_] + [t for t in ts if t.requires_grad]
        grads = torch.autograd.grad((scale, shift), wanted, (gsc.view_as(scale), gsh.view_as(shift)), allow_unused=True, retain_graph=True)
        gm = (grads[0] if grads[0] is not None else torch.zeros_like(sc)).contiguous().float()
        gv = (grads[1] if grads[1] is not None else torch.zeros_like(sc)).contiguous().float()
        gh = torch.empty_like(h)
        with torch.cuda.device(h.device), _op("row_norm_act_bwd", 1, 12 * h.numel()):
            check(lib.snb_row_norm_act_bwd(ptr(gy), ptr(h), ptr(sc), ptr(sh), ptr(m_.detach()), ptr(gm), ptr(gv), R, L, ctx.slope, ptr(gh),
                                           None, stream_ptr()), "row_norm_act_bwd")
        it = iter(grads[2:])
        gts = [next(it) if t.requires_grad else None for t in ts]
        return (gh, None, None, None, *gts)


def conv_row_reduce_backward(x, W, mean, imax, imin, gmean, gvar, gmax, gmin, need_x=True, need_w=True, split_row_term=False):
    """Adjoint of (x [B,Ci,N], W [Co,Ci]) -> per-row mean / biased var / max / min of h = W x WITHOUT h.
    The statistics' gradient is dense in h but affine in it, gh = a*h + b per row (a = 2 gvar/N, b = gmean/N - a*mean), so with
    h = W x:   gx = (W^T diag(a) W) x + W^T b 1^T,   gW = sum_b diag(a_b) W (x_b x_b^T) + b_b (sum_n x_b)^T
    -- Ci x Ci Gram matrices instead of two GEMMs over the [B,Co,N] tensor; the max/min gradients touch one column each.
    Device-agnostic torch (the small matrices in float64); verified against autograd in tests/.
    split_row_term: return (gx without the row-constant term W^T b 1^T, gW, that term as [B,Ci]) so that the consumer can add it on
    the fly instead of a full-size broadcast add here."""
    B, Ci, N = x.shape
    dt = x.dtype
    z = torch.zeros_like(mean)
    gmean = z if gmean is None else gmean
    gvar = z if gvar is None else gvar
    a = gvar.double() * (2.0 / N)                                  # [B,Co]
    b = gmean.double() / N - a * mean.double()
    Wd = W.double()
    WA = a.unsqueeze(-1) * Wd                                     # [B,Co,Ci]
    gx = gW = None
    if need_x:
        M = torch.matmul(Wd.t(), WA).to(dt)                       # [B,Ci,Ci] = W^T diag(a_b) W
        own = x.is_cuda and dt == torch.float32 and x.is_contiguous() and N % 32 == 0 and Ci % 32 == 0 and tf32_matmul.enabled
        if own:                                                   # one "weight" per sample on the tcgen05 GEMM (TF32 like the reference's
            gx, _ = gemm.conv_fwd(x, M.contiguous())              # cuDNN data gradient)
        else:
            with tf32_matmul():                                   # the reference's arithmetic here is cuDNN's TF32 data gradient
                gx = torch.bmm(M, x)
        row_term = torch.matmul(b, Wd).to(dt)                      # [B,Ci]
        if not split_row_term:
            gx += row_term.unsqueeze(-1)
    if need_w:
        if x.is_cuda and dt == torch.float32 and x.is_contiguous() and N % 32 == 0 and Ci % 32 == 0 and tf32_matmul.enabled:
            G = gemm.conv_wgrad(x, x, batched=True).double()      # [B,Ci,Ci] = x x^T: the weight-gradient arrangement of the tcgen05 GEMM
        else:
            with tf32_matmul():                                   # ... and its TF32 weight gradient
                G = torch.bmm(x, x.transpose(1, 2)).double()      # [B,Ci,Ci]
        gW = torch.bmm(WA, G).sum(0) + torch.matmul(b.t(), x.sum(2).double())
    if x.is_cuda and dt == torch.float32 and (gmax is not None or gmin is not None) and (need_x or need_w):
        # both extrema in one launch (csrc/rowops.cu: snb_conv_extrema_bwd), accumulating into gx and the fp32 weight gradient
        gW32 = gW.to(dt) if need_w else None
        gmx, gmn = (None if g is None else g.contiguous().float() for g in (gmax, gmin))
        with torch.cuda.device(x.device), _op("conv_extrema_bwd", 1):
            check(_lib.load().snb_conv_extrema_bwd(ptr(x), ptr(W.contiguous()), ptr(imax.contiguous()), ptr(imin.contiguous()), ptr(gmx), ptr(gmn),
                                                   B, Ci, W.shape[0], N, ptr(gx) if need_x else None, ptr(gW32), stream_ptr()), "conv_extrema_bwd")
        gW = gW32
        gmax = gmin = None
    for g, idx in ((gmax, imax), (gmin, imin)):
        if g is None:
            continue
        col = idx.long().unsqueeze(1).expand(B, Ci, -1)           # column of x each (sample, out-channel) extremum sits in
        if need_x:
            gx.scatter_add_(2, col, (g.unsqueeze(-1) * W).transpose(1, 2))
        if need_w:
            gW += torch.einsum("bc,bic->ci", g.double(), x.gather(2, col).double())
    if split_row_term:
        return gx, (gW.to(dt) if gW is not None else None), (row_term if need_x else None)
    return gx, (gW.to(dt) if gW is not None else None)


class ConvRowReduce(torch.autograd.Function):
    """x [B,Ci,N], W [Co,Ci] -> (mean, biased var, max, min) over N of h = W x, each [B,Co].  h is produced by the library
    1x1 convolution, reduced by ONE pass of snb_row_stats_minmax and dropped; the backward is conv_row_reduce_backward."""
    @staticmethod
    def forward(ctx, x, W):
        x = x.contiguous()
        W2 = W.reshape(W.size(0), -1)
        h = torch.nn.functional.conv1d(x, W2.unsqueeze(-1))
        B, Co, N = h.shape
        dev = h.device
        mean, var, vmax, vmin = (torch.empty(B, Co, device=dev, dtype=torch.float32) for _ in range(4))
        imax, imin = (torch.empty(B, Co, device=dev, dtype=torch.int32) for _ in range(2))
        with torch.cuda.device(dev), _op("row_stats_minmax", 1, 4 * h.numel()):
            check(_lib.load().snb_row_stats_minmax(ptr(h), B * Co, N, ptr(mean), ptr(var), ptr(vmax), ptr(vmin), ptr(imax), ptr(imin), stream_ptr()),
                  "row_stats_minmax")
        ctx.save_for_backward(x, W2, mean, imax, imin)
        ctx.wshape = W.shape
        return mean, var, vmax, vmin

    @staticmethod
    def backward(ctx, gmean, gvar, gmax, gmin):
        x, W2, mean, imax, imin = ctx.saved_tensors
        gx, gW = conv_row_reduce_backward(x, W2, mean, imax, imin, gmean, gvar, gmax, gmin, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return gx, (gW.view(ctx.wshape) if gW is not None else None)


class RowNormActPool(torch.autograd.Function):
    """(max over the row, mean over the row) of leaky_relu(scale*h + shift) with (scale, shift) = fn(row statistics, *tensors): the
    encoder's pooled output WITHOUT storing the activated tensor.  Same two-phase backward as RowNormAct, with the upstream gradient
    being (gmean/L everywhere) + (gmax at the arg-max position) instead of a full tensor."""
    @staticmethod
    def forward(ctx, h, fn, slope, stats, *tensors):
        h = h.contiguous()
        L = h.shape[-1]
        R = h.numel() // L
        lib = _lib.load()
        if stats is not None:
            mean, var = (t.detach().reshape(h.shape[:-1]).contiguous().float() for t in stats)
        else:
            mean, var = row_stats_nograd(h)
        with torch.enable_grad():
            m_, v_ = mean.requires_grad_(True), var.requires_grad_(True)
            ts = [t.detach().requires_grad_(t.requires_grad) for t in tensors]
            scale, shift = fn(m_, v_, *ts)
        assert scale.shape == mean.shape and shift.shape == mean.shape, "fn must return per-row scale/shift"
        sc, sh = scale.detach().contiguous().float(), shift.detach().contiguous().float()
        vmax, vmean = torch.empty_like(sc), torch.empty_like(sc)
        imax = torch.empty(sc.shape, dtype=torch.int32, device=h.device)
        with torch.cuda.device(h.device), _op("row_act_pool_fwd", 1, 4 * h.numel()):
            check(lib.snb_row_act_pool_fwd(ptr(h), ptr(sc), ptr(sh), R, L, float(slope), ptr(vmax), ptr(imax), ptr(vmean), stream_ptr()), "row_act_pool_fwd")
        ctx.save_for_backward(h, sc, sh, imax)
        ctx.graph = (m_, v_, ts, scale, shift)
        ctx.slope = float(slope)
        return vmax, vmean

    @staticmethod
    def backward(ctx, gmax, gmean):
        h, sc, sh, imax = ctx.saved_tensors
        m_, v_, ts, scale, shift = ctx.graph
        L = h.shape[-1]
        R = h.numel() // L
        lib = _lib.load()
        gmax = (gmax if gmax is not None else torch.zeros_like(sc)).contiguous().float()
        gmean = (gmean if gmean is not None else torch.zeros_like(sc)).contiguous().float()
        gsc, gsh = torch.empty_like(sc), torch.empty_like(sh)
        with torch.cuda.device(h.device), _op("row_act_pool_bwd_reduce", 1, 4 * h.numel()):
            check(lib.snb_row_act_pool_bwd_reduce(ptr(h), ptr(sc), ptr(sh), ptr(gmax), ptr(gmean), ptr(imax), R, L, ctx.slope, ptr(gsc), ptr(gsh),
                                                  stream_ptr()), "row_act_pool_bwd_reduce")
        wanted = [m_, v_] + [t for t in ts if t.requires_grad]
        grads = torch.autograd.grad((scale, shift), wanted, (gsc.view_as(scale), gsh.view_as(shift)), allow_unused=True, retain_graph=True)
        gm = (grads[0] if grads[0] is not None else torch.zeros_like(sc)).contiguous().float()
        gv = (grads[1] if grads[1] is not None else torch.zeros_like(sc)).contiguous().float()
        gh = torch.empty_like(h)
        with torch.cuda.device(h.device), _op("row_act_pool_bwd", 1, 8 * h.numel()):
            check(lib.snb_row_act_pool_bwd(ptr(h), ptr(sc), ptr(sh), ptr(m_.detach()), ptr(gmax), ptr(gmean), ptr(imax), ptr(gm), ptr(gv), R, L,
                                           ctx.slope, ptr(gh), stream_ptr()), "row_act_pool_bwd")
        it = iter(grads[2:])
        gts = [next(it) if t.requires_grad else None for t in ts]
        return (gh, None, None, None, *gts)


def row_norm_act_pool(h, fn, tensors, slope=0.0, stats=None):
    return RowNormActPool.apply(h, fn, slope, stats, *tensors)


def edge_reduce(a, c, idx):
    return EdgeReduce.apply(a, c, idx)


def edge_reduce_sel_stacked(ac, idx, sel_max):
    return EdgeReduceSelStacked.apply(ac, idx, sel_max)


def edge_reduce_sel(a, c, idx, sel_max):
    return EdgeReduceSel.apply(a, c, idx, sel_max)


def row_norm_act(h, fn, tensors, slope=0.0, stats=None):
    return RowNormAct.apply(h, fn, slope, stats, *tensors)


def row_stats_nograd(h):
    """(mean, biased var) over the last dim WITHOUT an autograd edge: for consumers that own the statistics' gradient
    (Prologue / RowNormAct fold it into their single backward pass over h)."""
    h = h.detach().contiguous()
    L = h.shape[-1]
    mean = torch.empty(h.shape[:-1], device=h.device, dtype=torch.float32)
    var = torch.empty_like(mean)
    with torch.cuda.device(h.device), _op("row_stats", 1, 4 * h.numel()):
        check(_lib.load().snb_row_stats(ptr(h), h.numel() // L, L, ptr(mean), ptr(var), stream_ptr()), "row_stats")
    return mean, var


# ---------------------------------------------------------------------------- tensor-core 1x1 convolutions (csrc/gemm_tc.cu)
class Conv1x1(torch.autograd.Function):
    """y = W x on [G, Cin, *pos] (W [Cout, Cin] or one per batch entry [G, Cout, Cin]) through snb_gemm_tf32: forward, data
    gradient and weight gradient are the three operand arrangements of the same tcgen05 kernel.  With stats_seg the epilogue
    also returns the mean / biased variance of every output row segment, as NON-differentiable tensors: their consumer
    (Prologue or row_norm_act(stats=...)) owns that gradient."""
    @staticmethod
    def forward(ctx, x, W, stats_seg):
        x = x.contiguous()
        y, st = gemm.conv_fwd(x, W, stats_seg=stats_seg)
        ctx.save_for_backward(x, W)
        if stats_seg is None:
            return y
        mean, var = st["mean"].reshape(y.shape[:-1]), st["var"].reshape(y.shape[:-1])
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, gy, *_):
        x, W = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gemm.conv_dgrad(gy, W) if ctx.needs_input_grad[0] else None
        gW = gemm.conv_wgrad(gy, x, batched=W.dim() == 3) if ctx.needs_input_grad[1] else None
        return gx, gW, None


class SmallBatchLinear(torch.autograd.Function):
    """y = x W^T + bias for x [B <= 32, K] in exact fp32 (csrc/linear.cu: snb_linear_fwd / _dgrad / _wgrad) -- the three fully connected
    layers of the encoder -> decoder bridge (models/sparenet_generator.py:85-120, 289-330), which the reference runs as fp32 cuBLAS GEMMs.
    With 32 rows the product is a stream over the 64 MB weight; the library's tiles reach a tenth of that rate."""
    @staticmethod
    def forward(ctx, x, W, bias):
        x, W = x.contiguous(), W.contiguous()
        B, K = x.shape
        O = W.shape[0]
        lib = _lib.load()
        y = torch.empty(B, O, device=x.device, dtype=torch.float32)
        ws = torch.empty(max(lib.snb_linear_workspace_floats(B, K, O), 4), device=x.device, dtype=torch.float32)
        b_ = None if bias is None else bias.detach().contiguous()
        with torch.cuda.device(x.device), _op("linear_fwd", 2):
            check(lib.snb_linear_fwd(ptr(x), ptr(W), ptr(b_), B, K, O, ptr(y), ptr(ws), stream_ptr()), "linear_fwd")
        ctx.save_for_backward(x, W)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        gy = gy.contiguous()
        B, K = x.shape
        O = W.shape[0]
        lib = _lib.load()
        gx = gW = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            ws = torch.empty(max(lib.snb_linear_workspace_floats(B, K, O), 4), device=x.device, dtype=torch.float32)
            with torch.cuda.device(x.device), _op("linear_dgrad", 2):
                check(lib.snb_linear_dgrad(ptr(gy), ptr(W), B, K, O, ptr(gx), ptr(ws), stream_ptr()), "linear_dgrad")
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gW = torch.empty_like(W)
            gb = torch.empty(O, device=x.device, dtype=torch.float32) if ctx.has_bias else None
            with torch.cuda.device(x.device), _op("linear_wgrad", 1):
                check(lib.snb_linear_wgrad(ptr(gy), ptr(x), B, K, O, ptr(gW), ptr(gb), stream_ptr()), "linear_wgrad")
        return gx, gW, gb


def linear(x, W, bias=None):
    """nn.Linear's arithmetic (fp32) for a small batch; shapes the kernels do not serve (more than 32 rows, K % 4, non-CUDA / non-fp32
    tensors) go to torch.nn.functional.linear, explicitly."""
    if x.is_cuda and x.dtype == torch.float32 and W.dtype == torch.float32 and x.dim() == 2 and x.shape[0] <= 32 and x.shape[1] % 4 == 0:
        return SmallBatchLinear.apply(x, W, bias)
    return torch.nn.functional.linear(x, W, bias)


class CatConv1x1(torch.autograd.Function):
    """y = W . cat(xs, dim=1) with the data gradient computed PER INPUT (gx_i = W[:, slice_i]^T gy, one GEMM each, written straight
    into its own contiguous tensor): autograd's cat backward hands out channel slices of one [G, sum C_i, N] gradient, which every
    consumer then had to copy to make contiguous (0.39 ms per step for the encoder's conv5 over [x1|x2|x3|x4],
    models/sparenet_generator.py:234-236).  Forward and weight gradient run on the concatenated operand as before."""
    @staticmethod
    def forward(ctx, W, stats_seg, *xs):
        xcat = torch.cat(xs, dim=1)
        y, st = gemm.conv_fwd(xcat, W, stats_seg=stats_seg)
        ctx.save_for_backward(xcat, W)
        ctx.splits = [int(x.shape[1]) for x in xs]
        if stats_seg is None:
            return y
        mean, var = st["mean"].reshape(y.shape[:-1]), st["var"].reshape(y.shape[:-1])
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, gy, *_):
        xcat, W = ctx.saved_tensors
        gy = gy.contiguous()
        gW = gemm.conv_wgrad(gy, xcat, batched=False) if ctx.needs_input_grad[0] else None
        gxs, off = [], 0
        for i, c in enumerate(ctx.splits):
            gxs.append(gemm.conv_dgrad(gy, W[:, off:off + c]) if ctx.needs_input_grad[2 + i] else None)
            off += c
        return (gW, None, *gxs)


def cat_conv1x1(xs, W, stats_seg=None):
    """W [Cout, sum C_i] shared by the batch; every C_i a multiple of 32."""
    return CatConv1x1.apply(W, stats_seg, *xs)


class Conv1x1AddInto(torch.autograd.Function):
    """acc += W x, IN PLACE: the GEMM's epilogue reduce-adds its tiles into `acc` (TMA cp.reduce), so a residual branch costs neither
    a second [G,Cout,N] tensor nor an elementwise add pass (EdgeConvResFeat's `x + resconv(x_prev)`, models/sparenet_generator.py:
    196-232).  `acc` must be a non-leaf whose producer does not need its own output in backward (row_affine_act does not)."""
    @staticmethod
    def forward(ctx, acc, x, W):
        x = x.contiguous()
        gemm.conv_fwd(x, W, out=acc)
        ctx.mark_dirty(acc)
        ctx.save_for_backward(x, W)
        return acc

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        g = g.contiguous()
        gx = gemm.conv_dgrad(g, W) if ctx.needs_input_grad[1] else None
        gW = gemm.conv_wgrad(g, x, batched=W.dim() == 3) if ctx.needs_input_grad[2] else None
        return g, gx, gW


def conv1x1_add_into(acc, x, W):
    return Conv1x1AddInto.apply(acc, x, W)


class BnSeTail(torch.autograd.Function):
    """(scale, shift) [B,C] of the folded BatchNorm1d . SELayer1D . ReLU tail from row statistics -- ONE launch per direction
    (csrc/tails.cu) instead of ~37 + ~48 microsecond-sized PyTorch launches.  m_bc, v_bc [B,C]: row mean / biased row variance of
    the pre-activation h [B,C,L]; rb: conv bias [C] or per-sample bias [B,C] (only shifts the statistics and folds into the shift);
    bn: the nn.BatchNorm1d whose running statistics are advanced IN the kernel in train mode; w1 [H,C], w2 [C,H]: the SE layer."""
    @staticmethod
    def forward(ctx, m_bc, v_bc, rb, g, beta, w1, w2, bn, L):
        m_bc, v_bc = m_bc.contiguous().float(), v_bc.contiguous().float()
        B, C = m_bc.shape
        H = w1.shape[0]
        g, beta, w1, w2 = (t.detach().contiguous().float() for t in (g, beta, w1, w2))
        rb = None if rb is None else rb.detach().contiguous().float()
        assert (rb is None or rb.shape in ((C,), (B, C))) and w1.shape == (H, C) and w2.shape == (C, H)
        dev = m_bc.device
        lib = _lib.load()
        S, T = torch.empty(B, C, device=dev), torch.empty(B, C, device=dev)
        save = torch.empty(lib.snb_bn_se_tail_save_floats(B, C, H), device=dev)
        training = bool(bn.training)
        track = training and bn.track_running_stats
        count = B * L
        if bn.momentum is None and training and bn.track_running_stats:
            raise NotImplementedError("momentum=None (cumulative moving average) is not served by the fused tail kernels")
        mom = bn.momentum if bn.momentum is not None else 0.1
        rm, rv, nbt = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if (track or not training) else (None, None, None)
        with torch.cuda.device(dev), _op("bn_se_tail_fwd", 1):
            check(lib.snb_bn_se_tail_fwd(ptr(m_bc), ptr(v_bc), ptr(rb), int(rb is not None and rb.dim() == 2), ptr(g), ptr(beta), ptr(w1), ptr(w2), B, C, H, float(bn.eps),
                                         int(training), float(mom), float(count / max(count - 1, 1)), ptr(rm), ptr(rv),
                                         ptr(nbt) if track else None, ptr(S), ptr(T), ptr(save), stream_ptr()), "bn_se_tail_fwd")
        ctx.save_for_backward(m_bc, g, w1, w2, save, *(() if rb is None else (rb,)))
        ctx.dims = (B, C, H, training)
        return S, T

    @staticmethod
    def backward(ctx, gS, gT):
        m_bc, g, w1, w2, save, *rest = ctx.saved_tensors
        rb = rest[0] if rest else None
        B, C, H, training = ctx.dims
        dev = m_bc.device
        lib = _lib.load()
        gS, gT = gS.contiguous().float(), gT.contiguous().float()
        gm, gv, grb = torch.empty_like(m_bc), torch.empty_like(m_bc), (None if rb is None else torch.empty_like(rb))
        gg, gbeta, gw1, gw2 = torch.empty_like(g), torch.empty_like(g), torch.empty_like(w1), torch.empty_like(w2)
        scratch = torch.empty(lib.snb_bn_se_tail_scratch_floats(B, C, H), device=dev)
        with torch.cuda.device(dev), _op("bn_se_tail_bwd", 1):
            check(lib.snb_bn_se_tail_bwd(ptr(gS), ptr(gT), ptr(m_bc), ptr(rb), int(rb is not None and rb.dim() == 2), ptr(g), ptr(w1), ptr(w2), B, C, H, int(training),
                                         ptr(save), ptr(scratch), ptr(gm), ptr(gv), ptr(grb), ptr(gg), ptr(gbeta), ptr(gw1), ptr(gw2), stream_ptr()),
                  "bn_se_tail_bwd")
        return gm, gv, grb, gg, gbeta, gw1, gw2, None, None


def bn_se_tail(m_bc, v_bc, rb, g, beta, w1, w2, bn, L):
    return BnSeTail.apply(m_bc, v_bc, rb, g, beta, w1, w2, bn, L)


class BnMaxTail(torch.autograd.Function):
    """glob [B,C] = max over the points of BatchNorm1d(h + bias) from the row statistics and extrema of h alone (PointNetRes conv3 ->
    bn3 -> max, models/sparenet_generator.py:626-629) -- ONE launch per direction (csrc/tails.cu: snb_bn_max_tail_*) instead of ~25 +
    ~45 PyTorch launches.  m_bc, v_bc, hmax, hmin [B,C]: row mean / biased row variance / max / min of h [B,C,L] (without the conv
    bias); bn: the nn.BatchNorm1d (running statistics advanced in the kernel in train mode)."""
    @staticmethod
    def forward(ctx, m_bc, v_bc, hmax, hmin, bias, g, beta, bn, L):
        m_bc, v_bc, hmax, hmin = (t.contiguous().float() for t in (m_bc, v_bc, hmax, hmin))
        B, C = m_bc.shape
        bias_, g_, beta_ = (t.detach().contiguous().float() for t in (bias, g, beta))
        dev = m_bc.device
        lib = _lib.load()
        glob = torch.empty(B, C, device=dev)
        save = torch.empty(2 * C, device=dev)
        training = bool(bn.training)
        track = training and bn.track_running_stats
        count = B * L
        if bn.momentum is None and training and bn.track_running_stats:
            raise NotImplementedError("momentum=None (cumulative moving average) is not served by the fused tail kernels")
        mom = bn.momentum if bn.momentum is not None else 0.1
        rm, rv, nbt = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if (track or not training) else (None, None, None)
        with torch.cuda.device(dev), _op("bn_max_tail_fwd", 1):
            check(lib.snb_bn_max_tail_fwd(ptr(m_bc), ptr(v_bc), ptr(hmax), ptr(hmin), ptr(bias_), ptr(g_), ptr(beta_), B, C, float(bn.eps),
                                          int(training), float(mom), float(count / max(count - 1, 1)), ptr(rm), ptr(rv),
                                          ptr(nbt) if track else None, ptr(glob), ptr(save), stream_ptr()), "bn_max_tail_fwd")
        ctx.save_for_backward(m_bc, hmax, hmin, bias_, g_, save)
        ctx.training = training
        return glob

    @staticmethod
    def backward(ctx, gglob):
        m_bc, hmax, hmin, bias_, g_, save = ctx.saved_tensors
        B, C = m_bc.shape
        gglob = gglob.contiguous().float()
        gm, gv, gmax, gmin = (torch.empty_like(m_bc) for _ in range(4))
        gg, gbeta, gbias = (torch.empty_like(g_) for _ in range(3))
        with torch.cuda.device(m_bc.device), _op("bn_max_tail_bwd", 1):
            check(_lib.load().snb_bn_max_tail_bwd(ptr(gglob), ptr(m_bc), ptr(hmax), ptr(hmin), ptr(bias_), ptr(g_), ptr(save), B, C,
                                                  int(ctx.training), ptr(gm), ptr(gv), ptr(gmax), ptr(gmin), ptr(gg), ptr(gbeta), ptr(gbias),
                                                  stream_ptr()), "bn_max_tail_bwd")
        return gm, gv, gmax, gmin, gbias, gg, gbeta, None, None


def bn_max_tail(m_bc, v_bc, hmax, hmin, bias, g, beta, bn, L):
    return BnMaxTail.apply(m_bc, v_bc, hmax, hmin, bias, g, beta, bn, L)


class AdainTail(torch.autograd.Function):
    """(scale, shift) [P,Cpad,B] of the decoders' folded instance-norm . AdaIN . BatchNorm . SE . ReLU tail from the rows' statistics
    -- ONE launch per direction for all P primitives (csrc/tails.cu: snb_adain_tail_*), train-mode BatchNorm.  mean, var [P,Cpad,B];
    wsty, bsty [B,C] (the sample's AdaIN weight / bias); gam, bet [P,C(,1)]; w1 [P,H,C], w2 [P,C,H].  Also returns the BatchNorm batch
    statistics (mu, q) [P,C], non-differentiable, for the caller's running-statistics bookkeeping."""
    @staticmethod
    def forward(ctx, mean, var, wsty, bsty, gam, bet, w1, w2, eps):
        mean, var = mean.contiguous().float(), var.contiguous().float()
        P, Cp, B = mean.shape
        C, H = wsty.shape[1], w1.shape[1]
        wsty, bsty, w1, w2 = (t.detach().contiguous().float() for t in (wsty, bsty, w1, w2))
        gshape = gam.shape
        gam2, bet2 = gam.detach().reshape(P, C).contiguous().float(), bet.detach().reshape(P, C).contiguous().float()
        assert bsty.shape == (B, C) and w1.shape == (P, H, C) and w2.shape == (P, C, H) and Cp >= C and B <= 32
        dev = mean.device
        lib = _lib.load()
        sc, sh = torch.empty(P, Cp, B, device=dev), torch.empty(P, Cp, B, device=dev)
        mu, q = torch.empty(P, C, device=dev), torch.empty(P, C, device=dev)
        save = torch.empty(lib.snb_adain_tail_save_floats(P, C, B, H), device=dev)
        with torch.cuda.device(dev), _op("adain_tail_fwd", 1):
            check(lib.snb_adain_tail_fwd(ptr(mean), ptr(var), ptr(wsty), ptr(bsty), ptr(gam2), ptr(bet2), ptr(w1), ptr(w2), P, C, Cp, B, H, float(eps),
                                         ptr(sc), ptr(sh), ptr(save), ptr(mu), ptr(q), stream_ptr()), "adain_tail_fwd")
        ctx.save_for_backward(mean, var, wsty, bsty, gam2, w1, w2, save)
        ctx.meta = (P, C, Cp, B, H, float(eps), gshape)
        ctx.mark_non_differentiable(mu, q)
        return sc, sh, mu, q

    @staticmethod
    def backward(ctx, gsc, gsh, *_):
        mean, var, wsty, bsty, gam2, w1, w2, save = ctx.saved_tensors
        P, C, Cp, B, H, eps, gshape = ctx.meta
        dev = mean.device
        lib = _lib.load()
        gsc, gsh = gsc.contiguous().float(), gsh.contiguous().float()
        gmean, gvar = torch.empty_like(mean), torch.empty_like(var)
        gws, gbs = torch.empty(P, B, C, device=dev), torch.empty(P, B, C, device=dev)
        ggam, gbet = torch.empty(P, C, device=dev), torch.empty(P, C, device=dev)
        gw1, gw2 = torch.empty_like(w1), torch.empty_like(w2)
        scratch = torch.empty(lib.snb_adain_tail_scratch_floats(P, C, B, H), device=dev)
        with torch.cuda.device(dev), _op("adain_tail_bwd", 1):
            check(lib.snb_adain_tail_bwd(ptr(gsc), ptr(gsh), ptr(mean), ptr(var), ptr(wsty), ptr(bsty), ptr(gam2), ptr(w1), ptr(w2), P, C, Cp, B, H, eps,
                                         ptr(save), ptr(scratch), ptr(gmean), ptr(gvar), ptr(gws), ptr(gbs), ptr(ggam), ptr(gbet), ptr(gw1), ptr(gw2),
                                         stream_ptr()), "adain_tail_bwd")
        return gmean, gvar, gws.sum(0), gbs.sum(0), ggam.view(gshape), gbet.view(gshape), gw1, gw2, None


def adain_tail(mean, var, wsty, bsty, gam, bet, w1, w2, eps):
    return AdainTail.apply(mean, var, wsty, bsty, gam, bet, w1, w2, eps)


class Prologue:
    """The folded normalisation tail of a dense layer, y = leaky_relu(scale*h + shift, slope) with
    (scale, shift) = fn(row_mean(h), row_var(h), *tensors), held as DATA instead of being applied: the next layer's GEMM applies it
    to its operand in shared memory (ActConv), so y never exists in HBM.  fn runs once, here (it may advance BatchNorm running
    statistics), on detached leaves; every consumer pulls its gradient through the recorded small graph in `backward`."""

    def __init__(self, h, mean, var, fn, tensors, slope=0.0):
        self.h = h.detach().contiguous()
        self.tensors = tuple(tensors)
        self.slope = float(slope)
        rows = self.h.shape[:-1]
        with torch.enable_grad():
            m_ = mean.detach().reshape(rows).contiguous().float().requires_grad_(True)
            v_ = var.detach().reshape(rows).contiguous().float().requires_grad_(True)
            ts = [t.detach().requires_grad_(t.requires_grad) for t in self.tensors]
            scale, shift = fn(m_, v_, *ts)
        assert scale.shape == rows and shift.shape == rows, "fn must return per-row scale/shift"
        self.sc, self.sh = scale.detach().contiguous().float(), shift.detach().contiguous().float()
        self.graph = (m_, v_, ts, scale, shift)

    def backward(self, gy, gy_row=None):
        """gy (+ gy_row [rows], a term constant along each row, added on the fly) = gradient w.r.t. the activated tensor ->
        (gradient w.r.t. h, gradients of `tensors`): the two-phase row backward of RowNormAct (reduce, small graph, one write of gh)."""
        h, sc, sh = self.h, self.sc, self.sh
        m_, v_, ts, scale, shift = self.graph
        L = h.shape[-1]
        R = h.numel() // L
        lib = _lib.load()
        gy = gy.contiguous()
        if gy_row is not None:
            gy_row = gy_row.reshape(-1).contiguous().float()
            assert gy_row.numel() == R
        gsc, gsh = torch.empty_like(sc), torch.empty_like(sh)
        with torch.cuda.device(h.device), _op("row_act_bwd_reduce", 1, 8 * h.numel()):
            check(lib.snb_row_act_bwd_reduce(ptr(gy), ptr(h), ptr(sc), ptr(sh), R, L, self.slope, ptr(gsc), ptr(gsh), ptr(gy_row), stream_ptr()),
                  "row_act_bwd_reduce")
        wanted = [m_, v_] + [t for t in ts if t.requires_grad]
        grads = torch.autograd.grad((scale, shift), wanted, (gsc.view_as(scale), gsh.view_as(shift)), allow_unused=True, retain_graph=True)
        gm = (grads[0] if grads[0] is not None else torch.zeros_like(sc)).contiguous().float()
        gv = (grads[1] if grads[1] is not None else torch.zeros_like(sc)).contiguous().float()
        gh = torch.empty_like(h)
        with torch.cuda.device(h.device), _op("row_norm_act_bwd", 1, 12 * h.numel()):
            check(lib.snb_row_norm_act_bwd(ptr(gy), ptr(h), ptr(sc), ptr(sh), ptr(m_.detach()), ptr(gm), ptr(gv), R, L, self.slope, ptr(gh),
                                           ptr(gy_row), stream_ptr()), "row_norm_act_bwd")
        it = iter(grads[2:])
        return gh, [next(it) if t.requires_grad else None for t in ts]

    def materialise(self):
        """The activated tensor itself (backward of the layers that need it as a plain operand)."""
        h = self.h
        L = h.shape[-1]
        y = torch.empty_like(h)
        with torch.cuda.device(h.device), _op("row_affine_act_fwd", 1, 8 * h.numel()):
            check(_lib.load().snb_row_affine_act_fwd(ptr(h), ptr(self.sc), ptr(self.sh), h.numel() // L, L, 1, self.slope, ptr(y), stream_ptr()),
                  "row_affine_act_fwd")
        return y


class ActConv(torch.autograd.Function):
    """y = W . leaky_relu(scale*h + shift) with the activation applied to the GEMM operand in shared memory (`pro` is the
    Prologue of h).  Backward: data gradient (tcgen05 DGRAD), weight gradient with the activation re-applied on the fly (WGRAD
    prologue: the activated tensor is never stored), then the Prologue's row backward.  `h` and `tensors` are passed so that
    autograd routes their gradients; the statistics outputs are non-differentiable (see Conv1x1)."""
    @staticmethod
    def forward(ctx, h, W, pro, stats_seg, *tensors):
        L = pro.h.shape[-1]
        y, st = gemm.conv_fwd(pro.h, W, scale=pro.sc, shift=pro.sh, slope=pro.slope, seg=L, stats_seg=stats_seg)
        ctx.pro = pro
        ctx.save_for_backward(W)
        if stats_seg is None:
            return y
        mean, var = st["mean"].reshape(y.shape[:-1]), st["var"].reshape(y.shape[:-1])
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, gy, *_):
        (W,) = ctx.saved_tensors
        pro = ctx.pro
        gy = gy.contiguous()
        L = pro.h.shape[-1]
        gW = gemm.conv_wgrad(gy, pro.h, batched=W.dim() == 3, scale=pro.sc, shift=pro.sh, slope=pro.slope, seg=L) if ctx.needs_input_grad[1] else None
        gh, gts = pro.backward(gemm.conv_dgrad(gy, W))
        return (gh, gW, None, None, *gts)


class ActConvRowReduce(torch.autograd.Function):
    """(mean, biased var, max, min) over the positions of W . leaky_relu(scale*h + shift), each [G, Cout], WITHOUT storing the
    product: statistics and extrema come out of the GEMM epilogue (PointNetRes conv3 -> bn3 -> max over points,
    models/sparenet_generator.py:626-629).  Backward: conv_row_reduce_backward (Gram matrices) on the activated operand.
    MATERIALISE (default): the activated operand is written once in the forward and kept for the backward (which needs it anyway), and
    the GEMM runs WITHOUT a prologue -- with Cout / 128 = 8 row tiles per operand tile and only 4 k-blocks, the in-shared-memory
    transform was repeated 8 times per tile and dominated (524 -> ~360 us per call, and the backward's re-materialisation goes)."""
    MATERIALISE = True

    @staticmethod
    def forward(ctx, h, W, pro, *tensors):
        W2 = W.reshape(W.size(0), -1)
        N = pro.h.shape[-1]
        if ActConvRowReduce.MATERIALISE:
            x = pro.materialise()
            _, st = gemm.conv_fwd(x, W2, stats_seg=N, minmax=True, store=False)
            ctx.x = x
        else:
            _, st = gemm.conv_fwd(pro.h, W2, scale=pro.sc, shift=pro.sh, slope=pro.slope, seg=N, stats_seg=N, minmax=True, store=False)
            ctx.x = None
        mean, var = st["mean"].reshape(pro.h.shape[0], -1), st["var"].reshape(pro.h.shape[0], -1)
        ctx.pro = pro
        ctx.save_for_backward(W2, mean, st["imax"], st["imin"])
        ctx.wshape = W.shape
        return mean, var, st["max"], st["min"]

    @staticmethod
    def backward(ctx, gmean, gvar, gmax, gmin):
        W2, mean, imax, imin = ctx.saved_tensors
        pro = ctx.pro
        x = ctx.x if ctx.x is not None else pro.materialise()
        gx, gW, row_term = conv_row_reduce_backward(x, W2, mean, imax, imin, gmean, gvar, gmax, gmin, True, ctx.needs_input_grad[1],
                                                    split_row_term=True)
        gh, gts = pro.backward(gx, gy_row=row_term)               # the row-constant term is added inside the two row kernels
        return (gh, gW.view(ctx.wshape) if gW is not None else None, None, *gts)


class BcastActConv(torch.autograd.Function):
    """y[p, :, b] = W[p] . leaky_relu(A[p,:,b] * xhat[p] + D[p,:,b]) for every sample b, with xhat [P, C, L] SHARED by the samples
    (the decoders' first layer: the lattice response is batch independent, only the AdaIN.BN.SE scale/shift A, D [P, C, B] differ).
    The [P, C, B, L] activated tensor is never formed: the GEMM reads xhat (L2 resident) once per sample tile and applies (A, D) in
    its prologue; the weight gradient does the same.  Returns y [P, Cout, B, L] and its per-(p, c, b) row statistics."""
    @staticmethod
    def forward(ctx, xhat, A, D, W, slope):
        xhat, A, D = xhat.contiguous(), A.contiguous().float(), D.contiguous().float()
        L, B = xhat.shape[-1], A.shape[-1]
        y, st = gemm.conv_fwd(xhat, W, scale=A, shift=D, slope=slope, seg=L, stats_seg=L, x_repeat=B)
        ctx.save_for_backward(xhat, A, D, W)
        ctx.slope = float(slope)
        mean, var = st["mean"].reshape(y.shape[:-1]), st["var"].reshape(y.shape[:-1])
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, gy, *_):
        xhat, A, D, W = ctx.saved_tensors
        gy = gy.contiguous()
        L, B = xhat.shape[-1], A.shape[-1]
        gW = gemm.conv_wgrad(gy, xhat, batched=True, scale=A, shift=D, slope=ctx.slope, seg=L, x_repeat=B) if ctx.needs_input_grad[3] else None
        gx = gemm.conv_dgrad(gy, W)                                   # [P, C, B, L]: gradient w.r.t. the activated tensor
        R = A.numel()
        gh = torch.empty_like(xhat)
        gsc, gsh = torch.empty_like(A), torch.empty_like(D)
        with torch.cuda.device(xhat.device), _op("row_affine_act_bwd", 1):
            check(_lib.load().snb_row_affine_act_bwd(ptr(gx), ptr(xhat), ptr(A), ptr(D), R, L, B, ctx.slope, ptr(gh), ptr(gsc), ptr(gsh),
                                                     stream_ptr()), "row_affine_act_bwd")
        return gh, gsc, gsh, gW, None


def bcast_act_conv(xhat, A, D, W, slope=0.0):
    return BcastActConv.apply(xhat, A, D, W, slope)


def conv1x1(x, W, stats_seg=None):
    return Conv1x1.apply(x, W, stats_seg)


def act_conv(W, pro, stats_seg=None, h=None):
    """h: the autograd-connected tensor the Prologue was built from (pro.h is its detached copy)."""
    return ActConv.apply(h, W, pro, stats_seg, *pro.tensors)


def act_conv_row_reduce(W, pro, h):
    return ActConvRowReduce.apply(h, W, pro, *pro.tensors)


def conv_row_reduce(x, W):
    return ConvRowReduce.apply(x, W)


def row_stats(h):
    return RowStats.apply(h)


def row_affine_act(h, scale, shift, slope=0.0, in_div=1, out_shape=None):
    return RowAffineAct.apply(h, scale, shift, in_div, slope, tuple(h.shape) if out_shape is None else tuple(out_shape))


def row_minmax(h):
    return RowMinMax.apply(h)
