"""Autograd Functions over the fused generator kernels (snb_edge_reduce_*, snb_row_*): host plumbing only.

Used by sparenet_b200/dropin/models/sparenet_generator.py.  CUDA float32 only; no fallback.
"""
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .functional import _op


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


class RowStats(torch.autograd.Function):
    """h [..., L] -> (mean, biased var) over the last dim, shape h.shape[:-1]."""
    @staticmethod
    def forward(ctx, h):
        h = h.contiguous()
        L = h.shape[-1]
        R = h.numel() // L
        mean = torch.empty(h.shape[:-1], device=h.device, dtype=torch.float32)
        var = torch.empty_like(mean)
        with torch.cuda.device(h.device), _op("row_stats", 1):
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
        with torch.cuda.device(h.device), _op("row_stats_bwd", 1):
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


def edge_reduce(a, c, idx):
    return EdgeReduce.apply(a, c, idx)


def row_stats(h):
    return RowStats.apply(h)


def row_affine_act(h, scale, shift, slope=0.0, in_div=1, out_shape=None):
    return RowAffineAct.apply(h, scale, shift, in_div, slope, tuple(h.shape) if out_shape is None else tuple(out_shape))


def row_minmax(h):
    return RowMinMax.apply(h)
