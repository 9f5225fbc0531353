"""Host plumbing of the tcgen05 TF32 GEMM (csrc/gemm_tc.cu, C ABI snb_gemm_tf32): 1x1 convolutions on channel-major activations.

Everything here is shape bookkeeping: torch CUDA tensors in, one C-ABI call, torch CUDA tensors out.  No fallback: a shape the kernel
does not serve raises (thin layers -- fewer than 32 channels on the contracted / transposed side -- are routed to a batched library
GEMM by the CALLER, explicitly, in dropin/models/sparenet_generator.py).

Activations are [G, C, *pos] with the positions contiguous (any trailing shape, e.g. [P, C, B, 512] for the folding decoders);
weights are [Cout, Cin] (shared) or [G, Cout, Cin] (one per batch entry), last dimension contiguous, row stride a multiple of 4.
"""
import ctypes

import torch

from . import _lib
from ._lib import GemmDesc, check, stream_ptr
from .functional import SnbValueError, _op

FWD, DGRAD, WGRAD = 0, 1, 2


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _act3(x, name):
    """[G, C, *pos] -> (G, C, Npos, row stride, batch stride) of a tensor whose positions are contiguous."""
    if not x.is_cuda or x.dtype != torch.float32:
        raise SnbValueError(f"{name} must be a CUDA float32 tensor (sparenet_b200 has no CPU path)")
    if x.dim() < 3:
        raise SnbValueError(f"{name} must be [G, C, *positions]")
    if not x.is_contiguous():
        x = x.contiguous()
    G, C = x.shape[0], x.shape[1]
    n = x.numel() // (G * C) if G * C else 0
    return x, G, C, n


def _weight(W, G, name):
    if not W.is_cuda or W.dtype != torch.float32:
        raise SnbValueError(f"{name} must be a CUDA float32 tensor")
    if W.dim() == 3 and W.shape[0] != G:
        raise SnbValueError(f"{name}: batched weight needs {G} entries, got {W.shape[0]}")
    if W.dim() not in (2, 3):
        raise SnbValueError(f"{name} must be [Cout, Cin] or [G, Cout, Cin]")
    if W.stride(-1) != 1 or W.stride(-2) % 4 != 0 or (W.dim() == 3 and W.stride(0) % 4 != 0) or W.data_ptr() % 16 != 0:
        W = W.contiguous()
        if W.stride(-2) % 4 != 0:
            raise SnbValueError(f"{name}: the contiguous dimension must be a multiple of 4 (got {tuple(W.shape)})")
    return W


def stat_tiles(N, block_n=0):
    """(number of statistics tiles along N, their width) of a forward call."""
    lib = _lib.load()
    return lib.snb_gemm_tf32_tiles(int(N), int(block_n)), lib.snb_gemm_tf32_block_n(int(N), int(block_n)) // 2


SHAPE_LOG = None   # development: a list that receives (label, start_event, end_event, flops) per call (tools/gemm_shapes.py)


def _run(desc, dev, what, nbytes=None):
    flops = 2 * int(desc.G) * int(desc.BI) * int(desc.M) * int(desc.N) * int(desc.K)
    with torch.cuda.device(dev), _op(what, 1, nbytes, flops):
        if SHAPE_LOG is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        check(_lib.load().snb_gemm_tf32(ctypes.byref(desc), stream_ptr()), what)
        if SHAPE_LOG is not None:
            b.record()
            label = (f"{what[5:]:5s} G={desc.G:<3d} BI={desc.BI:<3d} M={desc.M:<5d} N={desc.N:<6d} K={desc.K:<6d} "
                     f"{'prologue' if desc.scale else 'plain   '} store={desc.store} {'stats' if desc.pmean else ''}{' minmax' if desc.pmax else ''}")
            SHAPE_LOG.append((label, a, b, flops))


def conv_fwd(x, W, scale=None, shift=None, slope=0.0, seg=None, stats_seg=None, minmax=False, store=True, x_repeat=1, out=None):
    """y[g] = W . T(x[g]),  T(x) = leaky_relu(scale*x + shift, slope) per (g, channel, segment of `seg` positions) when scale is given.
    x_repeat = R > 1: x [G, Cin, n] is tiled R times along the position axis (N = R*n output positions; scale/shift still per segment of
    the long axis) -- the decoder's first layer, whose input is the same lattice response for every sample.
    out: an existing contiguous [G, Cout, *pos] tensor the product is ADDED to (the epilogue's TMA reduce-add; no statistics then).
    Returns (y or None, stats) with stats = {} or {"mean","var": [G, Cout, Npos/stats_seg]} (biased variance of every segment of
    `stats_seg` positions of the OUTPUT rows) and, with minmax, {"max","min","imax","imin": [G, Cout]} over all positions."""
    x, G, Cin, n_in = _act3(x, "x")
    N = n_in * int(x_repeat)
    W = _weight(W, G, "W")
    Cout = W.shape[-2]
    if W.shape[-1] != Cin:
        raise SnbValueError(f"weight {tuple(W.shape)} does not match {Cin} input channels")
    if n_in % 32 != 0 or Cin % 4 != 0:
        raise SnbValueError(f"conv_fwd needs positions % 32 == 0 and Cin % 4 == 0 (got {n_in}, {Cin})")
    dev = x.device
    d = GemmDesc()
    d.mode, d.G, d.BI, d.M, d.N, d.K = FWD, G, 1, Cout, N, Cin
    d.A, d.lda, d.a_batch_stride = _p(W), W.stride(-2), (W.stride(0) if W.dim() == 3 else 0)
    d.B, d.ldb, d.b_batch_stride = _p(x), n_in, Cin * n_in
    if x_repeat > 1:
        d.b_pos_mod = n_in
    y = None
    if out is not None:
        if stats_seg is not None or minmax or not store:
            raise SnbValueError("conv_fwd(out=...) accumulates into `out`: no statistics, store must stay on")
        if not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == G * Cout * N and out.shape[1] == Cout):
            raise SnbValueError(f"out must be a contiguous CUDA float32 [G, {Cout}, *positions] tensor")
        y = out
        d.D, d.ldd, d.d_batch_stride = _p(y), N, Cout * N
        d.store = 2
    elif store:
        yshape = (G, Cout) + ((int(x_repeat),) + tuple(x.shape[2:]) if x_repeat > 1 else tuple(x.shape[2:]))
        y = torch.empty(yshape, device=dev, dtype=torch.float32)
        d.D, d.ldd, d.d_batch_stride = _p(y), N, Cout * N
        d.store = 1
    else:
        d.store = 0
    keep = [x, W, y]
    if scale is not None:
        seg = N if seg is None else int(seg)
        scale, shift = scale.contiguous().float(), shift.contiguous().float()
        if scale.numel() != G * Cin * (N // seg) or shift.numel() != scale.numel() or N % seg != 0:
            raise SnbValueError(f"scale/shift must hold G*Cin*(N/seg) = {G * Cin * (N // seg)} entries")
        d.scale, d.shift, d.slope, d.seg = _p(scale), _p(shift), float(slope), seg
        keep += [scale, shift]
    T, w = stat_tiles(N)
    pm = p2 = px = pn = ix = in_ = None
    if stats_seg is not None:
        if N % (2 * w) != 0 or stats_seg % w != 0 or N % stats_seg != 0:
            raise SnbValueError(f"statistics need positions % {2 * w} == 0 and segment % {w} == 0 (got {N}, {stats_seg})")
        pm, p2 = (torch.empty(G, Cout, T, device=dev, dtype=torch.float32) for _ in range(2))
        d.pmean, d.pm2 = _p(pm), _p(p2)
    if minmax:
        if N % (2 * w) != 0:
            raise SnbValueError(f"extrema need positions % {2 * w} == 0 (got {N})")
        px, pn = (torch.empty(G, Cout, T, device=dev, dtype=torch.float32) for _ in range(2))
        ix, in_ = (torch.empty(G, Cout, T, device=dev, dtype=torch.int32) for _ in range(2))
        d.pmax, d.pmin, d.pimax, d.pimin = _p(px), _p(pn), _p(ix), _p(in_)
    _run(d, dev, "gemm_fwd", 4 * (x.numel() + (y.numel() if store else 0)))
    stats = {}
    lib = _lib.load()
    if stats_seg is not None:
        # Chan's merge of the tiles of every segment (within-tile + between-tile second moments: no cancellation), one launch
        tps = stats_seg // w                                  # tiles per segment
        nseg = N // stats_seg
        mean, var = (torch.empty(G, Cout, nseg, device=dev, dtype=torch.float32) for _ in range(2))
        with torch.cuda.device(dev), _op("gemm_stats_merge", 1):
            check(lib.snb_gemm_stats_merge(_p(pm), _p(p2), G * Cout * nseg, tps, w, _p(mean), _p(var), stream_ptr()), "gemm_stats_merge")
        stats["mean"], stats["var"] = mean, var
    if minmax:
        vmax, vmin = (torch.empty(G, Cout, device=dev, dtype=torch.float32) for _ in range(2))
        imax, imin = (torch.empty(G, Cout, device=dev, dtype=torch.int32) for _ in range(2))
        with torch.cuda.device(dev), _op("gemm_stats_merge", 1):
            check(lib.snb_gemm_minmax_merge(_p(px), _p(pn), _p(ix), _p(in_), G * Cout, T, _p(vmax), _p(vmin), _p(imax), _p(imin), stream_ptr()),
                  "gemm_minmax_merge")
        stats["max"], stats["min"], stats["imax"], stats["imin"] = vmax, vmin, imax, imin
    return y, stats


def conv_dgrad(gy, W):
    """gx[g] = W^T . gy[g]: gy [G, Cout, *pos] -> gx [G, Cin, *pos].  Cin % 32 == 0."""
    gy, G, Cout, N = _act3(gy, "gy")
    W = _weight(W, G, "W")
    Cin = W.shape[-1]
    if W.shape[-2] != Cout:
        raise SnbValueError(f"weight {tuple(W.shape)} does not match {Cout} output channels")
    if N % 32 != 0 or Cin % 32 != 0:
        raise SnbValueError(f"conv_dgrad needs positions % 32 == 0 and Cin % 32 == 0 (got {N}, {Cin})")
    gx = torch.empty((G, Cin) + tuple(gy.shape[2:]), device=gy.device, dtype=torch.float32)
    d = GemmDesc()
    d.mode, d.G, d.BI, d.M, d.N, d.K = DGRAD, G, 1, Cin, N, Cout
    d.A, d.lda, d.a_batch_stride = _p(W), W.stride(-2), (W.stride(0) if W.dim() == 3 else 0)
    d.B, d.ldb, d.b_batch_stride = _p(gy), N, Cout * N
    d.D, d.ldd, d.d_batch_stride = _p(gx), N, Cin * N
    d.store = 1
    _run(d, gy.device, "gemm_dgrad", 4 * (gy.numel() + gx.numel()))
    return gx


def conv_wgrad(gy, x, batched, scale=None, shift=None, slope=0.0, seg=None, x_repeat=1):
    """gW = sum over (batch,) positions of gy . T(x)^T: gy [G, Cout, *pos], x [G, Cin, *pos] -> [Cout, Cin] (batched=False: the
    batch is reduced) or [G, Cout, Cin] (batched=True: one weight per batch entry).  T as in conv_fwd (the forward's prologue);
    x_repeat as in conv_fwd (batched only)."""
    gy, G, Cout, N = _act3(gy, "gy")
    x, G2, Cin, N2 = _act3(x, "x")
    if G2 != G or N2 * int(x_repeat) != N or (x_repeat > 1 and not batched):
        raise SnbValueError("gy and x must agree in batch and positions")
    if N % 4 != 0:
        raise SnbValueError("conv_wgrad needs positions % 4 == 0")
    dev = x.device
    GO, BI = (G, 1) if batched else (1, G)
    gW = torch.zeros((GO, Cout, Cin) if batched else (Cout, Cin), device=dev, dtype=torch.float32)
    d = GemmDesc()
    d.mode, d.G, d.BI, d.M, d.N, d.K = WGRAD, GO, BI, Cout, Cin, N
    d.A, d.lda, d.a_batch_stride = _p(gy), N, Cout * N
    d.B, d.ldb, d.b_batch_stride = _p(x), N2, Cin * N2
    if x_repeat > 1:
        d.b_pos_mod = N2
    d.D, d.ldd, d.d_batch_stride = _p(gW), Cin, Cout * Cin
    d.store, d.split = 2, 0                                   # TMA reduce-add into the zeroed gradient; split-K chosen by the library
    keep = [gy, x, gW]
    if scale is not None:
        seg = N if seg is None else int(seg)
        scale, shift = scale.contiguous().float(), shift.contiguous().float()
        if scale.numel() != G * Cin * (N // seg) or N % seg != 0 or seg % 32 != 0:
            raise SnbValueError("scale/shift must hold G*Cin*(N/seg) entries, seg % 32 == 0")
        d.scale, d.shift, d.slope, d.seg = _p(scale), _p(shift), float(slope), seg
        keep += [scale, shift]
    if Cin % 4 != 0:
        raise SnbValueError("conv_wgrad needs Cin % 4 == 0")
    _run(d, dev, "gemm_wgrad", 4 * (gy.numel() + x.numel()))
    return gW
