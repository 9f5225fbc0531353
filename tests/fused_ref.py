"""Plain PyTorch definitions of the fused generator kernels (sparenet_b200/fused.py) -- TEST INFRASTRUCTURE ONLY.
They specify what snb_edge_reduce_* / snb_row_* must compute: used (a) to monkeypatch the kernels away so the
generator algebra can be verified on CPU, (b) as the fp32/fp64 reference the CUDA kernels are compared with on the GPU."""
import torch
import torch.nn.functional as F


def edge_reduce(a, c, idx):
    B, C, N = a.shape
    k = idx.shape[2]
    u = torch.gather(a, 2, idx.long().reshape(B, 1, N * k).expand(-1, C, -1)).view(B, C, N, k) + c.unsqueeze(-1)
    ud = u.double()
    return u.amax(-1), u.amin(-1), ud.sum((2, 3)), (ud * ud).sum((2, 3))


def edge_reduce_sel(a, c, idx, sel_max):
    umax, umin, S1, S2 = edge_reduce(a, c, idx)
    return torch.where(sel_max.view(1, -1, 1), umax, umin), S1, S2


def edge_reduce_sel_stacked(ac, idx, sel_max):
    C = ac.shape[1] // 2
    return edge_reduce_sel(ac[:, :C], ac[:, C:], idx, sel_max)


def row_stats(h):
    var, mean = torch.var_mean(h, dim=-1, unbiased=False)
    return mean, var


def row_affine_act(h, scale, shift, slope=0.0, in_div=1, out_shape=None):
    L = h.shape[-1]
    hin = h.reshape(-1, L)
    if in_div > 1:
        hin = hin.repeat_interleave(in_div, 0)
    y = F.leaky_relu(hin * scale.reshape(-1, 1) + shift.reshape(-1, 1), slope)
    return y.view(tuple(out_shape) if out_shape is not None else h.shape)


def row_minmax(h):
    return h.amax(-1), h.amin(-1)


def row_norm_act(h, fn, tensors, slope=0.0, stats=None):   # stats (from a GEMM epilogue) are recomputed here: plain autograd
    mean, var = row_stats(h)
    scale, shift = fn(mean, var, *tensors)
    return F.leaky_relu(h * scale.unsqueeze(-1) + shift.unsqueeze(-1), slope)


def row_norm_act_pool(h, fn, tensors, slope=0.0, stats=None):
    y = row_norm_act(h, fn, tensors, slope)
    return y.amax(-1), y.mean(-1)


def conv_row_reduce(x, W):
    h = torch.matmul(W.reshape(W.size(0), -1), x)
    mean, var = row_stats(h)
    return mean, var, h.amax(-1), h.amin(-1)


def thin_conv(x, W):
    return torch.matmul(W, x)


def row_stats_nograd(h):
    return row_stats(h.detach())


def conv1x1(x, W, stats_seg=None):
    """y = W x on [G, Cin, *pos], W [Cout, Cin] or [G, Cout, Cin]; with stats_seg also the row statistics of y's last dim."""
    G, Cin = x.shape[0], x.shape[1]
    y = torch.matmul(W, x.reshape(G, Cin, -1)).view((G, W.shape[-2]) + tuple(x.shape[2:]))
    if stats_seg is None:
        return y
    assert stats_seg == y.shape[-1]
    return (y,) + row_stats(y)


def cat_conv1x1(xs, W, stats_seg=None):
    return conv1x1(torch.cat(tuple(xs), dim=1), W, stats_seg)


def bcast_act_conv(xhat, A, D, W, slope=0.0):
    P, C, L = xhat.shape
    B = A.shape[-1]
    x = row_affine_act(xhat, A, D, slope=slope, in_div=B, out_shape=(P, C, B, L))
    return conv1x1(x, W, stats_seg=L)


class Prologue:
    """Reference of fused.Prologue: the activated tensor is simply formed, with plain autograd through the row statistics (the
    statistics handed in by a GEMM epilogue are ignored and recomputed from h)."""

    def __init__(self, h, mean, var, fn, tensors, slope=0.0):
        self.tensors = tuple(tensors)
        m, v = row_stats(h)
        scale, shift = fn(m, v, *tensors)
        self.y = F.leaky_relu(h * scale.unsqueeze(-1) + shift.unsqueeze(-1), slope)


def act_conv(W, pro, stats_seg=None, h=None):
    return conv1x1(pro.y, W, stats_seg)


def act_conv_row_reduce(W, pro, h):
    return conv_row_reduce(pro.y, W)


def patch(monkeypatch):
    from sparenet_b200 import fused
    for name in ("edge_reduce", "edge_reduce_sel", "edge_reduce_sel_stacked", "row_stats", "row_affine_act", "row_minmax", "row_norm_act", "conv_row_reduce",
                 "row_stats_nograd", "conv1x1", "Prologue", "act_conv", "act_conv_row_reduce", "bcast_act_conv", "thin_conv", "row_norm_act_pool", "cat_conv1x1"):
        monkeypatch.setattr(fused, name, globals()[name])
