"""Drop-in for the reference's cuda/p2i_op/__init__.py: p2i(points, point_features, batch_inds, background,
kernel_radius, kernel_kind_str="cos", reduce="sum") (reference :99-131) and the two autograd Functions
(:22-93).  points are in [-1,1]^2 as (row, col); the affine map to pixel space stays in Python so autograd
differentiates it exactly as in the reference (:116-121).  float32 and float64 are both served (the latter
is the reference's gradcheck path, p2i_test.py:24-35).
"""
import torch
from torch.autograd import Function

from sparenet_b200 import functional as F_

__all__ = ["p2i"]


class P2ISumFunction(Function):
    @staticmethod
    def forward(ctx, points, point_features, batch_inds, background, kernel_kind, kernel_radius):
        ctx.save_for_backward(points, point_features, batch_inds)
        ctx.kernel_kind, ctx.kernel_radius = kernel_kind, kernel_radius
        return F_.p2i_sum_forward(points, point_features, batch_inds, background, kernel_kind, kernel_radius)

    @staticmethod
    def backward(ctx, out_grad):
        points, point_features, batch_inds = ctx.saved_tensors
        gp, gf = F_.p2i_sum_backward(out_grad, points, point_features, batch_inds, ctx.kernel_kind, ctx.kernel_radius)
        return gp, gf, None, out_grad, None, None


class P2IMaxFunction(Function):
    @staticmethod
    def forward(ctx, points, point_features, batch_inds, background, kernel_kind, kernel_radius):
        out, ids = F_.p2i_max_forward(points, point_features, batch_inds, background, kernel_kind, kernel_radius)
        ctx.save_for_backward(points, point_features, ids)
        ctx.kernel_kind, ctx.kernel_radius = kernel_kind, kernel_radius
        return out

    @staticmethod
    def backward(ctx, out_grad):
        points, point_features, ids = ctx.saved_tensors
        gp, gf, gb = F_.p2i_max_backward(out_grad, ids, points, point_features, ctx.kernel_kind, ctx.kernel_radius)
        return gp, gf, None, gb, None, None


_kernel_kind_dict = {"cos": 0}


def p2i(points, point_features, batch_inds, background, kernel_radius, kernel_kind_str="cos", reduce="sum"):
    kernel_kind = _kernel_kind_dict[kernel_kind_str]
    out_h, out_w = background.shape[2:]
    scale = torch.tensor([out_h - 1, out_w - 1], dtype=points.dtype, device=points.device).view(1, 2)
    points = (points + 1) / 2 * scale
    if reduce == "sum":
        return P2ISumFunction.apply(points, point_features, batch_inds, background, kernel_kind, kernel_radius)
    if reduce == "max":
        return P2IMaxFunction.apply(points, point_features, batch_inds, background, kernel_kind, kernel_radius)
    raise RuntimeError(f"Invalid reduce value: {reduce}")


custom_fun = P2ISumFunction.apply
