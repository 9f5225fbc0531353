"""Drop-in for the reference's cuda/chamfer_dist/__init__.py (GRNet-flavoured Chamfer API).

Same names, argument order and return arities: ChamferFunction.apply(xyz1, xyz2) -> (dist1, dist2)
(reference :6-18), ChamferDistance(ignore_zeros)(xyz1, xyz2) -> scalar mean(d1)+mean(d2) (:21-35),
ChamferDistanceSeperate(...) -> (mean(d1), mean(d2)) (:38-52).  The arithmetic runs in
libsparenet_b200.so (snb_chamfer_fwd / snb_chamfer_bwd).
"""
import torch

from sparenet_b200 import functional as F_


class ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        d1, d2, i1, i2 = F_.chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, i1, i2)
        ctx.mark_non_differentiable(i1, i2)
        return d1, d2

    @staticmethod
    def backward(ctx, g1, g2):
        xyz1, xyz2, i1, i2 = ctx.saved_tensors
        return F_.chamfer_backward(xyz1, xyz2, i1, i2, g1.contiguous(), g2.contiguous())


def _drop_zero_rows(xyz1, xyz2):
    # reference :28-32 -- only for batch size 1, rows whose coordinate SUM is zero are removed
    keep1 = xyz1.sum(dim=2).ne(0)
    keep2 = xyz2.sum(dim=2).ne(0)
    return xyz1[keep1].unsqueeze(0), xyz2[keep2].unsqueeze(0)


class ChamferDistance(torch.nn.Module):
    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1, xyz2 = _drop_zero_rows(xyz1, xyz2)
        d1, d2 = ChamferFunction.apply(xyz1, xyz2)
        return d1.mean() + d2.mean()


class ChamferDistanceSeperate(torch.nn.Module):
    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1, xyz2 = _drop_zero_rows(xyz1, xyz2)
        d1, d2 = ChamferFunction.apply(xyz1, xyz2)
        return d1.mean(), d2.mean()
