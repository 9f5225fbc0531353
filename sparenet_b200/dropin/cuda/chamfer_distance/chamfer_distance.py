"""Drop-in for the reference's cuda/chamfer_distance/chamfer_distance.py (MSN-flavoured Chamfer API, the
one every runner and Metrics imports: runners/sparenet_runner.py:10, utils/misc.py:10).

ChamferDistanceFunction.apply(xyz1, xyz2) -> (dist1, dist2) (reference :18-61); ChamferDistance()(a, b)
returns the two per-point tensors (:64-66); ChamferDistanceMean()(a, b) the scalar (:69-72).  Unlike the
reference there is no JIT build at import, no CPU-allocate-then-.cuda() of the outputs (:25-37) and no CPU
code path: CPU tensors raise.
"""
import torch

from sparenet_b200 import functional as F_


class ChamferDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
        d1, d2, i1, i2 = F_.chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, i1, i2)
        return d1, d2

    @staticmethod
    def backward(ctx, graddist1, graddist2):
        xyz1, xyz2, i1, i2 = ctx.saved_tensors
        return F_.chamfer_backward(xyz1, xyz2, i1, i2, graddist1.contiguous(), graddist2.contiguous())


class ChamferDistance(torch.nn.Module):
    def forward(self, xyz1, xyz2):
        return ChamferDistanceFunction.apply(xyz1, xyz2)


class ChamferDistanceMean(torch.nn.Module):
    def forward(self, xyz1, xyz2):
        d1, d2 = ChamferDistanceFunction.apply(xyz1, xyz2)
        return d1.mean() + d2.mean()
