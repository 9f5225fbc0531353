"""Drop-in for the reference's cuda/MDS/MDS_module.py: minimum_density_sample and gather_operation.

minimum_density_sample(xyz [B,n,3], npoint, mean_mst_length [B]) -> idx [B,npoint] int32, non-differentiable
(reference :7-38); gather_operation(features [B,C,n], idx) -> [B,C,npoint] with a scatter-add backward
(:44-75).  Dtype/contiguity requirements follow MDS.cpp:54-59 (float32 / int32, contiguous) and raise.
"""
import torch
from torch.autograd import Function

from sparenet_b200 import functional as F_


class MinimumDensitySampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint, mean_mst_length):
        if not xyz.is_contiguous():
            raise RuntimeError("points must be a contiguous tensor")  # MDS.cpp:116
        idx = F_.mds_sample(xyz, npoint, mean_mst_length.contiguous())
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, grad_idx=None):
        return None, None, None


minimum_density_sample = MinimumDensitySampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        if not features.is_contiguous() or not idx.is_contiguous():
            raise RuntimeError("points / idx must be contiguous tensors")  # MDS.cpp:56-57
        _, C, N = features.size()
        ctx.for_backwards = (idx, C, N)
        return F_.gather_forward(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        return F_.gather_backward(grad_out.contiguous(), idx, N), None


gather_operation = GatherOperation.apply
