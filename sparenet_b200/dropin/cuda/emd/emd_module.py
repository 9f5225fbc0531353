"""Drop-in for the reference's cuda/emd/emd_module.py: emdFunction / emdModule.

emdModule()(xyz1, xyz2, eps, iters) -> (dist [B,N] squared distances, assignment [B,N] int32)
(reference :29-95).  Same limits, failing the same way: n == m, n % 1024 == 0, B <= 512 are asserted
(:36-39); inputs are forced to float32 CUDA (:41-42); only xyz1 receives a gradient, xyz2 gets zeros
(:79-87).  The 12 scratch tensors of the reference (:43-54) are one workspace sized by
snb_emd_workspace_bytes; all `iters` rounds run inside one kernel launch (snb_emd_fwd).
"""
import torch
from torch import nn
from torch.autograd import Function

from sparenet_b200 import functional as F_


class emdFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        batchsize, n, _ = xyz1.size()
        _, m, _ = xyz2.size()
        assert n == m
        assert xyz1.size()[0] == xyz2.size()[0]
        assert n % 1024 == 0
        assert batchsize <= 512
        xyz1 = xyz1.contiguous().float().cuda()
        xyz2 = xyz2.contiguous().float().cuda()
        dist, assignment = F_.emd_forward(xyz1, xyz2, eps, iters)
        ctx.save_for_backward(xyz1, xyz2, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx, graddist, gradidx):
        xyz1, xyz2, assignment = ctx.saved_tensors
        gradxyz1 = F_.emd_backward(xyz1, xyz2, graddist.contiguous(), assignment)
        return gradxyz1, torch.zeros_like(xyz2), None, None


class emdModule(nn.Module):
    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1, input2, eps, iters)
