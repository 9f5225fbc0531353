"""Drop-in for the reference's cuda/expansion_penalty/expansion_penalty_module.py.

expansionPenaltyModule()(xyz, primitive_size, alpha) -> (dist [B,N], assignment [B,N] int32,
mean_mst_length [B]) (reference :24-56).  Asserts primitive_size <= 512 and n % primitive_size == 0 like the
reference (:27,:29); additionally primitive_size must be a power of two (the reference's tree reductions
silently ignore part of the primitive otherwise, expansion_penalty_cuda.cu:64-73).  No B*n*512 scratch.
"""
import torch
from torch import nn
from torch.autograd import Function

from sparenet_b200 import functional as F_


class expansionPenaltyFunction(Function):
    @staticmethod
    def forward(ctx, xyz, primitive_size, alpha):
        assert primitive_size <= 512
        batchsize, n, _ = xyz.size()
        assert n % primitive_size == 0
        assert primitive_size >= 2 and (primitive_size & (primitive_size - 1)) == 0, "primitive_size must be a power of two"
        xyz = xyz.contiguous().float().cuda()
        dist, assignment, mean_mst_length = F_.expansion_forward(xyz, primitive_size, alpha)
        ctx.save_for_backward(xyz, assignment)
        ctx.mark_non_differentiable(assignment, mean_mst_length)
        return dist, assignment, mean_mst_length

    @staticmethod
    def backward(ctx, grad_dist, grad_idx, grad_mml):
        xyz, assignment = ctx.saved_tensors
        return F_.expansion_backward(xyz, grad_dist.contiguous(), assignment), None, None


class expansionPenaltyModule(nn.Module):
    def forward(self, input, primitive_size, alpha):
        return expansionPenaltyFunction.apply(input, primitive_size, alpha)
