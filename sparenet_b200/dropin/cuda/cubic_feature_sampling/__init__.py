"""Drop-in for the reference's cuda/cubic_feature_sampling/__init__.py (GRNet): CubicFeatureSamplingFunction (:13-35) and
CubicFeatureSampling()(ptcloud [B,n,3] in [-1,1], cubic_features [B,C,S,S,S], neighborhood_size=1) -> [B, n, (2 ns)^3, C] (:38-45):
the cloud is mapped to grid units (p * S/2 + S/2) and every point gathers the features of the (2 ns)^3 vertices around it (zeros
outside the grid).  Only the feature volume receives a gradient; the cloud's is zero, as in the reference (:165-170 of the .cu)."""
import torch

from sparenet_b200 import functional as F_


class CubicFeatureSamplingFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ptcloud, cubic_features, neighborhood_size=1):
        scale = cubic_features.size(2)
        point_features, grid_pt_indexes = F_.cubic_sampling_forward(ptcloud.contiguous(), cubic_features.contiguous(), neighborhood_size)
        ctx.save_for_backward(grid_pt_indexes)
        ctx.meta = (int(scale), int(neighborhood_size), ptcloud.shape)
        return point_features

    @staticmethod
    def backward(ctx, grad_point_features):
        (grid_pt_indexes,) = ctx.saved_tensors
        scale, neighborhood_size, pshape = ctx.meta
        grad_cubic_features = F_.cubic_sampling_backward(grad_point_features.contiguous(), grid_pt_indexes, scale, neighborhood_size)
        return grad_point_features.new_zeros(pshape), grad_cubic_features, None


class CubicFeatureSampling(torch.nn.Module):
    def forward(self, ptcloud, cubic_features, neighborhood_size=1):
        h_scale = cubic_features.size(2) / 2
        ptcloud = ptcloud * h_scale + h_scale
        return CubicFeatureSamplingFunction.apply(ptcloud, cubic_features, neighborhood_size)
