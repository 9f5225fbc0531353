"""Drop-in for the reference's cuda/gridding/__init__.py (GRNet): Gridding(scale)(ptcloud [B,n,3]) -> [B, scale^3]
and GriddingReverse(scale)(grid [B,scale,scale,scale]) -> [B, scale^3, 3] (reference :13-75).  Gridding scales the cloud by
scale // 2, drops per sample the rows whose coordinate SUM is zero (:41-47, hence the per-sample loop: the kept count
differs per sample) and uses the bounds [-s, s-1] on every axis (:16-18).  GriddingReverse rescales by 2/scale (:75)."""
import torch

from sparenet_b200 import functional as F_


class GriddingFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scale, ptcloud):
        grid, w, ix = F_.gridding_forward(ptcloud.contiguous(), (-scale, scale - 1, -scale, scale - 1, -scale, scale - 1))
        ctx.save_for_backward(w, ix)
        return grid

    @staticmethod
    def backward(ctx, grad_grid):
        w, ix = ctx.saved_tensors
        return None, F_.gridding_backward(w, ix, grad_grid.contiguous())


class Gridding(torch.nn.Module):
    def __init__(self, scale=1):
        super().__init__()
        self.scale = scale // 2

    def forward(self, ptcloud):
        ptcloud = ptcloud * self.scale
        grids = []
        for p in torch.split(ptcloud, 1, dim=0):
            keep = torch.sum(p, dim=2).ne(0)
            grids.append(GriddingFunction.apply(self.scale, p[keep].unsqueeze(dim=0)))
        return torch.cat(grids, dim=0).contiguous()


class GriddingReverseFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scale, grid):
        grid = grid.contiguous()
        ptcloud = F_.gridding_reverse_forward(grid.view(grid.size(0), -1), scale)
        ctx.save_for_backward(grid, ptcloud)
        ctx.scale = scale
        return ptcloud

    @staticmethod
    def backward(ctx, grad_ptcloud):
        grid, ptcloud = ctx.saved_tensors
        s = ctx.scale
        g = F_.gridding_reverse_backward(ptcloud, grid.view(grid.size(0), -1), grad_ptcloud.contiguous(), s)
        return None, g.view(-1, s, s, s)


class GriddingReverse(torch.nn.Module):
    def __init__(self, scale=1):
        super().__init__()
        self.scale = scale

    def forward(self, grid):
        return GriddingReverseFunction.apply(self.scale, grid) / self.scale * 2
