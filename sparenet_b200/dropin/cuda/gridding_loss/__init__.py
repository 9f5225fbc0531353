"""Drop-in for the reference's cuda/gridding_loss/__init__.py (GRNet): GriddingDistanceFunction (:13-47), GriddingDistance(scale)
(pred, gt) -> (pred_grid, gt_grid) [B, V, 8] (:50-99) and GriddingLoss(scales, alphas)(pred, gt) -> scalar (:102-123).

The clouds are scaled by scale / 2; the bounds are floor(min) - 1 / ceil(max) + 1 over BOTH clouds of the whole batch (:62-81);
per sample the rows whose coordinate SUM is zero are dropped (:88-92, hence the per-sample loop: the kept count differs).  The grid
keeps eight accumulators per vertex, one per corner role (snb_gridding_dist_*).  The six bounds come from twelve device
reductions like in the reference and are read on the host once per call (the reference passes them to its extension as Python
floats, i.e. it synchronises too)."""
import torch

from sparenet_b200 import functional as F_


class GriddingDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, min_x, max_x, min_y, max_y, min_z, max_z, pred_cloud, gt_cloud):
        bounds = tuple(float(v) for v in (min_x, max_x, min_y, max_y, min_z, max_z))
        pred_grid, pw, pi = F_.gridding_dist_forward(pred_cloud.contiguous(), bounds)
        gt_grid, gw, gi = F_.gridding_dist_forward(gt_cloud.contiguous(), bounds)
        ctx.save_for_backward(pw, pi, gw, gi)
        return pred_grid, gt_grid

    @staticmethod
    def backward(ctx, grad_pred_grid, grad_gt_grid):
        pw, pi, gw, gi = ctx.saved_tensors
        grad_pred = F_.gridding_dist_backward(pw, pi, grad_pred_grid.contiguous())
        grad_gt = F_.gridding_dist_backward(gw, gi, grad_gt_grid.contiguous())
        return None, None, None, None, None, None, grad_pred, grad_gt


class GriddingDistance(torch.nn.Module):
    def __init__(self, scale=1):
        super().__init__()
        self.scale = scale

    def forward(self, pred_cloud, gt_cloud):
        pred_cloud = pred_cloud * self.scale / 2
        gt_cloud = gt_cloud * self.scale / 2
        lo = torch.minimum(pred_cloud.amin((0, 1)), gt_cloud.amin((0, 1)))
        hi = torch.maximum(pred_cloud.amax((0, 1)), gt_cloud.amax((0, 1)))
        b = torch.stack((torch.floor(lo) - 1, torch.ceil(hi) + 1), 1).reshape(-1).tolist()     # min_x, max_x, min_y, max_y, min_z, max_z
        pred_grids, gt_grids = [], []
        for pc, gc in zip(torch.split(pred_cloud, 1, dim=0), torch.split(gt_cloud, 1, dim=0)):
            pc = pc[torch.sum(pc, dim=2).ne(0)].unsqueeze(dim=0)
            gc = gc[torch.sum(gc, dim=2).ne(0)].unsqueeze(dim=0)
            pg, gg = GriddingDistanceFunction.apply(*b, pc, gc)
            pred_grids.append(pg)
            gt_grids.append(gg)
        return torch.cat(pred_grids, dim=0).contiguous(), torch.cat(gt_grids, dim=0).contiguous()


class GriddingLoss(torch.nn.Module):
    def __init__(self, scales=[], alphas=[]):
        super().__init__()
        self.scales, self.alphas = scales, alphas
        self.gridding_dists = [GriddingDistance(scale=s) for s in scales]
        self.l1_loss = torch.nn.L1Loss()

    def forward(self, pred_cloud, gt_cloud):
        gridding_loss = None
        for alpha, gdist in zip(self.alphas, self.gridding_dists):
            pred_grid, gt_grid = gdist(pred_cloud, gt_cloud)
            term = alpha * self.l1_loss(pred_grid, gt_grid)
            gridding_loss = term if gridding_loss is None else gridding_loss + term
        return gridding_loss
