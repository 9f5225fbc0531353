"""Drop-in for the reference's utils/p2i_utils.py: the differentiable multi-view depth-map renderer.

ComputeDepthMaps(projection, eyepos_scale, image_size)(data [B,N,3], view_id=0, radius_list=[10.0]) ->
[B, len(radius_list), S, S] (reference :168-252); returns None for view_id >= 8 (:212-213).  Eight fixed look-at
views from the cube corners (:173-182), "orthorgonal" (scale 1.5, near 0.1, far 10) or "perspective" (fovy pi/4)
projection (:185-198), pixel row = -y, col = x (:225), feature = 1 - (z - zmin)/(zmax - zmin) with min/max over
the WHOLE call (:226), one p2i(max) per radius (:230-251).

float32 CUDA clouds take the FUSED renderer (snb_depthmaps_fwd/bwd, sparenet_b200/csrc/depthmaps.cu): view transform, depth
normalisation (with the gradient through zmin/zmax) and the lock-free max-splat in one C-ABI call per (view, radius); the 4x4
matrix travels as 16 kernel arguments instead of a [B*N,4,4] tensor expanded on the host and uploaded (33 MB per call, :217).
float64 (the reference's gradcheck dtype) keeps the step-by-step route through cuda.p2i_op.p2i (snb_p2i_max_*).
"""
import math

import torch

from sparenet_b200 import functional as F_
from sparenet_b200.dropin.cuda.p2i_op import p2i


class DepthMapFunction(torch.autograd.Function):
    """data [B,N,3] -> depth map [B,1,S,S] for one view matrix and one radius (snb_depthmaps_fwd / _bwd)."""
    @staticmethod
    def forward(ctx, data, view_matrix, size, radius):
        data = data.contiguous()
        out, ids, ws = F_.depthmaps_forward(data, view_matrix, size, size, radius)
        ctx.save_for_backward(data, ids, ws)
        ctx.meta = (view_matrix, float(radius))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        data, ids, ws = ctx.saved_tensors
        view_matrix, radius = ctx.meta
        return F_.depthmaps_backward(grad_out.contiguous(), ids, data, view_matrix, radius, ws), None, None, None

N_VIEWS_PREDEFINED = 8


def normalize(x, dim):
    n = x.norm(None, dim=dim, keepdim=True)
    return x / torch.max(n, torch.tensor(1e-6, dtype=x.dtype, device=x.device))


def look_at(eyes, centers, ups):
    """[batch,3] eye / target / up -> [batch,4,4] view matrices (rotate after translating the eye to the origin)."""
    z = normalize(eyes - centers, dim=1)
    x = normalize(torch.cross(ups, z, dim=1), dim=1)
    y = torch.cross(z, x, dim=1)
    n = eyes.size(0)
    rot = torch.zeros(n, 4, 4, dtype=eyes.dtype, device=eyes.device)
    rot[:, 0, :3], rot[:, 1, :3], rot[:, 2, :3] = x, y, z
    rot[:, 3, 3] = 1
    trans = torch.eye(4, dtype=eyes.dtype, device=eyes.device).repeat(n, 1, 1)
    trans[:, :3, 3] = -eyes
    return rot @ trans


def perspective(fovy, aspect, z_near, z_far):
    t = torch.tan(fovy / 2.0)
    m = torch.zeros(fovy.size(0), 4, 4, dtype=fovy.dtype, device=fovy.device)
    m[:, 0, 0] = 1.0 / aspect / t
    m[:, 1, 1] = 1.0 / t
    m[:, 2, 2] = -(z_far + z_near) / (z_far - z_near)
    m[:, 2, 3] = -2.0 * z_far * z_near / (z_far - z_near)
    m[:, 3, 2] = -1.0
    return m


def orthorgonal(scalex, scaley, z_near, z_far):
    m = torch.zeros(z_near.size(0), 4, 4, dtype=z_near.dtype, device=z_near.device)
    m[:, 0, 0] = scalex
    m[:, 1, 1] = scaley
    m[:, 2, 2] = -2.0 / (z_far - z_near)
    m[:, 2, 3] = (z_far + z_near) / (z_far - z_near)
    m[:, 3, 3] = 1.0
    return m


def transform(matrix, points):
    """matrix [n,4,4] (or [4,4]), points [n,3] -> perspective-divided [n,3]."""
    if matrix.dim() == 2:
        h = points @ matrix[:, :3].t() + matrix[:, 3]
    else:
        h = (matrix[:, :, :3] @ points.unsqueeze(-1)).squeeze(-1) + matrix[:, :, 3]
    return h[:, :3] / h[:, 3:4]


class ComputeDepthMaps(torch.nn.Module):
    def __init__(self, projection: str = "orthorgonal", eyepos_scale: float = 1.0, image_size: int = 256):
        super().__init__()
        self.image_size = image_size
        self.eyes_pos_list = [[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
        self.num_views = len(self.eyes_pos_list)
        assert projection in {"perspective", "orthorgonal"}
        one = torch.tensor([1.0])
        if projection == "perspective":
            self.projection_matrix = perspective(fovy=one * (math.pi / 4), aspect=one, z_near=one * 0.1, z_far=one * 10.0)
        else:
            self.projection_matrix = orthorgonal(scalex=one * 1.5, scaley=one * 1.5, z_near=one * 0.1, z_far=one * 10.0)
        eyes = torch.tensor(self.eyes_pos_list, dtype=torch.float32) * eyepos_scale
        views = look_at(eyes, torch.zeros(self.num_views, 3), torch.tensor([[0.0, 0.0, 1.0]]).expand(self.num_views, 3))
        pre = self.projection_matrix @ views                                   # [8,4,4]
        self.pre_matrix_list = [pre[i:i + 1] for i in range(self.num_views)]
        # the reference re-registers one buffer per view, so its state_dict holds the LAST view's matrix (:208)
        self.register_buffer("_pre_matrix", pre[-1:].clone())
        self.register_buffer("_all_pre", pre.clone(), persistent=False)
        self._view_rows = [tuple(pre[i].reshape(-1).tolist()) for i in range(self.num_views)]   # 16 floats per view, host side
        self._cache = {}

    def _aux(self, B, N, dtype, device):
        key = (B, N, dtype, device)
        if key not in self._cache:
            binds = torch.arange(B, dtype=torch.int32, device=device).repeat_interleave(N)
            bg = torch.zeros(B, 1, self.image_size, self.image_size, dtype=dtype, device=device)
            self._cache = {key: (binds, bg)}
        return self._cache[key]

    def forward(self, data, view_id=0, radius_list=[10.0]):
        if view_id >= self.num_views:
            return None
        B, N = data.size(0), data.size(1)
        if data.dtype == torch.float32 and data.is_cuda:     # (CPU tensors fall through to p2i, which rejects them: no CPU path)
            maps = [DepthMapFunction.apply(data, self._view_rows[view_id], self.image_size, float(r)) for r in radius_list]
            return maps[0] if len(maps) == 1 else torch.cat(maps, dim=1)
        m = self._all_pre[view_id].to(device=data.device, dtype=data.dtype)
        batch_inds, background = self._aux(B, N, data.dtype, data.device)
        pos = transform(m, data.reshape(-1, 3))
        pos_xs, pos_ys, pos_zs = pos[:, 0:1], pos[:, 1:2], pos[:, 2:3]
        pos_ijs = torch.cat([-pos_ys, pos_xs], dim=1)
        zmin, zmax = pos_zs.min(), pos_zs.max()
        point_features = 1.0 - (pos_zs - zmin) / (zmax - zmin)
        maps = [p2i(pos_ijs, point_features, batch_inds, background, kernel_radius=r, kernel_kind_str="cos", reduce="max")
                for r in radius_list]
        return maps[0] if len(maps) == 1 else torch.cat(maps, dim=1)
