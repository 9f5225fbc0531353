"""Generates tests/golden/depthmaps_ref.npz from the REAL reference utils/p2i_utils.py (ComputeDepthMaps host math:
look_at / projection / transform / depth feature) on CPU in the build container.  The reference's cuda.p2i_op JIT-builds a
CUDA extension at import, so it is replaced here by a stub that applies the reference's pixel mapping
(cuda/p2i_op/__init__.py:116-121) and calls the CPU oracle's p2i max."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def p2i_stub(points, point_features, batch_inds, background, kernel_radius, kernel_kind_str="cos", reduce="sum"):
    h, w = background.shape[2:]
    pts = (points + 1) / 2 * torch.tensor([h - 1, w - 1], dtype=points.dtype).view(1, 2)
    assert reduce == "max" and kernel_kind_str == "cos"
    return oracle.p2i_max_fwd(pts.contiguous(), point_features.contiguous(), batch_inds, background, float(kernel_radius))[0]


stub = types.ModuleType("cuda.p2i_op")
stub.p2i = p2i_stub
sys.modules["cuda.p2i_op"] = stub
sys.path.insert(0, "/root/reference")
import utils.p2i_utils as R  # noqa: E402

out = {}
torch.manual_seed(3)
data = (torch.rand(2, 300, 3) - 0.5)
out["data"] = data.numpy()
for proj in ("orthorgonal", "perspective"):
    r = R.ComputeDepthMaps(projection=proj, eyepos_scale=1.0, image_size=32).float()
    out[f"{proj}_pre"] = torch.cat(r.pre_matrix_list, 0).numpy()
    for view in (0, 5):
        out[f"{proj}_v{view}"] = r(data, view_id=view, radius_list=[3.0, 5.0]).numpy()
    assert r(data, view_id=8) is None
np.savez_compressed(os.path.join(os.path.dirname(__file__), "depthmaps_ref.npz"), **out)
print({k: v.shape for k, v in out.items()})
