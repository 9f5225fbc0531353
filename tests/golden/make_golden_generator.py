"""Generates tests/golden/generator_ref.npz by running the REAL reference classes of
/root/reference/models/sparenet_generator.py on CPU (build container only): EdgeConvResFeat (kNN via the
reference's own CPU fallback :871-875), SpareNetEncode, StyleBasedAdaIn/GridDecoder, PointNetRes.
Weights come from oracle.generator_ref.deterministic_fill (a closed-form fill keyed on parameter names), so no
checkpoint has to be shipped.  The GPU box never runs this."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.build_ref import load_ref  # noqa: E402
from oracle.generator_ref import deterministic_fill  # noqa: E402

load_ref("MDS")
load_ref("expansion_penalty")
sys.path.insert(0, "/root/reference")
_real_avail = torch.cuda.is_available
torch.cuda.is_available = lambda: False     # force the reference's CPU kNN branch (no knn_cuda wheel here)
import models.sparenet_generator as R  # noqa: E402

out = {}


def run(tag, mod, inputs, grad_wrt=0):
    mod.train()
    deterministic_fill(mod)
    ins = [t.clone().requires_grad_(i == grad_wrt) for i, t in enumerate(inputs)]
    y = mod(*ins)
    w = torch.sin(torch.arange(y.numel(), dtype=torch.float32) * 0.7).view_as(y)
    (y * w).sum().backward()
    out[f"{tag}_out"] = y.detach().numpy()
    out[f"{tag}_gin"] = ins[grad_wrt].grad.numpy()
    for i, t in enumerate(inputs):
        out[f"{tag}_in{i}"] = t.numpy()
    # one running-stat buffer and one weight gradient as extra pins
    for name, b in mod.named_buffers():
        if name.endswith("running_var") and "adain" not in name:
            out[f"{tag}_rv"] = b.numpy().copy()
            out[f"{tag}_rv_name"] = np.array(name)
            break
    name, p = next((n, p) for n, p in mod.named_parameters() if p.grad is not None and p.dim() >= 2)
    out[f"{tag}_gw"] = p.grad.numpy().copy()
    out[f"{tag}_gw_name"] = np.array(name)


def det(shape, k):
    n = int(np.prod(shape))
    return torch.sin(torch.arange(n, dtype=torch.float64) * (0.91 + 0.07 * k) + k).view(*shape).float()


run("edge_small", R.EdgeConvResFeat(use_SElayer=True, k=8, output_size=64, hide_size=256), [det((2, 3, 96), 1) * 0.5])
run("edge_full", R.EdgeConvResFeat(use_SElayer=True, k=8, output_size=128, hide_size=4096), [det((2, 3, 64), 2) * 0.5])
run("encode", R.SpareNetEncode(hide_size=64, bottleneck_size=32, use_SElayer=True, encode="Residualnet"), [det((3, 3, 80), 3) * 0.5])
run("pnres", R.PointNetRes(use_SElayer=True), [det((2, 4, 64), 4) * 0.5])


class AdaWrap(torch.nn.Module):  # StyleBasedAdaIn.forward(content, style, adain_params) (:420-422)
    def __init__(self):
        super().__init__()
        self.m = R.StyleBasedAdaIn(input_dim=2, style_dim=16, use_SElayer=True)

    def forward(self, params, content):
        return self.m(content, None, params)


run("adain", AdaWrap(), [det((2, 3590), 5) * 0.7 + 0.1, det((2, 2, 32), 6)])
out["grid"] = np.array(R.grid_generation(16384, 32)[0], dtype=np.float32)
torch.cuda.is_available = _real_avail
np.savez_compressed(os.path.join(os.path.dirname(__file__), "generator_ref.npz"), **out)
print({k: v.shape for k, v in out.items() if k.endswith("_out")})
