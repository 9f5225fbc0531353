"""Pins oracle/generator_ref.py (the plain restatement of the reference generator) against goldens produced by
the REAL reference classes on CPU (tests/golden/make_golden_generator.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import generator_ref as G

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "generator_ref.npz"))


class AdaWrap(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.m = G.StyleBasedAdaIn(2, 16)

    def forward(self, params, content):
        return self.m(content, None, params)


CASES = {
    "edge_small": lambda: G.EdgeConvResFeat(True, 8, hide_size=256, output_size=64),
    "edge_full": lambda: G.EdgeConvResFeat(True, 8, hide_size=4096, output_size=128),
    "encode": lambda: G.SpareNetEncode(bottleneck_size=32, hide_size=64),
    "pnres": lambda: G.PointNetRes(),
    "adain": lambda: AdaWrap(),
}


def run_case(tag, mod, device="cpu"):
    mod = mod.to(device).train()
    G.deterministic_fill(mod)
    ins = []
    i = 0
    while f"{tag}_in{i}" in GOLD:
        ins.append(torch.from_numpy(GOLD[f"{tag}_in{i}"]).to(device))
        i += 1
    ins[0].requires_grad_()
    y = mod(*ins)
    w = torch.sin(torch.arange(y.numel(), dtype=torch.float32) * 0.7).view_as(y).to(device)
    (y * w).sum().backward()
    params = dict(mod.named_parameters())
    bufs = dict(mod.named_buffers())
    res = {"out": y.detach().cpu(), "gin": ins[0].grad.cpu(), "gw": params[str(GOLD[f"{tag}_gw_name"])].grad.cpu()}
    if f"{tag}_rv" in GOLD:
        res["rv"] = bufs[str(GOLD[f"{tag}_rv_name"])].cpu()
    return res


def compare(tag, res, rtol, atol):
    for key, val in res.items():
        ref = torch.from_numpy(GOLD[f"{tag}_{key}"])
        scale = ref.abs().max().item() + 1e-12
        err = (val - ref).abs().max().item()
        assert err <= atol * scale + rtol * scale, f"{tag}.{key}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("tag", sorted(CASES))
def test_restatement_matches_reference_classes(tag):
    compare(tag, run_case(tag, CASES[tag]()), rtol=2e-5, atol=2e-6)


def test_state_dict_names_match_reference_layout():
    g = G.SpareNetGenerator(n_primitives=2, hide_size=32, bottleneck_size=32, num_points=64)
    keys = set(g.state_dict())
    for k in ["conv1.weight", "encoder.feat_extractor.conv1.weight", "encoder.feat_extractor.se4.fc.2.weight",
              "encoder.feat_extractor.resconv3.weight", "encoder.feat_extractor.bn5.running_var", "encoder.linear.bias", "encoder.bn.weight",
              "decoder.mlp.0.weight", "decoder.mlp.2.bias", "decoder.decoder.1.dec.conv3.bias", "decoder.decoder.0.dec.adain2.running_mean",
              "decoder.decoder.0.dec.se1.fc.0.weight", "decoder.decoder.1.dec.bn3.weight", "refine.residual.conv7.weight",
              "refine.residual.bn7.weight", "refine.residual.se6.fc.2.weight"]:
        assert k in keys, k
    assert not any("se3" in k for k in keys if k.startswith("refine."))   # PointNetRes has no se3 (:602-607)


def test_grid_matches_reference():
    g = G.grid_points(16384, 32)
    ref = (torch.from_numpy(GOLD["grid"]) - 0.5) * 2
    assert torch.equal(g, ref.t().contiguous())


def test_full_generator_cpu_runs_with_oracle_ops():
    torch.manual_seed(0)
    g = G.SpareNetGenerator(n_primitives=4, hide_size=64, bottleneck_size=64, num_points=256).train()
    g.apply(G.init_weights)
    data = {"partial_cloud": torch.rand(2, 128, 3) - 0.5}
    coarse, middle, refine, loss_mst = g(data)
    assert coarse.shape == middle.shape == refine.shape == (2, 256, 3)
    (refine.sum() + loss_mst).backward()
    assert all(p.grad is not None for n, p in g.named_parameters() if not n.startswith("conv1.") and "bn7" not in n)
