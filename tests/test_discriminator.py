"""The drop-in discriminators (sparenet_b200/dropin/models/sparenet_discriminator.py) against the REAL reference classes run in the
build container (tests/golden/discriminator_ref.npz, written by tests/golden/make_golden_discriminator.py): identical state_dict
keys, outputs / feature maps / input gradient / a weight gradient within fp32 tolerance (1e-5 relative), and the spectral-norm
power-iteration state carried between calls.  CPU test: the module is plain PyTorch over cuDNN/ATen."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from _fill import name_fill  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "discriminator_ref.npz"))


def _close(a, b, tol=1e-5):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return (a - b).abs().max().item() <= tol * (b.abs().max().item() + 1e-30)


@pytest.mark.parametrize("tag", ["proj", "patch"])
def test_discriminator_matches_reference(tag):
    from sparenet_b200.dropin.models import sparenet_discriminator as D
    net = D.ProjectionD(num_classes=8, img_shape=(16, 32, 32)) if tag == "proj" else D.PatchDiscriminator(img_shape=(16, 64, 64))
    assert sorted(net.state_dict().keys()) == list(G[f"{tag}_keys"])
    name_fill(net)
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    net.train()
    img = torch.from_numpy(G[f"{tag}_img"])
    y = torch.from_numpy(G["proj_y"]) if tag == "proj" else None
    x = img.clone().requires_grad_()
    val, feats = net(x, feat=True, y=y)
    assert len(feats) == 4
    loss = (val ** 2).mean() + sum((f * f).mean() for f in feats)
    loss.backward()
    assert _close(val.detach(), G[f"{tag}_val"])
    for i, f in enumerate(feats):
        assert _close(f.detach(), G[f"{tag}_feat{i}"]), i
    assert _close(x.grad, G[f"{tag}_gimg"], 1e-4)
    gw = dict(net.named_parameters())[str(G[f"{tag}_gw_name"])].grad
    assert _close(gw, G[f"{tag}_gw"], 1e-4)
    assert _close(net(img, y=y).detach(), G[f"{tag}_val2"])     # u, v advanced by the first call exactly like the reference's


def test_spectral_norm_weight_is_not_a_parameter():
    from sparenet_b200.dropin.models.sparenet_discriminator import SpectralNorm
    sn = SpectralNorm(torch.nn.Conv2d(3, 4, 3))
    names = [n for n, _ in sn.named_parameters()]
    assert sorted(names) == ["module.bias", "module.weight_bar", "module.weight_u", "module.weight_v"]
    sn(torch.rand(1, 3, 8, 8)).sum().backward()
    assert sn.module.weight_bar.grad is not None and sn.module.weight_u.grad is None
