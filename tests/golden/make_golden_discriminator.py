"""Generates tests/golden/discriminator_ref.npz by running the REAL reference classes of
/root/reference/models/sparenet_discriminator.py on CPU (build container only): ProjectionD (with the class-projection term) and
PatchDiscriminator on small images, every parameter and buffer (the spectral-norm u, v vectors included) filled by tests/golden/_fill.py:name_fill, so no state dict
is stored.  The input, the outputs, the gradient w.r.t. the input and one weight gradient are stored; dropout is switched off (p = 0) so that train-mode BatchNorm (eps = 0.8 in ProjectionD) is pinned too.
The GPU box never runs this."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import models.sparenet_discriminator as R  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _fill import name_fill  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}


def run(tag, net, img, y):
    name_fill(net)
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    net.train()
    out[f"{tag}_keys"] = np.array(sorted(net.state_dict().keys()))
    x = img.clone().requires_grad_()
    val, feats = net(x, feat=True, y=y)
    loss = (val ** 2).mean() + sum((f * f).mean() for f in feats)
    loss.backward()
    out[f"{tag}_img"] = img.numpy()
    out[f"{tag}_val"] = val.detach().numpy()
    for i, f in enumerate(feats):
        out[f"{tag}_feat{i}"] = f.detach().numpy()
    out[f"{tag}_gimg"] = x.grad.numpy()
    name, p = next((n, p) for n, p in net.named_parameters() if n.endswith("weight_bar"))
    out[f"{tag}_gw"] = p.grad.numpy().copy()
    out[f"{tag}_gw_name"] = np.array(name)
    # second forward: u, v have advanced once more (state carried between calls)
    out[f"{tag}_val2"] = net(img, y=y).detach().numpy()


torch.manual_seed(0)
img = torch.rand(3, 16, 32, 32)
y = torch.tensor([1, 5, 2])
out["proj_y"] = y.numpy()
run("proj", R.ProjectionD(num_classes=8, img_shape=(16, 32, 32)), img, y)
torch.manual_seed(1)
img2 = torch.rand(2, 16, 64, 64)
run("patch", R.PatchDiscriminator(img_shape=(16, 64, 64)), img2, None)
np.savez_compressed(os.path.join(HERE, "discriminator_ref.npz"), **out)
print("wrote", len(out), "arrays", sum(v.nbytes for v in out.values()) / 1e6, "MB")
