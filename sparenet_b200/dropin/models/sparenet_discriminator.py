"""Drop-in for the reference's models/sparenet_discriminator.py: the image discriminators of the GAN step.

ProjectionD(num_classes, img_shape)(img [B, 2*views, S, S], feat=False, y=None) -> validity [B,1] (and the four feature maps with
feat=True) -- reference :84-149 -- and PatchDiscriminator (:13-81), both over the reference's hand-rolled SpectralNorm wrapper
(:156-211), with IDENTICAL state_dict keys (conv1.0.module.weight_bar / weight_u / weight_v, adv_layer.weight_orig, l_y.weight_orig ...)
so reference checkpoints load unchanged.  The 0.4 M-parameter discriminator is 3x3 / 4x4 strided convolutions on 256x256 images:
its dense math stays on cuDNN (SURVEY.md 2 #13 -- not a kernel target); what is ours on this side of the GAN step is the renderer
that feeds it (snb_depthmaps_*, utils/p2i_utils.py).

SpectralNorm keeps the reference's mechanics: one power iteration per forward through `.data` (no autograd edge, no version
bump), sigma = u . (W v) differentiated w.r.t. W, the normalised weight set as a plain attribute of the wrapped module.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init, utils


def l2normalize(v, eps=1e-12):
    return v / (v.norm() + eps)


class SpectralNorm(nn.Module):  # reference :156-211
    def __init__(self, module, name="weight", power_iterations=1):
        super().__init__()
        self.module, self.name, self.power_iterations = module, name, power_iterations
        if not hasattr(module, name + "_bar"):
            w = getattr(module, name)
            height = w.shape[0]
            width = w.view(height, -1).shape[1]
            u = nn.Parameter(l2normalize(w.data.new(height).normal_(0, 1)), requires_grad=False)
            v = nn.Parameter(l2normalize(w.data.new(width).normal_(0, 1)), requires_grad=False)
            w_bar = nn.Parameter(w.data)
            del module._parameters[name]
            module.register_parameter(name + "_u", u)
            module.register_parameter(name + "_v", v)
            module.register_parameter(name + "_bar", w_bar)

    def normalised_weight(self):
        u, v, w = (getattr(self.module, self.name + s) for s in ("_u", "_v", "_bar"))
        w2 = w.view(w.shape[0], -1)
        # like the reference (:171-173) the iterates REPLACE u.data / v.data: an in-place update would invalidate the sigma graph
        # of an earlier forward that is still waiting for its backward (the D step runs the discriminator twice before backward)
        for _ in range(self.power_iterations):
            v.data = l2normalize(torch.mv(w2.data.t(), u.data))
            u.data = l2normalize(torch.mv(w2.data, v.data))
        sigma = u.dot(w2.mv(v))
        return w / sigma.expand_as(w)

    def forward(self, *args):
        setattr(self.module, self.name, self.normalised_weight())   # like the reference: a plain attribute, not a Parameter
        return self.module.forward(*args)


def _block(cin, cout, kernel, bn_first, bn):
    """One down-sampling block; the two discriminators order normalisation / activation / dropout differently."""
    conv = SpectralNorm(nn.Conv2d(cin, cout, kernel, stride=2, padding=1))
    if bn_first:    # PatchDiscriminator :31-41: conv, [BatchNorm2d], LeakyReLU
        return nn.Sequential(*([conv] + ([nn.BatchNorm2d(cout)] if bn else []) + [nn.LeakyReLU(0.2, inplace=True)]))
    # ProjectionD :103-112: conv, LeakyReLU, Dropout2d(0.25), [BatchNorm2d(cout, eps=0.8)] (0.8 is the positional eps)
    return nn.Sequential(*([conv, nn.LeakyReLU(0.2, inplace=True), nn.Dropout2d(0.25)] + ([nn.BatchNorm2d(cout, 0.8)] if bn else [])))


class PatchDiscriminator(nn.Module):  # reference :13-81
    def __init__(self, img_shape: tuple = (2, 256, 256)):
        super().__init__()
        widths = [img_shape[0], 16, 32, 64, 128, 256, 512]
        for i in range(6):
            setattr(self, f"conv{i + 1}", _block(widths[i], widths[i + 1], 4, True, i > 0))
        self.adv_layer = SpectralNorm(nn.Conv2d(512, 1, 3, padding=1, bias=False))

    def forward(self, img, feat=False, y=None):
        feats, x = [], img
        for i in range(6):
            x = getattr(self, f"conv{i + 1}")(x)
            feats.append(x)
        validity = self.adv_layer(x)
        validity = F.avg_pool2d(validity, validity.size()[2:]).view(validity.size(0), -1)
        return (validity, feats[:4]) if feat else validity


class ProjectionD(nn.Module):  # reference :84-149
    def __init__(self, num_classes: int = 0, img_shape: tuple = (2, 256, 256)):
        super().__init__()
        widths = [img_shape[0], 16, 32, 64, 128]
        for i in range(4):
            setattr(self, f"conv{i + 1}", _block(widths[i], widths[i + 1], 3, False, i > 0))
        ds_size = img_shape[1] // 2 ** 4
        self.adv_layer = utils.spectral_norm(nn.Linear(widths[-1] * ds_size ** 2, 1))
        if num_classes > 0:
            self.l_y = utils.spectral_norm(nn.Embedding(num_classes, widths[-1] * ds_size ** 2))
        self._initialize()

    def _initialize(self):
        init.xavier_uniform_(self.adv_layer.weight.data)
        if getattr(self, "l_y", None) is not None:
            init.xavier_uniform_(self.l_y.weight.data)

    def forward(self, img, feat=False, y=None):
        feats, x = [], img
        for i in range(4):
            x = getattr(self, f"conv{i + 1}")(x)
            feats.append(x)
        out = x.view(x.shape[0], -1)
        validity = self.adv_layer(out)
        if y is not None:
            validity = validity + torch.sum(self.l_y(y) * out, dim=1, keepdim=True)   # projection term (:143-144)
        return (validity, feats) if feat else validity
