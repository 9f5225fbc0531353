"""Drop-in for the part of the reference's utils/model_init.py the generator path needs: `init_weights`
(reference utils/model_init.py:137-159), the initialisation recipe `generator_init` applies with `net.apply(init_weights)`.
Host-side parameter initialisation only: Conv2d/3d (+transposed) kaiming-normal with zero bias, Conv1d N(0, 0.02),
BatchNorm2d/3d weight 1 bias 0, BatchNorm1d weight N(1, 0.02) bias 0, Linear N(0, 0.01) with zero bias.
`init_weights_D` (:162-178) is the discriminator's recipe (Conv N(0, 0.02); BatchNorm weight N(1, 0.02), bias 0).
The builders around them (optimizers, schedulers, config plumbing) are out of scope."""
import torch.nn as nn

_KAIMING = (nn.Conv2d, nn.ConvTranspose2d, nn.Conv3d, nn.ConvTranspose3d)


def init_weights(m):
    t = type(m)
    if t in _KAIMING and hasattr(m, "weight"):
        nn.init.kaiming_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    if t is nn.Conv1d:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif t in (nn.BatchNorm2d, nn.BatchNorm3d):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)
    elif t is nn.BatchNorm1d:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0.0)
    elif t is nn.Linear and hasattr(m, "weight"):
        nn.init.normal_(m.weight, 0, 0.01)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)


def init_weights_D(m):
    """Reference utils/model_init.py:162-178: matched by class NAME (so the SpectralNorm-wrapped convs, whose `weight` attribute is
    deleted in favour of weight_bar, are skipped by the hasattr test exactly like in the reference)."""
    name = m.__class__.__name__
    if ("Conv2d" in name or "Conv1d" in name) and hasattr(m, "weight"):
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    if "BatchNorm2d" in name or "BatchNorm1d" in name:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0.0)
