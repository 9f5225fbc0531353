"""Drop-in for the part of the reference's utils/model_init.py the generator path needs: `init_weights`
(reference utils/model_init.py:137-159), the initialisation recipe `generator_init` applies with `net.apply(init_weights)`.
Host-side parameter initialisation only: Conv2d/3d (+transposed) kaiming-normal with zero bias, Conv1d N(0, 0.02),
BatchNorm2d/3d weight 1 bias 0, BatchNorm1d weight N(1, 0.02) bias 0, Linear N(0, 0.01) with zero bias.
The builders around it (optimizers, schedulers, the discriminator) are out of scope."""
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
