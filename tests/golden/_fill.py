"""Seed-free parameter/buffer fill keyed on the tensor NAME: the real reference classes (golden scripts, build container) and
the drop-in modules (tests) receive bit-identical values without a stored checkpoint."""
import math

import torch


def name_fill(module):
    with torch.no_grad():
        for name, p in sorted(list(module.named_parameters()) + list(module.named_buffers())):
            if "num_batches_tracked" in name or not p.is_floating_point():
                continue
            h = sum((i + 1) * ord(ch) for i, ch in enumerate(name)) % 9973
            v = torch.sin(torch.arange(p.numel(), dtype=torch.float64) * (0.37 + 0.001 * (h % 211)) + h)
            if name.endswith("running_var"):
                v = 1.0 + 0.2 * v
            elif p.dim() <= 1 and not name.endswith(("_u", "_v")):
                v = (1.0 + 0.2 * v) if name.endswith("weight") else 0.1 * v     # norm scales around 1, biases small
            elif p.dim() >= 2:
                v = v * (1.5 / math.sqrt(p[0].numel()))
            p.copy_(v.view_as(p).to(p.dtype))
