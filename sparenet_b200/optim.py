"""FlatAdam: torch.optim.Adam's update on ONE flat parameter arena, one launch per step (csrc/optim.cu, C ABI snb_adam_flat).

The reference trains with torch.optim.Adam (runners/sparenet_runner.py:31-60 via utils/model_init.py).  Its ~800 parameter tensors
make even the fused multi-tensor implementation 40 launches at a third of the HBM rate; here the parameters are MOVED into one
16-byte-aligned flat buffer with the layout of sparenet_b200.dist.GradArena (every p.data / p.grad becomes a view), and both moments
are flat too.  Same update rule, same hyper-parameters; `state_dict()` is not torch.optim.Adam's (moments are two flat tensors).
"""
import ctypes

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .dist import GradArena
from .functional import _op


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, arena=None):
        self.arena = arena if arena is not None else GradArena(list(params))
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.step_count = 0
        self.flat_params, self.exp_avg, self.exp_avg_sq = [], [], []
        groups = {}
        for p in self.arena.params:
            groups.setdefault((p.dtype, p.device), []).append(p)
        for gflat, ((dt, dev), ps) in zip(self.arena.flats, groups.items()):
            if dt != torch.float32 or dev.type != "cuda":
                raise ValueError("FlatAdam serves CUDA float32 parameters (sparenet_b200 has no CPU path)")
            flat = torch.zeros_like(gflat)
            o = 0
            with torch.no_grad():
                for p in ps:
                    n = p.numel()
                    assert self.arena.views[id(p)].data_ptr() == gflat.data_ptr() + 4 * o, "parameter order must match the gradient arena"
                    flat[o:o + n].copy_(p.detach().reshape(-1))
                    p.data = flat[o:o + n].view_as(p)          # the parameter now lives in the arena
                    o += (n + 3) // 4 * 4
            self.flat_params.append(flat)
            self.exp_avg.append(torch.zeros_like(flat))
            self.exp_avg_sq.append(torch.zeros_like(flat))

    def zero_grad(self, set_to_none=True):
        if self.arena.own_grads:
            self.arena.zero()                                # the gradients are views into the arena: never set to None
        else:
            for p in self.arena.params:                      # fresh gradient tensors every backward; arena.pack() gathers them
                p.grad = None

    @torch.no_grad()
    def step(self, packed=False):
        """packed=True: the caller has already run arena.pack() (e.g. captured at the end of the step's CUDA graph)."""
        if not self.arena.own_grads and not packed:
            self.arena.pack()
        self.step_count += 1
        lib = _lib.load()
        for p, g, m, v in zip(self.flat_params, self.arena.flats, self.exp_avg, self.exp_avg_sq):
            with torch.cuda.device(p.device), _op("adam_flat", 1, 28 * p.numel()):
                check(lib.snb_adam_flat(ptr(p), ptr(g), ptr(m), ptr(v), ctypes.c_size_t(p.numel()), self.lr, self.betas[0], self.betas[1], self.eps,
                                        self.weight_decay, self.step_count, stream_ptr()), "adam_flat")

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "hyper": {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        for dst, src in zip(self.exp_avg, sd["exp_avg"]):
            dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, sd["exp_avg_sq"]):
            dst.copy_(src)
