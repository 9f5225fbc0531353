"""The reference's own GPU path timed beside ours on the same B200 (TEST / MEASUREMENT INFRASTRUCTURE, never the product path).

"Reference" = the reference's CUDA extensions rebuilt UNMODIFIED for sm_100a (oracle/_ref/*.so, recipe oracle/build_ref.py; p2i with
the 4-token .type() -> .scalar_type() shim) called the way the reference's Python wrappers call them (tests/refcalls.py), under the
plain-PyTorch restatement of its generator (oracle/generator_ref.py: unfused, per-edge convs, materialised [B,2C,N,k] graph
features, cuDNN/cuBLAS for the dense math).  The un-vendored knn_cuda wheel is stood in for by the reference's in-repo fallback
formula (models/sparenet_generator.py:872-875) on the GPU.  /root/reference is NOT read at run time (it does not exist on the GPU box).

  step  : configs[1] -- generator forward + 3 x ChamferDistanceMean + 0.1 * expansion + 0.5 * consistency CD, backward, Adam
          (runners/sparenet_runner.py:84-105), B=32, 2048 -> 16384 points; CUDA events, 2 warm-up + 3 timed steps.
  ops   : CD fwd+bwd, EMD fwd+bwd (eps 0.005, 50 iterations), the 8-view 256x256 ComputeDepthMaps render fwd+bwd at radius 5 --
          the "CD+EMD+p2i ms/batch" half of BASELINE.json's metric -- through the reference extensions and the reference's own
          host-side protocol (utils/p2i_utils.py:211-252 restated below, including its per-call expand of a CPU matrix).
bench.py runs this module in a subprocess (`bench.py --impl reference-gpu`) and prints its numbers beside ours.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402
from oracle import generator_ref as G  # noqa: E402
from tests import refcalls  # noqa: E402

N_OUT, N_PARTIAL, N_PRIM = 16384, 2048, 32


def _ext(name):
    if not build_ref.available(name):
        raise RuntimeError(f"oracle/_ref/{name}.so is missing: run oracle/build_ref.py where /root/reference exists")
    return build_ref.load_ref(name)


# ------------------------------------------------------------------------------------------------ the step
def make_ref_step(dev, B):
    e_exp, e_mds, e_ch = _ext("expansion_penalty"), _ext("MDS"), _ext("chamfer")

    class RefOps:
        knn = staticmethod(G.CpuOps.knn)   # the reference's fallback formula (:872-875), on the GPU

        @staticmethod
        def expansion(xyz, p, alpha):      # cuda/expansion_penalty/expansion_penalty_module.py:24-48
            class _Fn(torch.autograd.Function):
                @staticmethod
                def forward(ctx, xyz):
                    dist, idx, mml = refcalls.expansion_fwd(e_exp, xyz.detach().contiguous(), p, alpha)
                    ctx.save_for_backward(xyz.detach(), idx)
                    ctx.mark_non_differentiable(idx, mml)
                    return dist, idx, mml

                @staticmethod
                def backward(ctx, g, _a, _b):
                    xyz, idx = ctx.saved_tensors
                    return refcalls.expansion_bwd(e_exp, xyz.contiguous(), g.contiguous(), idx)
            return _Fn.apply(xyz)

        @staticmethod
        def mds(xyz, npoint, mml):         # cuda/MDS/MDS_module.py:9-34
            return refcalls.mds(e_mds, xyz.detach().contiguous(), npoint, mml.detach().contiguous())

        @staticmethod
        def gather(features, idx):         # cuda/MDS/MDS_module.py:44-75
            class _Fn(torch.autograd.Function):
                @staticmethod
                def forward(ctx, f, idx):
                    ctx.save_for_backward(idx)
                    ctx.n = f.size(2)
                    return e_mds.gather_points(f.contiguous(), idx)

                @staticmethod
                def backward(ctx, g):
                    (idx,) = ctx.saved_tensors
                    return e_mds.gather_points_grad(g.contiguous(), idx, ctx.n), None
            return _Fn.apply(features, idx)

    class RefChamfer(torch.autograd.Function):   # cuda/chamfer_dist/__init__.py:6-18
        @staticmethod
        def forward(ctx, a, b):
            d1, d2, i1, i2 = e_ch.forward(a, b)
            ctx.save_for_backward(a, b, i1, i2)
            return d1, d2

        @staticmethod
        def backward(ctx, g1, g2):
            a, b, i1, i2 = ctx.saved_tensors
            ga, gb = e_ch.backward(a, b, i1, i2, g1.contiguous(), g2.contiguous())
            return ga, gb

    torch.manual_seed(0)
    net = G.SpareNetGenerator(n_primitives=N_PRIM, hide_size=4096, bottleneck_size=4096, num_points=N_OUT, ops=RefOps)
    net.apply(G.init_weights)
    net = net.to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.0, 0.9))

    def cdm(a, b):
        d1, d2 = RefChamfer.apply(a.contiguous(), b)
        return d1.mean() + d2.mean()

    def step(partial, gt):
        coarse, middle, refine, loss_mst = net({"partial_cloud": partial})
        loss = cdm(coarse, gt) + cdm(middle, gt) + cdm(refine, gt) + loss_mst.mean() * 0.1
        d1, _ = RefChamfer.apply(refine.contiguous(), gt)
        loss = loss + d1.mean() * 0.5
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss
    step.net = net
    return step


def time_step(step, partial, gt, warm, reps):
    losses = []
    for _ in range(warm):
        losses.append(float(step(partial, gt)))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        loss = step(partial, gt)
    b.record()
    torch.cuda.synchronize()
    losses.append(float(loss))
    return a.elapsed_time(b) / reps, losses


# ------------------------------------------------------------------------------------------------ the loss ops
class RefDepthMaps:
    """utils/p2i_utils.py:168-252 + cuda/p2i_op/__init__.py:59-131 restated over oracle/_ref/ext.so.  The eight pre-matrices are the
    reference's own (tests/golden/depthmaps_ref.npz, written by tests/golden/make_golden_depthmaps.py from the real module) and stay on
    the CPU like the reference's `pre_matrix_list` (:208-209), so every call expands one to [B*N,4,4] on the host and uploads it (:217)."""

    def __init__(self, ext, image_size=256):
        pre = np.load(os.path.join(ROOT, "tests", "golden", "depthmaps_ref.npz"))["orthorgonal_pre"]
        self.pre = [torch.from_numpy(pre[i:i + 1].copy()) for i in range(8)]
        self.ext, self.S = ext, image_size

    def __call__(self, data, view_id, radius):
        ext, S = self.ext, self.S
        B, N = data.size(0), data.size(1)
        matrix = self.pre[view_id].expand(B * N, 4, 4).to(data.device)                      # :217
        background = torch.zeros(B, 1, S, S, dtype=data.dtype, device=data.device)          # :218
        binds = torch.arange(0, B, dtype=torch.int32, device=data.device).unsqueeze(1).expand(B, N).reshape(-1)   # :219-220
        pcds = data.view(-1, 3)
        out = torch.cat([pcds, torch.ones_like(pcds[:, [0]])], dim=1).view(-1, 4, 1)        # transform :153-165
        out = matrix @ out
        pos = out[:, :3, 0] / out[:, [3], 0]
        xs, ys, zs = pos.split(dim=1, split_size=1)
        ijs = torch.cat([-ys, xs], dim=1)                                                   # :225
        feat = 1.0 - (zs - zs.min()) / (zs.max() - zs.min())                                # :226
        pts = (ijs + 1) / 2 * torch.tensor([S - 1, S - 1], dtype=ijs.dtype, device=ijs.device).view(1, 2)   # p2i_op/__init__.py:116-121

        class _Max(torch.autograd.Function):                                                # P2IMaxFunction :59-93
            @staticmethod
            def forward(ctx, points, point_features, batch_inds, bg):
                o, ids = ext.p2i_max_forward_gpu(points.contiguous(), point_features.contiguous(), batch_inds.contiguous(), bg.contiguous(), 0, radius)
                ctx.save_for_backward(points, point_features, ids)
                return o

            @staticmethod
            def backward(ctx, og):
                points, point_features, ids = ctx.saved_tensors
                gp, gf, gb = ext.p2i_max_backward_gpu(og.contiguous(), ids, points, point_features, 0, radius)
                return gp, gf, None, gb
        return _Max.apply(pts, feat, binds, background)


def _timed(fn, warm=2, reps=3):
    ts = []
    for i in range(warm + reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def ref_ops_ms(dev, B):
    """Same inputs as bench.py's aux_ops_ms (seed 4, iid U[0,1)^3 clouds; the render on the centred cloud)."""
    e_ch, e_emd, e_p2i = _ext("chamfer"), _ext("emd"), _ext("ext")
    g = torch.Generator(device=dev).manual_seed(4)
    x = torch.rand(B, N_OUT, 3, device=dev, generator=g)
    y = torch.rand(B, N_OUT, 3, device=dev, generator=g)
    out = {"B": B, "N": N_OUT}

    def f_cd():   # ChamferDistanceMean fwd + bwd (cuda/chamfer_dist/__init__.py:8-18)
        d1, d2, i1, i2 = refcalls.chamfer_fwd(e_ch, x, y)
        g1, g2 = torch.full_like(d1, 1.0 / d1.numel()), torch.full_like(d2, 1.0 / d2.numel())
        refcalls.chamfer_bwd(e_ch, x, y, i1, i2, g1, g2)
    out["cd_fwd_bwd_ms"] = _timed(f_cd)

    def f_emd():  # emdModule fwd + bwd (cuda/emd/emd_module.py:31-87), loss sqrt(dist).mean(1).mean()
        d, a = refcalls.emd_fwd(e_emd, x, y, 0.005, 50)
        gd = 0.5 / torch.sqrt(d) / d.numel()
        refcalls.emd_bwd(e_emd, x, y, gd.contiguous(), a)
    out["emd_fwd_bwd_ms"] = _timed(f_emd, warm=1, reps=3)

    render = RefDepthMaps(e_p2i)
    xc = (x - 0.5).requires_grad_()

    def f_p2i():
        xc.grad = None
        torch.cat([render(xc, v, 5.0) for v in range(8)], 1).mean().backward()
    out["p2i_8view_fwd_bwd_ms"] = _timed(f_p2i)
    return out


def run(B=32, warm=2, reps=3, ops=True):
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    # the reference runs its convolutions through cuDNN with TF32 allowed (torch default) and its nn.Linear layers in fp32
    torch.backends.cudnn.allow_tf32 = True
    gp, gg = torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)
    partial = (torch.rand(B, N_PARTIAL, 3, generator=gp) - 0.5).to(dev)
    gt = (torch.rand(B, N_OUT, 3, generator=gg) - 0.5).to(dev)
    torch.cuda.reset_peak_memory_stats()
    step = make_ref_step(dev, B)
    ms, losses = time_step(step, partial, gt, warm, reps)
    res = {"ms_per_step": ms, "value": B / ms * 1e3, "unit": "completions/s", "B": B, "warmup": warm, "steps": reps, "losses": losses,
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
           "what": "the reference's CUDA extensions rebuilt for sm_100a (oracle/_ref/*.so) under the plain-PyTorch restatement of its generator "
                   "(cuDNN/cuBLAS dense math, kNN = its in-repo matmul+topk fallback), same inputs / losses / Adam as our arm"}
    del step
    torch.cuda.empty_cache()
    if ops:
        try:
            res["ops_ms_per_batch"] = ref_ops_ms(dev, B)
        except Exception as e:   # the step number stands on its own
            res["ops_ms_per_batch"] = {"error": repr(e)[:200]}
    return res


if __name__ == "__main__":
    print(json.dumps(run(B=int(os.environ.get("PB", "32")))), flush=True)
