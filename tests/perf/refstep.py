"""The reference's step on the SAME GPU, next to ours (development measurement; writes gpurun_out/refstep.json).

"Reference" here = the reference's own CUDA extensions rebuilt for sm_100a (oracle/_ref/*.so: expansion_penalty, MDS incl.
gather_points, chamfer) called the way its Python wrappers call them (tests/refcalls.py), under the plain-PyTorch
restatement of its generator (oracle/generator_ref.py: unfused, per-edge convs, [B,2C,N,k] graph features, cuDNN/cuBLAS
for the dense math) with the in-repo kNN fallback formula (models/sparenet_generator.py:872-875) standing in for the
un-vendored knn_cuda wheel.  Same workload as bench.py (configs[1]): B=32, 2048 -> 16384 points, 3 x ChamferDistanceMean
+ 0.1 * expansion + 0.5 * consistency CD, backward, Adam.  CUDA events, 2 warm-up + 3 timed steps.
Test infrastructure: imports oracle/; never part of the product path or of bench.py's GPU arm."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import build_ref  # noqa: E402
from oracle import generator_ref as G  # noqa: E402
from tests import refcalls  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
B = int(os.environ.get("PB", "32"))
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
e_exp, e_mds, e_ch = build_ref.load_ref("expansion_penalty"), build_ref.load_ref("MDS"), build_ref.load_ref("chamfer")


class RefOps:
    knn = staticmethod(G.CpuOps.knn)   # the reference's own fallback formula, on the GPU

    @staticmethod
    def expansion(xyz, p, alpha):
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
    def mds(xyz, npoint, mml):
        return refcalls.mds(e_mds, xyz.detach().contiguous(), npoint, mml.detach().contiguous())

    @staticmethod
    def gather(features, idx):
        class _Fn(torch.autograd.Function):   # cuda/MDS/MDS_module.py:44-75
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


def make_ref():
    torch.manual_seed(0)
    net = G.SpareNetGenerator(n_primitives=32, hide_size=4096, bottleneck_size=4096, num_points=16384, ops=RefOps)
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
    return step


def timed(step, p, g, warm=2, reps=3):
    for _ in range(warm):
        step(p, g)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        loss = step(p, g)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, float(loss)


class A:
    batch = B


gp = torch.Generator().manual_seed(1)
gg = torch.Generator().manual_seed(2)
partial = (torch.rand(B, 2048, 3, generator=gp) - 0.5).to(dev)
gt = (torch.rand(B, 16384, 3, generator=gg) - 0.5).to(dev)
out = {"B": B}
ref_ms, ref_loss = timed(make_ref(), partial, gt)
out.update(ref_ms_per_step=ref_ms, ref_completions_per_s=B / ref_ms * 1e3, ref_first_losses=ref_loss, ref_peak_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
print(json.dumps(out), flush=True)
torch.cuda.empty_cache()
torch.cuda.reset_peak_memory_stats()
ours, _, _ = bench.build_gpu(A, dev, 0)
our_ms, our_loss = timed(ours, partial, gt, warm=3, reps=5)
out.update(ours_eager_ms_per_step=our_ms, ours_completions_per_s=B / our_ms * 1e3, ours_loss=our_loss, speedup_eager=ref_ms / our_ms,
           ours_peak_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
print(json.dumps(out), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "refstep.json"), "w"), indent=1)
