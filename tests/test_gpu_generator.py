"""GPU parity of the drop-in generator (fused algebra + sm_100a point ops + tcgen05 TF32 GEMMs) against the plain restatement
(oracle/generator_ref.py, itself pinned to the real reference classes) with the SAME state_dict.

Numerics contract.  The reference runs its 1x1 convolutions through cuDNN with TF32 allowed (torch's default for convolutions) and
its nn.Linear layers in fp32; ours runs the same convolutions on its own TF32 tensor-core GEMM (same operand precision, different
accumulation order and fused normalisation algebra) and the Linear layers in fp32.  The algebra itself is verified exactly
(float64, CPU) in tests/test_generator_algebra.py.  Here the tolerance is CALIBRATED instead of guessed: the restatement is run twice,
with cuDNN TF32 on (the reference as shipped) and off (fp32), and the distance between those two runs -- the reference's own TF32
band -- bounds how far ours may be from the fp32 run (factor 3, floor 2e-3 of the coordinate scale)."""
import pytest
import torch

from oracle import generator_ref as G

pytestmark = pytest.mark.gpu


class GpuOps:
    """Point ops for the restatement on the GPU: our kernels, so only the dense algebra differs between the two models."""
    @staticmethod
    def knn(x, k):
        from sparenet_b200 import functional as F_
        return F_.knn_indices(x.contiguous(), k).long()

    @staticmethod
    def expansion(xyz, p, alpha):
        from sparenet_b200.dropin.cuda.expansion_penalty.expansion_penalty_module import expansionPenaltyModule
        return expansionPenaltyModule()(xyz, p, alpha)

    @staticmethod
    def mds(xyz, m, mml):
        from sparenet_b200 import functional as F_
        return F_.mds_sample(xyz.contiguous(), m, mml.contiguous())

    @staticmethod
    def gather(f, idx):
        from sparenet_b200.dropin.cuda.MDS.MDS_module import gather_operation
        return gather_operation(f.contiguous(), idx)


class _tf32:
    """cuDNN TF32 on/off for the restatement; matmul stays fp32 like the reference's nn.Linear layers."""
    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = self.on

    def __exit__(self, *a):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.old


def _band(a_fp32, a_tf32, ours, what, factor=3.0, floor=5e-3):
    """RMS distance to the reference's fp32 run, against the reference's own TF32-vs-fp32 RMS band.  (The maximum over the ~10^5..10^6
    coordinates is printed but not asserted: a TF32-sized perturbation of the features flips a few k-nearest-neighbour sets and
    ReLU / max-pool switches, so single points move by 10-30 % of the cloud's extent between the reference's OWN fp32 and TF32 runs.)"""
    scale = a_fp32.abs().max().item()
    rms = lambda t: t.pow(2).mean().sqrt().item()
    band, err = rms(a_tf32 - a_fp32) / scale, rms(ours - a_fp32) / scale
    print(f"[{what}] scale {scale:.3g}: reference TF32-vs-fp32 band rms {band:.2e} (max {(a_tf32 - a_fp32).abs().max().item() / scale:.2e}), "
          f"ours-vs-fp32 rms {err:.2e} (max {(ours - a_fp32).abs().max().item() / scale:.2e})")
    assert err <= max(factor * band, floor), what
    return band, err


def test_generator_wiring_exact_in_fp32_with_library_gemm(cuda, monkeypatch):
    """The fused algebra + point kernels + module wiring with the dense convolutions on fp32 library GEMMs (the LIBRARY_GEMM
    measurement switch, TF32 off everywhere): agreement with the restatement at fp32 re-association level."""
    from sparenet_b200 import fused
    from sparenet_b200.dropin.models import sparenet_generator as M
    monkeypatch.setattr(M, "LIBRARY_GEMM", True)
    monkeypatch.setattr(fused.tf32_matmul, "enabled", False)     # thin convolutions and Gram products in fp32 too
    kw = dict(n_primitives=8, hide_size=256, bottleneck_size=256, num_points=8 * 512)
    ref = G.SpareNetGenerator(ops=GpuOps, **kw)
    torch.manual_seed(0)
    ref.apply(G.init_weights)
    ref = ref.to(cuda).train()
    mine = M.SpareNetGenerator(use_SElayer=True, use_AdaIn="share", encode="Residualnet", **kw).to(cuda).train()
    mine.load_state_dict(ref.state_dict())
    torch.manual_seed(1)
    data = {"partial_cloud": (torch.rand(4, 1024, 3, device=cuda) - 0.5)}
    with _tf32(False):
        c1, m1, r1, l1 = ref(data)
        c2, m2, r2, l2 = mine(data)
    # B=4 batch-norms amplify fp32 re-association noise (float64 CPU test: 1e-13); hold to 1% of the coordinate scale
    assert torch.allclose(c1, c2, rtol=1e-2, atol=2e-3), (c1 - c2).abs().max()
    assert abs(l1.item() - l2.item()) <= 1e-2 * abs(l1.item()) + 1e-8
    # running statistics advanced identically (sample a few)
    b1, b2 = dict(ref.named_buffers()), dict(mine.named_buffers())
    for k in ("encoder.feat_extractor.bn3.running_var", "decoder.decoder.5.dec.bn2.running_mean", "refine.residual.bn4.running_var",
              "encoder.bn.running_mean", "decoder.decoder.0.dec.bn1.num_batches_tracked"):
        assert torch.allclose(b1[k].float(), b2[k].float(), rtol=2e-2, atol=1e-4), k   # fp32 noise through B=4 batch norms


def test_generator_matches_restatement(cuda):
    from sparenet_b200.dropin.models import sparenet_generator as M
    kw = dict(n_primitives=8, hide_size=256, bottleneck_size=256, num_points=8 * 512)
    ref = G.SpareNetGenerator(ops=GpuOps, **kw)
    torch.manual_seed(0)
    ref.apply(G.init_weights)
    ref = ref.to(cuda).train()
    mine = M.SpareNetGenerator(use_SElayer=True, use_AdaIn="share", encode="Residualnet", **kw).to(cuda).train()
    assert sorted(mine.state_dict()) == sorted(ref.state_dict())
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    mine.load_state_dict(sd0)
    torch.manual_seed(1)
    data = {"partial_cloud": (torch.rand(4, 1024, 3, device=cuda) - 0.5)}
    with _tf32(False):
        c1, m1, r1, l1 = ref(data)
    b1 = {k: v.clone() for k, v in ref.named_buffers()}
    ref.load_state_dict(sd0)
    with _tf32(True):
        c1t, _, _, l1t = ref(data)
    with _tf32(False):
        c2, m2, r2, l2 = mine(data)
    # A batch of FOUR samples: the BatchNorms over the batch amplify a 1e-7 perturbation to 1e-2 (see the fp32 test above), so a
    # TF32-sized one saturates -- the reference's own TF32 run already sits 10 % (max) from its fp32 run.  This case therefore only
    # guards against gross errors; the calibrated comparison is tests/test_gpu_baseline_config.py at B=32.
    _band(c1, c1t, c2, "coarse cloud, n_primitives=8 hide=256 B=4", factor=8.0, floor=0.1)
    assert abs(l1.item() - l2.item()) <= max(3 * abs(l1.item() - l1t.item()), 1e-2 * abs(l1.item())) + 1e-8   # MST edges flip across the alpha threshold
    # running statistics advanced identically (sample a few)
    b2 = dict(mine.named_buffers())
    for k in ("encoder.feat_extractor.bn3.running_var", "decoder.decoder.5.dec.bn2.running_mean", "refine.residual.bn4.running_var",
              "encoder.bn.running_mean", "decoder.decoder.0.dec.bn1.num_batches_tracked"):
        assert torch.allclose(b1[k].float(), b2[k].float(), rtol=1e-1, atol=5e-3), k   # TF32 noise through B=4 batch norms (sanity)
    (r2.mean() + l2).backward()
    assert all(torch.isfinite(p.grad).all() for p in mine.parameters() if p.grad is not None)


def test_refiner_matches_restatement_given_same_inputs(cuda):
    from sparenet_b200.dropin.models import sparenet_generator as M
    ref = G.SpareNetRefine(n_primitives=4, num_points=2048, ops=GpuOps).to(cuda).train()
    torch.manual_seed(2)
    ref.apply(G.init_weights)
    mine = M.SpareNetRefine(n_primitives=4, num_points=2048, use_SElayer=True).to(cuda).train()
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    mine.load_state_dict(sd0)
    coarse = (torch.rand(3, 2048, 3, device=cuda) - 0.5) * 0.8
    partial = (torch.rand(3, 3, 512, device=cuda) - 0.5)
    c1, c1t, c2 = (coarse.clone().requires_grad_() for _ in range(3))
    w = torch.randn(3, 2048, 3, device=cuda)
    with _tf32(False):
        o1, l1 = ref(c1.transpose(1, 2).contiguous(), partial, c1)
        ((o1 * w).sum() + l1).backward()
    ref.load_state_dict(sd0)          # (in-place: only after the first graph has been used)
    ref.zero_grad()
    with _tf32(True):
        o1t, _ = ref(c1t.transpose(1, 2).contiguous(), partial, c1t)
        (o1t * w).sum().backward()
    o2, l2 = mine(c2.transpose(1, 2).contiguous(), partial, c2)
    assert torch.equal(l1, l2)
    _band(o1.detach(), o1t.detach(), o2.detach(), "refined cloud B=3 N=2048")
    ((o2 * w).sum() + l2).backward()
    # 7 BatchNorms over a batch of 3 clouds and ReLU / max-pool switches: a TF32-sized perturbation flips a few of them, so the
    # gradient is compared as a relative L2 distance against the same distance between the reference's own TF32 and fp32 runs
    def rel(a, b):
        return ((a - b).norm() / b.norm()).item()
    band, err = rel(c1t.grad, c1.grad), rel(c2.grad, c1.grad)
    print(f"[refiner input gradient] reference TF32-vs-fp32 relative L2 {band:.2e}, ours-vs-fp32 {err:.2e}")
    assert err <= max(3 * band, 2e-2)


def test_training_step_runs_and_learns(cuda):
    """A few Adam steps of the BASELINE config-2 loss at reduced size: the loss must go down."""
    from sparenet_b200.dropin.cuda.chamfer_distance import ChamferDistance, ChamferDistanceMean
    from sparenet_b200.dropin.models import sparenet_generator as M
    torch.manual_seed(0)
    net = M.SpareNetGenerator(n_primitives=8, hide_size=256, bottleneck_size=256, num_points=4096, use_SElayer=True, use_AdaIn="share",
                              encode="Residualnet")
    net.apply(G.init_weights)
    net = net.to(cuda).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.0, 0.9))
    gt = torch.rand(4, 4096, 3, device=cuda) - 0.5
    partial = gt[:, :1024].contiguous()
    cdm, cd = ChamferDistanceMean(), ChamferDistance()
    losses = []
    for _ in range(6):
        coarse, middle, refine, lm = net({"partial_cloud": partial})
        loss = cdm(coarse, gt) + cdm(middle, gt) + cdm(refine, gt) + lm * 0.1 + cd(refine, gt)[0].mean() * 0.5
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(map(lambda v: v == v, losses)) and losses[-1] < losses[0]
