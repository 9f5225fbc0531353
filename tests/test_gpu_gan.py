"""GPU smoke + invariants of the GAN training step drop-in (sparenet_b200/dropin/runners/sparenet_gan_runner.py) at reduced size:
both optimizers move their parameters, the D step does not touch the generator, every logged loss is finite, and the step is
deterministic given the radius and the dropout seed."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(cuda, seed=0):
    from oracle import generator_ref as G
    from sparenet_b200.dropin.models.sparenet_discriminator import ProjectionD
    from sparenet_b200.dropin.models.sparenet_generator import SpareNetGenerator
    from sparenet_b200.dropin.runners.sparenet_gan_runner import sparenetGANStep
    from sparenet_b200.dropin.utils.model_init import init_weights_D
    from sparenet_b200.dropin.utils.p2i_utils import ComputeDepthMaps
    torch.manual_seed(seed)
    net = SpareNetGenerator(n_primitives=8, hide_size=256, bottleneck_size=256, num_points=4096, use_SElayer=True, use_AdaIn="share",
                            encode="Residualnet")
    net.apply(G.init_weights)
    net = net.to(cuda).train()
    net_D = ProjectionD(num_classes=8, img_shape=(16, 64, 64))
    net_D.apply(init_weights_D)
    net_D = net_D.to(cuda).train()
    renderer = ComputeDepthMaps("orthorgonal", 1.0, 64).to(cuda)
    oG = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.0, 0.9))
    oD = torch.optim.Adam([p for p in net_D.parameters() if p.requires_grad], lr=1e-4, betas=(0.0, 0.9))
    return net, net_D, sparenetGANStep(net, net_D, renderer, oG, oD)


@pytest.mark.parametrize("metric", ["emd", "chamfer"])
def test_gan_step_updates_both_networks(cuda, metric):
    net, net_D, gan = _build(cuda)
    gan.metric = metric
    torch.manual_seed(5)
    gt = torch.rand(2, 4096, 3, device=cuda) - 0.5
    data = {"partial_cloud": gt[:, :1024].contiguous(), "gtcloud": gt}
    labels = torch.tensor([1, 6], device=cuda)
    g0 = [p.detach().clone() for p in net.parameters()]
    d0 = [p.detach().clone() for p in net_D.parameters()]
    loss = gan.train_step(data, labels, radius=5.0)
    assert set(loss) == {"coarse_loss", "refine_loss", "rec_loss", "errG", "errG_D", "errD_real", "errD_fake"}
    assert all(torch.isfinite(v).all() for v in loss.values())
    assert gan.fake_imgs.shape == (2, 8, 64, 64) and gan.real_imgs.shape == (2, 8, 64, 64) and gan.input_imgs.shape == (2, 8, 64, 64)
    assert gan.fake_imgs.requires_grad and 0 <= float(gan.real_imgs.min()) and float(gan.real_imgs.max()) <= 1.0 + 1e-6
    moved_g = sum(int(not torch.equal(a, b)) for a, b in zip(g0, net.parameters()))
    moved_d = sum(int(not torch.equal(a, b)) for a, b in zip(d0, net_D.parameters()) if b.requires_grad)
    assert moved_g > 100 and moved_d >= 8
    loss2 = gan.train_step(data, labels, radius=7.0)
    assert all(torch.isfinite(v).all() for v in loss2.values())


def test_gan_step_is_deterministic(cuda):
    outs = []
    for _ in range(2):
        net, net_D, gan = _build(cuda, seed=3)
        gan.metric = "chamfer"
        torch.manual_seed(9)
        gt = torch.rand(2, 4096, 3, device=cuda) - 0.5
        data = {"partial_cloud": gt[:, :1024].contiguous(), "gtcloud": gt}
        loss = gan.train_step(data, torch.tensor([0, 3], device=cuda), radius=5.0)
        outs.append({k: float(v) for k, v in loss.items()})
    for k in outs[0]:
        assert abs(outs[0][k] - outs[1][k]) <= 2e-3 * abs(outs[0][k]) + 1e-7, (k, outs)   # float-atomic summation order only
