"""Plain PyTorch restatement of the reference's style-based generator -- TEST INFRASTRUCTURE ONLY.

Follows models/sparenet_generator.py of the reference module by module (citations below), unfused and
device-agnostic, with IDENTICAL parameter / buffer names so a state_dict moves freely between the
reference, this restatement and the product (sparenet_b200/dropin/models/sparenet_generator.py).
It exists to (1) be pinned against goldens produced by the real reference classes on CPU
(tests/golden/make_golden_generator.py), (2) check the fused GPU generator, (3) serve as the CPU arm of
bench.py.  The point ops (expansion penalty, MDS, gather, kNN) are injected: the CPU oracle's by default.

Only the configuration the shipped YAML uses is restated (configs/sparenet.yaml:18-24):
encode="Residualnet", use_AdaIn="share", use_SElayer=True.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---- injected point ops (CPU oracle by default) -------------------------------------------------------
class CpuOps:
    @staticmethod
    def knn(x, k):  # models/sparenet_generator.py:871-875 (the reference's own CPU branch)
        inner = -2 * torch.matmul(x.transpose(2, 1), x)
        xx = torch.sum(x ** 2, dim=1, keepdim=True)
        return (-xx - inner - xx.transpose(2, 1)).topk(k=k, dim=-1)[1]

    @staticmethod
    def expansion(xyz, primitive_size, alpha):
        import oracle

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, xyz):
                dist, idx, mml = oracle.expansion_fwd(xyz.detach().contiguous(), primitive_size, alpha)
                ctx.save_for_backward(xyz.detach(), idx)
                ctx.mark_non_differentiable(idx, mml)
                return dist, idx, mml

            @staticmethod
            def backward(ctx, g, _gi, _gm):
                xyz, idx = ctx.saved_tensors
                return oracle.expansion_bwd(xyz.contiguous(), g.contiguous(), idx)
        return _Fn.apply(xyz)

    @staticmethod
    def mds(xyz, npoint, mml):
        import oracle
        return oracle.mds(xyz.detach().contiguous(), npoint, mml.detach().contiguous())

    @staticmethod
    def gather(features, idx):
        return torch.gather(features, 2, idx.long().unsqueeze(1).expand(-1, features.size(1), -1))


# ---- building blocks ------------------------------------------------------------------------------------
class SELayer(nn.Module):  # :741-764 (2-D) and :767-790 (1-D) share this body up to the pooled dims
    def __init__(self, channel, reduction=16):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())

    def forward(self, x):
        b, c = x.shape[:2]
        y = x.reshape(b, c, -1).mean(-1)
        y = self.fc(y).view(b, c, *([1] * (x.dim() - 2)))
        return x * y


SELayer1D = SELayer


def get_graph_feature(x, k, knn_fn):  # :880-906
    B, C, N = x.shape
    idx = knn_fn(x, k).long()                                   # [B, N, k]
    xt = x.transpose(2, 1).contiguous()                         # [B, N, C]
    flat = (idx + torch.arange(B, device=x.device).view(B, 1, 1) * N).view(-1)    # :893-899: batch offset, flat row gather
    nb = xt.view(B * N, C)[flat].view(B, N, k, C)
    ctr = xt.unsqueeze(2).expand(-1, -1, k, -1)
    return torch.cat((nb - ctr, ctr), dim=3).permute(0, 3, 1, 2).contiguous()  # [B, 2C, N, k]


class EdgeConvResFeat(nn.Module):  # :123-242 (use_SElayer=True branch :191-210)
    def __init__(self, use_SElayer=True, k=8, hide_size=2048, output_size=4096, ops=CpuOps):
        super().__init__()
        assert use_SElayer
        self.k, self.output_size, self.ops = k, output_size, ops
        h = hide_size
        self.conv1 = nn.Conv2d(6, h // 16, 1, bias=False)
        self.conv2 = nn.Conv2d(h // 8, h // 16, 1, bias=False)
        self.conv3 = nn.Conv2d(h // 8, h // 8, 1, bias=False)
        self.conv4 = nn.Conv2d(h // 4, h // 4, 1, bias=False)
        self.conv5 = nn.Conv1d(h // 2, output_size // 2, 1, bias=False)
        self.se1, self.se2, self.se3, self.se4 = SELayer(h // 16), SELayer(h // 16), SELayer(h // 8), SELayer(h // 4)
        self.bn1, self.bn2, self.bn3, self.bn4 = nn.BatchNorm2d(h // 16), nn.BatchNorm2d(h // 16), nn.BatchNorm2d(h // 8), nn.BatchNorm2d(h // 4)
        self.bn5 = nn.BatchNorm1d(output_size // 2)
        self.resconv1 = nn.Conv1d(h // 16, h // 16, 1, bias=False)
        self.resconv2 = nn.Conv1d(h // 16, h // 8, 1, bias=False)
        self.resconv3 = nn.Conv1d(h // 8, h // 4, 1, bias=False)

    def _block(self, x, conv, bn, se):
        f = get_graph_feature(x, self.k, self.ops.knn)
        return F.leaky_relu(se(bn(conv(f))), 0.2).max(dim=-1)[0]

    def forward(self, x):
        B = x.size(0)
        x1 = self._block(x, self.conv1, self.bn1, self.se1)
        x2 = self._block(x1, self.conv2, self.bn2, self.se2) + self.resconv1(x1)
        x3 = self._block(x2, self.conv3, self.bn3, self.se3) + self.resconv2(x2)
        x4 = self._block(x3, self.conv4, self.bn4, self.se4) + self.resconv3(x3)
        x = F.leaky_relu(self.bn5(self.conv5(torch.cat((x1, x2, x3, x4), dim=1))), 0.2)
        return torch.cat((x.max(dim=2)[0], x.mean(dim=2)), 1).view(B, self.output_size)


class SpareNetEncode(nn.Module):  # :85-120
    def __init__(self, bottleneck_size=4096, hide_size=4096, ops=CpuOps):
        super().__init__()
        self.feat_extractor = EdgeConvResFeat(use_SElayer=True, k=8, output_size=hide_size, hide_size=4096, ops=ops)
        self.linear = nn.Linear(hide_size, bottleneck_size)
        self.bn = nn.BatchNorm1d(bottleneck_size)

    def forward(self, x):
        return F.relu(self.bn(self.linear(self.feat_extractor(x))))


class AdaptiveInstanceNorm1d(nn.Module):  # :909-959
    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = None
        self.bias = None
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    def forward(self, x):
        b, c = x.shape[:2]
        out = F.batch_norm(x.contiguous().view(1, b * c, -1), self.running_mean.repeat(b), self.running_var.repeat(b),
                           self.weight, self.bias, True, self.momentum, self.eps)
        return out.view_as(x)


class GridDecoder(nn.Module):  # :962-1062 (use_SElayer=True, no sine)
    def __init__(self, input_dim=2, bottleneck_size=1026):
        super().__init__()
        bs = bottleneck_size
        self.conv1 = nn.Conv1d(input_dim, bs, 1)
        self.conv2 = nn.Conv1d(bs, bs // 2, 1)
        self.conv3 = nn.Conv1d(bs // 2, bs // 4, 1)
        self.conv4 = nn.Conv1d(bs // 4, 3, 1)
        self.adain1, self.adain2, self.adain3 = AdaptiveInstanceNorm1d(bs), AdaptiveInstanceNorm1d(bs // 2), AdaptiveInstanceNorm1d(bs // 4)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(bs), nn.BatchNorm1d(bs // 2), nn.BatchNorm1d(bs // 4)
        self.se1, self.se2, self.se3 = SELayer1D(bs), SELayer1D(bs // 2), SELayer1D(bs // 4)

    def forward(self, x):
        x = F.relu(self.se1(self.bn1(self.adain1(self.conv1(x)))))
        x = F.relu(self.se2(self.bn2(self.adain2(self.conv2(x)))))
        x = F.relu(self.se3(self.bn3(self.adain3(self.conv3(x)))))
        return torch.tanh(self.conv4(x))


class StyleBasedAdaIn(nn.Module):  # :394-422 + assign_adain_params :831-849
    def __init__(self, input_dim=2, style_dim=1024, bottleneck_size=1026):
        super().__init__()
        self.dec = GridDecoder(input_dim, bottleneck_size)

    def forward(self, content, style, adain_params):
        for m in (self.dec.adain1, self.dec.adain2, self.dec.adain3):
            nf = m.num_features
            m.bias = adain_params[:, :nf].contiguous().view(-1)
            m.weight = adain_params[:, nf:2 * nf].contiguous().view(-1)
            adain_params = adain_params[:, 2 * nf:]
        return self.dec(content)


def grid_points(num_points, n_primitives):  # grid_generation :793-812 -> [2, pts] in [-1, 1]
    per = num_points / n_primitives
    gx = 2 ** math.floor(math.log2(per) / 2) - 1
    gy = 2 ** math.ceil(math.log2(per) / 2) - 1
    verts = [[i / gx, j / gy] for i in range(int(gx + 1)) for j in range(int(gy + 1))]
    return ((torch.tensor(verts, dtype=torch.float32) - 0.5) * 2).t().contiguous()


class SpareNetDecode(nn.Module):  # :289-391, use_AdaIn == "share"
    def __init__(self, num_points=16384, n_primitives=32, bottleneck_size=4096):
        super().__init__()
        self.num_points, self.n_primitives = num_points, n_primitives
        self.decoder = nn.ModuleList([StyleBasedAdaIn(2, bottleneck_size) for _ in range(n_primitives)])
        n_adain = 2 * (1026 + 513 + 256)
        self.mlp = nn.Sequential(nn.Linear(bottleneck_size, bottleneck_size), nn.ReLU(), nn.Linear(bottleneck_size, n_adain))

    def forward(self, style, partial_x):
        adain_params = self.mlp(style)
        grid = grid_points(self.num_points, self.n_primitives).to(style.device)
        grid = grid.unsqueeze(0).expand(style.size(0), -1, -1).contiguous()
        return torch.cat([d(grid, style, adain_params) for d in self.decoder], 2).contiguous()


class PointNetRes(nn.Module):  # :582-646 (use_SElayer=True)
    def __init__(self):
        super().__init__()
        self.conv1, self.conv2, self.conv3 = nn.Conv1d(4, 64, 1), nn.Conv1d(64, 128, 1), nn.Conv1d(128, 1024, 1)
        self.conv4, self.conv5, self.conv6, self.conv7 = nn.Conv1d(1088, 512, 1), nn.Conv1d(512, 256, 1), nn.Conv1d(256, 128, 1), nn.Conv1d(128, 3, 1)
        self.se1, self.se2, self.se4, self.se5, self.se6 = SELayer1D(64), SELayer1D(128), SELayer1D(512), SELayer1D(256), SELayer1D(128)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(128), nn.BatchNorm1d(1024)
        self.bn4, self.bn5, self.bn6, self.bn7 = nn.BatchNorm1d(512), nn.BatchNorm1d(256), nn.BatchNorm1d(128), nn.BatchNorm1d(3)

    def forward(self, x):
        n = x.size(2)
        x = F.relu(self.se1(self.bn1(self.conv1(x))))
        pointfeat = x
        x = F.relu(self.se2(self.bn2(self.conv2(x))))
        x = self.bn3(self.conv3(x)).max(dim=2)[0]
        x = torch.cat([x.view(-1, 1024, 1).repeat(1, 1, n), pointfeat], 1)
        x = F.relu(self.se4(self.bn4(self.conv4(x))))
        x = F.relu(self.se5(self.bn5(self.conv5(x))))
        x = F.relu(self.se6(self.bn6(self.conv6(x))))
        return torch.tanh(self.conv7(x))


class SpareNetRefine(nn.Module):  # :530-579
    def __init__(self, n_primitives=32, num_points=16384, ops=CpuOps):
        super().__init__()
        self.num_points, self.n_primitives, self.ops = num_points, n_primitives, ops
        self.residual = PointNetRes()

    def forward(self, inps, partial, coarse):
        dist, _, mean_mst_dis = self.ops.expansion(coarse, self.num_points // self.n_primitives, 1.5)
        loss_mst = torch.mean(dist)
        inps = torch.cat((inps, torch.zeros_like(inps[:, :1])), 1)
        partial = torch.cat((partial, torch.ones_like(partial[:, :1])), 1)
        base = torch.cat((inps, partial), 2)
        idx = self.ops.mds(base[:, 0:3, :].transpose(1, 2).contiguous(), coarse.shape[1], mean_mst_dis)
        base = self.ops.gather(base.contiguous(), idx)
        delta = self.residual(base)
        outs = base[:, 0:3, :] + delta
        return outs.transpose(2, 1).contiguous(), loss_mst


class SpareNetGenerator(nn.Module):  # :12-82
    def __init__(self, n_primitives=32, hide_size=4096, bottleneck_size=4096, num_points=16384, use_SElayer=True,
                 use_AdaIn="share", encode="Residualnet", ops=CpuOps):
        super().__init__()
        assert use_SElayer and use_AdaIn == "share" and encode == "Residualnet", "only the shipped configuration is restated"
        self.conv1 = nn.Conv1d(3, 64, 1)  # unused in forward, kept for checkpoint compatibility (:43)
        self.encoder = SpareNetEncode(bottleneck_size=bottleneck_size, hide_size=hide_size, ops=ops)
        self.decoder = SpareNetDecode(num_points=num_points, n_primitives=n_primitives, bottleneck_size=bottleneck_size)
        self.refine = SpareNetRefine(num_points=num_points, n_primitives=n_primitives, ops=ops)

    def forward(self, data):
        partial = data["partial_cloud"].transpose(1, 2).contiguous()
        style = self.encoder(partial)
        outs = self.decoder(style, partial)
        coarse = outs.transpose(1, 2).contiguous()
        middle, loss_mst = self.refine(outs, partial, coarse)
        refine, _ = self.refine(middle.transpose(1, 2).contiguous(), partial, middle)
        return coarse, middle, refine, loss_mst


def init_weights(m):  # utils/model_init.py:137-159
    if isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    if type(m) == nn.Conv1d:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)
    elif type(m) == nn.BatchNorm1d:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0.0)
    elif type(m) == nn.Linear:
        nn.init.normal_(m.weight, 0, 0.01)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)


def deterministic_fill(module, scale_w=0.05):
    """Seed-free parameter fill keyed on the parameter NAME, so the real reference classes, this restatement and
    the product receive bit-identical weights without shipping a checkpoint (used by the golden script and tests)."""
    with torch.no_grad():
        for name, p in sorted(list(module.named_parameters()) + list(module.named_buffers())):
            if "num_batches_tracked" in name:
                continue
            h = sum((i + 1) * ord(ch) for i, ch in enumerate(name)) % 9973
            t = torch.arange(p.numel(), dtype=torch.float64)
            v = torch.sin(t * (0.37 + 0.001 * (h % 211)) + h)
            if name.endswith("running_var"):
                v = 1.0 + 0.2 * v
            elif name.endswith("running_mean"):
                v = 0.1 * v
            elif p.dim() <= 1 and ("bn" in name.split(".")[-2] or name.split(".")[-2] == "bn"):
                v = (0.2 + 1.2 * v) if name.endswith("weight") else 0.1 * v      # MIXED-SIGN gamma (exercises the max/min fusion), small beta
            elif p.dim() <= 1:
                v = 0.1 * v
            else:
                fan_in = p[0].numel()
                v = v * (scale_w if fan_in < 16 else 1.5 / math.sqrt(fan_in))
            p.copy_(v.view_as(p).to(p.dtype))
