"""Drop-in for the per-batch part of the reference's runners/sparenet_gan_runner.py: the GAN training step.

`sparenetGANStep` keeps the reference's method names and arithmetic -- completion (:131-190), discriminator_backward (:192-268),
generator_backward (:270-346), train_step (:69-118) -- but is constructed from the modules and optimizers themselves instead of an
EasyDict config: the reference's control plane (config files, logging, checkpoints, DataParallel wrappers, data loaders) is out of
scope; what is here is the path between a device batch and the two optimizer steps:

  generator forward -> rec loss (3 x EMD or Chamfer + 0.1 expansion + 0.5 consistency CD) -> 24 depth-map renders (8 views x
  {ground truth, completion, partial input}, ONE radius drawn per step) -> D step on detached images (MSE to real / fake labels)
  -> G step: weight_l2 * rec + weight_gan * MSE(D(fake), real) + weight_fm * feature matching + weight_im * L1(fake, real images).

Every point-cloud op on that path is an sm_100a kernel of this repository (generator, EMD / Chamfer, the fused renderer
snb_depthmaps_*); the 0.4 M-parameter image discriminator runs on cuDNN.  Multi-GPU: one process per GPU, `allreduce` hooks let
the caller average the G and D gradients over NCCL before each optimizer step (sparenet_b200.dist.allreduce_gradients).
"""
import random

import torch

from sparenet_b200 import functional as F_
from sparenet_b200.dropin.cuda.chamfer_distance import ChamferDistance, ChamferDistanceMean
from sparenet_b200.dropin.cuda.emd import emd_module as emd
from sparenet_b200.dropin.utils.p2i_utils import N_VIEWS_PREDEFINED


class sparenetGANStep:
    def __init__(self, models, models_D, renderer, optimizers, optimizers_D, metric="emd", use_consist_loss=True, use_cgan=True,
                 use_fm=True, use_im=True, weight_gan=0.1, weight_l2=200.0, weight_im=1.0, weight_fm=1.0, radius_list=(5.0, 7.0, 10.0),
                 allreduce_G=None, allreduce_D=None):
        self.models, self.models_D, self.renderer = models, models_D, renderer
        self.optimizers, self.optimizers_D = optimizers, optimizers_D
        self.metric, self.use_consist_loss = metric, use_consist_loss          # configs/sparenet_gan.yaml:18-24
        self.use_cgan, self.use_fm, self.use_im = use_cgan, use_fm, use_im     # :35-37
        self.weight_gan, self.weight_l2, self.weight_im, self.weight_fm = weight_gan, weight_l2, weight_im, weight_fm   # :38-41
        self.radius_list = list(radius_list)                                    # :27-31
        self.allreduce_G, self.allreduce_D = allreduce_G, allreduce_D
        self.chamfer_dist, self.chamfer_dist_mean, self.emd_dist = ChamferDistance(), ChamferDistanceMean(), emd.emdModule()
        self.criterionD = torch.nn.MSELoss()
        self.loss = {}

    # ------------------------------------------------------------------------------------------ reference :131-190
    def completion(self, data):
        coarse_ptcloud, middle_ptcloud, refine_ptcloud, expansion_penalty = self.models(data)
        gt = data["gtcloud"]
        if self.metric == "chamfer":
            coarse_loss = self.chamfer_dist_mean(coarse_ptcloud, gt).mean()
            middle_loss = self.chamfer_dist_mean(middle_ptcloud, gt).mean()
            refine_loss = self.chamfer_dist_mean(refine_ptcloud, gt).mean()
        elif self.metric == "emd":
            emd_coarse, _ = self.emd_dist(coarse_ptcloud, gt, eps=0.005, iters=50)
            emd_middle, _ = self.emd_dist(middle_ptcloud, gt, eps=0.005, iters=50)
            emd_refine, _ = self.emd_dist(refine_ptcloud, gt, eps=0.005, iters=50)
            coarse_loss = torch.sqrt(emd_coarse).mean(1).mean()
            refine_loss = torch.sqrt(emd_refine).mean(1).mean()
            middle_loss = torch.sqrt(emd_middle).mean(1).mean()
        else:
            raise Exception("unknown training metric")
        _loss = coarse_loss + middle_loss + refine_loss + expansion_penalty.mean() * 0.1
        if self.use_consist_loss:
            with F_.chamfer_reuse():      # metric "chamfer": Chamfer(refine, gt) was searched a moment ago -- reuse it (identical values)
                dist1, _ = self.chamfer_dist(refine_ptcloud, gt)
            _loss = _loss + torch.mean(dist1).mean() * 0.5
        return _loss, refine_ptcloud, middle_ptcloud, coarse_ptcloud, refine_loss, coarse_loss

    # ------------------------------------------------------------------------------------------ the 24 renders, :207-238
    def render(self, ptcloud, radius):
        """[B,N,3] -> [B, 8, S, S]: the eight predefined views concatenated on the channel axis."""
        return torch.cat([self.renderer(ptcloud, view_id=v, radius_list=[radius]) for v in range(N_VIEWS_PREDEFINED)], dim=1)

    # ------------------------------------------------------------------------------------------ reference :192-268
    def discriminator_backward(self, data, labels, rendered_ptcloud, radius=None):
        self.optimizers_D.zero_grad()
        random_radius = random.sample(self.radius_list, 1)[0] if radius is None else radius
        self.real_imgs = self.render(data["gtcloud"], random_radius)
        self.fake_imgs = self.render(rendered_ptcloud, random_radius)
        self.input_imgs = self.render(data["partial_cloud"], random_radius)
        y = labels if self.use_cgan else None
        D_real_pred = self.models_D(torch.cat((self.input_imgs, self.real_imgs), dim=1).detach(), y=y)
        D_fake_pred = self.models_D(torch.cat((self.input_imgs, self.fake_imgs), dim=1).detach(), y=y)
        errD_real = self.criterionD(D_real_pred, self.real_label)
        errD_fake = self.criterionD(D_fake_pred, self.fake_label)
        (errD_real + errD_fake).backward()
        if self.allreduce_D is not None:
            self.allreduce_D()
        self.optimizers_D.step()
        return errD_real, errD_fake

    # ------------------------------------------------------------------------------------------ reference :270-346
    def generator_backward(self, data, labels, rec_loss):
        self.optimizers.zero_grad()
        y = labels if self.use_cgan else None
        loss_fm, loss_im = 0.0, 0.0
        fake_in = torch.cat((self.input_imgs, self.fake_imgs), dim=1)
        if self.use_fm:
            D_fake_pred, D_fake_features = self.models_D(fake_in, feat=True, y=y)
            _, D_real_features = self.models_D(torch.cat((self.input_imgs, self.real_imgs), dim=1), feat=True, y=y)
            map_nums = [f.shape[1] for f in D_fake_features]       # weighted by the number of feature maps (:311-318)
            for j, n in enumerate(map_nums):
                loss_fm = loss_fm + float(n) / sum(map_nums) * torch.mean((D_fake_features[j] - D_real_features[j].detach()) ** 2)
        else:
            D_fake_pred = self.models_D(fake_in, y=y)
        errG_D = self.criterionD(D_fake_pred, self.real_label)
        if self.use_im:
            loss_im = loss_im + torch.nn.L1Loss()(self.fake_imgs, self.real_imgs.detach())
        errG = self.weight_l2 * rec_loss + self.weight_gan * errG_D
        if self.use_fm:
            errG = errG + self.weight_fm * loss_fm
        if self.use_im:
            errG = errG + self.weight_im * loss_im
        errG.backward()
        if self.allreduce_G is not None:
            self.allreduce_G()
        self.optimizers.step()
        return errG, errG_D

    # ------------------------------------------------------------------------------------------ reference :69-118
    def train_step(self, data, labels, radius=None):
        """data: {"partial_cloud": [B,Np,3], "gtcloud": [B,N,3]} on the device, labels [B] int64 class ids (the cGAN condition).
        Returns the dict of loss tensors the reference logs (no host synchronisation here: the caller decides when to .item())."""
        B = data["partial_cloud"].size(0)
        dev = data["partial_cloud"].device
        self.real_label = torch.ones(B, 1, device=dev)
        self.fake_label = torch.zeros(B, 1, device=dev)
        _loss, _, middle_ptcloud, _, refine_loss, coarse_loss = self.completion(data)
        errD_real, errD_fake = self.discriminator_backward(data, labels, middle_ptcloud, radius)
        errG, errG_D = self.generator_backward(data, labels, _loss)
        self.loss = {"coarse_loss": coarse_loss * 1000, "refine_loss": refine_loss * 1000, "rec_loss": _loss, "errG": errG, "errG_D": errG_D,
                     "errD_real": errD_real, "errD_fake": errD_fake}
        return self.loss
