"""Drop-in for the point-cloud part of the reference's datasets/data_transforms.py: Compose (:11-40), ToTensor (:43-53),
RandomSamplePoints (:162-175: random subset, ZERO-padded up to n_points -- the duplicate points every op's tie handling is tested
against), RandomClipPoints (:178-187), RandomRotatePoints / RandomScalePoints / RandomMirrorPoints (:190-232).

The numpy global RNG is consumed in exactly the reference's order (one uniform draw per transform in Compose, then the transform's own
draws), so a fixed np.random.seed reproduces the reference's batches bit for bit (tests/golden/transforms_ref.npz, produced by the real
module).  transforms3d (un-vendored, absent here) is needed only for two 3x3 matrices, restated from its published definitions:
zooms.zfdir2mat(f, d) = I + (f - 1) d d^T and axangles.axangle2mat (Rodrigues).  The image transforms of the reference (crops, colour
jitter for the GRNet image branch) are not on this path and are not mirrored."""
import math

import numpy as np
import torch


def zfdir2mat(factor, direction=None):
    """transforms3d.zooms.zfdir2mat: zoom by `factor` along `direction` (None: isotropic)."""
    if direction is None:
        return np.eye(3) * factor
    d = np.asarray(direction, dtype=np.float64)
    d = d / math.sqrt((d ** 2).sum())
    return np.eye(3) + (factor - 1.0) * np.outer(d, d)


def axangle2mat(axis, angle):
    """transforms3d.axangles.axangle2mat (unit axis assumed after normalisation)."""
    x, y, z = np.asarray(axis, dtype=np.float64) / math.sqrt(sum(a * a for a in axis))
    c, s = math.cos(angle), math.sin(angle)
    C = 1 - c
    return np.array([[x * x * C + c, x * y * C - z * s, x * z * C + y * s],
                     [y * x * C + z * s, y * y * C + c, y * z * C - x * s],
                     [z * x * C - y * s, z * y * C + x * s, z * z * C + c]])


class ToTensor(object):
    def __init__(self, parameters):
        pass

    def __call__(self, arr):
        if len(arr.shape) == 3:
            arr = arr.transpose(2, 0, 1)
        return torch.from_numpy(arr.copy()).float()


class RandomSamplePoints(object):
    def __init__(self, parameters):
        self.n_points = parameters["n_points"]

    def __call__(self, ptcloud):
        choice = np.random.permutation(ptcloud.shape[0])
        ptcloud = ptcloud[choice[: self.n_points]]
        if ptcloud.shape[0] < self.n_points:
            ptcloud = np.concatenate([ptcloud, np.zeros((self.n_points - ptcloud.shape[0], 3))])
        return ptcloud


class RandomClipPoints(object):
    def __init__(self, parameters):
        self.sigma = parameters["sigma"] if "sigma" in parameters else 0.01
        self.clip = parameters["clip"] if "clip" in parameters else 0.05

    def __call__(self, ptcloud):
        ptcloud += np.clip(self.sigma * np.random.randn(*ptcloud.shape), -self.clip, self.clip).astype(np.float32)
        return ptcloud


class RandomRotatePoints(object):
    def __init__(self, parameters):
        pass

    def __call__(self, ptcloud, rnd_value):
        trfm_mat = np.dot(axangle2mat([0, 1, 0], 2 * math.pi * rnd_value), zfdir2mat(1))
        ptcloud[:, :3] = np.dot(ptcloud[:, :3], trfm_mat.T)
        return ptcloud


class RandomScalePoints(object):
    def __init__(self, parameters):
        self.scale = parameters["scale"]

    def __call__(self, ptcloud, rnd_value):
        scale = np.random.uniform(1.0 / self.scale * rnd_value, self.scale * rnd_value)
        trfm_mat = np.dot(zfdir2mat(scale), zfdir2mat(1))
        ptcloud[:, :3] = np.dot(ptcloud[:, :3], trfm_mat.T)
        return ptcloud


class RandomMirrorPoints(object):
    def __init__(self, parameters):
        pass

    def __call__(self, ptcloud, rnd_value):
        trfm_mat = zfdir2mat(1)
        trfm_mat_x = np.dot(zfdir2mat(-1, [1, 0, 0]), trfm_mat)
        trfm_mat_z = np.dot(zfdir2mat(-1, [0, 0, 1]), trfm_mat)
        if rnd_value <= 0.25:
            trfm_mat = np.dot(trfm_mat_z, np.dot(trfm_mat_x, trfm_mat))
        elif rnd_value <= 0.5:
            trfm_mat = np.dot(trfm_mat_x, trfm_mat)
        elif rnd_value <= 0.75:
            trfm_mat = np.dot(trfm_mat_z, trfm_mat)
        ptcloud[:, :3] = np.dot(ptcloud[:, :3], trfm_mat.T)
        return ptcloud


_WITH_RND = (RandomRotatePoints, RandomScalePoints, RandomMirrorPoints)
_REGISTRY = {c.__name__: c for c in (ToTensor, RandomSamplePoints, RandomClipPoints, RandomRotatePoints, RandomScalePoints, RandomMirrorPoints)}


class Compose(object):
    def __init__(self, transforms):
        self.transformers = []
        for tr in transforms:
            cb = tr["callback"]
            transformer = _REGISTRY[cb] if isinstance(cb, str) else cb       # the reference eval()s the name (:15)
            self.transformers.append({"callback": transformer(tr["parameters"] if "parameters" in tr else None), "objects": tr["objects"]})

    def __call__(self, data):
        for tr in self.transformers:
            transform, objects = tr["callback"], tr["objects"]
            rnd_value = np.random.uniform(0, 1)                              # ONE draw per transform, shared by its objects (:24)
            for k, v in data.items():
                if k in objects and k in data:
                    data[k] = transform(v, rnd_value) if isinstance(transform, _WITH_RND) else transform(v)
        return data
