"""Generates tests/golden/transforms_ref.npz by running the REAL reference datasets/data_transforms.py (build container only) under
fixed numpy seeds: RandomSamplePoints (with zero padding), RandomMirrorPoints for the four rnd_value branches, RandomClipPoints,
RandomRotatePoints, RandomScalePoints and the ShapeNet training Compose (sample 3000 / 16384, mirror, ToTensor).
transforms3d is un-vendored and absent here: a two-function stand-in (zfdir2mat, axangle2mat, restated from its published definitions)
is injected before the reference module is imported -- the only thing the reference takes from it.  The GPU box never runs this."""
import math
import os
import sys
import types

import numpy as np

t3d = types.ModuleType("transforms3d")
t3d.zooms, t3d.axangles = types.ModuleType("transforms3d.zooms"), types.ModuleType("transforms3d.axangles")


def zfdir2mat(factor, direction=None):
    if direction is None:
        return np.eye(3) * factor
    d = np.asarray(direction, dtype=np.float64)
    d = d / math.sqrt((d ** 2).sum())
    return np.eye(3) + (factor - 1.0) * np.outer(d, d)


def axangle2mat(axis, angle, is_normalized=False):
    x, y, z = np.asarray(axis, dtype=np.float64) / math.sqrt(sum(a * a for a in axis))
    c, s = math.cos(angle), math.sin(angle)
    C = 1 - c
    return np.array([[x * x * C + c, x * y * C - z * s, x * z * C + y * s], [y * x * C + z * s, y * y * C + c, y * z * C - x * s],
                     [z * x * C - y * s, z * y * C + x * s, z * z * C + c]])


t3d.zooms.zfdir2mat, t3d.axangles.axangle2mat = zfdir2mat, axangle2mat
sys.modules.update({"transforms3d": t3d, "transforms3d.zooms": t3d.zooms, "transforms3d.axangles": t3d.axangles})
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("ref_data_transforms", "/root/reference/datasets/data_transforms.py")
R = importlib.util.module_from_spec(_spec)     # by path: a `datasets` package from site-packages shadows the reference's directory
_spec.loader.exec_module(R)

out = {}
rng = np.random.RandomState(7)
cloud_small = rng.rand(100, 3).astype(np.float32) - 0.5
cloud_big = rng.rand(500, 3).astype(np.float32) - 0.5
out["cloud_small"], out["cloud_big"] = cloud_small, cloud_big
np.random.seed(11)
out["sample_pad"] = R.RandomSamplePoints({"n_points": 128})(cloud_small.copy())      # 100 -> 128: zero padded
np.random.seed(12)
out["sample_sub"] = R.RandomSamplePoints({"n_points": 64})(cloud_big.copy())
for i, rv in enumerate((0.1, 0.4, 0.7, 0.9)):
    out[f"mirror_{i}"] = R.RandomMirrorPoints(None)(cloud_small.copy(), rv)
np.random.seed(13)
out["clip"] = R.RandomClipPoints({"sigma": 0.02, "clip": 0.03})(cloud_small.copy())
out["rotate"] = R.RandomRotatePoints(None)(cloud_small.copy(), 0.3)
np.random.seed(14)
out["scale"] = R.RandomScalePoints({"scale": 1.2})(cloud_small.copy(), 0.8)
comp = R.Compose([{"callback": "RandomSamplePoints", "parameters": {"n_points": 300}, "objects": ["partial_cloud"]},
                  {"callback": "RandomSamplePoints", "parameters": {"n_points": 600}, "objects": ["gtcloud"]},
                  {"callback": "RandomMirrorPoints", "objects": ["partial_cloud", "gtcloud"]},
                  {"callback": "ToTensor", "objects": ["partial_cloud", "gtcloud"]}])
np.random.seed(15)
res = comp({"partial_cloud": cloud_small.copy(), "gtcloud": cloud_big.copy()})
out["compose_partial"], out["compose_gt"] = res["partial_cloud"].numpy(), res["gtcloud"].numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "transforms_ref.npz"), **out)
print("wrote", sorted(out))
