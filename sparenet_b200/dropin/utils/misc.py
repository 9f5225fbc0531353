"""Drop-in for the evaluation-metrics part of the reference's utils/misc.py: class Metrics (reference :133-260).

Same interface -- Metrics.get(pred, gt) -> [F-Score, ChamferDistance, EMD], items()/names(), Metrics(metric_name, values),
state_dict(), better_than() -- but everything stays on the GPU: the F-score (@ threshold 0.01, :180-190) comes from the
nearest-neighbour distances of ONE Chamfer call instead of two open3d KD-tree queries on the CPU, and that same call also
yields the Chamfer metric (:201-203).  EMD: eps 0.005, 50 iterations, x100 (:206-211).  The rest of the reference module
(tensorboard writers, checkpoint I/O, logging) is host bookkeeping outside the hot path and is not provided here.
"""
import logging

import torch

from sparenet_b200.dropin.cuda.chamfer_distance import ChamferDistance
from sparenet_b200.dropin.cuda.emd import emd_module as emd

logger = logging.getLogger()


class Metrics(object):
    ITEMS = [
        {"name": "F-Score", "enabled": True, "eval_func": "cls._get_f_score", "is_greater_better": True, "init_value": 0},
        {"name": "ChamferDistance", "enabled": True, "eval_func": "cls._get_chamfer_distance", "eval_object": ChamferDistance(),
         "is_greater_better": False, "init_value": 32767},
        {"name": "EMD", "enabled": True, "eval_func": "cls._get_emd", "eval_object": emd.emdModule(), "is_greater_better": False,
         "init_value": 32767},
    ]

    @classmethod
    def get(cls, pred, gt):
        return [getattr(cls, item["eval_func"].split(".")[1])(pred, gt) for item in cls.items()]

    @classmethod
    def items(cls):
        return [i for i in cls.ITEMS if i["enabled"]]

    @classmethod
    def names(cls):
        return [i["name"] for i in cls.items()]

    @classmethod
    def _nn(cls, pred, gt):
        pred = pred.reshape(-1, pred.shape[-2], 3).float().contiguous()
        gt = gt.reshape(-1, gt.shape[-2], 3).float().contiguous()
        return cls.ITEMS[1]["eval_object"](pred, gt)            # squared NN distances, both directions (memoised per input pair)

    @classmethod
    def _get_f_score(cls, pred, gt, th=0.01):
        d1, d2 = cls._nn(pred, gt)                              # open3d's compute_point_cloud_distance is the Euclidean NN distance
        precision = (d1.sqrt() < th).float().mean().item()
        recall = (d2.sqrt() < th).float().mean().item()
        return 2 * recall * precision / (recall + precision) if recall + precision else 0

    @classmethod
    def _get_chamfer_distance(cls, pred, gt):
        d1, d2 = cls._nn(pred, gt)
        return (d1.mean() + d2.mean()).item() * 1000

    @classmethod
    def _get_emd(cls, pred, gt):
        dist, _ = cls.ITEMS[2]["eval_object"](pred, gt, 0.005, 50)
        return torch.sqrt(dist).mean(1).mean().item() * 100

    def __init__(self, metric_name, values):
        self._items = Metrics.items()
        self._values = [item["init_value"] for item in self._items]
        self.metric_name = metric_name
        if isinstance(values, dict):
            index = {item["name"]: i for i, item in enumerate(self._items)}
            for k, v in values.items():
                if k not in index:
                    logger.warning("Ignore Metric[Name=%s] due to disability." % k)
                    continue
                self._values[index[k]] = v
        elif isinstance(values, list):
            self._values = values
        else:
            raise Exception("Unsupported value type: %s" % type(values))

    def state_dict(self):
        return {item["name"]: self._values[i] for i, item in enumerate(self._items)}

    def __repr__(self):
        return str(self.state_dict())

    def better_than(self, other):
        if other is None:
            return True
        for i, item in enumerate(self._items):
            if item["name"] == self.metric_name:
                a, b = self._values[i], other._values[i]
                return a > b if item["is_greater_better"] else a < b
        raise Exception("Invalid metric name to compare.")
