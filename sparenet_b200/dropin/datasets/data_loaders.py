"""Drop-in for the batch-assembly part of the reference's datasets/data_loaders.py: DatasetSubset (:65-69), collate_fn (:72-93),
Dataset (:96-127), the ShapeNet file list and transform recipes (:130-240) -- built from explicit arguments instead of the EasyDict
config -- plus the piece the reference leaves to `torch.utils.data.DataLoader(pin_memory=True)` and `.to(gpu)` in the runner
(runners/sparenet_gan_runner.py:73-75): DeviceBatches, which stages every collated batch in PINNED host memory and issues the
host->device copies on a copy stream one batch ahead of the consumer, so the step never waits for PCIe.
"""
import json
import random
from enum import Enum, unique

import numpy as np
import torch
import torch.utils.data.dataset

from . import data_transforms
from .io import IO


@unique
class DatasetSubset(Enum):
    TRAIN = 0
    TEST = 1
    VAL = 2


def collate_fn(batch):
    taxonomy_ids, model_ids, labels, data = [], [], [], {}
    for sample in batch:
        taxonomy_ids.append(sample[0])
        labels.append(sample[1])
        model_ids.append(sample[2])
        for k, v in sample[3].items():
            data.setdefault(k, []).append(v)
    for k, v in data.items():
        data[k] = torch.stack(v, 0)
    return taxonomy_ids, labels, model_ids, data


class Dataset(torch.utils.data.dataset.Dataset):
    def __init__(self, options, file_list, transforms=None):
        self.options, self.file_list, self.transforms = options, file_list, transforms

    def __len__(self):
        return len(self.file_list)

    def __getitem__(self, idx):
        sample = self.file_list[idx]
        data = {}
        rand_idx = -1
        if "n_renderings" in self.options:
            rand_idx = random.randint(0, self.options["n_renderings"] - 1) if self.options["shuffle"] else 0
        for ri in self.options["required_items"]:
            file_path = sample["%s_path" % ri]
            if type(file_path) == list:
                file_path = file_path[rand_idx]
            data[ri] = IO.get(file_path).astype(np.float32)
        if self.transforms is not None:
            data = self.transforms(data)
        return sample["taxonomy_id"], sample["label"], sample["model_id"], data


def shapenet_transforms(subset, n_outpoints=16384, n_partial=3000):
    """ShapeNetDataLoader._get_transforms (:155-190): sample 3000 / n_outpoints points (zero-padded), mirror (training only), tensors."""
    tr = [{"callback": "RandomSamplePoints", "parameters": {"n_points": n_partial}, "objects": ["partial_cloud"]},
          {"callback": "RandomSamplePoints", "parameters": {"n_points": n_outpoints}, "objects": ["gtcloud"]}]
    if subset == DatasetSubset.TRAIN:
        tr.append({"callback": "RandomMirrorPoints", "objects": ["partial_cloud", "gtcloud"]})
    tr.append({"callback": "ToTensor", "objects": ["partial_cloud", "gtcloud"]})
    return data_transforms.Compose(tr)


def shapenet_file_list(category_file_path, partial_points_path, complete_points_path, subset="train", n_renderings=1, version="GRnet"):
    """ShapeNetDataLoader._get_file_list (:200-240): one entry per model with its n_renderings partial views ("GRnet") or one entry per
    (model, view) ("ShapeNet"); label = index of the taxonomy in the category file (the cGAN class id)."""
    with open(category_file_path) as f:
        categories = json.loads(f.read())
    file_list = []
    for label, dc in enumerate(categories):
        for s in dc[subset]:
            if version == "GRnet":
                file_list.append({"taxonomy_id": dc["taxonomy_id"], "label": label, "model_id": s,
                                  "partial_cloud_path": [partial_points_path % (subset, dc["taxonomy_id"], s, i) for i in range(n_renderings)],
                                  "gtcloud_path": complete_points_path % (subset, dc["taxonomy_id"], s)})
            else:
                for i in range(n_renderings):
                    file_list.append({"taxonomy_id": dc["taxonomy_id"], "label": label, "model_id": s + str(i),
                                      "partial_cloud_path": partial_points_path % (subset, dc["taxonomy_id"], s, i),
                                      "gtcloud_path": complete_points_path % (subset, dc["taxonomy_id"], s)})
    return file_list


class DeviceBatches:
    """Iterates a DataLoader (or any iterable of collate_fn batches) and yields (taxonomy_ids, labels [B] int64 on the device,
    model_ids, data on the device): every tensor goes through a reusable PINNED staging buffer and an asynchronous copy on a dedicated
    stream, issued one batch ahead; the consumer's stream waits on the copy event only."""

    def __init__(self, loader, device, n_sampling_points=None):
        self.loader, self.device = loader, torch.device(device)
        self.n_sampling_points = n_sampling_points      # runners sub-sample the 3000-point partial cloud to 2048 on the device side
        self.stream = torch.cuda.Stream(device=self.device)
        self._pinned = [{}, {}]
        self._events = [None, None]

    def _stage(self, slot, batch):
        tax, labels, mids, data = batch
        out = {}
        if self._events[slot] is not None:
            self._events[slot].synchronize()            # the copies that last read this slot's pinned buffers have completed
        with torch.cuda.stream(self.stream):
            for k, v in list(data.items()) + [("__labels__", torch.as_tensor(labels, dtype=torch.long))]:
                buf = self._pinned[slot].get(k)
                if buf is None or buf.shape != v.shape or buf.dtype != v.dtype:
                    buf = torch.empty(v.shape, dtype=v.dtype).pin_memory()
                    self._pinned[slot][k] = buf
                buf.copy_(v)
                out[k] = buf.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self._events[slot] = ev
        return tax, out.pop("__labels__"), mids, out, ev

    def __iter__(self):
        it = iter(self.loader)
        slot, pending = 0, None
        try:
            pending = self._stage(slot, next(it))
        except StopIteration:
            return
        while pending is not None:
            cur = pending
            slot ^= 1
            try:
                pending = self._stage(slot, next(it))       # the next batch's copies run under the consumer's current step
            except StopIteration:
                pending = None
            tax, labels, mids, data, ev = cur
            torch.cuda.current_stream(self.device).wait_event(ev)
            for t in list(data.values()) + [labels]:
                t.record_stream(torch.cuda.current_stream(self.device))
            yield tax, labels, mids, data
