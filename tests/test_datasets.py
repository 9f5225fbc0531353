"""Cloud ingestion drop-in (sparenet_b200/dropin/datasets): the transforms against goldens produced by the REAL reference module under
fixed numpy seeds (bit-exact: same RNG consumption order), the PCD container (ASCII / binary / LZF binary_compressed, field order, NaN
rows) against the format's definition and round trips, collate_fn, the ShapeNet file list / recipe, and -- on the GPU -- the pinned,
double-buffered host->device staging."""
import json
import os
import struct

import numpy as np
import pytest
import torch

from sparenet_b200.dropin.datasets import data_loaders as L
from sparenet_b200.dropin.datasets import data_transforms as T
from sparenet_b200.dropin.datasets.io import IO, _lzf_decompress, read_pcd, write_pcd

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "transforms_ref.npz"))


def test_transforms_match_reference_bit_for_bit():
    small, big = G["cloud_small"], G["cloud_big"]
    np.random.seed(11)
    a = T.RandomSamplePoints({"n_points": 128})(small.copy())
    assert a.shape == (128, 3) and np.array_equal(a, G["sample_pad"]) and np.all(a[100:] == 0)          # zero padding (:170-173)
    np.random.seed(12)
    assert np.array_equal(T.RandomSamplePoints({"n_points": 64})(big.copy()), G["sample_sub"])
    for i, rv in enumerate((0.1, 0.4, 0.7, 0.9)):
        assert np.array_equal(T.RandomMirrorPoints(None)(small.copy(), rv), G[f"mirror_{i}"])
    np.random.seed(13)
    assert np.array_equal(T.RandomClipPoints({"sigma": 0.02, "clip": 0.03})(small.copy()), G["clip"])
    assert np.array_equal(T.RandomRotatePoints(None)(small.copy(), 0.3), G["rotate"])
    np.random.seed(14)
    assert np.array_equal(T.RandomScalePoints({"scale": 1.2})(small.copy(), 0.8), G["scale"])
    comp = T.Compose([{"callback": "RandomSamplePoints", "parameters": {"n_points": 300}, "objects": ["partial_cloud"]},
                      {"callback": "RandomSamplePoints", "parameters": {"n_points": 600}, "objects": ["gtcloud"]},
                      {"callback": "RandomMirrorPoints", "objects": ["partial_cloud", "gtcloud"]},
                      {"callback": "ToTensor", "objects": ["partial_cloud", "gtcloud"]}])
    np.random.seed(15)
    res = comp({"partial_cloud": small.copy(), "gtcloud": big.copy()})
    assert res["partial_cloud"].dtype == torch.float32 and res["partial_cloud"].shape == (300, 3)
    assert np.array_equal(res["partial_cloud"].numpy(), G["compose_partial"]) and np.array_equal(res["gtcloud"].numpy(), G["compose_gt"])


def _lzf_literal(data: bytes) -> bytes:      # a valid (uncompressed-literal) LZF stream: runs of <= 32 bytes
    out = bytearray()
    for i in range(0, len(data), 32):
        chunk = data[i:i + 32]
        out.append(len(chunk) - 1)
        out += chunk
    return bytes(out)


def test_pcd_reader_formats(tmp_path):
    rng = np.random.RandomState(3)
    pts = (rng.rand(257, 3).astype(np.float32) - 0.5)
    for binary in (True, False):
        p = str(tmp_path / f"c{int(binary)}.pcd")
        write_pcd(p, pts, binary=binary)
        got = IO.get(p)
        assert got.dtype == np.float64 and got.shape == (257, 3)
        assert np.array_equal(got.astype(np.float32), pts)
    # field order other than x y z, an extra field, a NaN row (dropped like open3d does), float64 z
    n = 5
    rec = np.dtype([("intensity", np.float32), ("z", np.float64), ("x", np.float32), ("y", np.float32)])
    arr = np.zeros(n, dtype=rec)
    arr["x"], arr["y"], arr["z"], arr["intensity"] = np.arange(n), np.arange(n) * 2, np.arange(n) * 3, 9
    arr["y"][2] = np.nan
    head = f"VERSION .7\nFIELDS intensity z x y\nSIZE 4 8 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH {n}\nHEIGHT 1\nPOINTS {n}\nDATA binary\n"
    p = str(tmp_path / "order.pcd")
    open(p, "wb").write(head.encode() + arr.tobytes())
    got = read_pcd(p)
    assert got.shape == (4, 3) and np.array_equal(got[:, 0], [0, 1, 3, 4]) and np.array_equal(got[:, 2], [0, 3, 9, 12])
    # binary_compressed: field-major block behind an (compressed size, uncompressed size) header, LZF
    soa = np.concatenate([pts[:, 0], pts[:, 1], pts[:, 2]]).astype(np.float32).tobytes()
    comp = _lzf_literal(soa)
    head = f"VERSION .7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 257\nHEIGHT 1\nPOINTS 257\nDATA binary_compressed\n"
    p = str(tmp_path / "lzf.pcd")
    open(p, "wb").write(head.encode() + struct.pack("<II", len(comp), len(soa)) + comp)
    assert np.array_equal(read_pcd(p).astype(np.float32), pts)
    # LZF back references (overlapping copy = run-length): literal 'ab', then copy 6 bytes from 2 back
    assert _lzf_decompress(bytes([1, ord("a"), ord("b"), (4 << 5) | 0, 1]), 8) == b"abababab"
    with pytest.raises(Exception):
        IO.get(str(tmp_path / "x.obj"))


def test_collate_filelist_and_dataset(tmp_path):
    cat = [{"taxonomy_id": "02691156", "taxonomy_name": "airplane", "train": ["m1", "m2"], "test": ["m3"]},
           {"taxonomy_id": "02958343", "taxonomy_name": "car", "train": ["c1"], "test": []}]
    cf = tmp_path / "cats.json"
    cf.write_text(json.dumps(cat))
    root = tmp_path / "data"
    rng = np.random.RandomState(5)
    pp, cp = str(root / "%s" / "partial" / "%s" / "%s" / "%02d.pcd"), str(root / "%s" / "complete" / "%s" / "%s.pcd")
    fl = L.shapenet_file_list(str(cf), pp, cp, subset="train", n_renderings=2)
    assert [f["label"] for f in fl] == [0, 0, 1] and len(fl[0]["partial_cloud_path"]) == 2
    fl2 = L.shapenet_file_list(str(cf), pp, cp, subset="train", n_renderings=2, version="ShapeNet")
    assert len(fl2) == 6 and fl2[1]["model_id"] == "m11"
    for f in fl:
        for p in f["partial_cloud_path"] + [f["gtcloud_path"]]:
            os.makedirs(os.path.dirname(p), exist_ok=True)
            write_pcd(p, rng.rand(150 if "partial" in p else 700, 3) - 0.5)
    ds = L.Dataset({"n_renderings": 2, "required_items": ["partial_cloud", "gtcloud"], "shuffle": True}, fl,
                   L.shapenet_transforms(L.DatasetSubset.TRAIN, n_outpoints=512, n_partial=200))
    tax, labels, mids, data = L.collate_fn([ds[i] for i in range(3)])
    assert tax == ["02691156", "02691156", "02958343"] and labels == [0, 0, 1] and mids == ["m1", "m2", "c1"]
    assert data["partial_cloud"].shape == (3, 200, 3) and data["gtcloud"].shape == (3, 512, 3) and data["gtcloud"].dtype == torch.float32
    assert (data["partial_cloud"][:, 150:] == 0).all()                      # 150 points padded with zeros to 200


@pytest.mark.gpu
def test_device_batches_pinned_double_buffered(cuda):
    torch.manual_seed(0)
    batches = [(["t"] * 4, [i, i + 1, i + 2, i + 3], ["m"] * 4, {"partial_cloud": torch.rand(4, 300, 3), "gtcloud": torch.rand(4, 900, 3)})
               for i in range(5)]
    seen = 0
    for (tax, labels, mids, data), ref in zip(L.DeviceBatches(batches, cuda), batches):
        assert labels.device.type == "cuda" and labels.dtype == torch.long and labels.tolist() == ref[1]
        for k in ("partial_cloud", "gtcloud"):
            assert data[k].device.type == "cuda" and torch.equal(data[k].cpu(), ref[3][k])
        seen += 1
    assert seen == 5
