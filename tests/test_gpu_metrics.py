"""Evaluation metrics on the GPU: the drop-in utils/misc.py:Metrics against a float64 brute-force
restatement of the reference's definitions (F-score@0.01 from Euclidean NN distances as open3d computes them,
utils/misc.py:180-190; ChamferDistanceMean x 1000, :201-203; EMD x 100, :206-211)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_metrics_match_reference_definitions(cuda):
    from sparenet_b200.dropin.utils.misc import Metrics
    torch.manual_seed(0)
    gt = torch.rand(1, 2048, 3, device=cuda) - 0.5
    pred = gt[:, torch.randperm(2048, device=cuda)] + 0.004 * torch.randn(1, 2048, 3, device=cuda)   # NN distances straddle th = 0.01
    f, cd, emd = Metrics.get(pred, gt)
    assert Metrics.names() == ["F-Score", "ChamferDistance", "EMD"]
    D = torch.cdist(pred[0].double(), gt[0].double())
    d1, d2 = D.min(1)[0], D.min(0)[0]
    precision, recall = (d1 < 0.01).double().mean().item(), (d2 < 0.01).double().mean().item()
    f_ref = 2 * recall * precision / (recall + precision)
    assert 0.05 < f_ref < 0.999 and abs(f - f_ref) < 2e-3                  # a point within fp32 rounding of the threshold may flip
    cd_ref = (d1.pow(2).mean() + d2.pow(2).mean()).item() * 1000
    assert abs(cd - cd_ref) <= 1e-5 * cd_ref
    assert 0 < emd < 100 * 0.02                                              # the matching cannot cost more than a few noise sigmas
    a, b = Metrics("F-Score", [f, cd, emd]), Metrics("F-Score", {"F-Score": f - 0.1, "ChamferDistance": cd, "EMD": emd})
    assert a.better_than(b) and not b.better_than(a) and a.better_than(None)
    assert Metrics("EMD", [f, cd, emd - 1]).better_than(Metrics("EMD", [f, cd, emd]))
    assert a.state_dict()["ChamferDistance"] == cd
