"""Stand-in for the third-party `knn_cuda` wheel (KNN_CUDA 0.2, setup_env.sh:5) at the one call site the
generator has: KNN(k=k, transpose_mode=True)(ref, query) -> (dist, idx) with ref == query
(models/sparenet_generator.py:865-869).  Served by snb_knn; idx is int64 like the wheel's.  Only the
self-query case (ref is query) is implemented -- that is the only one on the path.
"""
import torch

from sparenet_b200 import functional as F_


class KNN(torch.nn.Module):
    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k, self.transpose_mode = k, transpose_mode

    def forward(self, ref, query):
        if ref.data_ptr() != query.data_ptr() or ref.shape != query.shape:
            raise NotImplementedError("sparenet_b200.knn_cuda serves the self-query case only (ref is query)")
        x = ref.transpose(1, 2).contiguous() if self.transpose_mode else ref.contiguous()  # -> [B, C, N]
        idx = F_.knn_indices(x, self.k).long()  # [B, N, k]
        xt = x.transpose(1, 2)  # [B, N, C]
        nb = torch.gather(xt.unsqueeze(1).expand(-1, xt.size(1), -1, -1), 2, idx.unsqueeze(-1).expand(-1, -1, -1, xt.size(2)))
        dist = (nb - xt.unsqueeze(2)).pow(2).sum(-1).sqrt()
        if not self.transpose_mode:
            dist, idx = dist.transpose(1, 2).contiguous(), idx.transpose(1, 2).contiguous()
        return dist, idx
