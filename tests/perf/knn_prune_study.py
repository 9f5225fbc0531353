"""Feasibility study (CPU, development tool; imports the oracle, hence under tests/): kNN for the C >= 256 EdgeConv layers with a
TF32 tensor-core Gram matrix as a PRUNING filter followed by exact fp32 re-evaluation ("MMA is admissible only as a
pruning filter").  Rule: with d~(i,j) = |a_i|^2 + |a_j|^2 - 2 G~(i,j) (G~ from TF32-rounded operands, fp32 accumulation) and
eps_i = 2^-9 |a_i| max_j|a_j| (bound on |d~ - d|), the candidates of row i are {j : d~(i,j) <= kth_smallest_j d~(i,j) + 2 eps_i};
they provably contain the exact top-k.  This script checks that on the encoder's real features (random-init network, the bench's
input distribution) and reports how many candidates per row survive -- the cost of the exact re-evaluation.
    python tests/perf/knn_prune_study.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import generator_ref as G  # noqa: E402


def tf32(x):
    """TRUNCATION to 10 explicit mantissa bits: the worst case of what a tensor core does to fp32 operands (the bound in
    csrc/knn_prune.cu is sized for it; round-to-nearest is twice as accurate)"""
    u = x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def study(name, x, k=8):
    C, N = x.shape
    a = x.T.astype(np.float32)                                  # [N, C]
    d_exact = ((a[:, None, :].astype(np.float64) - a[None, :, :].astype(np.float64)) ** 2).sum(-1)
    nrm = (a.astype(np.float64) ** 2).sum(1)
    g = tf32(a) @ tf32(a).T                                     # fp32 accumulation of TF32-rounded operands
    d_apx = nrm[:, None] + nrm[None, :] - 2.0 * g.astype(np.float64)
    err = np.abs(d_apx - d_exact)
    an = np.sqrt(nrm)
    eps = 2.0 ** -7.5 * an * an.max() + 2.0 ** -13 * (nrm + nrm.max())       # the rule of csrc/knn_prune.cu
    assert (err <= eps[:, None]).all(), "error bound violated"
    kth = np.partition(d_apx, k - 1, axis=1)[:, k - 1]
    cand = d_apx <= (kth + np.abs(kth) * 2.0 ** -10 + 2 * eps)[:, None]
    top = np.argsort(d_exact, axis=1, kind="stable")[:, :k]
    assert np.take_along_axis(cand, top, 1).all(), "a true neighbour was pruned"
    cnt = cand.sum(1)
    dk = np.sort(d_exact, axis=1)[:, k - 1]
    print(f"{name}: C={C} N={N}  |a|^2 mean {nrm.mean():9.3g}  k-th distance mean {dk.mean():9.3g}  max |d~-d| {err.max():8.2e} (bound {eps.max():8.2e})"
          f"  candidates/row: mean {cnt.mean():7.1f}  p50 {np.median(cnt):6.0f}  p99 {np.percentile(cnt, 99):6.0f}  max {cnt.max()}")
    return cnt.mean()


torch.manual_seed(0)
enc = G.EdgeConvResFeat(use_SElayer=True, k=8, output_size=4096, hide_size=4096).train()
enc.apply(G.init_weights)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = (torch.rand(2, N, 3, generator=torch.Generator().manual_seed(1)) - 0.5).transpose(1, 2).contiguous()
with torch.no_grad():
    x1 = enc._block(x, enc.conv1, enc.bn1, enc.se1)
    x2 = enc._block(x1, enc.conv2, enc.bn2, enc.se2) + enc.resconv1(x1)
    x3 = enc._block(x2, enc.conv3, enc.bn3, enc.se3) + enc.resconv2(x2)
for nm, t in (("x1 -> kNN of layer 2", x1), ("x2 -> kNN of layer 3", x2), ("x3 -> kNN of layer 4", x3)):
    study(nm, t[0].numpy())
z = x3[1].numpy().copy()
z[:, 900:] = 0                       # zero-padded tail and duplicated points: the degenerate rows
z[:, 500:560] = z[:, 100:160]
study("x3 with zero / duplicate points", z)
