"""GPU diagnostics #2: who wins the reference's GetMax race (emd_cuda.cu:181-194)?  Crafted duplicates make two
bidders produce identical bids/increments in round 0 (iters=1), then max_idx shows the hardware's winner."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402

dev = torch.device("cuda:0")
EMD = build_ref.load_ref("emd")
n, B = 8192, 1
pairs = [(5, 17), (5, 31), (33, 40), (5, 100), (5, 1000), (700, 1023), (5, 2000), (1023, 1024), (3000, 7000), (100, 8191), (4096, 4097),
         (10, 5000), (6000, 6100), (2047, 2048), (64, 96)]
for trial in range(2):
    torch.manual_seed(40 + trial)
    x, y = torch.rand(B, n, 3, device=dev), torch.rand(B, n, 3, device=dev)
    for (a, b) in pairs:
        x[0, b] = x[0, a]

    def z(*s, dt=torch.float32):
        return torch.zeros(*s, device=dev, dtype=dt)

    dist = z(B, n)
    assignment = z(B, n, dt=torch.int32) - 1
    assignment_inv = z(B, n, dt=torch.int32) - 1
    price, bid, bid_inc, max_inc = z(B, n), z(B, n, dt=torch.int32), z(B, n), z(B, n)
    unass_idx, max_idx = z(B * n, dt=torch.int32), z(B * n, dt=torch.int32)
    c1, c2, c3 = z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32)
    EMD.forward(x, y, dist, assignment, price, assignment_inv, bid, bid_inc, max_inc, unass_idx, c1, c2, c3, max_idx, 0.005, 1)
    torch.cuda.synchronize()
    out = []
    for (a, b) in pairs:
        o = int(bid[0, a])
        assert int(bid[0, b]) == o and float(bid_inc[0, a]) == float(bid_inc[0, b])
        # is (a,b) the top bidder pair on o?
        others = ((bid[0] == o).nonzero().flatten().tolist())
        top = float(bid_inc[0][bid[0] == o].max())
        qual = abs(float(bid_inc[0, a]) - top) <= 1e-6
        out.append(((a, b), int(max_idx[o]), "qual" if qual else "notmax", len(others)))
    print(f"[emd-race] trial {trial}:", out)
