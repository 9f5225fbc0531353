"""Design study for round 2 (CPU, numpy; development tool): how much of the EMD auction's Bid work would a price-aware spatial grid
prune?  Bid needs, per unassigned bidder, the best and second-best value v_k = 3 - |x_i - y_k| - price_k over ALL objects.  With
the objects binned into a uniform grid, v_k <= 3 - dist(x_i, cell box) - min price(cell) bounds a whole cell, so cells whose bound is
below the bidder's current second-best value need not be visited (exactly: the result is unchanged).  The script runs a plain
auction (eps 0.005, 50 rounds) on iid clouds and reports, per round, the number of bidders and the fraction of the (bidder, object)
pairs that lie in cells that survive the bound when cells are visited nearest-first.
    python tools/emd_grid_study.py [N] [cells per axis]"""
import sys

import numpy as np

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
G = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rng = np.random.default_rng(4)
x = rng.random((N, 3), dtype=np.float32)
y = rng.random((N, 3), dtype=np.float32)
price = np.zeros(N, dtype=np.float64)
assign = np.full(N, -1)
inv = np.full(N, -1)
cell = np.minimum((y * G).astype(np.int64), G - 1)
cid = (cell[:, 0] * G + cell[:, 1]) * G + cell[:, 2]
order = np.argsort(cid, kind="stable")
counts = np.bincount(cid, minlength=G ** 3)
ii, jj, kk = np.meshgrid(np.arange(G), np.arange(G), np.arange(G), indexing="ij")
clo = np.stack([ii, jj, kk], -1).reshape(-1, 3).astype(np.float32) / G
chi = clo + np.float32(1.0 / G)
tot_pairs = tot_kept = 0
for it in range(50):
    un = np.nonzero(assign < 0)[0]
    if len(un) == 0:
        break
    d = np.sqrt(((x[un, None, :] - y[None, :, :]) ** 2).sum(-1))
    v = 3.0 - d - price[None, :]
    best_i = v.argmax(1)
    best = v[np.arange(len(un)), best_i]
    v2 = v.copy()
    v2[np.arange(len(un)), best_i] = -np.inf
    better = v2.max(1)
    # grid bound per (bidder, cell)
    pmin = np.full(G ** 3, np.inf)
    np.minimum.at(pmin, cid, price)
    dd = np.maximum(np.maximum(clo[None] - x[un, None, :], x[un, None, :] - chi[None]), 0)
    bound = 3.0 - np.sqrt((dd * dd).sum(-1)) - pmin[None, :]
    keep = bound >= better[:, None]                     # cells that could still hold the best or second best
    kept_pairs = (keep * counts[None, :]).sum()
    tot_pairs += len(un) * N
    tot_kept += kept_pairs
    if it < 6 or it % 10 == 9:
        print(f"round {it:2d}: bidders {len(un):6d}  pairs kept by the grid bound {kept_pairs / (len(un) * N):7.2%}  (cells visited per bidder {keep.sum(1).mean():6.1f} of {G ** 3})")
    inc = best - better + 0.005
    # GetMax / Assign: the largest increment wins each object (ties: larger bidder index)
    win = {}
    for b, o, c in zip(un, best_i, inc):
        if o not in win or (c, b) > win[o]:
            win[o] = (c, b)
    for o, (c, b) in win.items():
        if inv[o] >= 0:
            assign[inv[o]] = -1
        inv[o] = b
        assign[b] = o
        price[o] += c
print(f"N={N} grid {G}^3: {tot_kept / tot_pairs:.2%} of the {tot_pairs:.3g} (bidder, object) pairs survive the bound -> {tot_pairs / tot_kept:.1f}x less Bid work")
