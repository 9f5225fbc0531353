"""Design study for round 2 (CPU, numpy; development tool): how much of the MDS update work could warps SKIP exactly, if every
group of 32 points held in the same register slot of a warp were spatially compact (Morton order)?

A (pick, point) update adds w = fac * exp(-d/t) to the point's density; it is a no-op in fp32 when w < ulp(density)/2.  A group can
be skipped for a pick when that holds for ALL its live points -- decidable from the group's bounding box and its minimum live
density without touching the points.  This script replays the sampler on one sample of the dumped bench inputs
(tools/dump_mds_inputs.py) and reports, per phase of the run, the fraction of (pick, group) pairs that are skippable
  (a) by the exact per-point criterion (upper limit for any layout),
  (b) by the conservative box criterion with Morton-sorted groups of 32,
  (c) by the box criterion with the current index-order groups (k mod 4 CTA deal, then consecutive).
    python tools/mds_cull_study.py [call 0|1] [sample] [picks]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
call = int(sys.argv[1]) if len(sys.argv) > 1 else 0
sample = int(sys.argv[2]) if len(sys.argv) > 2 else 0
npick = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
x, m, mml = torch.load(os.path.join(ROOT, "tools", "_data", "mds_inputs.pt"))[call]
xyz = x[sample].numpy().astype(np.float32)
n = len(xyz)
t = np.float32(5.0 * float(mml[sample]) ** 2)
fac = np.where(np.arange(n) < 8192, np.float32(1), np.float32(2))


def morton(p):
    q = ((p - p.min(0)) / (p.max(0) - p.min(0) + 1e-9) * 1023).astype(np.uint64)
    code = np.zeros(len(p), dtype=np.uint64)
    for b in range(10):
        for a in range(3):
            code |= ((q[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
    return code


def groups_of(order):
    g = np.full(n, -1, dtype=np.int64)
    g[order] = np.arange(n) // 32
    return g


g_morton = groups_of(np.argsort(morton(xyz), kind="stable"))
g_index = groups_of(np.argsort(np.arange(n) % 4, kind="stable"))     # the kernel's deal: CTA = k mod 4, then index order
ng = (n + 31) // 32


def boxes(g):
    lo = np.full((ng, 3), np.inf, dtype=np.float32)
    hi = np.full((ng, 3), -np.inf, dtype=np.float32)
    np.minimum.at(lo, g, xyz)
    np.maximum.at(hi, g, xyz)
    return lo, hi


box = {"morton": boxes(g_morton), "index": boxes(g_index)}
gid = {"morton": g_morton, "index": g_index}
temp = np.zeros(n, dtype=np.float32)
live = np.ones(n, dtype=bool)
live[0] = False
last = 0
stats = {k: [] for k in ("exact", "morton", "index", "points")}
key = np.arange(n)
for j in range(1, npick + 1):
    d = xyz - xyz[last]
    d2 = (d[:, 2] * d[:, 2] + (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])).astype(np.float32)
    w = (np.exp(-(d2 / t).astype(np.float32)).astype(np.float32) * fac).astype(np.float32)
    new = (temp + w).astype(np.float32)
    changed = (new != temp) & live                                          # per-point: the update is not a no-op
    stats["points"].append(changed.sum() / max(live.sum(), 1))
    for name in ("morton", "index"):
        g = gid[name]
        lo, hi = box[name]
        # exact skippability of a group: none of its live points changes
        if name == "morton":
            ch = np.zeros(ng, dtype=bool)
            np.logical_or.at(ch, g[live], changed[live])
            has = np.zeros(ng, dtype=bool)
            has[g[live]] = True
            stats["exact"].append(1.0 - ch[has].mean())
        # conservative box criterion: max weight over the box < half an ulp of the group's minimum live density
        dd = np.maximum(np.maximum(lo - xyz[last], xyz[last] - hi), 0)
        dmin = (dd * dd).sum(1) * np.float32(0.9999)
        wmax = 2.0 * np.exp(-dmin / t)
        tmin = np.full(ng, np.inf, dtype=np.float32)
        np.minimum.at(tmin, g[live], temp[live])
        has = np.isfinite(tmin)
        skip = wmax < tmin * np.float32(2.0 ** -25)                         # < ulp/2 for every density >= tmin
        stats[name].append(skip[has].mean())
        chg = np.zeros(ng, dtype=bool)
        np.logical_or.at(chg, g[live], changed[live])
        assert not (skip & chg).any(), "the box criterion skipped a group with a point whose density changes"   # exactness
    temp = new
    cand = np.where(live, temp, np.float32(np.inf))
    last = int(np.lexsort((key, cand))[0])
    live[last] = False
for a, b in ((0, npick // 4), (npick // 4, npick // 2), (npick // 2, npick)):
    print(f"call {call} sample {sample} picks {a:5d}..{b:5d}: points really changed {np.mean(stats['points'][a:b]):6.1%} | groups skippable: exact (Morton groups) "
          f"{np.mean(stats['exact'][a:b]):6.1%}, box test Morton {np.mean(stats['morton'][a:b]):6.1%}, box test index order {np.mean(stats['index'][a:b]):6.1%}")
