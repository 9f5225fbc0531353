"""Real per-kernel timeline of ONE graphed bench step via torch.profiler (CUPTI kernel activity records: true start/duration inside the
graph replay, unlike ncu's serialised cold-cache replays).  Writes gpurun_out/timeline_<tag>.csv (name,start_us,dur_us) and prints
busy time, idle gaps and the top kernels.  python tools/timeline_step.py [tag]"""
import os
import re
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "step"


class A:
    batch = 32


dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
step, hp, hg = bench.build_gpu(A, dev, 0)
p, g = hp.to(dev), hg.to(dev)
for _ in range(4):
    step(p, g)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(p, g)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
rows = sorted(((e.name, e.time_range.start, e.time_range.end - e.time_range.start) for e in ev if "memcpy" not in e.name.lower() or True), key=lambda r: r[1])
if not rows:
    print("no CUDA activity records (CUPTI unavailable?)")
    sys.exit(0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"timeline_{tag}.csv"), "w") as f:
    f.write("name,start_us,dur_us\n")
    for n, s, d in rows:
        f.write('"%s",%.3f,%.3f\n' % (n.replace('"', "'")[:160], s, d))
t0, t1 = rows[0][1], max(s + d for _, s, d in rows)
busy = 0.0
cur_end = t0
for _, s, d in rows:          # union of intervals
    if s + d > cur_end:
        busy += s + d - max(s, cur_end)
        cur_end = s + d


def cat(n):
    if "snb::" in n:
        m = re.search(r"snb::(?:\(anonymous namespace\)::|<unnamed>::)?([a-z_0-9]+)", n)
        return "snb::" + (m.group(1) if m else "?")
    if n.startswith("void at::") or "at::native" in n:
        return "at::*"
    if "cutlass" in n or "gemm" in n.lower() or "cublas" in n.lower():
        return "library gemm"
    return "other"


agg = defaultdict(lambda: [0, 0.0])
for n, s, d in rows:
    a = agg[cat(n)]
    a[0] += 1
    a[1] += d
print(f"span {(t1 - t0) / 1e3:.3f} ms, {len(rows)} records, busy (union) {busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms")
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"  {k:40s} {c:5d} launches {d / 1e3:8.3f} ms")
