"""torch.profiler breakdown of one bench step on the GPU box (development tool): top CUDA kernels by total time."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class A:
    batch = int(os.environ.get("PB", "32"))


dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
step, hp, hg = bench.build_gpu(A, dev, 0)
p, g = hp.to(dev), hg.to(dev)
for _ in range(3):
    step(p, g)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile, record_function  # noqa: E402

BY_SHAPE = os.environ.get("GROUP", "") == "shape"
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=BY_SHAPE) as prof:
    for _ in range(2):
        step(p, g)
    torch.cuda.synchronize()
if BY_SHAPE:  # aten ops by input shape: which tensors the remaining torch time is spent on
    evs = sorted(prof.key_averages(group_by_input_shape=True), key=lambda e: -e.self_device_time_total)
    for e in evs[:int(os.environ.get("ROWS", "60"))]:
        print(f"{e.self_device_time_total / 2e3:9.3f} ms/step  x{e.count // 2:<5d} {e.key[:48]:48s} {str(e.input_shapes)[:110]}")
    sys.exit(0)
tab = prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=int(os.environ.get("ROWS", "45")), max_name_column_width=70)
keep = []
for line in tab.splitlines():
    keep.append(line[:72] + " | " + " ".join(line[72:].split()[-12:]))
print("\n".join(keep))
print("peak memory GB", torch.cuda.max_memory_allocated() / 2 ** 30)
