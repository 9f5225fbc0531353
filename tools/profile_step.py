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
if os.environ.get("GROUP", "") == "kernels":  # every device kernel: launches per step, average and total time, plus a size histogram
    ks = [e for e in prof.key_averages() if getattr(e, "device_type", None) is not None and "CUDA" in str(e.device_type) and e.self_device_time_total > 0]
    ks.sort(key=lambda e: -e.self_device_time_total)
    nk = sum(e.count for e in ks) / 2
    tt = sum(e.self_device_time_total for e in ks) / 2e3
    print(f"{nk:.0f} kernel launches / step, {tt:.2f} ms kernel time / step")
    for lim in (3, 6, 20, 100, 1e9):
        sel = [e for e in ks if e.self_device_time_total / e.count < lim]
        print(f"  avg < {lim:g} us: {sum(e.count for e in sel) / 2:.0f} launches, {sum(e.self_device_time_total for e in sel) / 2e3:.2f} ms")
    for e in ks:
        print(f"{e.self_device_time_total / 2e3:9.3f} ms/step  x{e.count / 2:<7.1f} avg {e.self_device_time_total / e.count:9.1f} us  {e.key[:140]}")
    sys.exit(0)
if os.environ.get("GROUP", "") == "ops":  # host-side ops by launch count (where do the tiny kernels come from?)
    es = sorted(prof.key_averages(group_by_stack_n=0), key=lambda e: -e.count)
    for e in es[:int(os.environ.get("ROWS", "80"))]:
        print(f"x{e.count / 2:<7.1f} self-device {e.self_device_time_total / 2e3:8.3f} ms/step  cpu {e.self_cpu_time_total / 2e3:8.3f} ms/step  {e.key[:100]}")
    sys.exit(0)
tab = prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=int(os.environ.get("ROWS", "45")), max_name_column_width=70)
keep = []
for line in tab.splitlines():
    keep.append(line[:72] + " | " + " ".join(line[72:].split()[-12:]))
print("\n".join(keep))
print("peak memory GB", torch.cuda.max_memory_allocated() / 2 ** 30)
