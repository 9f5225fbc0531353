"""MDS cycle accounting (development tool): builds csrc/mds.cu with -DSNB_MDS_STATS into a private .so and prints where block 0's
replay warp and worker warp 0 spend their cycles.  `python tools/mds_stats.py [bench|uniform] [n] [m] [B]`"""
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
extra = os.environ.get("MDS_STATS_FLAGS", "").split()
so = os.path.join(ROOT, "sparenet_b200", "build", "libmds_stats" + "".join(f.replace("-D", "_") for f in extra) + ".so")
if not os.path.exists(so) or "--build" in sys.argv:
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
                    "-Xcompiler", "-fPIC", "-DSNB_MDS_STATS", *extra, "-shared", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "sparenet_b200", "csrc", "mds.cu"), "-o", so], check=True)
if "--build" in sys.argv:
    sys.exit(0)
lib = ctypes.CDLL(so)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
kind = args[0] if args else "uniform"
n = int(args[1]) if len(args) > 1 else 18432
m = int(args[2]) if len(args) > 2 else 16384
B = int(args[3]) if len(args) > 3 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
if kind.startswith("file"):   # file0 / file1: the MDS calls of a bench step dumped by tools/dump_mds_inputs.py
    calls = torch.load(os.path.join(ROOT, "tools", "_data", "mds_inputs.pt"))
    x, m, mml = calls[int(kind[4:] or 0)]
    x, mml = x.to(dev).contiguous(), mml.to(dev).contiguous()
    B, n = x.shape[0], x.shape[1]
elif kind == "uniform":
    x = torch.rand(B, n, 3, device=dev) - 0.5
    mml = torch.full((B,), 0.01, device=dev)
else:  # surface-like: points on a sphere + noise
    x = torch.nn.functional.normalize(torch.randn(B, n, 3, device=dev), dim=-1) * 0.5 + 0.002 * torch.randn(B, n, 3, device=dev)
    mml = torch.full((B,), 0.006, device=dev)
idx = torch.empty(B, m, dtype=torch.int32, device=dev)
P = ctypes.c_void_p
lib.snb_mds_sample.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P, P, ctypes.c_size_t, P]
out = (ctypes.c_ulonglong * 16)()


def run():
    rc = lib.snb_mds_sample(x.data_ptr(), B, n, m, mml.data_ptr(), idx.data_ptr(), None, 0, None)
    assert rc == 0, rc


run()
torch.cuda.synchronize()
lib.snb_mds_debug_stats(out, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
run()
b.record()
torch.cuda.synchronize()
lib.snb_mds_debug_stats(out, 1)
s = list(out)
ms = a.elapsed_time(b)
gens, picks = max(s[0], 1), max(s[1], 1)
print(f"[{kind}] n={n} m={m} B={B} layout={os.environ.get('SNB_MDS_LAYOUT', 'default')} M={os.environ.get('SNB_MDS_M', 'default')}: {ms:.3f} ms, "
      f"{gens} generations, {picks / gens:.1f} picks/gen, pool<theta {s[5] / gens:.1f}")
print(f"  replay warp cycles/gen: wait-pool {s[2] / gens:.0f}  compact {s[3] / gens:.0f}  replay {s[4] / gens:.0f} ({s[4] / picks:.0f}/pick)   total {sum(s[2:5]) / 1e6:.2f} Mcycles")
print(f"  dead time/gen: replay END -> worker 0 applied all {s[11] / gens:.0f}; worker 0 published -> pool complete at the replay warp {s[6] / gens:.0f} (last worker warp of this CTA: {s[7] / gens:.0f})")
print("  pool complete at block 0 - last publish of rank r (ns/gen): " + "  ".join(f"r{i}: {s[12 + i] / gens:.0f}" for i in range(4)))
print(f"  worker warp0 cycles/gen: select+publish {s[8] / gens:.0f}  apply {s[9] / gens:.0f} ({s[9] / picks:.0f}/pick)  wait-picks {s[10] / gens:.0f}   total {sum(s[8:11]) / 1e6:.2f} Mcycles")
