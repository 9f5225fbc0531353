"""EMD pruned-auction accounting (development tool): builds csrc/emd.cu + chamfer_bvh.cu with -DSNB_EMD_STATS into a private .so and
prints rounds, bidders, warp passes, super-box / leaf visits and the cycle split of sample 0.  python tools/emd_stats.py [kind] [N] [B]"""
import ctypes
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "sparenet_b200", "build", "libemd_stats.so")
if not os.path.exists(so) or "--build" in sys.argv:
    os.makedirs(os.path.dirname(so), exist_ok=True)
    csrc = os.path.join(ROOT, "sparenet_b200", "csrc")
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
                    "-Xcompiler", "-fPIC", "-DSNB_EMD_STATS", "-shared", "-I", os.path.join(ROOT, "include"),
                    os.path.join(csrc, "emd.cu"), os.path.join(csrc, "chamfer_bvh.cu"), os.path.join(csrc, "chamfer.cu"), "-o", so], check=True)
if "--build" in sys.argv:
    sys.exit(0)
lib = ctypes.CDLL(so)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
kind = args[0] if args else "iid"
N = int(args[1]) if len(args) > 1 else 16384
B = int(args[2]) if len(args) > 2 else 32
dev = torch.device("cuda:0")
torch.manual_seed(4)
y = torch.rand(B, N, 3, device=dev) - 0.5
if kind == "iid":
    x = torch.rand(B, N, 3, device=dev) - 0.5
elif kind == "near":
    x = torch.stack([y[b, torch.randperm(N, device=dev)] for b in range(B)]) + 0.01 * torch.randn(B, N, 3, device=dev)
else:
    x = 0.05 * torch.randn(B, N, 3, device=dev)
P = ctypes.c_void_p
lib.snb_emd_workspace_bytes.restype = ctypes.c_size_t
lib.snb_emd_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int]
lib.snb_emd_fwd.argtypes = [P, P, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int, P, P, P, ctypes.c_size_t, P]
nb = lib.snb_emd_workspace_bytes(B, N)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
dist = torch.empty(B, N, device=dev)
ass = torch.empty(B, N, dtype=torch.int32, device=dev)
out = (ctypes.c_ulonglong * 16)()


def run():
    rc = lib.snb_emd_fwd(x.data_ptr(), y.data_ptr(), B, N, 0.005, 50, dist.data_ptr(), ass.data_ptr(), ws.data_ptr(), nb, None)
    assert rc == 0, rc


run()
torch.cuda.synchronize()
lib.snb_emd_debug_stats(out, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
run()
b.record()
torch.cuda.synchronize()
lib.snb_emd_debug_stats(out, 1)
s = list(out)
ms = a.elapsed_time(b)
rounds, sumU = s[0], s[1]
print(f"{kind} B={B} N={N}: {ms:.3f} ms (stats build); sample 0: {rounds} rounds, sum U = {sumU} = {sumU / N:.1f} n")
print(f"  all samples: warp passes {s[2]}, home leaves scanned {s[9] / max(s[2], 1):.1f}/pass, super-box visits {s[3] / max(s[2], 1):.1f}/pass, "
      f"leaves re-tested {s[8] / max(s[2], 1):.1f}/pass, leaves scanned {s[4] / max(s[2], 1):.1f}/pass of {N // 32}, "
      f"lanes needing a scanned leaf {s[5] / max(s[4], 1):.1f}/32, {s[12] / max(s[2], 1):.0f} cycles/pass")
r_ = max(rounds, 1)
print(f"  sample 0 thread 0, cycles/round: whole {s[7] / r_:.0f} = prices+compaction {s[10] / r_:.0f} | Bid {s[6] / r_:.0f} | wait+barrier {s[14] / r_:.0f} | "
      f"GetMax {s[11] / r_:.0f} | barrier+Assign+barrier {s[13] / r_:.0f}")
