"""gpurun_out/<file>.csv (ncu --metrics gpu__time_duration.sum --csv of a bench run) -> profiles/<tag>_launches_summary.csv.
Keeps the functor of ATen's elementwise/reduce kernels so that the PyTorch glue between our kernels can be attributed.
  python tools/launch_summary.py gpurun_out/launches_r2.csv r2 [top]"""
import csv
import io
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60


def clean(name):
    name = name.replace("void ", "")
    m = re.match(r"at::(native::)?([a-z_]*(elementwise|reduce)[a-z_]*kernel)<(.*)", name)
    if m:
        inner = m.group(4)
        f = re.search(r"(direct_copy_kernel_cuda|CUDAFunctor_[a-z]+|[A-Za-z_0-9]+Functor[A-Za-z_0-9]*|[a-z_0-9]+_kernel_cuda|[a-z_0-9]+_kernel_impl|[a-z_0-9]+_kernel(?=[<(])|"
                      r"(Mean|Max|Min|Sum|Norm|Welford)[A-Za-z]*Ops?|func_wrapper_t<[a-z]+, at::native::[A-Za-z]+|[A-Za-z]+Ops)", inner)
        return "at::" + m.group(2) + "<" + (f.group(1)[:60] if f else inner[:60]) + ">"
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    if "gemm_tf32_kernel" in name:
        return "snb::gemm_tf32_kernel<xform>" if re.search(r"gemm_tf32_kernel<(\(bool\))?(1|true)>", name) else "snb::gemm_tf32_kernel<plain>"
    return re.sub(r"<.*", "", name).split("(")[0].strip()


lines = [ln for ln in open(src, errors="replace") if ln.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    k = clean(r["Kernel Name"])
    agg[k][0] += 1
    agg[k][1] += ms
tot = sum(v[1] for v in agg.values())
out = os.path.join(ROOT, "profiles", f"{tag}_launches_summary.csv")
with open(out, "w") as f:
    f.write("kernel,launches,total_ms,share\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        f.write(f"\"{k}\",{n},{ms:.3f},{ms / tot:.4f}\n")
    f.write(f"TOTAL,{sum(v[0] for v in agg.values())},{tot:.3f},1.0\n")
print(open(out).read())
