"""Turn the raw ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (development tool).
  python tools/summarize_profiles.py r1
* gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum --csv of one eager bench step) -> profiles/<tag>_launches_summary.csv
* gpurun_out/full_<kernel>.ncu-rep (ncu --set full)                                           -> profiles/<tag>_ncu_full_<kernel>.csv
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_", "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_fmaheavy_cycles_active.avg.pct", "sm__pipe_alu_cycles_active.avg.pct",
        "sm__pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg", "sm__cycles_elapsed.max")


def launches():
    src = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if not os.path.exists(src):
        return
    lines = [ln for ln in open(src, errors="replace") if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<.*", "", r["Kernel Name"]).replace("void ", "").split("(")[0].strip()
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches_summary.csv"), "w") as f:
        f.write("kernel,launches,total_ms,share\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{k},{n},{ms:.3f},{ms / tot:.4f}\n")
        f.write(f"TOTAL,{sum(v[0] for v in agg.values())},{tot:.3f},1.0\n")
    print("launch list:", len(rows), "rows ->", f"{tag}_launches_summary.csv", f"(total {tot:.1f} ms)")


def full(rep):
    kern = os.path.basename(rep)[len("full_"):-len(".ncu-rep")]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(os.path.join(out_dir, f"{tag}_ncu_full_{kern}.csv"), "w") as f:
        f.write("metric,unit,value\n")
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name" or any(h.startswith(k) or k in h for k in KEEP):
                if v not in ("", "0", "0.000000") or h == "Kernel Name":
                    f.write(f'{h},{u},"{v}"\n' if "," in v else f"{h},{u},{v}\n")
    print("full capture:", kern)


launches()
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "full_*.ncu-rep"))):
    if len(sys.argv) > 2 and not any(k in rep for k in sys.argv[2:]):
        continue
    full(rep)
