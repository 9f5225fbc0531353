timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_s2c.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2c.log
timeout 200 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_s2c.json 2> gpurun_out/bench_s2c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_s2c.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
print({k:v for k,v in d["roofline"]["ops_ms_per_step"].items() if k in ("knn","row_stats","adain_tail_fwd","adain_tail_bwd","gemm_stats_merge","mds_sample","chamfer_fwd","expansion_fwd")})
PY
timeout 150 python tools/timeline_step.py s2c > gpurun_out/timeline_s2c.txt 2>&1; head -30 gpurun_out/timeline_s2c.txt | tail -28
