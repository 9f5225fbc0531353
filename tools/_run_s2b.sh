timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_generator.py tests/test_gpu_fused.py tests/test_gpu_graph.py -m gpu -q -x -k "knn or transpose or generator or decoder or tail or graph" > gpurun_out/pytest_s2b.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_s2b.log
timeout 200 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_s2b.json 2> gpurun_out/bench_s2b.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_s2b.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
print({k:v for k,v in d["roofline"]["ops_ms_per_step"].items() if k in ("knn","row_stats","adain_tail_fwd","adain_tail_bwd","row_affine_act_fwd","row_affine_act_bwd")})
PY
timeout 150 python tools/timeline_step.py s2b > gpurun_out/timeline_s2b.txt 2>&1; head -8 gpurun_out/timeline_s2b.txt | tail -6
