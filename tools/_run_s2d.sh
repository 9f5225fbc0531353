timeout 400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_generator.py tests/test_gpu_gemm.py tests/test_gpu_graph.py tests/test_gpu_baseline_config.py -m gpu -q -x > gpurun_out/pytest_s2d.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2d.log
for v in 1 0; do
SNB_KNN_PRUNE_CACHE=$v timeout 200 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_s2d_$v.json 2> gpurun_out/bench_s2d_$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s2d_$v.json").read().strip().splitlines()[-1])
print("cache=$v", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
print({k:v for k,v in d["roofline"]["ops_ms_per_step"].items() if k in ("knn","adain_tail_fwd","adain_tail_bwd","gemm_stats_merge","bn_max_tail_fwd","bn_max_tail_bwd","conv_extrema_bwd")})
PY
done
timeout 150 python tools/timeline_step.py s2d > gpurun_out/timeline_s2d.txt 2>&1; head -12 gpurun_out/timeline_s2d.txt | tail -10
