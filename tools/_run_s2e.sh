timeout 400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_generator.py tests/test_gpu_graph.py tests/test_gpu_baseline_config.py -m gpu -q -x > gpurun_out/pytest_s2e.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2e.log
timeout 200 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_s2e.json 2> gpurun_out/bench_s2e.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s2e.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
PY
timeout 150 python tools/timeline_step.py s2e > gpurun_out/timeline_s2e.txt 2>&1; head -12 gpurun_out/timeline_s2e.txt | tail -10
