timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_generator.py tests/test_gpu_baseline_config.py tests/test_gpu_ops.py -m gpu -q -x -k "edge or generator or knn or trajectory or graph" > gpurun_out/pytest_s2i.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2i.log
timeout 200 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_s2i.json 2> gpurun_out/bench_s2i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s2i.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"])
PY
tail -3 gpurun_out/bench_s2i.err
