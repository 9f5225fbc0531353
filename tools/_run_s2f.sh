timeout 400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_generator.py tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/pytest_s2f.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2f.log
timeout 200 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_s2f.json 2> gpurun_out/bench_s2f.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s2f.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
print({k:v for k,v in d["roofline"]["ops_ms_per_step"].items() if "tail" in k})
PY
