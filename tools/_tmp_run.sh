timeout 300 python -m pytest tests/test_gpu_fused.py tests/test_gpu_generator.py tests/test_gpu_baseline_config.py tests/test_gpu_ops.py -m gpu -q -x -k "linear or generator or trajectory or p2i" > gpurun_out/pytest_s2j.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2j.log; grep "\[linear B=32 K=4096 O=4096\]" gpurun_out/pytest_s2j.log | head -4
timeout 200 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_s2j.json 2> gpurun_out/bench_s2j.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s2j.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"])
print({k:v for k,v in d["roofline"]["ops_ms_per_step"].items() if "linear" in k})
print({k:v for k,v in d["roofline"]["hbm_bound_kernels"].items() if "linear" in k})
PY
tail -3 gpurun_out/bench_s2j.err
