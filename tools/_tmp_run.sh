timeout 400 python -m pytest tests/test_gpu_fused.py tests/test_gpu_generator.py tests/test_gpu_baseline_config.py tests/test_gpu_graph.py tests/test_gpu_gan.py -m gpu -q -x > gpurun_out/pytest_s2l.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2l.log
timeout 200 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_s2l.json 2> gpurun_out/bench_s2l.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s2l.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"])
print({k:v for k,v in d["roofline"]["hbm_bound_kernels"].items() if "thin" in k})
PY
tail -3 gpurun_out/bench_s2l.err
timeout 150 python tools/timeline_step.py s2l > gpurun_out/timeline_s2l.txt 2>&1; grep -E "span|thin|library gemm|at::" gpurun_out/timeline_s2l.txt | head
