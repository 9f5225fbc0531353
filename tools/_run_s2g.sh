timeout 500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_s2g.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2g.log
timeout 200 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_s2g.json 2> gpurun_out/bench_s2g.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s2g.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"])
PY
tail -3 gpurun_out/bench_s2g.err
