timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "knn" > gpurun_out/pytest_s2h.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_s2h.log
for v in 1 0; do
SNB_KNN_PRUNE_SPLIT=$v timeout 200 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_s2h_$v.json 2> gpurun_out/bench_s2h_$v.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s2h_$v.json").read().strip().splitlines()[-1])
print("split=$v", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["roofline"]["ops_ms_per_step"].get("knn"))
PY
done
