#!/bin/bash
# Round-end evidence: GPU parity tests, the bench line, the reference-extension step on the same GPU, the ncu launch list of one
# eager step and full captures of the top kernels (summarised into profiles/ by tools/summarize_profiles.py).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; head -c 300 gpurun_out/bench.json; echo
timeout 400 python tests/perf/refstep.py > gpurun_out/refstep.txt 2>&1; echo "refstep exit $?"; tail -2 gpurun_out/refstep.txt
NSTEPS=4 NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu launches exit $?"
for k in mds_cluster_kernel knn_dist_kernel chamfer_bvh_query_kernel row_norm_act_bwd_kernel edge_reduce_bwd_kernel expansion_kernel; do
  NSTEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/full_$k \
    python tools/ncu_step.py > gpurun_out/ncu_full_$k.log 2>&1; echo "ncu full $k exit $?"
done
