#!/bin/bash
# second evidence round: fixed tests, reference step, kernel-level profile, clean ncu launch list + full captures
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 500 python tests/perf/refstep.py > gpurun_out/refstep.txt 2>&1; echo "refstep exit $?"; tail -2 gpurun_out/refstep.txt
GROUP=kernels timeout 300 python tools/profile_step.py > gpurun_out/profile_kernels.txt 2>&1
GROUP=ops ROWS=120 timeout 300 python tools/profile_step.py > gpurun_out/profile_ops.txt 2>&1
# launch list of ONE eager step after 3 warm-up steps (tools/ncu_step.py runs NSTEPS steps; profile only the last via cudaProfilerApi range)
NSTEPS=4 NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu launches exit $?"
for k in mds_cluster_kernel knn_dist_kernel chamfer_bvh_query_kernel knn_topk_kernel; do
  NSTEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/full_$k \
    python tools/ncu_step.py > gpurun_out/ncu_full_$k.log 2>&1; echo "ncu full $k exit $?"
done
