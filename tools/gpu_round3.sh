#!/bin/bash
# Round-end evidence: GPU parity tests, smoke, the bench line, per-op timings beside the reference extensions, the ncu launch list of
# one eager step and full captures of the top kernels (summarised into profiles/ by tools/summarize_profiles.py).
# REFSTEP=1 adds the reference generator + extensions step on the same GPU (tests/perf/refstep.py, ~1 min).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; head -c 300 gpurun_out/bench.json; echo
timeout 400 python tests/perf/opbench.py > gpurun_out/opbench.txt 2>&1; echo "opbench exit $?"
if [ "${REFSTEP:-0}" = "1" ]; then
  timeout 400 python tests/perf/refstep.py > gpurun_out/refstep.txt 2>&1; echo "refstep exit $?"; tail -2 gpurun_out/refstep.txt
fi
NSTEPS=4 NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu launches exit $?"
NSTEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mds_cluster_kernel -c 1 -f -o gpurun_out/full_mds_cluster_kernel \
    python tools/ncu_step.py > gpurun_out/ncu_full_mds.log 2>&1; echo "ncu full mds exit $?"
NSTEPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_dist_kernel -s 3 -c 1 -f -o gpurun_out/full_knn_dist_kernel_c512 \
    python tools/ncu_step.py > gpurun_out/ncu_full_knn.log 2>&1; echo "ncu full knn exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:emd_auction -c 1 -f -o gpurun_out/full_emd_auction_kernel \
    python tools/emd_once.py 16384 50 > gpurun_out/ncu_full_emd.log 2>&1; echo "ncu full emd exit $?"
