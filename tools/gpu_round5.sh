#!/bin/bash
# Second-session evidence of round 2 in one gpurun call: GPU parity tests, smoke, the default bench line (with the reference-extension
# arm and the CPU arm), the GAN-config line, the CUPTI timeline of the graphed step and the ncu launch list of one eager step.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2s2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_r2s2.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_r2s2.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_r2s2.log
timeout 900 python bench.py > gpurun_out/bench_r2s2_final.json 2> gpurun_out/bench_r2s2_final.err; echo "bench exit $?"; head -c 300 gpurun_out/bench_r2s2_final.json; echo
timeout 600 python bench.py --config gan --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_r2s2_gan.json 2> gpurun_out/bench_r2s2_gan.err; echo "bench gan exit $?"; head -c 300 gpurun_out/bench_r2s2_gan.json; echo
timeout 150 python tools/timeline_step.py r2s2 > gpurun_out/timeline_r2s2.txt 2>&1; echo "timeline exit $?"
NSTEPS=4 NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu launches exit $?"
