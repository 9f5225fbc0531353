#!/bin/bash
# One gpurun call's worth of evidence (development tool): GPU parity tests, the bench line, a torch.profiler breakdown,
# per-op timings next to the reference extensions, and the ncu launch list of the bench command.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tests] [bench] [profile] [opbench] [ncu]'
set -u
mkdir -p gpurun_out
want() { [[ " $ARGS " == *" $1 "* ]]; }
ARGS="${*:-tests bench profile opbench ncu refstep}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
if want tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
if want bench; then
  timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
  tail -c 3000 gpurun_out/bench.json
fi
if want profile; then
  timeout 300 python tools/profile_step.py > gpurun_out/profile.txt 2>&1; echo "profile exit $?"
  GROUP=shape timeout 300 python tools/profile_step.py > gpurun_out/profile_shape.txt 2>&1
fi
if want opbench; then
  timeout 600 python tests/perf/opbench.py > gpurun_out/opbench.txt 2>&1; echo "opbench exit $?"
fi
if want ncu; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu exit $?"
fi
if want refstep; then
  timeout 400 python tests/perf/refstep.py > gpurun_out/refstep.txt 2>&1; echo "refstep exit $?"; tail -2 gpurun_out/refstep.txt
fi
