#!/bin/bash
# runs every correctness case (and with "perf" the layer-sized timing cases) of tools/_bin/gemm_test, one process each
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
read NC NP < <(tools/_bin/gemm_test count)
LOG=gpurun_out/gemm_test.log
: > $LOG
fail=0
for i in $(seq 0 $((NC-1))); do
  timeout 60 tools/_bin/gemm_test $i >> $LOG 2>&1 || { echo "case $i exit $?" >> $LOG; fail=$((fail+1)); }
done
if [ "$1" = "perf" ]; then
  for i in $(seq 0 $((NP-1))); do
    timeout 120 tools/_bin/gemm_test perf $i >> $LOG 2>&1 || { echo "perf $i exit $?" >> $LOG; fail=$((fail+1)); }
  done
fi
grep -E "^(PASS|FAIL)|exit|ours:|cuBLAS" $LOG
echo "failed: $fail"
