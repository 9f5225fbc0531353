#!/bin/bash
# Round-2 evidence in one gpurun call: GPU parity tests, smoke, the default bench line (with the reference-extension arm and the CPU
# arm), the GAN-config line, per-op timings beside the reference extensions, the ncu launch list of one eager step and full captures of
# the top kernels (summarised into profiles/ by tools/summarize_profiles.py r2).
set -u
mkdir -p gpurun_out
if [[ " ${*:-all} " =~ " all " || " $* " =~ " tests " ]]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_r2.log
  timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_r2.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_r2.log
fi
if [[ " ${*:-all} " =~ " all " || " $* " =~ " bench " ]]; then
  timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench exit $?"; head -c 400 gpurun_out/bench_r2_final.json; echo
  timeout 600 python bench.py --config gan --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_r2_gan.json 2> gpurun_out/bench_r2_gan.err; echo "bench gan exit $?"; head -c 300 gpurun_out/bench_r2_gan.json; echo
  timeout 400 python tests/perf/opbench.py > gpurun_out/opbench_r2.txt 2>&1; echo "opbench exit $?"
  timeout 300 python tools/emd_time.py > gpurun_out/emd_time_r2.txt 2>&1; echo "emd_time exit $?"
fi
if [[ " ${*:-all} " =~ " all " || " $* " =~ " ncu " ]]; then
  NSTEPS=4 NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu launches exit $?"
  # gemm_tf32_kernel launches of a step in order: #12 = encoder conv5 forward (plain, 2048^3 x 32), #13 = decoder conv2 forward (prologue)
  for spec in "mds_cluster_kernel:0:mds_cluster_kernel" "gemm_tf32_kernel:12:gemm_tf32_plain_enc_conv5" "gemm_tf32_kernel:13:gemm_tf32_prologue_dec_conv2" \
              "bn_se_tail_fwd_kernel:4:bn_se_tail_fwd_kernel" "chamfer_bvh_query_kernel:0:chamfer_bvh_query_kernel" "knn_prune_kernel:0:knn_prune_kernel"; do
    IFS=: read -r k skip name <<< "$spec"
    NSTEPS=1 timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$k" -s "$skip" -c 1 -f -o "gpurun_out/full_$name" \
        python tools/ncu_step.py > "gpurun_out/ncu_full_$name.log" 2>&1; echo "ncu full $name exit $?"
  done
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:emd_auction_tree -c 1 -f -o gpurun_out/full_emd_auction_tree_kernel \
      python tools/emd_once.py 16384 50 > gpurun_out/ncu_full_emd_tree.log 2>&1; echo "ncu full emd tree exit $?"
fi
