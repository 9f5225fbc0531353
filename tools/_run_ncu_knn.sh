for spec in "knn_prune_kernel:0:knn_prune_s2" "knn_small_kernel:0:knn_small_s2"; do
  IFS=: read -r k skip name <<< "$spec"
  NSTEPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s "$skip" -c 1 -f -o "gpurun_out/full_$name" \
      python tools/ncu_step.py > "gpurun_out/ncu_full_$name.log" 2>&1; echo "ncu full $name exit $?"
done
ls -la gpurun_out/*.ncu-rep | tail -3
