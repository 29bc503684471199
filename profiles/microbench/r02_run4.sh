#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_slice2.log; : > $L
for cfg in "1.0 2" "2.0 4" "c4 8"; do
  for mode in "" "RWR_DIST_LEGACY=1"; do
    env $mode timeout 300 python profiles/microbench/slice_probe.py $cfg 0 both >> $L 2>&1
  done
done
cat $L
timeout 600 python -m pytest tests/test_gpu_experiment.py tests/test_gpu_parity.py -m gpu -x -q -k "experiment or evaluate or hold_out or k_fold" 2>&1 | tail -5
