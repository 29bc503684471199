#!/bin/bash
# 2 GPUs: partitioned parity in every mode (partitioned build) + P=2 timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q > gpurun_out/r2_pytest_part2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_part2.log
tail -30 gpurun_out/r2_pytest_part2.log
L=gpurun_out/r2_part2b.log; : > $L
for mode in "" "RWR_DIST_OVERLAP=1" "RWR_TILE_LINKS=2048" "RWR_TILE_LINKS=1024"; do
  echo "== mode [$mode]" >> $L
  env $mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 \
      profiles/microbench/part_probe.py 1.0 >> $L 2>&1
done
grep -E "== mode|ms/iteration|Error|error|setup wall" $L | head -40
