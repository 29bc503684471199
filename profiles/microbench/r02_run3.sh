#!/bin/bash
# round 2, 2-GPU call: partitioned parity in every exchange mode + iteration time of the C2 graph at P=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q > gpurun_out/r2_pytest_part.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_part.log
tail -25 gpurun_out/r2_pytest_part.log
L=gpurun_out/r2_part2.log; : > $L
for mode in "" "RWR_DIST_LEGACY=1" "RWR_DIST_NO_P2P=1"; do
  echo "== mode [$mode]" >> $L
  env $mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 \
      profiles/microbench/part_probe.py 1.0 >> $L 2>&1
done
grep -E "== mode|ms/iteration|Error|error" $L | head -40
