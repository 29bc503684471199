#!/bin/bash
# 2 GPUs: partitioned tests (with the late-slice mode), build trace of C2 x 2 at P=2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_partitioned.py -m gpu -x -q 2>&1 | tail -15
RWR_BUILD_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 \
      profiles/microbench/part_probe.py 2.0 2>&1 | grep -E "rwr build r0|ms/iteration|setup wall|rror" | head -40
