#!/bin/bash
# round 2, GPU call 1: suite + slice probes (fake comm, one GPU) + ncu of the C4 slice kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
L=gpurun_out/r2_slice.log; : > $L
for xb in 1 4 8 16; do RWR_X_BLOCKS=$xb timeout 300 python profiles/microbench/slice_probe.py c4 8 0 both >> $L 2>&1; done
for xb in 1 8; do RWR_PART_NO_HUB=1 RWR_X_BLOCKS=$xb timeout 300 python profiles/microbench/slice_probe.py c4 8 0 both >> $L 2>&1; done
timeout 300 python profiles/microbench/slice_probe.py 1.0 2 0 both >> $L 2>&1
RWR_PART_NO_HUB=1 timeout 300 python profiles/microbench/slice_probe.py 1.0 2 0 both >> $L 2>&1
timeout 300 python profiles/microbench/slice_probe.py 2.0 4 0 both >> $L 2>&1
RWR_X_BLOCKS=2 timeout 300 python profiles/microbench/slice_probe.py 2.0 4 0 both >> $L 2>&1
RWR_X_BLOCKS=4 timeout 300 python profiles/microbench/slice_probe.py 2.0 4 0 both >> $L 2>&1
timeout 300 python profiles/microbench/slice_probe.py 1.0 1 0 both >> $L 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_ws --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r02_c4slice_spmv_fp64 python profiles/microbench/slice_probe.py c4 8 0 fp64 3 > gpurun_out/r2_ncu1.log 2>&1
RWR_X_BLOCKS=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_ws --launch-skip 4 --launch-count 1 \
   -o gpurun_out/r02_c4slice_spmv_fp64_xb8 python profiles/microbench/slice_probe.py c4 8 0 fp64 3 > gpurun_out/r2_ncu2.log 2>&1
tail -3 gpurun_out/r2_pytest1.log; cat $L
