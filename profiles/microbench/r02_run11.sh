#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
L=gpurun_out/r2_slice3.log; : > $L
RWR_DIST_OVERLAP=1 timeout 200 python profiles/microbench/slice_probe.py 1.0 2 0 both >> $L 2>&1
timeout 200 python profiles/microbench/slice_probe.py 2.0 4 0 both >> $L 2>&1
timeout 200 python profiles/microbench/slice_probe.py c4 8 0 both >> $L 2>&1
cut -c1-600 $L
