#!/bin/bash
# 8 GPUs: C4 with the overlapped exchange vs the peer-store epilogue, a graph beyond one GPU, then the bench as the driver runs it
mkdir -p gpurun_out
L=gpurun_out/r2_part8.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
echo "== C4 overlapped" >> $L; timeout 240 $TR 29801 profiles/microbench/part_big.py 1 >> $L 2>&1
echo "== C4 peer stores (RWR_DIST_LEGACY=1)" >> $L; RWR_DIST_LEGACY=1 timeout 240 $TR 29802 profiles/microbench/part_big.py 1 >> $L 2>&1
echo "== C4 fp32 overlapped" >> $L; timeout 240 $TR 29803 profiles/microbench/part_big.py 1 fp32 >> $L 2>&1
echo "== 4 x C4 (200 M nodes, 7.3 B links) overlapped" >> $L; timeout 400 $TR 29804 profiles/microbench/part_big.py 4 >> $L 2>&1
grep -E "^==|^rank|Error|error" $L | cut -c1-420
timeout 900 $TR 29805 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n8.json").read().strip().splitlines()[-1])
    print(json.dumps({k: d[k] for k in ("value", "e2e", "row_partitioned", "c5")}, indent=1)[:6000])
    print("batched", d["batched"]["fp64"], d["batched"]["fp32"])
except Exception as e:
    print("no bench line", e)
PY
