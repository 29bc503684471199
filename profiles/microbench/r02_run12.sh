#!/bin/bash
# 8 GPUs: C4 with the overlapped exchange (wait trace), then the bench with a reduced C5 leg
mkdir -p gpurun_out
L=gpurun_out/r2_part8b.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
echo "== C4 overlapped" >> $L; RWR_XCHG_TRACE=1 RWR_BUILD_TRACE=1 timeout 150 $TR 29801 profiles/microbench/part_big.py 1 >> $L 2>&1
grep -E "^==|^rank 0|^rank 5|rwr xchg r0|rwr xchg r5|rwr build r0|Error|error" $L | cut -c1-420
timeout 420 $TR 29805 bench.py --gpus 8 --steps 5 --warmup 3 --c5-users 16000 > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err; echo "bench rc=$?"
tail -c 800 gpurun_out/r2_bench_n8b.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n8b.json").read().strip().splitlines()[-1])
    rp = d["row_partitioned"]
    print("value", d["value"], "e2e", d["e2e"]["value"])
    print("row_partitioned", {k: rp[k] for k in ("gteps", "ms_per_iteration", "x_blocks", "hbm", "build", "parity")})
    print("c5", {k: d["c5"][k] for k in ("users", "seeds_per_s", "recall_at_10", "parity")})
    print("batched", d["batched"]["fp64"], d["batched"]["fp32"])
except Exception as e:
    print("no bench line", e)
PY
