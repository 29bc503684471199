#!/bin/bash
# ncu call: launch list of a short bench run + one full capture of k_spmv_ws on C2 (each after its plain run exited 0)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batched or spmm or fp32 or tile" 2>&1 | tail -3
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --batch-seeds 64 --no-c5"
timeout 300 $CMD > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_ncu_l.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_short.json").read().strip().splitlines()[-1])
print("short bench:", d["value"], d["batched"]["fp64"], d["batched"]["fp32"], d["batched"]["roofline"])
PY
P1="python profiles/microbench/prof_one.py 0 0"
timeout 200 $P1 > gpurun_out/r02_prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_ws -s 4 -c 1 -o gpurun_out/r02_spmv_ws_fp64 $P1 > gpurun_out/r02_ncu_f.log 2>&1
echo "full capture rc=$?"; cat gpurun_out/r02_prof_plain.log | tail -2
