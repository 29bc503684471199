#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["frac_iteration"], d["other_precision"]["value"])
print(d["parity"]["ok"], d["c5"]["seeds_per_s"], d["c5"]["parity"]["ok"], d["batched"]["fp64"], d["batched"]["fp32"], d["build"])
r = json.loads(open("gpurun_out/r02_bench_ref.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["config"] == d["config"])
PY
