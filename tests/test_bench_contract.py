"""CPU tests of bench.py's contract: the reference arm end to end on a scaled-down graph (it is CPU-only by design), the
byte formulas of SURVEY 8(d) behind `roofline` / `batched.roofline`, and that both arms print the same `config` object."""
import json
import os
import subprocess
import sys

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.004", "--steps", "2",
                        "--warmup", "1", "--gpus", "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                        # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["metric"].startswith("RWR GTEPS") and d["unit"] == "GTEPS" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["unit"] == "GTEPS" and 1 <= cb["cores"] <= 10 and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same `config` object as the GPU arm prints for this graph (workload keys only, no model keys)
    cfg = d["config"]
    assert cfg == bench.c2_config(cfg["n_nodes"], cfg["nnz"], 0.004, 1) and cfg["workload"].startswith("C2:")
    assert set(cfg) == {"workload", "n_nodes", "nnz", "scale", "iterations", "top_k", "parallelism", "l2"}
    # the reference's own Model.cs beside the port, where its compiled sources are here
    ri = cb["reference_itself"]
    if ri is not None:
        assert ri["kind"] == "reference" and "error" not in ri and ri["gteps"] > 0


def test_other_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "0.004", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_byte_formulas_of_survey_8d():
    n, nnz = 11_000_000, 205_156_156
    formula, actual = bench.algorithmic_bytes(n, nnz, 8, layout_index=True)
    assert formula == nnz * 12 + 4 * (n + 1) + 16 * n == 2_681_873_876        # the figure VERDICT r1 recomputed
    assert actual == nnz * 4 + 4 * (n + 1) + 16 * n                             # what the index-only layout moves
    assert bench.algorithmic_bytes(n, nnz, 4, layout_index=False) == (nnz * 8 + 4 * (n + 1) + 8 * n,) * 2
    # batched: E (4 + vb) + 4 (N + 1) + 2 N B vb per pass of B seeds; 128 seeds = 16 passes x 20 iterations
    peak = bench.measured_peak_gbs()[0]
    per_pass = nnz * 12 + 4 * (n + 1) + 2 * n * 8 * 8
    assert abs(bench.batched_frac(n, nnz, 8, 8, 128, 1.0) - per_pass * 16 * 20 / 1e9 / peak) < 1e-12
    assert bench.batched_frac(n, nnz, 8, 8, 0, 1.0) == 0.0
    spec = bench.scaled_spec(0.5)
    assert spec["n_users"] == 500_000 and spec["n_like"] == 38_000_000 and bench.scaled_spec(1.0) == bench.C2_SPEC
