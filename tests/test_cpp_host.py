"""The C++ host side of the drop-in (recommendersystems_b200/cpp/RWRBased.hpp) through ONE caller compiled twice
(tests/cpp/experiment_caller.cpp): against RWRBased.hpp over librwr_b200.so (caller_b200) and against the reference's own sources as
oracle/cs2cpp.py respells them (oracle/_ref/caller_reference, built by `make -C oracle ref`).  Same caller text, same graphs: the printed recommendation lists
must be identical and the scores agree to 1e-12 (exact zeros preserved).  Without a GPU caller_b200 must fail loudly."""
import os
import subprocess

import numpy as np
import pytest

from conftest import C1_SPEC, ROOT, bits, load_golden, unhex
import oracle as O

CPP = os.path.join(ROOT, "tests", "cpp")
B200 = os.path.join(CPP, "caller_b200")
REFERENCE = os.path.join(ROOT, "oracle", "_ref", "caller_reference")


def build():
    import ref as RF
    subprocess.check_call(["make", "-C", CPP, "-s"])
    RF.available()                  # `make -C oracle ref` where the reference's sources are: builds caller_reference, too


def write_graph(path, inp):
    w = np.asarray(inp["w"], np.float64)
    with open(path, "w") as f:
        f.write(f"{len(inp['node_id'])}\n")
        for i, t in zip(inp["node_id"], inp["node_type"]):
            f.write(f"{int(i)} {int(t)}\n")
        f.write(f"{len(inp['src'])}\n")
        for s, d, t, x in zip(inp["src"], inp["dst"], inp["etype"], w.tolist()):
            f.write(f"{int(s)} {int(d)} {int(t)} {float(x).hex()}\n")
    return path


def run(binary, graph_file, seed, n_iter, top_n=None):
    cmd = [binary, graph_file, str(seed), str(n_iter)] + ([] if top_n is None else [str(top_n)])
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    if p.returncode != 0:
        return p.returncode, p.stderr, None, None
    lines = p.stdout.split("\n")
    count = int(lines[0])
    rec = [(int(a), float.fromhex(b)) for a, b in (l.split() for l in lines[1:1 + count])]
    rank = np.array([float.fromhex(l.split()[2]) for l in lines[1 + count:] if l.startswith("rank ")])
    return 0, p.stderr, rec, rank


needs_reference = pytest.mark.skipif(not (os.path.exists(REFERENCE) or os.path.isdir("/root/reference")),
                                     reason="oracle/_ref/caller_reference: no reference sources and no prebuilt binary here")


@needs_reference
def test_the_caller_on_the_reference_reproduces_the_golden_vectors(tmp_path):
    build()
    for name in ("kat_8c", "small_b"):
        g = load_golden(name)
        path = write_graph(str(tmp_path / f"{name}.txt"), g["input"])
        for e in g["seeds"]:
            if e["recommendation"] == "KeyNotFoundException":
                rc, err, _, _ = run(REFERENCE, path, e["seed"], 10)
                assert rc == 3 and "KeyNotFoundException" in err
                continue
            want = e["recommendation"]
            rc, _, rec, rank = run(REFERENCE, path, e["seed"], want["n_iter"])
            assert rc == 0 and [p[0] for p in rec] == want["ids"]
            assert np.array_equal(bits([p[1] for p in rec]), bits(unhex(want["scores"])))
            assert np.array_equal(bits(rank), bits(unhex(e["ranks"][str(want["n_iter"])])))
            for k, top in e["top"].items():
                rc, _, rec, _ = run(REFERENCE, path, e["seed"], want["n_iter"], int(k))
                assert rc == 0 and [p[0] for p in rec] == top["ids"]


def test_the_caller_on_the_drop_in_fails_loudly_without_a_gpu(tmp_path):
    """No CPU fallback: on a machine without a CUDA device the drop-in's caller exits non-zero with the CUDA error."""
    import recommendersystems_b200 as rs
    if rs._native.lib().rwr_device_count() > 0:
        pytest.skip("a CUDA device is present")
    build()
    g = load_golden("kat_8c")
    rc, err, _, _ = run(B200, write_graph(str(tmp_path / "kat.txt"), g["input"]), 0, 10)
    assert rc == 1 and ("CUDA" in err or "device" in err), err


@pytest.mark.gpu
def test_one_caller_two_implementations(tmp_path):
    """caller_b200 against caller_reference (or, without the prebuilt reference binary, against the oracle) on small_b and on
    the C1-shaped ego network: full ranking, top-10, and the Model's rank vector."""
    build()
    # (every run of caller_b200 is a process with its own CUDA context: three runs in all)
    cases = [("small_b", load_golden("small_b")["input"], [e["seed"] for e in load_golden("small_b")["seeds"]
                                                            if e["recommendation"] != "KeyNotFoundException"][:1], 10, (None,))]
    s = O.synth_generate(C1_SPEC)
    cases.append(("c1", s, [int(np.flatnonzero(np.bincount(s["src"], minlength=len(s["node_id"]))[:1000] > 0)[0])], 3, (10,)))
    for name, inp, seeds, n_iter, tops in cases:
        path = write_graph(str(tmp_path / f"{name}.txt"), inp)
        og = O.OracleGraph(inp["node_id"], inp["node_type"], inp["src"], inp["dst"], inp["etype"], inp["w"])
        assert og.build() == 0
        for seed in seeds:
            for top_n in tops:
                rc, err, rec, rank = run(B200, path, seed, n_iter, top_n)
                assert rc == 0, err
                if os.path.exists(REFERENCE):
                    rc2, err2, want, want_rank = run(REFERENCE, path, seed, n_iter, top_n)
                    assert rc2 == 0, err2
                else:
                    ids, sc = og.recommend(seed, 0.15, n_iter, top_n=top_n)
                    want, want_rank = list(zip(ids.tolist(), sc.tolist())), og.run(seed, O.widen_float(0.15), n_iter=n_iter)[0]
                assert [p[0] for p in rec] == [p[0] for p in want], (name, seed, top_n)
                a, b = np.array([p[1] for p in rec]), np.array([p[1] for p in want])
                nz = b != 0
                assert np.all(a[~nz] == 0) and (not nz.any() or (np.abs(a[nz] - b[nz]) / b[nz]).max() <= 1e-12)
                nz = want_rank != 0
                assert np.all(rank[~nz] == 0) and (np.abs(rank[nz] - want_rank[nz]) / want_rank[nz]).max() <= 1e-12
        og.close()
    # the reference's exception for a seed without an `edges` entry, on the drop-in
    g = load_golden("kat_8c")
    missing = [e["seed"] for e in g["seeds"] if e["recommendation"] == "KeyNotFoundException"]
    if missing:
        rc, err, _, _ = run(B200, write_graph(str(tmp_path / "kat.txt"), g["input"]), missing[0], 10)
        assert rc == 3 and "KeyNotFoundException" in err
