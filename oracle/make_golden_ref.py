#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/ref_mid.json with the REFERENCE ITSELF.

    python oracle/make_golden_ref.py          (needs /root/reference: builds oracle/_ref/libref.so first)

The numbers come from the reference's own Graph.cs / Model.cs / Recommender.cs, compiled for this image through
oracle/cs2cpp.py (no C# toolchain exists here) -- not from any restatement.  The input is a mid-sized ego network from this
repo's deterministic generator (820 nodes, ~8 k links: fractional MENTION weights, UNDEFINED links, dangling rows, third-party
users).  Stored: the transition matrix (sizes + SHA-256 of row_ptr / col / val, it is bit-exact on every implementation), the
rank vectors of two seed users after 2 and 20 iterations (C99 hex doubles), iteration counts and rank hashes under four
thresholds, the full ranking (ids) and the top-10 (ids + scores) of `Recommendation(seed, 0.15f, 20)`.
tests/test_reference_pin.py requires the CPU oracle to reproduce the file bit for bit; tests/test_gpu_parity.py holds the CUDA
path to it (CSR hashes equal, scores <= 1e-12 relative with exact zeros, identical iteration counts and lists).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_mid.json")

SPEC = dict(seed=20260218, n_users=100, n_items=700, n_third=20, authorship_per_mille=800, n_like=2600, n_friend=700,
            n_follow=60, n_mention=60, undefined_per_mille=80, scramble=1, p1_byte=61, reserved=0)
N_ITERS = (2, 20)
THRESHOLDS = ("1e-3", "1e-6", "1e-9", "1e-12")
DAMPING_FLOAT = 0.15


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def seeds_of(links, n_users: int):
    """Two seed users with links: the first one, and the one with the most."""
    deg = np.bincount(links["src"], minlength=len(links["node_id"]))[:n_users]
    return [int(np.flatnonzero(deg > 0)[0]), int(np.argmax(deg))]


def generate(graph_class) -> dict:
    """The fixture's content as computed by `graph_class` (ref.ReferenceGraph to write it, oracle.OracleGraph to check it)."""
    links = O.synth_generate(SPEC)
    g = graph_class(links["node_id"], links["node_type"], links["src"], links["dst"], links["etype"], links["w"])
    assert g.build() == 0
    rp, col, val = g.csr()
    c = O.widen_float(DAMPING_FLOAT)
    out = {"spec": SPEC, "damping_float": DAMPING_FLOAT, "damping_double": float(c).hex(), "n_nodes": int(g.n), "n_links": int(len(links["src"])),
           "csr": {"nnz": int(len(col)), "dangling_rows": int((np.diff(rp) == 0).sum()), "row_ptr_sha256": sha(rp.astype(np.int64)),
                   "col_sha256": sha(col.astype(np.int32)), "val_sha256": sha(val.astype(np.float64))},
           "seeds": []}
    for seed in seeds_of(links, SPEC["n_users"]):
        e = {"seed": seed, "ranks": {}, "thresholds": {}}
        for n in N_ITERS:
            rank, _ = g.run(seed, c, n_iter=n)
            e["ranks"][str(n)] = [float(x).hex() for x in rank]
        for thr in THRESHOLDS:
            rank, iters = g.run(seed, c, threshold=float(thr))
            e["thresholds"][thr] = {"iters": int(iters), "rank_sha256": sha(rank)}
        ids, sc = g.recommend(seed, DAMPING_FLOAT, 20)
        e["recommendation"] = {"n_iter": 20, "ids": [int(x) for x in ids], "top10_scores": [float(x).hex() for x in sc[:10]],
                               "scores_sha256": sha(sc)}
        out["seeds"].append(e)
    g.close()
    return out


def main():
    import ref as RF
    if not RF.available():
        raise SystemExit("oracle/_ref/libref.so cannot be built here: /root/reference is missing")
    doc = generate(RF.ReferenceGraph)
    doc["made_by"] = "oracle/make_golden_ref.py: the reference's own sources through oracle/cs2cpp.py (oracle/_ref/libref.so)"
    doc["made_from"] = {rel: h for rel, h in RF.source_hashes().items() if "RWRBased" in rel}      # the three files of the hot path
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=0)
        f.write("\n")
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
