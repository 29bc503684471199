"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle.so (the C++ CPU restatement).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
The product package (recommendersystems_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORC_OK, ORC_E_INVALID, ORC_E_BADSEED, ORC_E_ALREADY_BUILT, ORC_E_BADINDEX, ORC_E_NOT_BUILT = 0, -1, -2, -3, -4, -5


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rwr_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.orc_graph_create.restype = vp
        L.orc_graph_create.argtypes = [i32, vp, vp, i64, vp, vp, vp, vp]
        L.orc_graph_destroy.argtypes = [vp]
        L.orc_graph_build.argtypes = [vp]
        L.orc_graph_nnz.restype = i64
        L.orc_graph_nnz.argtypes = [vp]
        L.orc_graph_get_csr.argtypes = [vp, vp, vp, vp]
        L.orc_model_run.argtypes = [vp, i32, f64, i32, i32, f64, i32, i64, vp, vp]
        L.orc_recommend.restype = i64
        L.orc_recommend.argtypes = [vp, i32, f64, i32, i32, i32, i32, vp, vp, i64]
        L.orc_rank_scores.restype = i64
        L.orc_rank_scores.argtypes = [vp, i32, vp, vp, vp, i64]
        L.orc_recommend_many.argtypes = [vp, vp, i32, f64, i32, i32, i32, vp, vp, vp]
        L.orc_run_many.argtypes = [vp, vp, i32, f64, i32, i32, vp]
        L.orc_evaluate.argtypes = [vp, i64, vp, i64, vp, vp]
        L.orc_synth_create.restype = vp
        L.orc_synth_create.argtypes = [vp]
        L.orc_synth_sizes.argtypes = [vp, vp, vp]
        L.orc_synth_copy.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.orc_synth_destroy.argtypes = [vp]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def widen_float(x: float) -> float:
    return float(np.float32(x))


class SynthSpec(C.Structure):
    """Mirror of rwr_synth_spec (include/rwr_b200.h)."""
    _fields_ = [("seed", C.c_uint64), ("n_users", C.c_int32), ("n_items", C.c_int32), ("n_third", C.c_int32),
                ("authorship_per_mille", C.c_int32), ("n_like", C.c_int64), ("n_friend", C.c_int64),
                ("n_follow", C.c_int64), ("n_mention", C.c_int64), ("undefined_per_mille", C.c_int32),
                ("scramble", C.c_int32), ("p1_byte", C.c_int32), ("reserved", C.c_int32)]


def synth_generate(spec_fields: dict):
    """CPU synthetic generator -> dict of numpy arrays (node_id, node_type, src, dst, etype, w)."""
    L = lib()
    spec = SynthSpec(**spec_fields)
    h = L.orc_synth_create(C.byref(spec))
    if not h:
        raise ValueError("orc_synth_create: invalid spec")
    try:
        n, e = C.c_int32(), C.c_int64()
        L.orc_synth_sizes(h, C.byref(n), C.byref(e))
        out = dict(node_id=np.empty(n.value, np.int64), node_type=np.empty(n.value, np.int32),
                   src=np.empty(e.value, np.int32), dst=np.empty(e.value, np.int32),
                   etype=np.empty(e.value, np.int32), w=np.empty(e.value, np.float64))
        L.orc_synth_copy(h, _p(out["node_id"]), _p(out["node_type"]), _p(out["src"]), _p(out["dst"]),
                         _p(out["etype"]), _p(out["w"]))
        return out
    finally:
        L.orc_synth_destroy(h)


class OracleGraph:
    """CPU oracle of Graph + Model + Recommender over the flattened SoA input."""

    def __init__(self, node_id, node_type, src, dst, etype, w):
        self.node_id = np.ascontiguousarray(node_id, np.int64)
        self.node_type = np.ascontiguousarray(node_type, np.int32)
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        etype = np.ascontiguousarray(etype, np.int32)
        w = np.ascontiguousarray(w, np.float64)
        self.n = int(self.node_id.shape[0])
        self._h = lib().orc_graph_create(self.n, _p(self.node_id), _p(self.node_type), int(src.shape[0]),
                                         _p(src), _p(dst), _p(etype), _p(w))
        if not self._h:
            raise ValueError("orc_graph_create failed (source index out of range?)")

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_graph_destroy(self._h)
            self._h = None

    __del__ = close

    def build(self) -> int:
        return lib().orc_graph_build(self._h)

    def nnz(self) -> int:
        return lib().orc_graph_nnz(self._h)

    def csr(self):
        nnz = self.nnz()
        if nnz < 0:
            raise RuntimeError("graph not built")
        rp = np.empty(self.n + 1, np.int64)
        col = np.empty(nnz, np.int32)
        val = np.empty(nnz, np.float64)
        rc = lib().orc_graph_get_csr(self._h, _p(rp), _p(col), _p(val))
        assert rc == 0
        return rp, col, val

    def run(self, seed: int, damping: float, n_iter: int | None = None, threshold: float | None = None,
            default_threshold: bool = False, literal: bool = False, max_iter: int = 0):
        """-> (rank[N], number of deliverRanks calls).  `damping` is a double (widen a float yourself)."""
        rank = np.empty(self.n, np.float64)
        iters = C.c_int64()
        if n_iter is not None:
            mode, ni, thr = 0, int(n_iter), 0.0
        elif default_threshold:
            mode, ni, thr = 2, 0, 0.0
        else:
            mode, ni, thr = 1, 0, float(threshold)
        rc = lib().orc_model_run(self._h, int(seed), float(damping), mode, ni, thr, int(literal), int(max_iter),
                                 _p(rank), C.byref(iters))
        if rc < 0:
            raise RuntimeError(f"orc_model_run rc={rc}")
        return rank, iters.value

    def recommend(self, seed: int, damping_float: float, n_iter: int, top_n: int | None = None, literal: bool = False):
        """Recommendation(idx, float c, nIter[, topN]) -> (ids, scores); raises KeyError like KeyNotFoundException."""
        cap = self.n
        ids = np.empty(cap, np.int64)
        sc = np.empty(cap, np.float64)
        cnt = lib().orc_recommend(self._h, int(seed), widen_float(damping_float), int(n_iter), int(literal),
                                  0 if top_n is None else 1, 0 if top_n is None else int(top_n), _p(ids), _p(sc), cap)
        if cnt == ORC_E_BADSEED:
            raise KeyError(seed)
        if cnt < 0:
            raise RuntimeError(f"orc_recommend rc={cnt}")
        return ids[:cnt].copy(), sc[:cnt].copy()

    def rank_scores(self, seed: int, rank):
        rank = np.ascontiguousarray(rank, np.float64)
        ids = np.empty(self.n, np.int64)
        sc = np.empty(self.n, np.float64)
        cnt = lib().orc_rank_scores(self._h, int(seed), _p(rank), _p(ids), _p(sc), self.n)
        if cnt == ORC_E_BADSEED:
            raise KeyError(seed)
        return ids[:cnt].copy(), sc[:cnt].copy()

    def recommend_many(self, seeds, damping_float: float, n_iter: int, top_n: int, n_threads: int):
        seeds = np.ascontiguousarray(seeds, np.int32)
        ids = np.zeros((len(seeds), top_n), np.int64)
        sc = np.zeros((len(seeds), top_n), np.float64)
        cnt = np.zeros(len(seeds), np.int32)
        rc = lib().orc_recommend_many(self._h, _p(seeds), len(seeds), widen_float(damping_float), int(n_iter),
                                      int(top_n), int(n_threads), _p(ids), _p(sc), _p(cnt))
        if rc == ORC_E_BADSEED:
            raise KeyError("seed without links")
        if rc < 0:
            raise RuntimeError(f"orc_recommend_many rc={rc}")
        return ids, sc, cnt


def _run_many(self, seeds, damping_float: float, n_iter: int, n_threads: int):
    """Model.run(n_iter) for each seed, n_threads seeds in flight -> rank[seed] per seed."""
    seeds = np.ascontiguousarray(seeds, np.int32)
    chk = np.zeros(len(seeds), np.float64)
    rc = lib().orc_run_many(self._h, _p(seeds), len(seeds), widen_float(damping_float), int(n_iter), int(n_threads), _p(chk))
    if rc < 0:
        raise RuntimeError(f"orc_run_many rc={rc}")
    return chk


OracleGraph.run_many = _run_many


def evaluate(ids, test_ids):
    """Experiment.cs:121-128 -> (nHits, average precision)."""
    ids = np.ascontiguousarray(ids, np.int64)
    t = np.sort(np.ascontiguousarray(test_ids, np.int64))
    hits, ap = C.c_int32(), C.c_double()
    lib().orc_evaluate(_p(ids), len(ids), _p(t), len(t), C.byref(hits), C.byref(ap))
    return hits.value, ap.value
