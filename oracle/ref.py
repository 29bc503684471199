"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/_ref/libref.so: the reference's OWN Graph.cs / Model.cs /
Recommender.cs, respelt for a C++ compiler by oracle/cs2cpp.py and built by `make -C oracle ref` where /root/reference is
present (this container).  The library travels to the GPU box prebuilt; the reference's sources do not.

`ReferenceGraph` has the interface of `oracle.OracleGraph` (build / nnz / csr / run / recommend), so a test can hold the
hand-written restatement to the reference itself, call for call.  Importable from tests/, __graft_entry__ and bench.py's CPU
legs only; the product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
HASHES_PATH = os.path.join(_HERE, "_ref", "sources.sha256")
REFERENCE_ROOT = "/root/reference"
_LIB = None

REF_OK, REF_E_INVALID, REF_E_BADSEED, REF_E_ALREADY_BUILT, REF_E_BADINDEX, REF_E_NOT_BUILT = 0, -1, -2, -3, -4, -5


def available(build: bool = True) -> bool:
    """True when libref.so exists (building it first when the reference's sources are on this machine)."""
    if build and os.path.isdir(os.path.join(REFERENCE_ROOT, "Recommenders", "RWRBased")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return os.path.exists(LIB_PATH)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not available():
            raise FileNotFoundError(LIB_PATH + ": neither the reference's sources nor a prebuilt library are here")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.ref_graph_create.restype = vp
        L.ref_graph_create.argtypes = [i32, vp, vp, i64, vp, vp, vp, vp, vp]
        L.ref_graph_destroy.argtypes = [vp]
        L.ref_graph_build.argtypes = [vp]
        L.ref_graph_nnz.restype = i64
        L.ref_graph_nnz.argtypes = [vp]
        L.ref_graph_get_csr.argtypes = [vp, vp, vp, vp, vp]
        L.ref_model_run.argtypes = [vp, i32, f64, i32, i32, f64, i64, vp, vp]
        L.ref_recommend.restype = i64
        L.ref_recommend.argtypes = [vp, i32, C.c_float, i32, i32, i32, vp, vp, i64]
        cp = C.c_char_p
        L.ref_db_reset.argtypes = [cp]
        L.ref_db_add.argtypes = [cp, cp, vp, vp, i64]
        L.ref_loader_validation.argtypes = [cp, i32, vp, vp, vp]
        L.ref_loader_run.restype = vp
        L.ref_loader_run.argtypes = [cp, i32, i32, i32]
        L.ref_loader_destroy.argtypes = [vp]
        L.ref_loader_sizes.argtypes = [vp, vp, vp, vp]
        L.ref_loader_copy.argtypes = [vp] + [vp] * 8
        L.ref_experiment_run.argtypes = [cp, i32, i32, i32, vp, vp, vp, vp]
        _LIB = L
    return _LIB


def source_hashes() -> dict:
    """{reference file: sha256} of the sources libref.so was made from (written next to it by `make -C oracle ref`)."""
    out = {}
    with open(HASHES_PATH) as f:
        for line in f:
            h, rel = line.split()
            out[rel] = h
    return out


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class ReferenceGraph:
    """`new Graph(nodes, edges)` + Model + Recommender of the reference, over the flattened SoA input of the C ABI."""

    def __init__(self, node_id, node_type, src, dst, etype, w, has_entry=None):
        self.node_id = np.ascontiguousarray(node_id, np.int64)
        node_type = np.ascontiguousarray(node_type, np.int32)
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        etype = np.ascontiguousarray(etype, np.int32)
        w = np.ascontiguousarray(w, np.float64)
        he = None if has_entry is None else np.ascontiguousarray(has_entry, np.int8)
        self.n = int(self.node_id.shape[0])
        if len(src) and (src.min() < 0 or src.max() >= self.n):
            raise ValueError("source index out of range")
        self._h = lib().ref_graph_create(self.n, _p(self.node_id), _p(node_type), int(src.shape[0]), _p(src), _p(dst),
                                         _p(etype), _p(w), _p(he))
        if not self._h:
            raise ValueError("ref_graph_create failed")

    def close(self):
        if getattr(self, "_h", None):
            lib().ref_graph_destroy(self._h)
            self._h = None

    __del__ = close

    def build(self) -> int:
        return lib().ref_graph_build(self._h)

    def nnz(self) -> int:
        return lib().ref_graph_nnz(self._h)

    def csr(self, with_types: bool = False):
        nnz = self.nnz()
        if nnz < 0:
            raise RuntimeError("graph not built")
        rp = np.empty(self.n + 1, np.int64)
        col = np.empty(nnz, np.int32)
        val = np.empty(nnz, np.float64)
        typ = np.empty(nnz, np.int32) if with_types else None
        assert lib().ref_graph_get_csr(self._h, _p(rp), _p(col), _p(val), _p(typ)) == 0
        return (rp, col, val, typ) if with_types else (rp, col, val)

    def run(self, seed: int, damping: float, n_iter: int | None = None, threshold: float | None = None,
            default_threshold: bool = False, literal: bool = True, max_iter: int = 0, own_loop: bool = False):
        """-> (rank[N], number of deliverRanks calls, or -1 when the reference's own run() / run(double) looped).
        `literal` is accepted for interface parity with OracleGraph: the reference only has its literal form."""
        rank = np.empty(self.n, np.float64)
        iters = C.c_int64()
        if n_iter is not None:
            mode, ni, thr = 0, int(n_iter), 0.0
        elif default_threshold:
            thr = (1.0 / 1.7976931348623157e308) * self.n                     # Model.cs:53, for the counted loop
            mode, ni = (2, 0) if own_loop else (1, 0)
        else:
            mode, ni, thr = (3 if own_loop else 1), 0, float(threshold)
        rc = lib().ref_model_run(self._h, int(seed), float(damping), mode, ni, thr, int(max_iter), _p(rank), C.byref(iters))
        if rc == REF_E_BADINDEX:
            raise IndexError("link target outside the graph (IndexOutOfRangeException, Model.cs:87)")
        if rc < 0:
            raise RuntimeError(f"ref_model_run rc={rc}")
        return rank, iters.value

    def recommend(self, seed: int, damping_float: float, n_iter: int, top_n: int | None = None, literal: bool = True):
        """Recommendation(idx, float c, nIter[, topN]) -> (ids, scores); raises KeyError like KeyNotFoundException."""
        cap = self.n
        ids = np.empty(cap, np.int64)
        sc = np.empty(cap, np.float64)
        cnt = lib().ref_recommend(self._h, int(seed), float(damping_float), int(n_iter), 0 if top_n is None else 1,
                                  0 if top_n is None else int(top_n), _p(ids), _p(sc), cap)
        if cnt == REF_E_BADSEED:
            raise KeyError(seed)
        if cnt < 0:
            raise RuntimeError(f"ref_recommend rc={cnt}")
        return ids[:cnt].copy(), sc[:cnt].copy()


# ---------------------------------------------------------------------------------------------- the callers
TABLES = {"follow": ("source", "target"), "tweet": ("id", "author"), "retweet": ("user", "tweet"), "quote": ("user", "tweet"),
          "favorite": ("user", "tweet"), "mention": ("source", "target")}


class ReferenceDb:
    """The tables of one ego network (schema of SQLiteAdapter.cs:27-125), handed to the reference's DataLoader in rowid order.
    `path` plays the role of the *.sqlite path: its file name is the ego user's id (DataLoader.cs:32)."""

    def __init__(self, path: str, tables: dict):
        self.path = path.encode()
        assert lib().ref_db_reset(self.path) == 0
        for name, rows in tables.items():
            assert name in TABLES, name
            a = np.ascontiguousarray([r[0] for r in rows], np.int64)
            b = np.ascontiguousarray([r[1] for r in rows], np.int64)
            assert lib().ref_db_add(self.path, name.encode(), _p(a), _p(b), len(a)) == 0

    @classmethod
    def from_sqlite(cls, db_path: str) -> "ReferenceDb":
        import sqlite3
        conn = sqlite3.connect(db_path)
        tables = {name: conn.execute(f"SELECT {a}, {b} FROM {name} ORDER BY rowid").fetchall() for name, (a, b) in TABLES.items()}
        conn.close()
        return cls(db_path, tables)

    def validation(self, n_folds: int):
        """DataLoader.checkEgoNetworkValidation -> (valid, cntLikes, cntFriends)."""
        v, l, f = C.c_int32(), C.c_int32(), C.c_int32()
        assert lib().ref_loader_validation(self.path, int(n_folds), C.byref(v), C.byref(l), C.byref(f)) == 0
        return bool(v.value), l.value, f.value

    def load(self, n_folds: int, methodology: int, fold: int) -> dict:
        """`new DataLoader(path, nFolds).graphConfiguration(methodology, fold)` -> allNodes / allLinks flattened + testSet."""
        h = lib().ref_loader_run(self.path, int(n_folds), int(methodology), int(fold))
        if not h:
            raise RuntimeError("the reference's DataLoader threw")
        try:
            n, e, t = C.c_int32(), C.c_int64(), C.c_int64()
            lib().ref_loader_sizes(h, C.byref(n), C.byref(e), C.byref(t))
            out = dict(node_id=np.empty(n.value, np.int64), node_type=np.empty(n.value, np.int32), has_entry=np.empty(n.value, np.int8),
                       src=np.empty(e.value, np.int32), dst=np.empty(e.value, np.int32), etype=np.empty(e.value, np.int32),
                       w=np.empty(e.value, np.float64), test_ids=np.empty(t.value, np.int64))
            lib().ref_loader_copy(h, *[_p(out[k]) for k in ("node_id", "node_type", "has_entry", "src", "dst", "etype", "w", "test_ids")])
            return out
        finally:
            lib().ref_loader_destroy(h)

    def experiment(self, n_folds: int, n_iter: int, methodology: int):
        """The k-fold loop of Experiment.runKFoldCrossValidation -> dict(valid, hit, avg_precision_sum, cnt_likes)."""
        hit, ap, likes, valid = C.c_double(), C.c_double(), C.c_int32(), C.c_int32()
        rc = lib().ref_experiment_run(self.path, int(n_folds), int(n_iter), int(methodology), C.byref(hit), C.byref(ap),
                                      C.byref(likes), C.byref(valid))
        if rc == REF_E_BADSEED:
            raise KeyError("the ego user has no links (KeyNotFoundException, Recommender.cs:21)")
        if rc < 0:
            raise RuntimeError(f"ref_experiment_run rc={rc}")
        return dict(valid=bool(valid.value), hit=hit.value, avg_precision_sum=ap.value, cnt_likes=likes.value)
