"""Experiment-style evaluation -- the caller of the hot path (TweetRecommender/Experiment.cs:31-155, DataLoader.cs:79-140).

Host driver only: the k-fold hold-out (`Graph.hold_out` -> rwr_graph_hold_out), the methodology masks
(rwr_methodology_masks), the graph build, the recommendations and the hit / average-precision arithmetic
(`evaluate_users` -> rwr_evaluate_users) all run on the device through the C ABI.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from .rwr import FP64, EdgeType, Graph, Methodology, NodeType, evaluate_users, methodology_options


def like_count(links: Dict[str, np.ndarray], user: int) -> int:
    """`loader.cntLikesOfEgoUser` (DataLoader.cs:94-109): the tweets the user likes, before any split."""
    sel = (links["src"] == user) & (links["etype"] == EdgeType.LIKE)
    return int((links["node_type"][links["dst"][sel]] == NodeType.ITEM).sum())


def ego_network_is_valid(links: Dict[str, np.ndarray], ego: int, n_folds: int) -> bool:
    """`DataLoader.checkEgoNetworkValidation` (DataLoader.cs:79-92): >= nFolds and >= 50 likes, >= 50 friends."""
    likes = like_count(links, ego)
    friends = int(((links["src"] == ego) & (links["etype"] == EdgeType.FRIENDSHIP)).sum())
    return not (likes < n_folds or likes < 50 or friends < 50)


def result_row(ego_id: int, methodology: int, n_folds: int, n_iter: int, hits: float, cnt_likes: int, sum_ap: float) -> str:
    """One line of result.dat (Experiment.cs:144-152; read back as 7 tab-separated tokens at Program.cs:41):
    ego \\t methodology \\t nFolds \\t nIterations \\t (int)HIT \\t cntLikes \\t AVGPRECISION / nFolds."""
    return "\t".join([str(int(ego_id)), str(int(methodology)), str(int(n_folds)), str(int(n_iter)), str(int(hits)),
                      str(int(cnt_likes)), dotnet_double_to_string(sum_ap / n_folds)])


def dotnet_double_to_string(x: float) -> str:
    """`"\\t" + double` on .NET Framework (Experiment.cs:150): double.ToString() is the "G15" format -- 15 significant digits,
    fixed notation while -5 < exponent < 15, else scientific with an upper-case E and at least two exponent digits."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    return ("%.15g" % x).replace("e", "E")      # C's %g switches notation at the same exponents and pads the exponent alike


def run_k_fold(links: Dict[str, np.ndarray], methodology: int = Methodology.ALL, n_folds: int = 10, n_iter: int = 20,
               ego: int = 0, precision: int = FP64, validate: bool = True, **graph_opts) -> Optional[dict]:
    """`Experiment.runKFoldCrossValidation` for one ego network and one methodology (Experiment.cs:46-155).  `links` is the
    network with every relation loaded (Methodology.ALL); the methodology's feature set is applied as link-type masks.
    Per fold: hold out the ego's fold (device), buildGraph (device), Recommendation(ego, 0.15f, nIterations) and the
    walk over the full ranking (device).  -> dict(row=<result.dat line>, hits, map, cnt_likes, folds=[...]) or None when
    the network fails `checkEgoNetworkValidation` (Experiment.cs:72-74)."""
    if validate and not ego_network_is_valid(links, ego, n_folds):
        return None
    cnt_likes = like_count(links, ego)
    opts = dict(methodology_options(methodology))
    opts.update(graph_opts)
    hits, sum_ap, folds = 0.0, 0.0, []
    for fold in range(n_folds):
        g = Graph.from_arrays(links["node_id"], links["node_type"], links["src"], links["dst"], links["etype"], links["w"], **opts)
        try:
            test = g.hold_out([ego], n_folds, fold)
            g.buildGraph()
            r = evaluate_users(g, None, None, 0.15, n_iter, k=10, precision=precision)
        finally:
            g.close()
        hits += int(r["hits"][0])                      # finalResult[HIT] += nHits           (Experiment.cs:134)
        sum_ap += float(r["avg_precision"][0])         # finalResult[AVGPRECISION] += ...    (Experiment.cs:136)
        folds.append(dict(fold=fold, n_test=len(test[int(ego)]), hits=int(r["hits"][0]), avg_precision=float(r["avg_precision"][0])))
    ego_id = int(links["node_id"][ego])
    return dict(row=result_row(ego_id, methodology, n_folds, n_iter, hits, cnt_likes, sum_ap), hits=int(hits),
                map=sum_ap / n_folds, cnt_likes=cnt_likes, folds=folds)


def evaluate_hold_out(graph: Graph, users: Sequence[int], n_folds: int = 10, fold: int = 9, n_iter: int = 20, k: int = 10,
                      precision: int = FP64) -> dict:
    """BASELINE config 5: one global hold-out for many test users of one graph (the newest tenth of every user's likes with
    the defaults), then top-k / full-ranking metrics for all of them.  `graph` must be created but not built.
    -> dict(recall_at_k, hits_at_k, users_counted, mean_avg_precision, hits, seeds_per_s, ...)."""
    graph.hold_out(users, n_folds, fold)
    graph.buildGraph()
    r = evaluate_users(graph, None, None, 0.15, n_iter, k=k, precision=precision)
    return summarize(r)


def summarize(r: dict) -> dict:
    has = r["n_test"] > 0
    counted = int(has.sum())
    recall = float((r["hits_at_k"][has] / r["n_test"][has]).mean()) if counted else 0.0
    return dict(recall_at_k=recall, hits_at_k=int(r["hits_at_k"].sum()), users_counted=counted,
                mean_avg_precision=float(r["avg_precision"][has].mean()) if counted else 0.0, hits=int(r["hits"].sum()),
                n_test=int(r["n_test"].sum()), total_ms=float(r["info"].total_ms), iterate_ms=float(r["info"].iterate_ms))
