"""Experiment-style evaluation (TweetRecommender/Experiment.cs:69-138, DataLoader.cs:122-140) on flattened link arrays:
hold out the newest likes of the test users, recommend on the remaining graph, score the lists.

Host-side data preparation only (numpy); every recommendation runs through the C ABI (`Recommender.RecommendationBatch`).
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np

from .rwr import EdgeType, NodeType


def hold_out_likes(links: Dict[str, np.ndarray], test_users: Sequence[int], fraction: float = 0.1):
    """For every test user the upper `fraction` (by tweet id, i.e. the newest -- DataLoader.cs:126-137 takes the last fold
    of the id-ordered likes) of the user's LIKE links leaves the graph in both directions and becomes the user's test set.

    -> (links without the held-out pairs, {user: int64 array of held-out tweet ids})."""
    src, dst, et = links["src"], links["dst"], links["etype"]
    node_id, node_type = links["node_id"], links["node_type"]
    n = len(node_id)
    drop = np.zeros(len(src), bool)
    test: Dict[int, np.ndarray] = {}
    like = et == EdgeType.LIKE
    order = np.argsort(src, kind="stable")
    starts = np.searchsorted(src[order], np.arange(n + 1))
    pair_key = src.astype(np.int64) * n + dst
    held_keys = []
    for u in test_users:
        rows = order[starts[u]:starts[u + 1]]
        rows = rows[like[rows] & (node_type[dst[rows]] == NodeType.ITEM)]
        if len(rows) == 0:
            test[int(u)] = np.zeros(0, np.int64)
            continue
        ids = node_id[dst[rows]]
        k = int(len(rows) * fraction)
        if k == 0:
            test[int(u)] = np.zeros(0, np.int64)
            continue
        newest = rows[np.argsort(ids, kind="stable")[-k:]]
        test[int(u)] = np.sort(node_id[dst[newest]])
        drop[newest] = True
        held_keys.append(dst[newest].astype(np.int64) * n + u)          # the reverse links tweet -> user
    if held_keys:
        hk = np.unique(np.concatenate(held_keys))
        rev = like & np.isin(pair_key, hk)
        drop |= rev
    keep = ~drop
    out = dict(links)
    for k_ in ("src", "dst", "etype", "w"):
        out[k_] = links[k_][keep]
    return out, test


def recall_at_k(ids: np.ndarray, counts: np.ndarray, users: Sequence[int], test: Dict[int, np.ndarray]) -> Tuple[float, int, int]:
    """-> (mean recall@k over the users with a non-empty test set, total hits, users counted)."""
    total, hits_all, counted = 0.0, 0, 0
    for i, u in enumerate(users):
        t = test.get(int(u))
        if t is None or len(t) == 0:
            continue
        hits = int(np.isin(ids[i, :counts[i]], t).sum())
        total += hits / len(t)
        hits_all += hits
        counted += 1
    return (total / counted if counted else 0.0), hits_all, counted
