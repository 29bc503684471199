"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy + the C++ oracle) of the reference's evaluation loop around the hot
path, used to check rwr_graph_hold_out / rwr_methodology_masks / rwr_evaluate_users.  Never imported by the product.

Follows TweetRecommender/DataLoader.cs:122-140 (splitLikeHistory), :142-219 (Methodology -> features), :221-254 (which
relations a feature set loads), :287-298 (the test fold's likes never become links), :398-436 (mention weights) and
TweetRecommender/Experiment.cs:84-101 (FRIENDSHIP retyped UNDEFINED), :104-109, :121-138, :144-152.
"""
from __future__ import annotations

import numpy as np

import oracle as O

LIKE, FRIENDSHIP, FOLLOW, MENTION, AUTHORSHIP = 1, 2, 3, 4, 5
ITEM = 2
F, T3, A, M = "FRIENDSHIP", "FOLLOWSHIP_ON_THIRDPARTY", "AUTHORSHIP", "MENTIONCOUNT"

# DataLoader.cs:144-214, one entry per `case`, features in the order they are added
FEATURES = {
    0: [], 1: [F], 2: [T3], 3: [A], 4: [F, M], 5: [F, T3], 6: [F, A], 7: [F, M], 8: [F, T3, A, M], 9: [F, T3, A, M],
    10: [F, A, M], 11: [F, T3, M], 12: [F, T3, A], 13: [T3, A], 14: [F, T3, M], 15: [A, M],
}
# Experiment.cs:84-86: methodologies whose FRIENDSHIP links are retyped UNDEFINED after loading
RETYPE_FRIENDSHIP = {4, 9, 14}


def split_like_history(like_ids, n_folds: int, fold: int):
    """DataLoader.cs:122-140 -> (trainSet, testSet) as sorted arrays of tweet ids."""
    ids = np.sort(np.asarray(like_ids, np.int64))             # likesList.Sort()
    unit = len(ids) // n_folds                                # int unitSize = likes.Count / nFolds
    lo = unit * fold
    hi = unit * (fold + 1) if fold < n_folds - 1 else len(ids)
    test = ids[lo:hi]
    train = np.concatenate([ids[:lo], ids[hi:]])
    return train, test


def hold_out(links: dict, users, n_folds: int, fold: int):
    """The graph DataLoader builds when the fold's likes of every user in `users` are test data: the links u -> t and
    t -> u of type LIKE are never added (DataLoader.cs:287-298 runs only over the training set), and a tweet nobody likes
    any more is no node at all.  -> (links without them, {user: testSet ids ascending})."""
    src, dst, et = links["src"], links["dst"], links["etype"]
    node_id, node_type = links["node_id"], links["node_type"]
    n = len(node_id)
    drop = np.zeros(len(src), bool)
    test = {}
    order = np.argsort(src, kind="stable")
    starts = np.searchsorted(src[order], np.arange(n + 1))
    held_pairs = []
    for u in users:
        rows = order[starts[u]:starts[u + 1]]
        rows = rows[(et[rows] == LIKE) & (node_type[dst[rows]] == ITEM)]
        _, tst = split_like_history(node_id[dst[rows]], n_folds, fold)
        test[int(u)] = tst
        if len(tst) == 0:
            continue
        held = rows[np.isin(node_id[dst[rows]], tst)]
        drop[held] = True
        held_pairs.append(dst[held].astype(np.int64) * n + int(u))
    if held_pairs:
        hk = np.unique(np.concatenate(held_pairs))
        key = src.astype(np.int64) * n + dst
        drop |= (et == LIKE) & np.isin(key, hk)
    # A held-out tweet that no user likes any more does not exist in the reference's graph at all: DataLoader creates a tweet
    # node only while it walks somebody's likes (:291-303), and addAuthorship skips tweets that are not nodes (:355-356).  Here
    # the node index stays (indices must not shift), but every link that still touches it goes and its type becomes UNDEFINED,
    # so that it is no candidate of Recommender.cs:29 and can never be a hit of Experiment.cs:124.
    node_type = node_type.copy()
    if held_pairs:
        held_tweets = np.unique(np.concatenate(held_pairs) // n)
        still = (et == LIKE) & ~drop & (node_type[dst] == ITEM)
        liked = np.zeros(n, bool)
        liked[dst[still]] = True
        orphans = held_tweets[~liked[held_tweets]]
        if len(orphans):
            gone = np.zeros(n, bool)
            gone[orphans] = True
            drop |= gone[src] | gone[dst]
            node_type[orphans] = 0
    keep = ~drop
    out = dict(links)
    out["node_type"] = node_type
    for k in ("src", "dst", "etype", "w"):
        out[k] = links[k][keep]
    return out, test


def apply_methodology(links: dict, methodology: int) -> dict:
    """What Experiment hands to `new Graph(nodes, edges)` for `methodology`, starting from the network with every relation
    loaded: relations of features that are not in the list are not there at all (DataLoader.cs:235-250), FRIENDSHIP links
    of methodologies 4 / 9 / 14 are retyped UNDEFINED (Experiment.cs:84-101), and MENTION weights are
    nFriendhips * ln(cnt) / sum (DataLoader.cs:431) -- exactly 0.0 when no FRIENDSHIP link was loaded, and absent for a member
    that owns no other link then (DataLoader.cs:403-405)."""
    feats = FEATURES[int(methodology)]
    et = links["etype"].copy()
    w = links["w"].copy()
    keep = np.ones(len(et), bool)
    if F not in feats:
        keep &= et != FRIENDSHIP
    if T3 not in feats:
        keep &= et != FOLLOW
    if A not in feats:
        keep &= et != AUTHORSHIP
    if M not in feats:
        keep &= et != MENTION
    elif F not in feats:
        # addMentionCount2 walks the members that already own an `allLinks` entry (DataLoader.cs:403-405): a member none of
        # whose other relations was loaded gets no mention links at all (a dangling row, not a row of zero weights); the
        # others get them with nFriendhips = 0, i.e. weight exactly 0.0
        carrier = np.zeros(len(links["node_id"]), bool)
        carrier[links["src"][keep & (et != MENTION) & (et != 0)]] = True
        keep &= ~((et == MENTION) & ~carrier[links["src"]])
        w[et == MENTION] = 0.0
    if int(methodology) in RETYPE_FRIENDSHIP:
        et[et == FRIENDSHIP] = 0
    out = dict(links)
    out["etype"], out["w"] = et[keep], w[keep]
    out["src"], out["dst"] = links["src"][keep], links["dst"][keep]
    return out


def evaluate_ranking(ids, test_ids, k: int):
    """Experiment.cs:121-128, :136 over a ranking -> (nHits, averagePrecision, hits among the first k)."""
    test = set(int(x) for x in test_ids)
    n_hits, sum_precision, at_k = 0, 0.0, 0
    for i, x in enumerate(ids):
        if int(x) in test:
            n_hits += 1
            sum_precision += n_hits / (i + 1)
            if i < k:
                at_k += 1
    return n_hits, (0.0 if n_hits == 0 else sum_precision / n_hits), at_k


def evaluate_users(links: dict, users, test: dict, n_iter: int, k: int = 10):
    """Recommendation(u, 0.15f, nIterations) + the evaluation loop for every user, on the CPU oracle."""
    og = O.OracleGraph(links["node_id"], links["node_type"], links["src"], links["dst"], links["etype"], links["w"])
    assert og.build() == 0
    out = []
    for u in users:
        ids, _ = og.recommend(int(u), 0.15, n_iter)
        out.append(evaluate_ranking(ids, test[int(u)], k))
    og.close()
    return out


def run_k_fold(links: dict, methodology: int, n_folds: int, n_iter: int, ego: int = 0):
    """Experiment.cs:69-138 for one ego network -> (HIT, AVGPRECISION sum, per-fold list)."""
    hits, sum_ap, folds = 0.0, 0.0, []
    for fold in range(n_folds):
        held, test = hold_out(links, [ego], n_folds, fold)
        cfg = apply_methodology(held, methodology)
        (h, ap, _), = evaluate_users(cfg, [ego], test, n_iter)
        hits += h
        sum_ap += ap
        folds.append((len(test[ego]), h, ap))
    return hits, sum_ap, folds
