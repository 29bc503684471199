"""Host-side evaluation helpers (no GPU): hold-out of the newest likes and recall@k, against a hand-checked case and
the oracle's evaluate() arithmetic (Experiment.cs:121-128)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from recommendersystems_b200.experiment import hold_out_likes, recall_at_k
from recommendersystems_b200.rwr import EdgeType, NodeType


def _tiny():
    # users 0, 1; items 2..7 with ids 102..107; user 0 likes items 2..6 (5 likes), user 1 likes 2 and 7; friendship 0<->1
    node_id = np.array([10, 11, 102, 103, 104, 105, 106, 107], np.int64)
    node_type = np.array([1, 1, 2, 2, 2, 2, 2, 2], np.int32)
    src, dst, et = [], [], []
    for u, items in ((0, [2, 3, 4, 5, 6]), (1, [2, 7])):
        for t in items:
            src += [u, t]; dst += [t, u]; et += [EdgeType.LIKE, EdgeType.LIKE]
    src += [0, 1]; dst += [1, 0]; et += [EdgeType.FRIENDSHIP, EdgeType.FRIENDSHIP]
    order = np.argsort(np.asarray(src), kind="stable")
    return dict(node_id=node_id, node_type=node_type, src=np.asarray(src, np.int32)[order], dst=np.asarray(dst, np.int32)[order],
                etype=np.asarray(et, np.int32)[order], w=np.ones(len(src)))


def test_hold_out_takes_the_newest_fraction_in_both_directions():
    links = _tiny()
    held, test = hold_out_likes(links, [0, 1], fraction=0.4)
    # user 0: 5 likes -> int(5 * 0.4) = 2 newest by tweet id: 105, 106; user 1: int(2 * 0.4) = 0 -> nothing held out
    assert test[0].tolist() == [105, 106] and len(test[1]) == 0
    pairs = set(zip(held["src"].tolist(), held["dst"].tolist()))
    for t in (5, 6):
        assert (0, t) not in pairs and (t, 0) not in pairs               # both directions left the graph
    for t in (2, 3, 4):
        assert (0, t) in pairs and (t, 0) in pairs
    assert (1, 2) in pairs and (2, 1) in pairs and (0, 1) in pairs and (1, 0) in pairs
    assert len(held["src"]) == len(links["src"]) - 4
    assert (np.diff(held["src"]) >= 0).all()                             # still grouped by source, insertion order kept


def test_recall_at_k_and_reference_evaluate_agree():
    import oracle as O
    test = {0: np.array([105, 106], np.int64), 1: np.zeros(0, np.int64)}
    ids = np.array([[106, 104, 105, 0], [107, 0, 0, 0]], np.int64)
    cnt = np.array([3, 1], np.int32)
    r, hits, counted = recall_at_k(ids, cnt, [0, 1], test)
    assert (r, hits, counted) == (1.0, 2, 1)                             # user 1 has no test set and is not counted
    h, ap = O.evaluate(ids[0, :3].tolist(), test[0].tolist())
    assert h == 2 and abs(ap - (1 / 1 + 2 / 3) / 2) < 1e-15               # Experiment.cs:121-128, :136
