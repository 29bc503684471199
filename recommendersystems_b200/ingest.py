"""SQLite ingest of one ego network: TweetRecommender/SQLiteAdapter.cs + DataLoader.cs restated for the
flattened input of `Graph.from_arrays`.

The reference reloads the database for every methodology and fold (`Experiment.cs:69-77`).  Here the network is loaded
ONCE with every relation -- what `DataLoader.graphConfiguration(Methodology.ALL, fold)` builds, but with all of the ego
user's likes -- and the methodology and the fold are applied on the device (`rs.methodology_options`, `Graph.hold_out`).

Schema (SQLiteAdapter.cs:30-120): follow(source, target), tweet(id, author), retweet(user, tweet), quote(user, tweet),
favorite(user, tweet), mention(source, target).  The ego user's id is the file name (`DataLoader.cs:31`).

Order matters for the drop-in (node indices, insertion order of every `edges[i]`): every step below follows the order in
which DataLoader.cs adds nodes and links; `HashSet<long>` enumeration is taken as insertion order (the sets are only ever
added to), i.e. the row order of the SELECTs.  Host-side data preparation only: nothing here computes scores.
"""
from __future__ import annotations

import math
import os
import sqlite3
from typing import Dict, List, Tuple

import numpy as np

from .rwr import EdgeType, NodeType


def _distinct(rows) -> List[int]:
    """HashSet<long> filled in row order: first occurrence wins, enumeration in insertion order."""
    return list(dict.fromkeys(int(r[0]) for r in rows))


class EgoNetwork:
    """`DataLoader` for one *.sqlite file (DataLoader.cs:7-36): `allNodes` / `allLinks` with every relation loaded."""

    def __init__(self, db_path: str):
        self.ego_id = int(os.path.splitext(os.path.basename(db_path))[0])        # DataLoader.cs:31
        self.conn = sqlite3.connect(db_path)
        self.node_id: List[int] = []
        self.node_type: List[int] = []
        self.links: List[List[Tuple[int, int, float]]] = []      # allLinks[i] = [(target, type, weight)], insertion order
        self.user_idx: Dict[int, int] = {}
        self.member_idx: Dict[int, int] = {}
        self.tweet_idx: Dict[int, int] = {}
        self._seen: List[set] = []
        self._friends = 0

    # ---- SQLiteAdapter.cs ---------------------------------------------------------------------------------------
    def following(self, user: int) -> List[int]:               # getFollowingUsers, SQLiteAdapter.cs:27-40
        return _distinct(self.conn.execute("SELECT target FROM follow WHERE source = ?", (user,)))

    def authored(self, user: int) -> List[int]:                # getAuthorship, :42-54
        return _distinct(self.conn.execute("SELECT id FROM tweet WHERE author = ?", (user,)))

    def liked(self, user: int) -> List[int]:
        """retweets, quotes, favorites in that order, duplicates dropped (DataLoader.cs:98-107, :276-285)."""
        out: Dict[int, None] = {}
        for table in ("retweet", "quote", "favorite"):        # getRetweets / getQuotedTweets / getFavoriteTweets, :56-96
            for t in _distinct(self.conn.execute(f"SELECT tweet FROM {table} WHERE user = ?", (user,))):
                out.setdefault(t)
        return list(out)

    # ---- DataLoader.cs: nodes and links --------------------------------------------------------------------------
    def _add_node(self, nid: int, ntype: int) -> int:
        self.node_id.append(nid); self.node_type.append(ntype); self.links.append([]); self._seen.append(set())
        return len(self.node_id) - 1

    def add_user(self, uid: int, ntype: int) -> None:          # addUserNode, DataLoader.cs:39-50
        if uid not in self.user_idx:
            self.user_idx[uid] = self._add_node(uid, ntype)
            if ntype == NodeType.USER:
                self.member_idx[uid] = self.user_idx[uid]

    def add_tweet(self, tid: int) -> None:                     # addTweetNode, :52-59
        if tid not in self.tweet_idx:
            self.tweet_idx[tid] = self._add_node(tid, NodeType.ITEM)

    def add_link(self, src: int, dst: int, etype: int, weight: float) -> None:     # addLink, :61-77: (target, type) unique
        if (dst, etype) not in self._seen[src]:
            self._seen[src].add((dst, etype))
            self.links[src].append((dst, int(etype), float(weight)))

    def load(self) -> "EgoNetwork":
        """graphConfiguration(ALL) (DataLoader.cs:221-254) with every like of the ego user kept."""
        ego = self.ego_id
        # addMemberNodes, :256-267: the ego user, then the followees that follow back
        self.add_user(ego, NodeType.USER)
        self._friends = 0
        for f in self.following(ego):
            if ego in set(self.following(f)):
                self.add_user(f, NodeType.USER)
                self._friends += 1                             # getFriendsCountOfEgoUser, :110-120 (a self-follow counts there, too)
        members = list(self.member_idx)
        # addTweetNodesAndLikeEdges, :269-307.  The ego user's likes are walked in ascending tweet id (the reference walks the
        # training part of the id-sorted list, :126-138, :289-295), every other member's in query order.
        for m in members:
            likes = self.liked(m)
            if self.user_idx[m] == 0:
                likes = sorted(likes)
            for t in likes:
                self.add_tweet(t)
                self.add_link(self.user_idx[m], self.tweet_idx[t], EdgeType.LIKE, 1.0)
                self.add_link(self.tweet_idx[t], self.user_idx[m], EdgeType.LIKE, 1.0)
        # addAllFollowship, :317-346: FRIENDSHIP between members, FOLLOW to third-party users (created on the way)
        for m in members:
            im = self.user_idx[m]
            for f in self.following(m):
                if f in self.member_idx:
                    self.add_link(im, self.user_idx[f], EdgeType.FRIENDSHIP, 1.0)
                    self.add_link(self.user_idx[f], im, EdgeType.FRIENDSHIP, 1.0)
                else:
                    self.add_user(f, NodeType.ETC)
                    self.add_link(im, self.user_idx[f], EdgeType.FOLLOW, 1.0)
                    self.add_link(self.user_idx[f], im, EdgeType.FOLLOW, 1.0)
        # addAuthorship, :348-363
        for m in members:
            im = self.user_idx[m]
            for t in self.authored(m):
                if t in self.tweet_idx:
                    self.add_link(im, self.tweet_idx[t], EdgeType.AUTHORSHIP, 1.0)
                    self.add_link(self.tweet_idx[t], im, EdgeType.AUTHORSHIP, 1.0)
        # addMentionCount2, :398-436.  getMentionCount(a, b) (SQLiteAdapter.cs:113-124) counts both directions; one GROUP BY
        # replaces the M^2 pairs of COUNT(*) queries
        count: Dict[Tuple[int, int], int] = {}
        for s, t, c in self.conn.execute("SELECT source, target, COUNT(*) FROM mention GROUP BY source, target"):
            count[(int(s), int(t))] = int(c)
        for m1 in members:
            i1 = self.user_idx[m1]
            if not self.links[i1]:                             # `if (!allLinks.ContainsKey(idxMember)) continue;`
                continue
            mention, sum_log = [], 0.0
            for m2 in members:
                if m1 == m2:
                    continue
                c = count.get((m1, m2), 0) + count.get((m2, m1), 0)
                if c > 1:
                    mention.append((self.user_idx[m2], c))
                    sum_log += math.log(c)
            n_friends = sum(1 for (_, t, _) in self.links[i1] if t == EdgeType.FRIENDSHIP)
            if sum_log > 1:
                for i2, c in mention:
                    self.add_link(i1, i2, EdgeType.MENTION, n_friends * math.log(c) / sum_log)     # :431
        self.conn.close()                                      # closeDB, :218
        return self

    # ---- what Experiment needs -----------------------------------------------------------------------------------
    def like_count(self) -> int:                               # getLikeCountOfEgoUser, :94-109
        return sum(1 for (_, t, _) in self.links[0] if t == EdgeType.LIKE)

    def friends_count(self) -> int:                            # getFriendsCountOfEgoUser, :110-120
        return self._friends

    def is_valid(self, n_folds: int) -> bool:                  # checkEgoNetworkValidation, :79-92
        likes, friends = self.like_count(), self.friends_count()
        return not (likes < n_folds or likes < 50 or friends < 50)

    def arrays(self) -> Dict[str, np.ndarray]:
        """`for i in 0..N-1: foreach l in allLinks[i]` flattened: the input of `Graph.from_arrays`."""
        src = [i for i, ls in enumerate(self.links) for _ in ls]
        flat = [l for ls in self.links for l in ls]
        return dict(node_id=np.asarray(self.node_id, np.int64), node_type=np.asarray(self.node_type, np.int32),
                    src=np.asarray(src, np.int32), dst=np.asarray([l[0] for l in flat], np.int32),
                    etype=np.asarray([l[1] for l in flat], np.int32), w=np.asarray([l[2] for l in flat], np.float64))


def load_ego_network(db_path: str) -> Tuple[Dict[str, np.ndarray], EgoNetwork]:
    """-> (flattened links with every relation loaded, the loader with its id maps and counts)."""
    net = EgoNetwork(db_path).load()
    return net.arrays(), net


def run_experiment(data_dir: str, methodologies, n_folds: int, n_iter: int, result_name: str = "result.dat") -> List[str]:
    """`Program.Main` + `Experiment.runKFoldCrossValidation` for every *.sqlite of a directory (Program.cs:30-69): skips the
    (ego, methodology) pairs result.dat already holds (7 tab-separated tokens, Program.cs:36-50), appends one row per pair
    (Experiment.cs:141-155).  One load per ego network; folds and methodologies are applied on the device."""
    from .experiment import result_row, run_k_fold
    path = os.path.join(data_dir, result_name)
    done = set()
    if os.path.exists(path):
        for line in open(path):
            tok = line.rstrip("\n").split("\t")
            if len(tok) == 7:
                done.add((int(tok[0]), int(tok[1])))
    rows = []
    for name in sorted(f for f in os.listdir(data_dir) if f.endswith(".sqlite")):
        links, net = load_ego_network(os.path.join(data_dir, name))
        if not net.is_valid(n_folds):                           # Experiment.cs:72-74
            continue
        for m in methodologies:
            if (net.ego_id, int(m)) in done:                    # Experiment.cs:53-58
                continue
            out = run_k_fold(links, int(m), n_folds, n_iter, ego=0, validate=False)
            with open(path, "a") as f:
                f.write(out["row"] + "\n")
            rows.append(out["row"])
    return rows
