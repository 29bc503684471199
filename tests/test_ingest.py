"""CPU tests of the SQLite ingest: a hand-built ego network in the schema of SQLiteAdapter.cs:30-120,
loaded in the order DataLoader.cs adds nodes and links, checked against the lists worked out by hand from DataLoader.cs."""
import math
import os
import sqlite3

import numpy as np

from recommendersystems_b200.ingest import EgoNetwork, load_ego_network, run_experiment
from recommendersystems_b200.rwr import EdgeType as E, NodeType as N


def make_db(path):
    c = sqlite3.connect(path)
    c.executescript("""
        CREATE TABLE follow(source INTEGER, target INTEGER);
        CREATE TABLE tweet(id INTEGER, author INTEGER);
        CREATE TABLE retweet(user INTEGER, tweet INTEGER);
        CREATE TABLE quote(user INTEGER, tweet INTEGER);
        CREATE TABLE favorite(user INTEGER, tweet INTEGER);
        CREATE TABLE mention(source INTEGER, target INTEGER);
    """)
    c.executemany("INSERT INTO follow VALUES (?, ?)", [(100, 101), (100, 102), (100, 104), (101, 100), (101, 103), (101, 102),
                                                       (104, 100), (104, 101)])
    c.executemany("INSERT INTO tweet VALUES (?, ?)", [(5001, 101), (5002, 100), (5003, 999), (5004, 104)])
    c.executemany("INSERT INTO retweet VALUES (?, ?)", [(100, 5002)])
    c.executemany("INSERT INTO quote VALUES (?, ?)", [(101, 5001)])
    c.executemany("INSERT INTO favorite VALUES (?, ?)", [(100, 5001), (101, 5003), (101, 5001), (104, 5002)])
    c.executemany("INSERT INTO mention VALUES (?, ?)", [(100, 101), (100, 101), (101, 100), (100, 104), (100, 104), (104, 100), (104, 100),
                                                        (101, 104)])
    c.commit()
    c.close()


def test_ego_network_follows_dataloader_order(tmp_path):
    db = str(tmp_path / "100.sqlite")
    make_db(db)
    links, net = load_ego_network(db)
    assert net.ego_id == 100
    # addMemberNodes: ego, then the followees that follow back (101, 104; 102 does not).  Tweets: the ego's likes in ascending id
    # (5001, 5002), then 101's (quote 5001, favorites 5003), then 104's (5002).  Third-party users as addAllFollowship meets them.
    assert links["node_id"].tolist() == [100, 101, 104, 5001, 5002, 5003, 102, 103]
    assert links["node_type"].tolist() == [N.USER, N.USER, N.USER, N.ITEM, N.ITEM, N.ITEM, N.ETC, N.ETC]
    adj = [[(d, t) for (d, t, _) in ls] for ls in net.links]
    # ego: likes, then friendship / follow in the order of its followees (101 member, 102 third party, 104 member), authorship, mentions
    assert adj[0] == [(3, E.LIKE), (4, E.LIKE), (1, E.FRIENDSHIP), (6, E.FOLLOW), (2, E.FRIENDSHIP), (4, E.AUTHORSHIP), (1, E.MENTION), (2, E.MENTION)]
    # 101: likes 5001, 5003; friendship with the ego (added while the ego's followees were walked), then its own followees:
    # 100 (duplicate, dropped by addLink), 103 and 102 third party; friendship from 104's walk; authorship of 5001; mention of the ego
    assert adj[1] == [(3, E.LIKE), (5, E.LIKE), (0, E.FRIENDSHIP), (7, E.FOLLOW), (6, E.FOLLOW), (2, E.FRIENDSHIP), (3, E.AUTHORSHIP), (0, E.MENTION)]
    assert adj[3] == [(0, E.LIKE), (1, E.LIKE), (1, E.AUTHORSHIP)]            # tweet 5001: liked by 100 and 101, written by 101
    assert adj[5] == [(1, E.LIKE)]                                           # tweet 5003: its author 999 is no member
    assert adj[6] == [(0, E.FOLLOW), (1, E.FOLLOW)]                          # third party 102
    # 5004 (written by 104) is liked by nobody: no node (addAuthorship skips tweets that are not in tweetIDs, DataLoader.cs:354)
    assert 5004 not in net.tweet_idx
    # mention weights, DataLoader.cs:431: nFriendhips * ln(cnt) / sum ln(cnt).  ego: pairs (101: 2 + 1 = 3), (104: 2 + 2 = 4), 2 friends
    w = {(i, d, t): x for i, ls in enumerate(net.links) for (d, t, x) in ls}
    s = math.log(3) + math.log(4)
    assert w[(0, 1, E.MENTION)] == 2 * math.log(3) / s and w[(0, 2, E.MENTION)] == 2 * math.log(4) / s
    # 101: partner ego (3); 104 has only 1 mention with 101 (not > 1).  sum = ln 3 > 1, 2 friends (ego, 104)
    assert w[(1, 0, E.MENTION)] == 2 * math.log(3) / math.log(3) and (1, 2, E.MENTION) not in w
    # flattened arrays: grouped by source, insertion order
    assert (np.diff(links["src"]) >= 0).all() and len(links["src"]) == sum(len(ls) for ls in net.links)
    assert links["dst"][links["src"] == 0].tolist() == [d for d, _ in adj[0]]
    assert net.like_count() == 2 and net.friends_count() == 2 and not net.is_valid(2)      # < 50 likes / friends: Experiment returns


def test_run_experiment_skips_invalid_networks_and_finished_pairs(tmp_path):
    make_db(str(tmp_path / "100.sqlite"))
    (tmp_path / "result.dat").write_text("100\t8\t2\t5\t0\t2\t0\nbroken line\n")
    assert run_experiment(str(tmp_path), [8], 2, 5) == []                    # invalid ego network: no GPU work, no row
    assert (tmp_path / "result.dat").read_text().count("\n") == 2
