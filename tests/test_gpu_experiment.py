"""GPU parity tests of the rows either side of the hot path: N2 rwr_graph_hold_out, N3 rwr_methodology_masks /
zero-weight links / link types of the CSR, N1 rwr_evaluate_users and the k-fold driver -- each against the CPU restatement
in oracle/experiment_ref.py (DataLoader.cs:122-140, :142-219, Experiment.cs:84-138)."""
import numpy as np
import pytest

from conftest import C1_SPEC, bits

pytestmark = pytest.mark.gpu

import experiment_ref as R
import oracle as O
import recommendersystems_b200 as rs
from recommendersystems_b200 import experiment as X
from recommendersystems_b200.rwr import run_fixed

SPEC = dict(seed=2031, n_users=3_000, n_items=24_000, n_third=300, authorship_per_mille=900, n_like=200_000, n_friend=30_000,
            n_follow=2_000, n_mention=1_500, undefined_per_mille=0, scramble=1, p1_byte=61)


def from_links(links, **opts):
    return rs.Graph.from_arrays(links["node_id"], links["node_type"], links["src"], links["dst"], links["etype"], links["w"], **opts)


def like_users(links, n_users, at_least, count):
    like = links["etype"] == 1
    deg = np.bincount(links["src"][like], minlength=len(links["node_id"]))
    return np.flatnonzero(deg[:n_users] >= at_least)[:count]


# ------------------------------------------------------------------------------------------ N2
@pytest.mark.parametrize("n_folds,fold", [(10, 9), (10, 0), (4, 2), (1, 0), (7, 6)])
def test_hold_out_matches_the_restatement(n_folds, fold):
    full = O.synth_generate(SPEC)
    users = like_users(full, SPEC["n_users"], 1, 40).tolist() + like_users(full, SPEC["n_users"], 30, 25).tolist()[-10:]
    users = list(dict.fromkeys(users + [int(np.flatnonzero(np.bincount(full["src"], minlength=len(full["node_id"]))[:SPEC["n_users"]] == 0)[0])]
                               if (np.bincount(full["src"], minlength=len(full["node_id"]))[:SPEC["n_users"]] == 0).any() else users))
    want_links, want_test = R.hold_out(full, users, n_folds, fold)
    g = from_links(full)
    test = g.hold_out(users, n_folds, fold)
    got = g.export_links()
    for k in ("src", "dst", "etype", "node_type"):
        assert np.array_equal(got[k], want_links[k]), k
    assert np.array_equal(bits(got["w"]), bits(want_links["w"]))
    # held-out tweets that nobody likes any more: no node of the reference's graph (retyped UNDEFINED here, their links gone)
    orphans = np.flatnonzero(got["node_type"] != full["node_type"])
    assert len(orphans) >= 10 and (got["node_type"][orphans] == 0).all() and (full["node_type"][orphans] == 2).all()
    assert not np.isin(got["src"], orphans).any() and not np.isin(got["dst"], orphans).any()
    for u in users:
        assert test[u].tolist() == want_test[u].tolist(), u
    # the edited graph builds like any other, bit-exact CSR against the oracle on the restated links
    g.buildGraph()
    og = O.OracleGraph(want_links["node_id"], want_links["node_type"], want_links["src"], want_links["dst"], want_links["etype"], want_links["w"])
    assert og.build() == 0
    rp, col, val = g.csr()
    orp, ocol, oval = og.csr()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol) and np.array_equal(bits(val), bits(oval))
    g.close()


def test_hold_out_rejects_a_built_graph_and_bad_arguments():
    full = O.synth_generate(C1_SPEC)
    g = from_links(full)
    with pytest.raises(rs.RwrError):
        g.hold_out([0], 0, 0)
    with pytest.raises(rs.RwrError):
        g.hold_out([0], 3, 3)
    with pytest.raises(KeyError):
        g.hold_out([len(full["node_id"])], 3, 0)
    with pytest.raises(rs.RwrError):
        g.hold_out([0, 5, 0], 3, 0)                    # a user listed twice
    g.buildGraph()
    with pytest.raises(ValueError):
        g.hold_out([0], 3, 0)
    g.close()


# ------------------------------------------------------------------------------------------ N3
@pytest.mark.parametrize("methodology", list(range(16)))
def test_methodology_masks_reproduce_the_configured_graph(methodology):
    """A graph with every relation + the methodology's masks == the graph Experiment.cs builds for that methodology:
    bit-exact transition matrix, scores within 1e-12, identical top-10."""
    full = O.synth_generate(SPEC)
    cfg = R.apply_methodology(full, methodology)
    og = O.OracleGraph(cfg["node_id"], cfg["node_type"], cfg["src"], cfg["dst"], cfg["etype"], cfg["w"])
    assert og.build() == 0
    g = from_links(full, **rs.methodology_options(methodology))
    g.buildGraph()
    rp, col, val = g.csr()
    orp, ocol, oval = og.csr()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)
    assert np.array_equal(bits(val), bits(oval))                       # NaN rows (all weights 0, methodology 15) included
    # graph[i][k].type
    t = g.csr_types()
    explicit = (cfg["etype"] != 0)
    assert np.array_equal(t, cfg["etype"][explicit])
    deg = np.bincount(cfg["src"][cfg["etype"] == 1], minlength=og.n)
    seed = int(np.flatnonzero(deg[:SPEC["n_users"]] >= 5)[3])
    c = rs.widen_float(0.15)
    res = run_fixed(g, [seed], c, 8)
    want, _ = og.run(seed, c, n_iter=8)
    got = res.scores(0)
    res.close()
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)
    nz = (want != 0) & ~nan
    assert np.all(got[(want == 0)] == 0) and (nz.sum() == 0 or (np.abs(got[nz] - want[nz]) / want[nz]).max() <= 1e-12)
    if not nan.any():
        top = rs.Recommender(g).Recommendation(seed, 0.15, 8, 10)
        oids, _ = og.recommend(seed, 0.15, 8, top_n=10)
        assert [p[0] for p in top] == oids.tolist()
    g.close()


def test_link_types_above_31_are_not_aliased_by_the_mask():
    """etype 34 must not alias onto bit 2 (FRIENDSHIP) of undefined_type_mask."""
    node_id = np.arange(4, dtype=np.int64) + 100
    node_type = np.array([1, 2, 2, 2], np.int32)
    src = np.array([0, 0, 1, 2], np.int32); dst = np.array([1, 2, 0, 0], np.int32)
    et = np.array([34, 2, 1, 1], np.int32); w = np.ones(4)
    g = rs.Graph.from_arrays(node_id, node_type, src, dst, et, w, undefined_types=[rs.EdgeType.FRIENDSHIP])
    g.buildGraph()
    rp, col, _ = g.csr()
    assert rp.tolist() == [0, 1, 2, 3, 3] and col.tolist() == [1, 0, 0]          # the type-34 link stays, the FRIENDSHIP link left
    assert g.csr_types().tolist() == [34, 1, 1]
    g.close()


# ------------------------------------------------------------------------------------------ N1
@pytest.mark.parametrize("precision", [rs.FP64, rs.FP32])
def test_evaluate_users_matches_the_reference_loop(precision):
    """Hits / average precision over the FULL ranking and hits@10 for 40 held-out users against Recommendation() +
    Experiment.cs:121-128 on the CPU oracle."""
    full = O.synth_generate(SPEC)
    users = like_users(full, SPEC["n_users"], 20, 40).tolist()
    held, test = R.hold_out(full, users, 10, 9)
    g = from_links(full)
    g.hold_out(users, 10, 9)
    g.buildGraph()
    r = rs.evaluate_users(g, None, None, 0.15, 12, k=10, precision=precision)
    want = R.evaluate_users(held, users, test, 12, 10)
    assert r["n_test"].tolist() == [len(test[u]) for u in users]
    assert r["hits"].tolist() == [w[0] for w in want]
    assert r["hits"].sum() > 50
    ap = np.array([w[1] for w in want])
    if precision == rs.FP64:
        assert r["hits_at_k"].tolist() == [w[2] for w in want]
        assert np.abs(r["avg_precision"] - ap).max() <= 1e-12
    else:
        assert np.abs(r["avg_precision"] - ap).max() <= 1e-3           # FP32 scores may permute near-ties of the ranking
    # explicit test sets, user list in another order, ids that are no candidates (unknown id, a USER node, a liked tweet)
    sub = users[5:13][::-1]
    liked = int(full["node_id"][held["dst"][(held["src"] == sub[0]) & (held["etype"] == 1)][0]])
    tsets = {u: list(test[u]) for u in sub}
    tsets[sub[0]] = tsets[sub[0]] + [999, int(full["node_id"][sub[1]]), liked]
    r2 = rs.evaluate_users(g, sub, tsets, 0.15, 12, k=10, precision=precision)
    idx = [users.index(u) for u in sub]
    assert r2["hits"].tolist() == r["hits"][idx].tolist()
    assert r2["n_test"][0] == len(test[sub[0]]) + 3
    if precision == rs.FP64:
        assert np.abs(r2["avg_precision"] - r["avg_precision"][idx]).max() <= 1e-15
    g.close()


def test_evaluate_users_counts_many_test_items_per_user():
    """More than EV_GROUP (8) test items per user: several counting passes over the rank tile."""
    full = O.synth_generate(SPEC)
    users = like_users(full, SPEC["n_users"], 120, 6).tolist()
    assert len(users) >= 3
    held, test = R.hold_out(full, users, 2, 1)              # half of >= 120 likes: >= 60 test items each
    g = from_links(full)
    g.hold_out(users, 2, 1)
    g.buildGraph()
    r = rs.evaluate_users(g, None, None, 0.15, 10, k=10)
    want = R.evaluate_users(held, users, test, 10, 10)
    assert r["hits"].tolist() == [w[0] for w in want] and min(r["hits"]) >= 60
    assert np.abs(r["avg_precision"] - np.array([w[1] for w in want])).max() <= 1e-12
    assert r["hits_at_k"].tolist() == [w[2] for w in want]
    g.close()


@pytest.mark.parametrize("methodology", [8, 4, 0])
def test_k_fold_driver_matches_the_reference_loop(methodology):
    """Experiment.runKFoldCrossValidation for one ego network (C1 size): HIT, MAP and the result.dat row."""
    full = O.synth_generate(C1_SPEC)
    like = full["etype"] == 1
    deg = np.bincount(full["src"][like], minlength=len(full["node_id"]))
    ego = int(np.argmax(deg[:C1_SPEC["n_users"]]))
    out = X.run_k_fold(full, methodology, n_folds=5, n_iter=10, ego=ego, validate=False)
    hits, sum_ap, folds = R.run_k_fold(full, methodology, 5, 10, ego)
    assert out["hits"] == int(hits) and [f["hits"] for f in out["folds"]] == [f[1] for f in folds]
    assert abs(out["map"] - sum_ap / 5) <= 1e-12
    tok = out["row"].split("\t")
    assert len(tok) == 7 and tok[:4] == [str(int(full["node_id"][ego])), str(methodology), "5", "10"] and tok[5] == str(int(deg[ego]))


# ------------------------------------------------------------------------------------------ N4 -> N3 -> N2 -> hot path -> N1
def _random_ego_db(path, seed=7, n_members=40, n_third=60, n_tweets=1500):
    """A random ego network in the schema of SQLiteAdapter.cs:30-120 (ego 1000, mutual followees = members)."""
    import sqlite3
    rng = np.random.default_rng(seed)
    c = sqlite3.connect(path)
    c.executescript("CREATE TABLE follow(source INTEGER, target INTEGER); CREATE TABLE tweet(id INTEGER, author INTEGER);"
                    "CREATE TABLE retweet(user INTEGER, tweet INTEGER); CREATE TABLE quote(user INTEGER, tweet INTEGER);"
                    "CREATE TABLE favorite(user INTEGER, tweet INTEGER); CREATE TABLE mention(source INTEGER, target INTEGER);")
    ego = 1000
    members = [ego] + [1001 + i for i in range(n_members)]
    third = [5000 + i for i in range(n_third)]
    follow = []
    for m in members[1:]:
        follow += [(ego, m), (m, ego)]
    for m in members:
        for o in rng.choice(members, 8, replace=False):
            if o != m:
                follow.append((m, int(o)))
        for t in rng.choice(third, 5, replace=False):
            follow.append((m, int(t)))
    c.executemany("INSERT INTO follow VALUES (?, ?)", follow)
    tweets = [900000 + i for i in range(n_tweets)]
    c.executemany("INSERT INTO tweet VALUES (?, ?)", [(t, int(rng.choice(members + third))) for t in tweets])
    for table, k in (("retweet", 25), ("quote", 5), ("favorite", 40)):
        rows = [(m, int(t)) for m in members for t in rng.choice(tweets, k, replace=False)]
        c.executemany(f"INSERT INTO {table} VALUES (?, ?)", rows)
    c.executemany("INSERT INTO mention VALUES (?, ?)", [(int(a), int(b)) for a, b in rng.choice(members, (600, 2)) if a != b])
    c.commit()
    c.close()


@pytest.mark.parametrize("methodology", [8, 15, 9])
def test_sqlite_ego_network_through_the_fold_loop(tmp_path, methodology):
    """SQLite file -> ingest (N4) -> methodology masks (N3) -> device hold-out (N2) -> RWR -> device evaluation (N1), against
    the same links through the CPU restatement of the fold loop; the MENTION weights make this the valued layout."""
    from recommendersystems_b200.ingest import load_ego_network
    db = str(tmp_path / "1000.sqlite")
    _random_ego_db(db)
    links, net = load_ego_network(db)
    assert net.like_count() >= 60 and (links["etype"] == R.MENTION).sum() > 10 and (links["node_type"] == 3).sum() > 10
    out = X.run_k_fold(links, methodology, n_folds=4, n_iter=12, ego=0, validate=False)
    hits, sum_ap, folds = R.run_k_fold(links, methodology, 4, 12, 0)
    assert out["hits"] == int(hits) and 0 < out["hits"] < net.like_count()  # held-out tweets nobody else likes are no nodes: no hit
    assert [f["hits"] for f in out["folds"]] == [f[1] for f in folds]
    assert abs(out["map"] - sum_ap / 4) <= 1e-12
    assert out["row"].split("\t")[:4] == ["1000", str(methodology), "4", "12"]


@pytest.mark.parametrize("methodology", [8, 15, 4, 0])
def test_k_fold_against_the_reference_itself(tmp_path, methodology):
    """The whole chain on the device -- SQLite ingest (N4), methodology masks (N3), hold-out (N2), RWR, evaluation (N1) -- against
    the reference's OWN DataLoader and k-fold loop (DataLoader.cs + Experiment.cs compiled from their sources, oracle/_ref):
    HIT identical, AVGPRECISION to 1e-12, the result.dat row.  Without the prebuilt library: against the restated loop."""
    import random
    import ego_db
    import ref as RF
    from recommendersystems_b200.ingest import load_ego_network
    tables = ego_db.random_tables(random.Random(32))      # 15 mention links lose their carrier under methodology 15
    db = ego_db.write_sqlite(str(tmp_path / "1000.sqlite"), tables)
    links, net = load_ego_network(db)
    assert net.is_valid(10)
    out = X.run_k_fold(links, methodology, n_folds=10, n_iter=6, ego=0)
    if RF.available(build=False):
        ref = RF.ReferenceDb(db, tables).experiment(10, 6, methodology)
        want_hits, want_ap, likes = ref["hit"], ref["avg_precision_sum"], ref["cnt_likes"]
        assert ref["valid"]
    else:
        want_hits, want_ap, _ = R.run_k_fold(links, methodology, 10, 6, 0)
        likes = net.like_count()
    assert out["hits"] == int(want_hits) and out["hits"] < likes == out["cnt_likes"]
    assert abs(out["map"] - want_ap / 10) <= 1e-12
    tok = out["row"].split("\t")
    assert tok[:6] == ["1000", str(methodology), "10", "6", str(int(want_hits)), str(likes)] and abs(float(tok[6]) - want_ap / 10) <= 1e-14


def test_evaluate_users_with_thousands_of_test_items():
    """More test items in one seed tile than the shared-memory histogram of k_ev_hist holds (2 048): the global-memory form."""
    full = O.synth_generate(SPEC)
    users = like_users(full, SPEC["n_users"], 20, 3).tolist()
    items = full["node_id"][full["node_type"] == 2]
    rng = np.random.default_rng(5)
    tsets = {u: rng.choice(items, 1500, replace=False).tolist() for u in users}       # 4 500 items in the tile
    g = from_links(full)
    g.buildGraph()
    r = rs.evaluate_users(g, users, tsets, 0.15, 10, k=10)
    og = O.OracleGraph(full["node_id"], full["node_type"], full["src"], full["dst"], full["etype"], full["w"])
    assert og.build() == 0
    for i, u in enumerate(users):
        ids, _ = og.recommend(int(u), 0.15, 10)
        h, ap, atk = R.evaluate_ranking(ids, tsets[u], 10)
        assert (int(r["hits"][i]), int(r["hits_at_k"][i])) == (h, atk) and abs(float(r["avg_precision"][i]) - ap) <= 1e-12
        assert 0 < h <= 1500                                  # liked tweets among the 1 500 are no candidates
    g.close()
