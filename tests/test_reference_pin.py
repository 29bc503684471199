"""CPU tests: the oracle is pinned to the reference's OWN sources.

No C# toolchain exists in this image, so `make -C oracle ref` compiles Recommenders/RWRBased/{Graph,Model,Recommender}.cs
from where they lie under /root/reference after oracle/cs2cpp.py has respelt their declarations for a C++ compiler
(oracle/_ref/libref.so; the respelt text lives in a temporary directory during the build, nothing of the reference stays).  Here:
  * the translator's rules are checked on a C# text written for this test (runs anywhere, needs no reference);
  * where libref.so exists, the transliterated reference must reproduce every committed golden vector bit for bit, and the
    hand-written oracle (literal and collapsed forms) must agree with it bit for bit on random graphs with the corner cases
    of SURVEY 8(a): dangling rows, multi-edges, UNDEFINED-only rows, fractional weights, zero weight sums (NaN / Inf),
    the exceptions of the drop-in boundary, and on the C1-shaped synthetic graph;
  * the callers' rows: the reference's own DataLoader (over in-memory tables instead of System.Data.SQLite)
    and the k-fold loop of Experiment.runKFoldCrossValidation against the product's ingest + the restated hold-out,
    methodology table and evaluation loop -- node and link order, test sets, HIT and AVGPRECISION of all 16 methodologies.
"""
import hashlib
import os
import random

import numpy as np
import pytest

from conftest import C1_SPEC, GOLDEN, bits, load_golden, unhex
import cs2cpp
from random_graphs import random_flat
import oracle as O
import ref as RF

CASES = ["kat_8c", "small_a", "small_b"]
HAVE_REF = RF.available()
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libref.so: no reference sources and no prebuilt library here")

# ---------------------------------------------------------------------------------------------- the translator's rules
CS_SAMPLE = """﻿using System.Collections.Generic;

namespace Demo.Things {
    public enum Colour { RED, GREEN }

    public struct Pair {
        public long key;
        public Colour colour;

        public Pair(long key) {
            this.key = key;
            this.colour = Colour.RED;
        }
    }

    public class Holder {
        public Dictionary<int, Pair[]> table;
        private double[] values;

        public Holder(int n) {
            this.table = new Dictionary<int, Pair[]>();
            values = new double[n];
            values[0] = 1d / n;   // } a brace in a comment
        }

        public int size() {
            return table.Count;
        }
    }

    public class User {
        private Holder holder;

        public User(Holder holder) {
            this.holder = holder;
        }

        public double total(float scale) {
            Holder other = new Holder(3);
            var list = new List<KeyValuePair<long, double>>();
            Pair[] row = null;
            foreach (Pair p in holder.table[0]) {
                if (p.colour == Colour.GREEN && row != null)
                    list.Add(new KeyValuePair<long, double>(p.key, scale * other.size()));
            }
            list.Sort((one, another) => {
                return one.Value.CompareTo(another.Value) * -1;
            });
            return list.Count + row.Length + double.MaxValue;
        }
    }
}
"""


def test_translator_rules():
    out = cs2cpp.transliterate({"Demo.cs": CS_SAMPLE})
    # order: enums, structs, classes by dependency (User names Holder)
    assert out.index("enum class Colour") < out.index("struct Pair") < out.index("struct Holder") < out.index("struct User")
    assert "namespace Demo_Things {" in out and "using System" not in out and "﻿" not in out
    assert "enum class Colour : int { RED, GREEN };" in out
    assert "Pair() = default;" in out and "long long key;" in out
    assert "this->colour = Colour::RED;" in out
    assert "public" not in out and "private" not in out
    assert "Dictionary<int, Array<Pair>> table;" in out and "Array<double> values;" in out
    assert "this->table = Dictionary<int, Array<Pair>>();" in out
    assert "values = Array<double>(n);" in out and "values[0] = 1.0 / n;" in out
    assert "return table.Count();" in out
    # class-typed variables are pointers inside the type that declares them; `new C(..)` stays
    assert "Holder* holder;" in out and "User(Holder* holder) {" in out and "this->holder = holder;" in out
    assert "Holder* other = new Holder(3);" in out and "other->size()" in out
    assert "auto list = List<KeyValuePair<long long, double>>();" in out
    assert "Array<Pair> row = nullptr;" in out and "row != nullptr" in out
    assert "for (Pair p : holder->table[0]) {" in out
    assert "p.colour == Colour::GREEN" in out
    assert "list.Add(KeyValuePair<long long, double>(p.key, scale * other->size()));" in out
    assert "list.Sort([&](auto one, auto another) {" in out
    assert "return CompareTo(one.Value, another.Value) * -1;" in out
    assert "return list.Count() + row.Length() + std::numeric_limits<double>::max();" in out
    # every line of a method body that needs no rule is passed through untouched
    assert "            values[0] = 1.0 / n;   // } a brace in a comment" in out


def test_translator_touches_no_statement_of_the_reference():
    """With the reference present: every line of the generated header is a line of the reference after the declared rules,
    and the lines that carry arithmetic (those with an arithmetic assignment) are character-for-character the reference's."""
    root = os.path.join(RF.REFERENCE_ROOT, "Recommenders", "RWRBased")
    if not os.path.isdir(root):
        pytest.skip("no reference sources on this machine")
    files = {}
    for rel in cs2cpp.SOURCES:
        with open(os.path.join(RF.REFERENCE_ROOT, rel), encoding="utf-8-sig") as f:
            files[rel] = f.read()
    out_lines = set(l.strip() for l in cs2cpp.transliterate(files).splitlines())
    arithmetic = [l.strip() for t in files.values() for l in t.splitlines()
                  if any(op in l for op in ("+=", "/=", "*=", " * ", " - ", " / ")) and "//" not in l.split("=")[0]]
    assert len(arithmetic) >= 10
    untouched = [l for l in arithmetic if l in out_lines]
    # the only arithmetic lines a rule respells: `1d` literals, `double.MaxValue`, `graph.size()` -> `graph->size()`, CompareTo
    respelt = [l for l in arithmetic if l not in out_lines]
    assert all(("1d" in l or "MaxValue" in l or "CompareTo" in l or "graph." in l) for l in respelt), respelt
    assert len(untouched) >= 8
    if HAVE_REF:                                            # libref.so was made from exactly these files
        for rel, text in files.items():
            assert RF.source_hashes()[rel] == hashlib.sha256(text.encode("utf-8")).hexdigest(), rel


# ---------------------------------------------------------------------------------------------- reference vs golden vectors
def make(g):
    i = g["input"]
    rg = RF.ReferenceGraph(i["node_id"], i["node_type"], i["src"], i["dst"], i["etype"], i["w"])
    assert rg.build() == 0
    return rg


@needs_ref
@pytest.mark.parametrize("name", CASES)
def test_reference_reproduces_the_golden_vectors(name):
    g = load_golden(name)
    rg = make(g)
    rp, col, val = rg.csr()
    assert rp.tolist() == g["csr"]["row_ptr"] and col.tolist() == g["csr"]["col"]
    assert np.array_equal(bits(val), bits(unhex(g["csr"]["val"])))
    assert rg.build() == RF.REF_E_ALREADY_BUILT           # graph.Add(i, ..) on an existing key: ArgumentException (Graph.cs:86)
    c = float.fromhex(g["damping_double"])
    for e in g["seeds"]:
        for n, want in e["ranks"].items():
            rank, it = rg.run(e["seed"], c, n_iter=int(n))
            assert it == int(n) and np.array_equal(bits(rank), bits(unhex(want))), (name, e["seed"], n)
        for thr, want in e["thresholds"].items():
            kw = dict(default_threshold=True) if thr == "default" else dict(threshold=float(thr))
            if want["iters"] is None:                     # no bitwise fixed point: the reference's run() never returns
                assert rg.run(e["seed"], c, max_iter=300, **kw)[1] == 300
                continue
            rank, it = rg.run(e["seed"], c, **kw)                        # counted loop over the public methods
            own, _ = rg.run(e["seed"], c, own_loop=True, **kw)           # Model.run() / Model.run(double) themselves
            assert it == want["iters"], (name, e["seed"], thr)
            assert np.array_equal(bits(rank), bits(unhex(want["rank"]))) and np.array_equal(bits(rank), bits(own))
        if e["recommendation"] == "KeyNotFoundException":
            with pytest.raises(KeyError):
                rg.recommend(e["seed"], g["damping_float"], 10)
            continue
        rec = e["recommendation"]
        ids, sc = rg.recommend(e["seed"], g["damping_float"], rec["n_iter"])
        assert ids.tolist() == rec["ids"] and np.array_equal(bits(sc), bits(unhex(rec["scores"])))
        for k, want in e["top"].items():
            ids, sc = rg.recommend(e["seed"], g["damping_float"], rec["n_iter"], top_n=int(k))
            assert ids.tolist() == want["ids"] and np.array_equal(bits(sc), bits(unhex(want["scores"])))
        ids0, _ = rg.recommend(e["seed"], g["damping_float"], rec["n_iter"], top_n=0)       # Recommender.cs:47
        assert ids0.tolist() == rec["ids"]
    for n, want in g["uniform"].items():
        rank, _ = rg.run(-1, c, n_iter=int(n))
        assert np.array_equal(bits(rank), bits(unhex(want)))


# ---------------------------------------------------------------------------------------------- oracle vs reference, differential
@needs_ref
@pytest.mark.parametrize("case", range(6))
def test_oracle_matches_the_reference_on_random_graphs(case):
    rng = random.Random(20260200 + case)
    poisoned = case >= 4                                   # cases 4, 5: NaN / Inf rows
    f = random_flat(rng, rng.randrange(3, 12), rng.randrange(5, 60), rng.randrange(0, 4), rng.randrange(20, 400),
                    zero_rows=1 if poisoned else 0, nan_rows=1 if poisoned else 0)
    args = (f["node_id"], f["node_type"], f["src"], f["dst"], f["etype"], f["w"])
    rg, og = RF.ReferenceGraph(*args), O.OracleGraph(*args)
    assert rg.build() == 0 and og.build() == 0
    for a, b in zip(rg.csr(), og.csr()):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    n = rg.n
    c = O.widen_float(0.15)
    has_links = np.bincount(np.asarray(f["src"], np.int64), minlength=n) > 0
    for seed in sorted(set([0, n // 2, n - 1] + [rng.randrange(n) for _ in range(3)])):
        for it in (0, 1, 2, 7, 20):
            want, _ = rg.run(seed, c, n_iter=it)
            for literal in (True, False):
                got, _ = og.run(seed, c, n_iter=it, literal=literal)       # NaN / Inf rows: `x * 0` poisons every node, in both forms
                assert np.array_equal(bits(got), bits(want)), (case, seed, it, literal)
        if not poisoned:
            for thr in (1e-3, 1e-9):
                want, wit = rg.run(seed, c, threshold=thr, max_iter=500)
                got, git = og.run(seed, c, threshold=thr, max_iter=500)
                assert wit == git and np.array_equal(bits(got), bits(want))
        if not has_links[seed]:
            with pytest.raises(KeyError):
                rg.recommend(seed, 0.15, 5)
            with pytest.raises(KeyError):
                og.recommend(seed, 0.15, 5)
            continue
        for top in (None, 1, 10, 0, -3):
            wi, ws = rg.recommend(seed, 0.15, 5, top_n=top)
            gi, gs = og.recommend(seed, 0.15, 5, top_n=top)
            assert wi.tolist() == gi.tolist() and np.array_equal(bits(ws), bits(gs)), (case, seed, top)
    u_want, _ = rg.run(-1, c, n_iter=4)                    # the uniform-restart constructor (Model.cs:14-31)
    u_got, _ = og.run(-1, c, n_iter=4, literal=True)
    assert np.array_equal(bits(u_got), bits(u_want))


@needs_ref
def test_oracle_matches_the_reference_on_many_small_graphs():
    """Breadth: 80 more random graphs (a third of them with NaN / Inf rows), one seed and three iteration counts each, both forms of
    the oracle, the full recommendation list -- all bit for bit."""
    c = O.widen_float(0.15)
    compared = 0
    for case in range(80):
        rng = random.Random(77000 + case)
        poisoned = case % 3 == 0
        f = random_flat(rng, rng.randrange(2, 8), rng.randrange(3, 30), rng.randrange(0, 3), rng.randrange(5, 150),
                        zero_rows=int(poisoned), nan_rows=int(poisoned and case % 2 == 0))
        args = (f["node_id"], f["node_type"], f["src"], f["dst"], f["etype"], f["w"])
        rg, og = RF.ReferenceGraph(*args), O.OracleGraph(*args)
        assert rg.build() == 0 and og.build() == 0
        for a, b in zip(rg.csr(), og.csr()):
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), case
        with_links = np.flatnonzero(np.bincount(np.asarray(f["src"], np.int64), minlength=rg.n) > 0)
        seed = int(with_links[rng.randrange(len(with_links))]) if len(with_links) else 0
        for it in (1, 3, 11):
            want, _ = rg.run(seed, c, n_iter=it)
            for literal in (True, False):
                assert np.array_equal(bits(og.run(seed, c, n_iter=it, literal=literal)[0]), bits(want)), (case, it, literal)
        if len(with_links):
            wi, ws = rg.recommend(seed, 0.15, 4)
            gi, gs = og.recommend(seed, 0.15, 4)
            assert wi.tolist() == gi.tolist() and np.array_equal(bits(ws), bits(gs)), case
            compared += 1
        rg.close(); og.close()
    assert compared >= 75


@needs_ref
def test_reference_exceptions_of_the_boundary():
    """IndexOutOfRangeException for a link target >= N (Model.cs:87), KeyNotFoundException before buildGraph()
    (Model.cs:79) and for a seed without an `edges` entry (Recommender.cs:21); an entry that exists but is empty is fine."""
    ids, types = [1, 2, 3], [1, 2, 2]
    rg = RF.ReferenceGraph(ids, types, [0, 1], [1, 7], [1, 1], [1.0, 1.0])
    rank = np.empty(3)
    assert RF.lib().ref_model_run(rg._h, 0, 0.15, 0, 1, 0.0, 0, rank.ctypes.data, None) == RF.REF_E_NOT_BUILT
    assert rg.build() == 0
    with pytest.raises(IndexError):
        rg.run(0, 0.15, n_iter=2)
    og = O.OracleGraph(ids, types, [0, 1], [1, 7], [1, 1], [1.0, 1.0])
    assert og.build() == O.ORC_E_BADINDEX or og.build() == 0          # the oracle may refuse the link at build time instead
    rg = RF.ReferenceGraph(ids, types, [0], [1], [1], [1.0], has_entry=[1, 0, 1])
    assert rg.build() == 0
    assert rg.recommend(2, 0.15, 3)[0].tolist() == [3, 2]          # empty entry: nothing liked, both items are candidates
    assert rg.recommend(0, 0.15, 3)[0].tolist() == [3]             # item 2 is liked by user 0: excluded
    with pytest.raises(KeyError):
        rg.recommend(1, 0.15, 3)


@needs_ref
def test_oracle_matches_the_reference_on_c1():
    """BASELINE configs[0]: the reference's CPU path as written, on the C1-shaped synthetic ego network."""
    s = O.synth_generate(C1_SPEC)
    args = (s["node_id"], s["node_type"], s["src"], s["dst"], s["etype"], s["w"])
    rg, og = RF.ReferenceGraph(*args), O.OracleGraph(*args)
    assert rg.build() == 0 and og.build() == 0
    for a, b in zip(rg.csr(), og.csr()):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    seed = int(np.flatnonzero(np.bincount(s["src"], minlength=og.n) > 0)[0])
    c = O.widen_float(0.15)
    want, _ = rg.run(seed, c, n_iter=4)
    assert np.array_equal(bits(og.run(seed, c, n_iter=4, literal=True)[0]), bits(want))
    assert np.array_equal(bits(og.run(seed, c, n_iter=4, literal=False)[0]), bits(want))
    wi, ws = rg.recommend(seed, 0.15, 3)
    gi, gs = og.recommend(seed, 0.15, 3)
    assert wi.tolist() == gi.tolist() and np.array_equal(bits(ws), bits(gs))
    assert len(wi) > 8000                                  # the whole ranking, as Experiment.cs:123 walks it


# ---------------------------------------------------------------------------------------------- the fixture the reference wrote
def test_reference_fixture():
    """tests/golden/ref_mid.json was written by the transliterated reference (oracle/make_golden_ref.py).  The oracle must
    reproduce it bit for bit; where libref.so is present the reference must still do so, too."""
    import make_golden_ref as MG
    with open(os.path.join(GOLDEN, "ref_mid.json")) as f:
        want = json_load(f)
    made_from = want.pop("made_from")
    assert "cs2cpp" in want.pop("made_by") and len(want["seeds"]) == 2 and want["csr"]["dangling_rows"] > 0
    assert MG.generate(O.OracleGraph) == want
    if HAVE_REF:
        assert MG.generate(RF.ReferenceGraph) == want
        now = RF.source_hashes()                        # the fixture and libref.so come from the same reference files
        assert all(now[rel] == h for rel, h in made_from.items()) and len(made_from) == 3


def json_load(f):
    import json
    return json.load(f)


# ---------------------------------------------------------------------------------------------- the callers: DataLoader, Experiment
def _ego(tmp_path, seed, **kw):
    import ego_db
    from recommendersystems_b200.ingest import load_ego_network
    tables = ego_db.random_tables(random.Random(seed), **kw)
    db = ego_db.write_sqlite(str(tmp_path / "1000.sqlite"), tables)
    links, net = load_ego_network(db)                                   # the product's ingest (N4), every relation loaded
    return RF.ReferenceDb(db, tables), links, net


def _by_id(L):
    """{source id: [(target id, type, weight) in insertion order]} and {node id: type} of a flattened link list."""
    ids = np.asarray(L["node_id"])
    per = {}
    for s, d, t, w in zip(np.asarray(L["src"]).tolist(), np.asarray(L["dst"]).tolist(), np.asarray(L["etype"]).tolist(),
                          np.asarray(L["w"]).tolist()):
        per.setdefault(int(ids[s]), []).append((int(ids[d]), t, w))
    return per, {int(i): int(t) for i, t in zip(ids, np.asarray(L["node_type"]))}


@needs_ref
@pytest.mark.parametrize("seed", [11, 12])
def test_k_fold_loop_matches_the_reference(tmp_path, seed):
    """Experiment.runKFoldCrossValidation (its own loop, compiled from Experiment.cs) against ingest + restated fold loop:
    HIT identical, AVGPRECISION to 1e-12, for every methodology."""
    import experiment_ref as R
    from recommendersystems_b200.experiment import result_row
    rdb, links, net = _ego(tmp_path, seed)
    n_folds, n_iter = 10, 6
    assert rdb.validation(n_folds) == (True, net.like_count(), net.friends_count()) and net.is_valid(n_folds)
    for m in range(16):
        ref = rdb.experiment(n_folds, n_iter, m)
        hits, sum_ap, folds = R.run_k_fold(links, m, n_folds, n_iter)
        assert ref["valid"] and ref["cnt_likes"] == net.like_count()
        assert ref["hit"] == hits and hits < net.like_count(), m          # some held-out tweets are liked by nobody else: no hit
        assert abs(ref["avg_precision_sum"] - sum_ap) <= 1e-12 * max(1.0, sum_ap), m
        row = result_row(net.ego_id, m, n_folds, n_iter, hits, net.like_count(), sum_ap).split("\t")
        assert row[:6] == ["1000", str(m), "10", "6", str(int(ref["hit"])), str(ref["cnt_likes"])]      # Experiment.cs:144-152
        assert abs(float(row[6]) - ref["avg_precision_sum"] / n_folds) <= 1e-14


@needs_ref
def test_invalid_ego_network_returns_like_the_reference(tmp_path):
    rdb, links, net = _ego(tmp_path, 5, n_friends=20, ego_likes=60)         # < 50 friends: Experiment.cs:72-74 returns
    assert rdb.validation(10) == (False, net.like_count(), net.friends_count()) and not net.is_valid(10)
    assert rdb.experiment(10, 3, 8)["valid"] is False
    rdb, links, net = _ego(tmp_path, 6, ego_likes=30)                        # < 50 likes
    assert rdb.validation(10)[0] is False and not net.is_valid(10)
    assert rdb.validation(40)[0] is False and rdb.validation(31)[0] is False  # cntLikes < nFolds


@needs_ref
@pytest.mark.parametrize("methodology,fold", [(8, 3), (0, 0), (4, 9), (2, 5), (15, 7), (13, 9), (9, 1)])
def test_loader_graph_matches_ingest_hold_out_and_masks(tmp_path, methodology, fold):
    """`new DataLoader(db, 10).graphConfiguration(methodology, fold)` (compiled from DataLoader.cs) against the product's way to
    the same graph: ONE ingest with every relation, then the fold's hold-out and the methodology's feature set.  Same test
    set; same nodes, except those the reference never creates (third-party users of an unloaded FOLLOW feature -- isolated
    here -- and held-out tweets nobody else likes -- isolated and retyped here); the same link list under every source, in
    the same order, with the same weights (MENTION weights through libm's log on both sides)."""
    import experiment_ref as R
    rdb, links, net = _ego(tmp_path, 21)
    ref = rdb.load(10, methodology, fold)
    held, test = R.hold_out(links, [0], 10, fold)
    cfg = R.apply_methodology(held, methodology)
    if methodology in R.RETYPE_FRIENDSHIP:                                # Experiment.cs:84-101 runs after the loader
        ref["etype"] = np.where(ref["etype"] == R.FRIENDSHIP, 0, ref["etype"])
    assert sorted(ref["test_ids"].tolist()) == test[0].tolist() and len(test[0]) == 7
    a, ta = _by_id(ref)
    b, tb = _by_id(cfg)
    assert a == b                                                         # link lists by source id, insertion order, weights
    assert set(ta) <= set(tb) and all(tb[i] == t for i, t in ta.items())
    extra = set(tb) - set(ta)
    assert all(i not in b and not any(i == d for ls in b.values() for (d, _, _) in ls) for i in extra)    # isolated
    assert all(tb[i] in (0, 3) for i in extra)                            # retyped orphans, unloaded third-party users
    # DataLoader creates an `edges` entry only by adding a link: "no entry" and "no link" are the same thing, which is what the
    # flattened input of the C ABI assumes (KeyNotFoundException for a seed without links, Recommender.cs:21)
    assert np.array_equal(ref["has_entry"].astype(bool), np.bincount(ref["src"], minlength=len(ref["node_id"])) > 0)
    # node order: the reference numbers members first, then tweets as the likes are walked; ingest does the same with all
    # of the ego's likes -- the members' prefix is identical
    k = int((np.asarray(ref["node_type"]) == 1).sum())
    assert ref["node_id"][:k].tolist() == links["node_id"][:k].tolist()


@needs_ref
@pytest.mark.parametrize("trial", range(10))
def test_loader_fuzz(tmp_path, trial):
    """Small and odd ego networks (few friends, 0..15 likes of the ego, 1..4 folds, self-follows, mentions of non-members, duplicate
    rows): checkEgoNetworkValidation's counts, the test set and the graph of EVERY methodology against the reference's DataLoader."""
    import ego_db
    import experiment_ref as R
    from recommendersystems_b200.ingest import load_ego_network
    rng = random.Random(5000 + trial)
    n_friends = rng.choice([8, 12, 20])
    tables = ego_db.random_tables(rng, n_friends=n_friends, n_nonfriend_followees=rng.choice([0, 3]), n_third=rng.choice([2, 6]),
                                  n_tweets=rng.choice([15, 60]), ego_likes=rng.choice([0, 3, 9, 15]))
    if trial % 2:
        ego_db.add_nasty_rows(tables, rng)
    db = ego_db.write_sqlite(str(tmp_path / "1000.sqlite"), tables)
    rdb = RF.ReferenceDb(db, tables)
    links, net = load_ego_network(db)
    n_folds = rng.choice([1, 2, 4])
    assert rdb.validation(n_folds) == (net.is_valid(n_folds), net.like_count(), net.friends_count())
    for methodology in range(16):
        fold = rng.randrange(n_folds)
        ref = rdb.load(n_folds, methodology, fold)
        held, test = R.hold_out(links, [0], n_folds, fold)
        cfg = R.apply_methodology(held, methodology)
        if methodology in R.RETYPE_FRIENDSHIP:
            ref["etype"] = np.where(ref["etype"] == R.FRIENDSHIP, 0, ref["etype"])
        assert sorted(ref["test_ids"].tolist()) == test[0].tolist(), (methodology, fold)
        a, ta = _by_id(ref)
        b, tb = _by_id(cfg)
        assert a == b, (methodology, fold)
        assert set(ta) <= set(tb) and all(tb[i] == t for i, t in ta.items()), (methodology, fold)
