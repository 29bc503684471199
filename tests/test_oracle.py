"""CPU tests: the C++ oracle against the golden fixtures (pure-Python literal restatement of the reference),
literal vs collapsed bit-identity, invariants, and the synthetic generator (C++ vs pure-Python)."""
import numpy as np
import pytest

from conftest import C1_SPEC, bits, load_golden, unhex
import oracle as O
import rwr_literal as R

CASES = ["kat_8c", "small_a", "small_b"]


def make(g):
    i = g["input"]
    og = O.OracleGraph(i["node_id"], i["node_type"], i["src"], i["dst"], i["etype"], i["w"])
    assert og.build() == 0
    return og


@pytest.mark.parametrize("name", CASES)
def test_csr_bit_exact(name):
    g = load_golden(name)
    og = make(g)
    rp, col, val = og.csr()
    assert rp.tolist() == g["csr"]["row_ptr"]
    assert col.tolist() == g["csr"]["col"]
    assert np.array_equal(bits(val), bits(unhex(g["csr"]["val"])))
    assert og.build() == O.ORC_E_ALREADY_BUILT          # Dictionary.Add twice -> ArgumentException


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("literal", [True, False])
def test_ranks_bit_exact(name, literal):
    g = load_golden(name)
    og = make(g)
    c = float.fromhex(g["damping_double"])
    assert c == O.widen_float(g["damping_float"]) == 0.15000000596046448
    for e in g["seeds"]:
        for n, want in e["ranks"].items():
            rank, it = og.run(e["seed"], c, n_iter=int(n), literal=literal)
            assert it == int(n)
            assert np.array_equal(bits(rank), bits(unhex(want))), (name, e["seed"], n)
        for thr, want in e["thresholds"].items():
            if want["iters"] is None:
                rank, it = og.run(e["seed"], c, default_threshold=True, literal=literal, max_iter=300)
                assert it == 300                          # no bitwise fixed point: the reference never returns
                continue
            if thr == "default":
                rank, it = og.run(e["seed"], c, default_threshold=True, literal=literal)
            else:
                rank, it = og.run(e["seed"], c, threshold=float(thr), literal=literal)
            assert it == want["iters"], (name, e["seed"], thr)
            assert np.array_equal(bits(rank), bits(unhex(want["rank"])))


@pytest.mark.parametrize("name", CASES)
def test_uniform_ctor(name):
    g = load_golden(name)
    og = make(g)
    c = float.fromhex(g["damping_double"])
    for n, want in g["uniform"].items():
        rank, _ = og.run(-1, c, n_iter=int(n))
        assert np.array_equal(bits(rank), bits(unhex(want)))


@pytest.mark.parametrize("name", CASES)
def test_recommendation(name):
    g = load_golden(name)
    og = make(g)
    for e in g["seeds"]:
        if e["recommendation"] == "KeyNotFoundException":
            with pytest.raises(KeyError):
                og.recommend(e["seed"], g["damping_float"], 10)
            continue
        rec = e["recommendation"]
        ids, sc = og.recommend(e["seed"], g["damping_float"], rec["n_iter"])
        assert ids.tolist() == rec["ids"]
        assert np.array_equal(bits(sc), bits(unhex(rec["scores"])))
        for k, want in e["top"].items():
            ids, sc = og.recommend(e["seed"], g["damping_float"], rec["n_iter"], top_n=int(k))
            assert ids.tolist() == want["ids"]
            assert np.array_equal(bits(sc), bits(unhex(want["scores"])))
        # Recommender.cs:47: `Count == topN` never fires for topN <= 0 -> whole list
        ids0, _ = og.recommend(e["seed"], g["damping_float"], rec["n_iter"], top_n=0)
        assert ids0.tolist() == rec["ids"]


def test_kat_8c_values():
    """The hand-derived known-answer vector of SURVEY.md section 8c, spelled out."""
    g = load_golden("kat_8c")
    og = make(g)
    rp, col, val = og.csr()
    assert np.diff(rp).tolist() == [2, 6, 3, 1, 0, 0, 0, 1]
    assert col.tolist() == [1, 2, 0, 2, 3, 4, 2, 0, 0, 1, 1, 1, 0]
    assert val[0].hex() == "0x1.0000000000000p-1" and val[2].hex() == "0x1.745d1745d1746p-3"
    assert val[7].hex() == "0x1.745d1745d1746p-4" and val[8].hex() == "0x1.5555555555555p-2"
    c = O.widen_float(0.15)
    r1, _ = og.run(0, c, n_iter=1)
    assert [x.hex() for x in r1[:3]] == ["0x1.3333340000000p+0", "0x1.b333330000000p+1", "0x1.b333330000000p+1"]
    r10, _ = og.run(0, c, n_iter=10)
    assert r10.tolist() == [2.685872555693182, 2.578891093966091, 1.938306208045018, 0.39846507114785384,
                            0.39846507114785384, 0.0, 0.0, 0.0]
    ids, sc = og.recommend(0, 0.15, 10)
    assert ids.tolist() == [5004, 5003, 5006, 5005] and sc.tolist()[2:] == [0.0, 0.0]
    for thr, n in [(1e-3, 11), (1e-6, 23), (1e-9, 35), (1e-12, 47)]:
        assert og.run(0, c, threshold=thr)[1] == n
    assert og.run(0, c, default_threshold=True)[1] == 61


def test_python_literal_matches_cpp_on_fresh_random_graphs():
    """Differential: a new random graph per run seed, python literal vs C++ literal vs C++ collapsed."""
    import random
    import make_golden as MG
    rng = random.Random(7)
    for _ in range(3):
        nodes, edges = MG.random_case(rng, 9, 20, 3, 90)
        flat = MG.flat_of(nodes, edges)
        og = O.OracleGraph(flat["node_id"], flat["node_type"], flat["src"], flat["dst"], flat["etype"], unhex(flat["w"]))
        assert og.build() == 0
        pg = R.Graph(nodes, edges)
        pg.buildGraph()
        c = O.widen_float(0.15)
        for seed in (0, 4, 10):
            m = R.Model(pg, c, seed)
            m.run(7)
            a, _ = og.run(seed, c, n_iter=7, literal=True)
            b, _ = og.run(seed, c, n_iter=7, literal=False)
            assert np.array_equal(bits(a), bits(np.array(m.rank))) and np.array_equal(bits(a), bits(b))


def test_invariants_c1():
    """Mass conservation, non-negativity, stochastic rows on the C1-shaped synthetic graph."""
    s = O.synth_generate(C1_SPEC)
    og = O.OracleGraph(s["node_id"], s["node_type"], s["src"], s["dst"], s["etype"], s["w"])
    assert og.build() == 0
    rp, col, val = og.csr()
    n = og.n
    deg = np.diff(rp)
    rows = np.repeat(np.arange(n), deg)
    sums = np.bincount(rows, weights=val, minlength=n)
    nz = deg > 0
    assert np.all(np.abs(sums[nz] - 1.0) <= 1e-15 * np.maximum(deg[nz], 1) * 4)
    assert (deg == 0).sum() >= n // 100                   # >= 1 % dangling rows
    assert np.any((s["etype"] == 4) & (s["w"] != 1.0))    # fractional MENTION weights present
    assert np.any(s["etype"] == 0)                        # UNDEFINED links present
    seed = int(np.flatnonzero(np.bincount(s["src"], minlength=n) > 0)[0])
    for it in (1, 5, 20):
        r, _ = og.run(seed, O.widen_float(0.15), n_iter=it)
        assert abs(r.sum() - n) < 1e-9 * n and r.min() >= 0.0
    # literal O(N^2) form == collapsed form on C1 (the reference's CPU path as written)
    a, _ = og.run(seed, O.widen_float(0.15), n_iter=3, literal=True)
    b, _ = og.run(seed, O.widen_float(0.15), n_iter=3, literal=False)
    assert np.array_equal(bits(a), bits(b))


def test_synth_generator_matches_python():
    g = load_golden("synth_tiny")
    s = O.synth_generate(g["spec"])
    w = g["graph"]
    for k in ("node_id", "node_type", "src", "dst", "etype"):
        assert s[k].tolist() == w[k], k
    assert np.array_equal(bits(s["w"]), bits(w["w"]))
    # canonical order: sorted by (src, class, dst), unique
    assert np.all(np.diff(s["src"]) >= 0)


def test_evaluate_matches_python():
    rec = [(50, .9), (40, .8), (30, .7), (20, .6), (10, .5)]
    for test in ({40, 10}, set(), {50}, {99}):
        want = R.evaluate(rec, test)
        got = O.evaluate([p[0] for p in rec], sorted(test))
        assert got[0] == want[0] and got[1] == want[1]


@pytest.mark.parametrize("name", CASES)
def test_fixed_point_is_the_closed_form(name):
    """An independent check of the iteration itself (not of its rounding): with A = (1-c) W^T restricted to nodes that
    have out-links, Model.cs:76-100 is r <- A r + q * 1^T (I - A) r, so the limit is proportional to (I - A)^-1 q and
    carries the initial mass N (Model.cs:41). 300 iterations leave 0.85^300 of the transient."""
    g = load_golden(name)
    og = make(g)
    rp, col, val = og.csr()
    n = len(rp) - 1
    c = float.fromhex(g["damping_double"])
    A = np.zeros((n, n))
    for i in range(n):
        for e in range(rp[i], rp[i + 1]):
            A[col[e], i] += (1 - c) * val[e]
    for e in g["seeds"]:
        q = np.zeros(n)
        q[e["seed"]] = 1.0
        r = np.linalg.solve(np.eye(n) - A, q)
        r *= n / r.sum()
        rank, _ = og.run(e["seed"], c, n_iter=300)
        assert abs(rank.sum() - n) < 1e-9 * n
        assert np.allclose(rank, r, rtol=1e-9, atol=1e-12), (name, e["seed"])
