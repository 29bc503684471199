"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI via the host mirror of
the reference API, against the golden fixtures and the CPU oracle on the same inputs.

Bars (BASELINE.json north_star):
  CSR structure, degrees and normalised weights ........ bit-exact
  iteration count under the same threshold ............. identical
  FP64 per-node scores ................................. <= 1e-12 relative (zeros exactly zero)
  FP32 scores .......................................... <= 1e-6 L1 on sum=1-normalised scores
  top-k / full ranking ................................. identical id lists (ties inside the tolerance excepted)
"""
import numpy as np
import pytest

from conftest import C1_SPEC, bits, load_golden, unhex

pytestmark = pytest.mark.gpu

import oracle as O
import recommendersystems_b200 as rs
from recommendersystems_b200 import _native as N
from recommendersystems_b200.rwr import run_fixed, run_threshold

REL = 1e-12            # FP64 tolerance stated by north_star
L1_FP32 = 1e-6         # FP32 tolerance stated by north_star
C015 = rs.widen_float(0.15)

CASES = ["kat_8c", "small_a", "small_b"]
OPTS = [dict(), dict(relabel=False), dict(hub_entries=0), dict(relabel=False, hub_entries=0, layout=N.LAYOUT_VALUED),
        dict(layout=N.LAYOUT_VALUED), dict(hot_min_degree=1), dict(hub_entries=64)]


def assert_close_fp64(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    zero = want == 0.0
    assert np.all(got[zero] == 0.0), f"{what}: nodes the reference leaves at exactly 0 must stay 0"
    rel = np.abs(got[~zero] - want[~zero]) / np.abs(want[~zero])
    assert rel.size == 0 or rel.max() <= REL, f"{what}: max rel err {rel.max():.3e}"


def assert_close_fp32(got, want, what=""):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    l1 = np.abs(got / got.sum() - want / want.sum()).sum()
    assert l1 <= L1_FP32, f"{what}: normalised L1 {l1:.3e}"


def same_ranking(ids, scores, want_ids, want_scores):
    """Identical lists, except that positions whose reference scores agree within REL may be permuted."""
    ids, want_ids = list(ids), list(want_ids)
    assert len(ids) == len(want_ids)
    if ids == want_ids:
        return
    want_scores = np.asarray(want_scores)
    i = 0
    while i < len(ids):
        j = i + 1
        while j < len(ids) and abs(want_scores[j] - want_scores[i]) <= REL * max(abs(want_scores[i]), 1e-300):
            j += 1
        assert sorted(ids[i:j]) == sorted(want_ids[i:j]), f"ranking differs outside a tie group at {i}"
        i = j


def gpu_graph(inp, **opts):
    g = rs.Graph.from_arrays(inp["node_id"], inp["node_type"], inp["src"], inp["dst"], inp["etype"], inp["w"], **opts)
    g.buildGraph()
    return g


def oracle_graph(inp):
    og = O.OracleGraph(inp["node_id"], inp["node_type"], inp["src"], inp["dst"], inp["etype"], inp["w"])
    assert og.build() == 0
    return og


# ------------------------------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("opts", OPTS)
def test_golden_csr_bit_exact(name, opts):
    g = load_golden(name)
    gg = gpu_graph(g["input"], **opts)
    rp, col, val = gg.csr()
    assert rp.tolist() == g["csr"]["row_ptr"]
    assert col.tolist() == g["csr"]["col"]
    assert np.array_equal(bits(val), bits(unhex(g["csr"]["val"])))
    assert gg.degrees().tolist() == g["outdeg"]
    with pytest.raises(ValueError):                      # ArgumentException, Graph.cs:86
        gg.buildGraph()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("opts", OPTS)
def test_golden_ranks_fp64(name, opts):
    g = load_golden(name)
    gg = gpu_graph(g["input"], **opts)
    for e in g["seeds"]:
        for n, want in e["ranks"].items():
            m = rs.Model(gg, C015, e["seed"])
            m.run(int(n))
            assert_close_fp64(m.rank, unhex(want), f"{name} seed {e['seed']} iter {n}")
            assert abs(m.rank.sum() - gg.size()) <= 1e-12 * gg.size()      # mass conservation
    # zero iterations == constructor state
    m = rs.Model(gg, C015, g["seeds"][0]["seed"])
    m.run(0)
    want = np.zeros(gg.size()); want[g["seeds"][0]["seed"]] = gg.size()
    assert np.array_equal(m.rank, want)


@pytest.mark.parametrize("name", CASES)
def test_golden_threshold_iteration_counts(name):
    g = load_golden(name)
    gg = gpu_graph(g["input"])
    for e in g["seeds"]:
        for thr, want in e["thresholds"].items():
            if thr == "default":
                continue          # subnormal threshold: stops only at a bitwise fixed point, order dependent
            m = rs.Model(gg, C015, e["seed"])
            m.run(float(thr))
            assert m.nIterations == want["iters"], (name, e["seed"], thr)
            assert_close_fp64(m.rank, unhex(want["rank"]), f"{name} seed {e['seed']} thr {thr}")


def test_default_threshold_terminates_on_fixed_point():
    g = load_golden("kat_8c")
    gg = gpu_graph(g["input"])
    m = rs.Model(gg, C015, 5)          # dangling seed: rank never moves, residual is exactly 0 < subnormal threshold
    m.run()
    assert m.nIterations == 1
    m = rs.Model(gg, C015, 0)
    m.run(None, max_iter=500)
    assert 40 <= m.nIterations <= 500


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("opts", OPTS[:2])
def test_golden_recommendation(name, opts):
    g = load_golden(name)
    gg = gpu_graph(g["input"], **opts)
    rec = rs.Recommender(gg)
    for e in g["seeds"]:
        if e["recommendation"] == "KeyNotFoundException":
            with pytest.raises(KeyError):
                rec.Recommendation(e["seed"], 0.15, 10)
            with pytest.raises(KeyError):
                rec.Recommendation(e["seed"], 0.15, 10, 3)
            continue
        want = e["recommendation"]
        full = rec.Recommendation(e["seed"], 0.15, want["n_iter"])
        same_ranking([p[0] for p in full], [p[1] for p in full], want["ids"], unhex(want["scores"]))
        assert_close_fp64([p[1] for p in full], unhex(want["scores"]))
        for k, w in e["top"].items():
            top = rec.Recommendation(e["seed"], 0.15, want["n_iter"], int(k))
            same_ranking([p[0] for p in top], [p[1] for p in top], w["ids"], unhex(w["scores"]))
        whole = rec.Recommendation(e["seed"], 0.15, want["n_iter"], 0)      # Recommender.cs:47 quirk
        assert len(whole) == len(want["ids"])
        big = rec.Recommendation(e["seed"], 0.15, want["n_iter"], 10 ** 6)
        assert [p[0] for p in big] == [p[0] for p in full]
        k20 = rec.Recommendation(e["seed"], 0.15, want["n_iter"], 20)        # k > 16: truncated full ranking
        assert [p[0] for p in k20] == [p[0] for p in full][:20]


def test_fixture_written_by_the_reference():
    """tests/golden/ref_mid.json holds what the reference's OWN Graph.cs / Model.cs / Recommender.cs compute on an 820-node ego
    network (oracle/make_golden_ref.py, through oracle/cs2cpp.py): the CUDA path against the reference itself, no oracle between."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "ref_mid.json")) as f:
        want = json.load(f)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    for from_generator in (False, True):               # the CPU generator's links uploaded / the device generator's own
        if from_generator:
            gg = rs.Graph.synthetic(want["spec"])
            gg.buildGraph()
        else:
            gg = gpu_graph(O.synth_generate(want["spec"]))
        assert gg.size() == want["n_nodes"]
        rp, col, val = gg.csr()
        assert (len(col), int((np.diff(rp) == 0).sum())) == (want["csr"]["nnz"], want["csr"]["dangling_rows"])
        assert sha(rp.astype(np.int64)) == want["csr"]["row_ptr_sha256"] and sha(col.astype(np.int32)) == want["csr"]["col_sha256"]
        assert sha(val.astype(np.float64)) == want["csr"]["val_sha256"]                      # A1-A3 bit for bit
        rec = rs.Recommender(gg)
        for e in want["seeds"]:
            for n, ranks in e["ranks"].items():
                m = rs.Model(gg, C015, e["seed"])
                m.run(int(n))
                assert_close_fp64(m.rank, unhex(ranks), f"seed {e['seed']} iter {n}")
            for thr, w in e["thresholds"].items():
                m = rs.Model(gg, C015, e["seed"])
                m.run(float(thr))
                assert m.nIterations == w["iters"], (e["seed"], thr)
            w = e["recommendation"]
            full = rec.Recommendation(e["seed"], want["damping_float"], w["n_iter"])
            assert len(full) == len(w["ids"])
            top10 = unhex(w["top10_scores"])
            assert_close_fp64([p[1] for p in full[:10]], top10)
            same_ranking([p[0] for p in full[:10]], [p[1] for p in full[:10]], w["ids"][:10], top10)
            top = rec.Recommendation(e["seed"], want["damping_float"], w["n_iter"], 10)
            same_ranking([p[0] for p in top], [p[1] for p in top], w["ids"][:10], top10)
            # beyond the stored scores: the same SET of candidates, and the long tail of exact zeros in id-descending order
            assert sorted(p[0] for p in full) == sorted(w["ids"])
            zeros = [p[0] for p in full if p[1] == 0.0]
            assert zeros == w["ids"][len(w["ids"]) - len(zeros):] and zeros == sorted(zeros, reverse=True)
        gg.close()


def test_kat_8c_spelled_out():
    g = load_golden("kat_8c")
    gg = gpu_graph(g["input"])
    rec = rs.Recommender(gg).Recommendation(0, 0.15, 10)
    assert [p[0] for p in rec] == [5004, 5003, 5006, 5005]
    assert rec[2][1] == 0.0 and rec[3][1] == 0.0
    assert abs(rec[0][1] - 0.39846507114785384) <= REL and rec[0][1] == rec[1][1]


@pytest.mark.parametrize("name", CASES)
def test_golden_uniform_ctor(name):
    g = load_golden(name)
    gg = gpu_graph(g["input"])
    for n, want in g["uniform"].items():
        m = rs.Model(gg, C015)
        m.run(int(n))
        want = unhex(want)
        assert np.abs(m.rank - want).max() <= 1e-12 * np.abs(want).max()


def test_golden_fp32():
    for name in CASES:
        g = load_golden(name)
        gg = gpu_graph(g["input"])
        for e in g["seeds"]:
            n = max(int(k) for k in e["ranks"])
            m = rs.Model(gg, C015, e["seed"], precision=rs.FP32)
            m.run(n)
            assert_close_fp32(m.rank, unhex(e["ranks"][str(n)]), f"{name} seed {e['seed']}")


def test_error_codes():
    g = load_golden("kat_8c")
    inp = dict(g["input"])
    gg = rs.Graph.from_arrays(inp["node_id"], inp["node_type"], inp["src"], inp["dst"], inp["etype"], inp["w"])
    with pytest.raises(KeyError):                      # Model before buildGraph: KeyNotFound at Model.cs:79
        run_fixed(gg, [0], C015, 1)
    et = list(inp["etype"])
    i_exp, i_und = et.index(1), et.index(0)
    bad = dict(inp); bad["dst"] = list(inp["dst"]); bad["dst"][i_exp] = 99
    g2 = rs.Graph.from_arrays(bad["node_id"], bad["node_type"], bad["src"], bad["dst"], bad["etype"], bad["w"])
    with pytest.raises(IndexError):                    # IndexOutOfRange at Model.cs:87
        g2.buildGraph()
    bad["dst"][i_exp] = inp["dst"][i_exp]; bad["dst"][i_und] = 99   # the UNDEFINED link: never dereferenced by the reference
    g3 = rs.Graph.from_arrays(bad["node_id"], bad["node_type"], bad["src"], bad["dst"], bad["etype"], bad["w"])
    g3.buildGraph()
    gg.buildGraph()
    with pytest.raises(KeyError):
        run_fixed(gg, [8], C015, 1)
    with pytest.raises(rs.RwrError):
        run_fixed(gg, [0], C015, 1, precision=7)


def test_empty_and_degenerate_graphs():
    g0 = rs.Graph.from_arrays([], [], [], [], [], [])
    g0.buildGraph()
    assert g0.size() == 0 and g0.csr()[0].tolist() == [0]
    # nodes without any link: every row dangling, rank stays put
    g1 = rs.Graph.from_arrays([10, 20, 30], [1, 2, 2], [], [], [], [])
    g1.buildGraph()
    m = rs.Model(g1, C015, 1); m.run(5)
    assert m.rank.tolist() == [0.0, 3.0, 0.0]
    # a row whose weights sum to 0 silently yields NaN (Graph.cs:81) -- no error
    g2 = rs.Graph.from_arrays([10, 20], [1, 2], [0, 0], [1, 1], [1, 5], [1.0, -1.0])
    g2.buildGraph()
    assert np.isinf(g2.csr()[2]).all()
    g3 = rs.Graph.from_arrays([10, 20], [1, 2], [0], [1], [1], [0.0])
    g3.buildGraph()
    assert np.isnan(g3.csr()[2]).all()


@pytest.mark.parametrize("case", [4, 5])
def test_nan_and_inf_rows_poison_every_node_like_the_reference(case):
    """A row whose weights sum to 0 has NaN / Inf weights (Graph.cs:81).  The reference's restart loops add `x * restart[r]` to
    EVERY node (Model.cs:92-93, :96-97): once a rank is NaN or Inf, `x * 0` makes the whole next vector NaN.  The oracle does
    exactly that (held to the reference's compiled sources in tests/test_reference_pin.py, same graphs); so must both device paths."""
    import random
    from random_graphs import random_flat
    rng = random.Random(20260200 + case)
    f = random_flat(rng, rng.randrange(3, 12), rng.randrange(5, 60), rng.randrange(0, 4), rng.randrange(20, 400), zero_rows=1, nan_rows=1)
    gg, og = gpu_graph(f), oracle_graph(f)
    rp, col, val = gg.csr()
    orp, ocol, oval = og.csr()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol) and np.array_equal(bits(val), bits(oval))
    assert np.isnan(oval).any() or np.isinf(oval).any()
    n = gg.size()
    has_links = np.bincount(np.asarray(f["src"], np.int64), minlength=n) > 0
    seeds = [s for s in sorted(set([0, n // 2, n - 1] + [rng.randrange(n) for _ in range(3)])) if has_links[s]]
    poisoned_runs = 0
    for seed in seeds:
        for it in (1, 2, 3, 8):
            want, _ = og.run(seed, C015, n_iter=it)
            m = rs.Model(gg, C015, seed)
            m.run(it)
            nan = np.isnan(want)
            assert np.array_equal(np.isnan(m.rank), nan), (seed, it)
            assert np.array_equal(np.isinf(m.rank), np.isinf(want)), (seed, it)
            ok = np.isfinite(want)
            assert_close_fp64(m.rank[ok], want[ok], f"seed {seed} iter {it}")
            poisoned_runs += int(nan.all())
    assert poisoned_runs > 0                                # some run ends with every score NaN, as in the reference
    # rankings: NaN scores sort below every number and tie among themselves -> id descending (Double.CompareTo, Recommender.cs:34-38)
    rec = rs.Recommender(gg)
    ids_b, sc_b, cnt_b = rec.RecommendationBatch(seeds, 0.15, 6, 10)           # the batched (SpMM) path
    for i, seed in enumerate(seeds):
        wi, ws = og.recommend(seed, 0.15, 6)
        full = rec.Recommendation(seed, 0.15, 6)
        assert [p[0] for p in full] == wi.tolist() and np.array_equal(np.isnan([p[1] for p in full]), np.isnan(ws)), seed
        k = min(10, len(wi))
        assert cnt_b[i] == k and ids_b[i, :k].tolist() == wi[:k].tolist() and np.array_equal(np.isnan(sc_b[i, :k]), np.isnan(ws[:k])), seed
    gg.close()


def test_unsorted_input_is_grouped_stably():
    g = load_golden("small_b")
    inp = g["input"]
    rng = np.random.default_rng(3)
    src = np.asarray(inp["src"])
    # interleave links of different sources but keep every source's own order
    perm = rng.permutation(len(src))
    slots = np.argsort(src[perm], kind="stable")            # output slots grouped by source, ascending
    order = np.empty(len(src), np.int64)
    order[slots] = np.argsort(src, kind="stable")           # ... receive that source's links in their own order
    sh = {k: np.asarray(inp[k])[order] for k in ("src", "dst", "etype", "w")}
    gg = rs.Graph.from_arrays(inp["node_id"], inp["node_type"], sh["src"], sh["dst"], sh["etype"], sh["w"])
    gg.buildGraph()
    rp, col, val = gg.csr()
    assert rp.tolist() == g["csr"]["row_ptr"] and col.tolist() == g["csr"]["col"]
    assert np.array_equal(bits(val), bits(unhex(g["csr"]["val"])))


# ------------------------------------------------------------------------------------------ synthetic generator
def test_synth_generator_bit_exact_tiny():
    g = load_golden("synth_tiny")
    gg = rs.Graph.synthetic(g["spec"])
    out = gg.export_links()
    w = g["graph"]
    for k in ("node_id", "node_type", "src", "dst", "etype"):
        assert out[k].tolist() == w[k], k
    assert np.array_equal(bits(out["w"]), bits(w["w"]))


@pytest.fixture(scope="module")
def c1():
    cpu = O.synth_generate(C1_SPEC)
    og = oracle_graph(cpu)
    gg = rs.Graph.synthetic(C1_SPEC)
    return cpu, og, gg


def test_c1_generator_matches_oracle(c1):
    cpu, og, gg = c1
    out = gg.export_links()
    for k in ("node_id", "node_type", "src", "dst", "etype"):
        assert np.array_equal(out[k], cpu[k]), k
    assert np.array_equal(bits(out["w"]), bits(cpu["w"]))
    assert 80_000 <= len(cpu["src"]) <= 120_000 and og.n == 10_200


def test_c1_csr_and_scores(c1):
    cpu, og, gg = c1
    if not gg.info().built:
        gg.buildGraph()
    rp, col, val = gg.csr()
    orp, ocol, oval = og.csr()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol) and np.array_equal(bits(val), bits(oval))
    info = gg.info()
    assert info.layout == N.LAYOUT_VALUED and info.n_dangling == int((np.diff(orp) == 0).sum())
    raw_deg = np.bincount(cpu["src"], minlength=og.n)
    seeds = [int(s) for s in np.flatnonzero((raw_deg > 0) & (np.arange(og.n) < C1_SPEC["n_users"]))[:3]]
    seeds.append(int(np.argmax(raw_deg)))                         # the biggest hub
    for seed in seeds:
        for it in (1, 2, 5, 10, 20):
            want, _ = og.run(seed, C015, n_iter=it)
            m = rs.Model(gg, C015, seed); m.run(it)
            assert_close_fp64(m.rank, want, f"C1 seed {seed} iter {it}")
        want, _ = og.run(seed, C015, n_iter=20)
        m32 = rs.Model(gg, C015, seed, precision=rs.FP32); m32.run(20)
        assert_close_fp32(m32.rank, want, f"C1 fp32 seed {seed}")
        ids, sc = og.recommend(seed, 0.15, 20, top_n=10)
        top = rs.Recommender(gg).Recommendation(seed, 0.15, 20, 10)
        same_ranking([p[0] for p in top], [p[1] for p in top], ids, sc)
        fids, fsc = og.recommend(seed, 0.15, 20)
        full = rs.Recommender(gg).Recommendation(seed, 0.15, 20)
        same_ranking([p[0] for p in full], [p[1] for p in full], fids, fsc)
        for thr in (1e-3 * og.n, 1e-6 * og.n, 1e-9 * og.n):
            _, want_it = og.run(seed, C015, threshold=thr)
            m = rs.Model(gg, C015, seed); m.run(thr)
            assert m.nIterations == want_it, (seed, thr)
    # batched request path
    bids, bsc, bcnt = rs.Recommender(gg).RecommendationBatch(seeds, 0.15, 20, 10)
    for i, seed in enumerate(seeds):
        ids, sc = og.recommend(seed, 0.15, 20, top_n=10)
        same_ranking(bids[i, :bcnt[i]], bsc[i, :bcnt[i]], ids, sc)


@pytest.mark.parametrize("opts", [dict(), dict(relabel=False), dict(hub_entries=0), dict(hub_entries=64), dict(layout=N.LAYOUT_VALUED)])
def test_unit_weight_graph_uses_index_layout(opts):
    """All raw weights 1.0 (the C2 shape): the row's weight folds into x, products stay bit-equal to the reference's."""
    spec = dict(C1_SPEC, n_mention=0, seed=77)
    cpu = O.synth_generate(spec)
    og = oracle_graph(cpu)
    gg = rs.Graph.synthetic(spec, **opts)
    gg.buildGraph()
    assert gg.info().layout == (N.LAYOUT_VALUED if opts.get("layout") == N.LAYOUT_VALUED else N.LAYOUT_INDEX)
    rp, col, val = gg.csr()
    orp, ocol, oval = og.csr()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol) and np.array_equal(bits(val), bits(oval))
    seed = int(np.argmax(np.bincount(cpu["src"], minlength=og.n)[:spec["n_users"]]))
    for it in (1, 3, 20):
        want, _ = og.run(seed, C015, n_iter=it)
        m = rs.Model(gg, C015, seed); m.run(it)
        assert_close_fp64(m.rank, want, f"unit seed {seed} iter {it} {opts}")
    m32 = rs.Model(gg, C015, seed, precision=rs.FP32); m32.run(20)
    assert_close_fp32(m32.rank, want)


def test_index_layout_refused_for_fractional_rows():
    gg = rs.Graph.synthetic(C1_SPEC, layout=N.LAYOUT_INDEX)
    with pytest.raises(rs.RwrError) as ei:
        gg.buildGraph()
    assert ei.value.code == N.RWR_E_UNSUPPORTED


# ------------------------------------------------------------------------------------------ hubs / chunk boundaries
def test_medium_graph_hubs_cross_chunks():
    """~1.2 M links, strong skew: hub rows span many merge-path chunks, lots of empty rows."""
    spec = dict(seed=5, n_users=20_000, n_items=150_000, n_third=5_000, authorship_per_mille=500, n_like=450_000,
                n_friend=120_000, n_follow=20_000, n_mention=3_000, undefined_per_mille=50, scramble=1, p1_byte=40)
    cpu = O.synth_generate(spec)
    og = oracle_graph(cpu)
    for opts in (dict(), dict(relabel=False, hub_entries=0), dict(hub_entries=256, hot_min_degree=3)):
        gg = rs.Graph.synthetic(spec, **opts)
        gg.buildGraph()
        info = gg.info()
        assert info.max_in_degree > 3 * 2044 and info.n_chunks > 300
        rp, col, val = gg.csr()
        orp, ocol, oval = og.csr()
        assert np.array_equal(rp, orp) and np.array_equal(col, ocol) and np.array_equal(bits(val), bits(oval))
        raw_deg = np.bincount(cpu["src"], minlength=og.n)
        for seed in (int(np.argmax(raw_deg)), int(np.flatnonzero(raw_deg[:spec["n_users"]] == 1)[0])):
            want, _ = og.run(seed, C015, n_iter=12)
            m = rs.Model(gg, C015, seed); m.run(12)
            assert_close_fp64(m.rank, want, f"medium seed {seed}")
            ids, sc = og.recommend(seed, 0.15, 12, top_n=10)
            top = rs.Recommender(gg).Recommendation(seed, 0.15, 12, 10)
            same_ranking([p[0] for p in top], [p[1] for p in top], ids, sc)
            thr = 1e-7 * og.n
            _, want_it = og.run(seed, C015, threshold=thr)
            m = rs.Model(gg, C015, seed); m.run(thr)
            assert m.nIterations == want_it


def test_determinism_and_full_size_properties():
    """Size-independent properties on a 22 M-link graph (a 1/9 scale C2): mass conservation, bit-identical reruns,
    FP32 vs FP64, ranking sortedness, index-only vs valued layouts bit-equal."""
    spec = dict(seed=11, n_users=120_000, n_items=1_100_000, n_third=0, authorship_per_mille=1000, n_like=7_800_000,
                n_friend=2_200_000, n_follow=0, n_mention=0, undefined_per_mille=0, scramble=1, p1_byte=61)
    gg = rs.Graph.synthetic(spec)
    gg.buildGraph()
    info = gg.info()
    n = info.n_nodes
    assert info.layout == N.LAYOUT_INDEX and info.nnz > 19_000_000
    deg = gg.degrees(raw=True)
    seed = int(np.flatnonzero(deg[:spec["n_users"]] > 50)[0])
    a = run_fixed(gg, [seed], C015, 20).scores(0)
    b = run_fixed(gg, [seed], C015, 20).scores(0)
    assert np.array_equal(bits(a), bits(b))
    assert abs(a.sum() - n) <= 1e-11 * n and a.min() >= 0.0
    f = run_fixed(gg, [seed], C015, 20, precision=rs.FP32).scores(0)
    assert_close_fp32(f, a)
    gv = rs.Graph.synthetic(spec, layout=N.LAYOUT_VALUED)
    gv.buildGraph()
    v = run_fixed(gv, [seed], C015, 20).scores(0)
    assert np.array_equal(bits(a), bits(v))            # folding the row weight into x keeps every product bit-equal
    res = run_fixed(gg, [seed], C015, 20)
    ids, sc, cnt = res.rank_all(0)
    assert cnt == len(ids) and np.all(np.diff(sc) <= 0)
    ties = np.diff(sc) == 0
    assert np.all(np.diff(ids)[ties] < 0)               # equal scores: id descending
    tid, tsc, tcnt = res.topk(10)
    assert tid[0, :tcnt[0]].tolist() == ids[:10].tolist()
    # spot parity against the oracle on this size (collapsed form, a few seconds of CPU)
    links = gg.export_links()
    og = oracle_graph(links)
    want, _ = og.run(seed, C015, n_iter=20)
    assert_close_fp64(a, want, "22M-link graph")
    oids, osc = og.recommend(seed, 0.15, 20, top_n=10)
    same_ranking(tid[0, :tcnt[0]], tsc[0, :tcnt[0]], oids, osc)


# ------------------------------------------------------------------------------------------ batched SpMM path (K8)
@pytest.mark.parametrize("unit", [False, True])
def test_batched_spmm_matches_oracle_and_single_column(unit):
    """rwr_recommend in SpMM tiles (8 FP64 / 16 FP32 columns): hubs crossing tiles, partial last tile, both layouts."""
    spec = dict(seed=9, n_users=6_000, n_items=50_000, n_third=1_000, authorship_per_mille=600, n_like=160_000,
                n_friend=40_000, n_follow=5_000, n_mention=0 if unit else 1_500, undefined_per_mille=30, scramble=1, p1_byte=45)
    cpu = O.synth_generate(spec)
    og = oracle_graph(cpu)
    gg = rs.Graph.synthetic(spec)
    gg.buildGraph()
    info = gg.info()
    assert info.layout == (N.LAYOUT_INDEX if unit else N.LAYOUT_VALUED) and info.max_in_degree > 2 * 1020
    raw_deg = np.bincount(cpu["src"], minlength=og.n)
    users = np.flatnonzero(raw_deg[:spec["n_users"]] > 0)
    seeds = [int(np.argmax(raw_deg))] + [int(u) for u in users[:: max(1, len(users) // 18)][:18]]      # 19 seeds: 3 FP64 tiles
    rec = rs.Recommender(gg)
    ids, sc, cnt = rec.RecommendationBatch(seeds, 0.15, 12, 10)
    assert rec.last_info.kernel_launches > 0
    for i, seed in enumerate(seeds):
        oids, osc = og.recommend(seed, 0.15, 12, top_n=10)
        same_ranking(ids[i, :cnt[i]], sc[i, :cnt[i]], oids, osc)
        assert_close_fp64(sc[i, :cnt[i]], osc, f"batched seed {seed}")
        single = rec.Recommendation(seed, 0.15, 12, 10)
        assert [p[0] for p in single] == ids[i, :cnt[i]].tolist()
    rec32 = rs.Recommender(gg, precision=rs.FP32)
    ids32, sc32, cnt32 = rec32.RecommendationBatch(seeds, 0.15, 12, 10)
    for i, seed in enumerate(seeds):
        oids, osc = og.recommend(seed, 0.15, 12, top_n=10)
        assert cnt32[i] == len(oids)
        assert np.abs(sc32[i, :cnt32[i]] - osc).max() <= 2e-6 * max(osc.max(), 1e-30) + 1e-30
    with pytest.raises(KeyError):
        rec.RecommendationBatch([seeds[0], int(np.flatnonzero(raw_deg == 0)[0])], 0.15, 3, 5)


# ------------------------------------------------------------------------------------------ Model re-run (buffers reused)
@pytest.mark.parametrize("prec", [rs.FP64, rs.FP32])
def test_rerun_reuses_the_model_and_matches_a_fresh_run(prec):
    g = load_golden("small_b")
    gg = gpu_graph(g["input"])
    deg = gg.degrees(raw=True)
    users = [int(u) for u in np.flatnonzero(deg > 0)[:4]]
    m = run_fixed(gg, [users[0]], C015, 7, prec)
    for seed, n_iter in ((users[1], 7), (users[2], 3), (users[0], 11)):
        m.rerun([seed], C015, n_iter)
        fresh = run_fixed(gg, [seed], C015, n_iter, prec)
        assert np.array_equal(m.scores(0), fresh.scores(0))          # same kernels, same order: bit-identical
        assert m.info().iterations == n_iter
        fresh.close()
    with pytest.raises(ValueError):
        m.rerun(users[:2], C015, 3)
    with pytest.raises(KeyError):
        m.rerun([gg.size()], C015, 3)
    m.close()


def test_rows_without_in_links_and_tile_cuts():
    """Edge-stream corner cases: rows with no in-links (padding link to the zero entry), a hub row spanning several
    4096-link tiles, rows ending exactly on a tile boundary; seeds inside each kind of row."""
    rng = np.random.default_rng(5)
    n = 30_000
    hub = 7
    src, dst = [], []
    # hub row of W^T: ~40k in-links (5 tiles); sources are distinct nodes each with a single out-link or two
    s = rng.choice(np.arange(100, n), size=20_000, replace=False)
    src += s.tolist(); dst += [hub] * len(s)
    # a band of nodes 8..60 that only have out-links (no in-links), pointing at random targets
    for u in range(8, 60):
        t = rng.choice(np.arange(1000, 2000), size=5, replace=False)
        src += [u] * 5; dst += t.tolist()
    # random background
    a = rng.integers(60, n, size=60_000); b = rng.integers(60, n, size=60_000)
    keep = a != b
    src += a[keep].tolist(); dst += b[keep].tolist()
    order = np.argsort(np.asarray(src), kind="stable")
    src = np.asarray(src, np.int32)[order]; dst = np.asarray(dst, np.int32)[order]
    key = src.astype(np.int64) * n + dst
    _, first = np.unique(key, return_index=True)
    first.sort()
    src, dst = src[first], dst[first]
    inp = dict(node_id=np.arange(n, dtype=np.int64) + 10_000, node_type=np.where(np.arange(n) % 3 == 0, 1, 2).astype(np.int32),
               src=src, dst=dst, etype=np.full(len(src), 2, np.int32), w=np.ones(len(src)))
    og = oracle_graph(inp)
    for opts in (dict(), dict(relabel=False), dict(layout=N.LAYOUT_VALUED), dict(hub_entries=0)):
        gg = gpu_graph(inp, **opts)
        for seed in (hub, 10, int(src[-1]), int(s[0])):
            want, _ = og.run(seed, C015, n_iter=6)
            r = run_fixed(gg, [seed], C015, 6)
            assert_close_fp64(r.scores(0), want, f"seed {seed} {opts}")
            r.close()
            r32 = run_fixed(gg, [seed], C015, 6, rs.FP32)
            assert_close_fp32(r32.scores(0), want, f"fp32 seed {seed} {opts}")
            r32.close()
        res, it = run_threshold(gg, [hub], C015, 1e-9 * n, max_iter=200)
        _, want_it = og.run(hub, C015, threshold=1e-9 * n)
        assert int(it[0]) == want_it
        res.close()
        gg.close()


# ------------------------------------------------------------------------------------------ Experiment-style evaluation (C5)
def test_experiment_style_holdout_evaluation():
    """Experiment.cs:69-138 in small (BASELINE config 5): hold out the newest tenth of the test users' likes on the device,
    recommend top-10 for every test user through the batched path, recall@10 against the oracle on the same graph."""
    import experiment_ref as R
    spec = dict(seed=2024, n_users=4_000, n_items=30_000, n_third=200, authorship_per_mille=900, n_like=260_000,
                n_friend=40_000, n_follow=1_000, n_mention=0, undefined_per_mille=0, scramble=1, p1_byte=61)
    full = O.synth_generate(spec)
    deg_like = np.bincount(full["src"][full["etype"] == 1], minlength=len(full["node_id"]))
    users = np.flatnonzero(deg_like[:spec["n_users"]] >= 20)[:96]
    links, test = R.hold_out(full, users, 10, 9)
    assert sum(len(t) for t in test.values()) > 200 and len(links["src"]) < len(full["src"])
    gg = rs.Graph.from_arrays(full["node_id"], full["node_type"], full["src"], full["dst"], full["etype"], full["w"])
    gtest = gg.hold_out(users, 10, 9)
    assert all(gtest[int(u)].tolist() == test[int(u)].tolist() for u in users)
    gg.buildGraph()
    og = oracle_graph(links)
    rec = rs.Recommender(gg)
    ids, sc, cnt = rec.RecommendationBatch(users, 0.15, 15, 10)
    hits_gpu = hits_cpu = 0
    for i, u in enumerate(users):
        a, b = og.recommend(int(u), 0.15, 15, top_n=10)
        same_ranking(ids[i, :cnt[i]], sc[i, :cnt[i]], a, b)
        hits_gpu += int(np.isin(ids[i, :cnt[i]], test[int(u)]).sum())
        hits_cpu += int(np.isin(a, test[int(u)]).sum())
    assert hits_gpu == hits_cpu
    # held-out tweets are recommendable again (they left the exclusion list with their LIKE link)
    assert hits_gpu > 0
    r = rs.evaluate_users(gg, None, None, 0.15, 15, k=10)
    assert int(r["hits_at_k"].sum()) == hits_gpu
    # full ranking of two users: hits / average precision as Experiment.cs:121-128 computes them
    for j, u in enumerate(users[:2]):
        full_rank = rec.Recommendation(int(u), 0.15, 15)
        a, b = og.recommend(int(u), 0.15, 15)
        same_ranking([p[0] for p in full_rank], [p[1] for p in full_rank], a, b)
        h1, ap1 = rs.evaluate(full_rank, test[int(u)])
        h2, ap2 = O.evaluate(a, test[int(u)])
        assert h1 == h2 == len(test[int(u)]) == int(r["hits"][j]) and abs(ap1 - ap2) <= 1e-12 and abs(r["avg_precision"][j] - ap2) <= 1e-12


# ------------------------------------------------------------------------------------------ concurrency (Program.cs:11, :61-66)
def test_concurrent_handles_from_several_threads():
    """The reference runs up to 10 ego networks at once, each thread on its own Graph / Recommender.  Every handle owns
    its stream, pools and cached iteration graph: concurrent create / build / recommend / destroy must not interfere."""
    import threading
    n_threads, rounds = 6, 3
    specs = [dict(seed=500 + t, n_users=300 + 40 * t, n_items=2500 + 300 * t, n_third=30, authorship_per_mille=800,
                  n_like=9000 + 1000 * t, n_friend=2000, n_follow=100, n_mention=0 if t % 2 else 200, undefined_per_mille=20,
                  scramble=1, p1_byte=61) for t in range(n_threads)]
    want = []
    for sp in specs:
        links = O.synth_generate(sp)
        og = oracle_graph(links)
        deg = np.bincount(links["src"], minlength=og.n)
        seeds = [int(s) for s in np.flatnonzero(deg[:sp["n_users"]] > 0)[:4]]
        want.append((links, seeds, [og.recommend(s, 0.15, 9, top_n=8)[0].tolist() for s in seeds]))
    errors = []

    def worker(t):
        try:
            links, seeds, lists = want[t]
            for _ in range(rounds):
                g = gpu_graph(links)
                rec = rs.Recommender(g)
                for s, w in zip(seeds, lists):
                    got = [p[0] for p in rec.Recommendation(s, 0.15, 9, 8)]
                    assert got == w, (t, s)
                ids, sc, cnt = rec.RecommendationBatch(seeds, 0.15, 9, 8)
                for i, w in enumerate(lists):
                    assert ids[i, :cnt[i]].tolist() == w, (t, i)
                g.close()
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


# ------------------------------------------------------------------------------------------ methodology switch (Experiment.cs:84-101)
def test_undefined_type_mask_equals_retyping_the_links():
    """`undefined_types=[FRIENDSHIP]` must give the graph the reference builds after rewriting every FRIENDSHIP link to
    UNDEFINED (Experiment.cs:84-101): same CSR, same ranks, same recommendation; raw links stay exportable."""
    g = load_golden("small_b")
    inp = {k: np.asarray(v) for k, v in g["input"].items()}
    et = inp["etype"].astype(np.int32)
    assert (et == rs.EdgeType.FRIENDSHIP).any()
    retyped = dict(inp)
    retyped["etype"] = np.where(et == rs.EdgeType.FRIENDSHIP, rs.EdgeType.UNDEFINED, et).astype(np.int32)
    og = oracle_graph(retyped)
    gm = gpu_graph(inp, undefined_types=[rs.EdgeType.FRIENDSHIP])
    rp, col, val = gm.csr()
    orp, ocol, oval = og.csr()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol) and np.array_equal(bits(val), bits(oval))
    assert np.array_equal(gm.export_links()["etype"], et)                    # the edge list itself is untouched
    deg = gm.degrees()
    seed = int(np.flatnonzero(deg > 0)[0])
    want, _ = og.run(seed, C015, n_iter=12)
    r = run_fixed(gm, [seed], C015, 12)
    assert_close_fp64(r.scores(0), want, "masked FRIENDSHIP")
    r.close()
    try:
        ids, sc = og.recommend(seed, 0.15, 12, top_n=10)
    except KeyError:
        return
    top = rs.Recommender(gm).Recommendation(seed, 0.15, 12, 10)
    same_ranking([p[0] for p in top], [p[1] for p in top], ids, sc)


def test_batched_topk_with_mass_ties_at_zero():
    """One iteration from a seed with few neighbours: fewer than k items have a score, the rest tie at exactly 0 and are
    ordered by id descending (Recommender.cs:34-38).  More than 65,536 such ties per column overflow the tile top-k's
    candidate list and must fall back to the exact per-column path."""
    spec = dict(seed=31, n_users=2_000, n_items=120_000, n_third=0, authorship_per_mille=300, n_like=60_000, n_friend=4_000,
                n_follow=0, n_mention=0, undefined_per_mille=0, scramble=1, p1_byte=100)
    cpu = O.synth_generate(spec)
    og = oracle_graph(cpu)
    gg = rs.Graph.synthetic(spec)
    gg.buildGraph()
    like_deg = np.bincount(cpu["src"][cpu["etype"] == 1], minlength=og.n)
    deg = np.bincount(cpu["src"], minlength=og.n)
    seeds = [int(u) for u in np.flatnonzero((deg[:2000] > 0) & (like_deg[:2000] <= 3))[:9]]
    assert len(seeds) == 9
    rec = rs.Recommender(gg)
    for n_iter in (1, 2):
        ids, sc, cnt = rec.RecommendationBatch(seeds, 0.15, n_iter, 10)
        for i, s in enumerate(seeds):
            oids, osc = og.recommend(s, 0.15, n_iter, top_n=10)
            assert cnt[i] == len(oids) == 10
            same_ranking(ids[i, :cnt[i]], sc[i, :cnt[i]], oids, osc)
    assert (sc[:, -1] == 0).any()          # the case under test did occur: a list padded with zero-score items
