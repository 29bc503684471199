"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/*.json.

    python oracle/make_golden.py

Every number is produced by the pure-Python literal restatement (oracle/rwr_literal.py) of
Graph.cs / Model.cs / Recommender.cs, i.e. by executing the reference's statements one by one in
IEEE double.  The reference has no tests or fixtures of its own and no C# runtime exists here; its sources
compiled for g++ (oracle/cs2cpp.py -> oracle/_ref/libref.so) reproduce every file written here bit for bit
(tests/test_reference_pin.py), and oracle/make_golden_ref.py writes one more fixture with the reference itself.
Tests then require the C++ oracle and the CUDA path to reproduce them.  Floats are stored as C99 hex strings (bit-exact).

The synthetic-generator fixture is produced by a third, pure-Python implementation of this repo's
generator spec (include/rwr_b200.h), independent of the C++ and CUDA ones.
"""
from __future__ import annotations

import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import rwr_literal as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
M64 = (1 << 64) - 1


def hexes(xs):
    return [float(x).hex() for x in xs]


def csr_of(g: R.Graph):
    row_ptr, col, val = [0], [], []
    for i in range(g.size()):
        links = g.graph[i]
        if links is not None:
            for l in links:
                col.append(l.targetNode)
                val.append(l.weight)
        row_ptr.append(len(col))
    return row_ptr, col, val


def flat_of(nodes, edges, global_order=None):
    """Flatten dictionaries to the C-ABI SoA in (source asc, insertion) order."""
    n = len(nodes)
    src, dst, et, w = [], [], [], []
    for i in range(n):
        for l in edges.get(i, []):
            src.append(i); dst.append(l.targetNode); et.append(l.type); w.append(l.weight)
    return dict(node_id=[nodes[i].id for i in range(n)], node_type=[nodes[i].type for i in range(n)],
                src=src, dst=dst, etype=et, w=hexes(w))


def run_threshold_capped(m: R.Model, threshold: float, cap: int):
    """Model.run(double) (Model.cs:57-66) with an iteration cap the reference does not have."""
    while True:
        m.deliverRanks()
        if m.checkConvergence(threshold):
            m.updateRanks()
            return m.nDeliver
        m.updateRanks()
        if m.nDeliver >= cap:
            return None


def case_from(nodes, edges, seeds, iters, thresholds, c_float=0.15, top_ns=(1, 3, 10), uniform_iters=(1, 3)):
    g = R.Graph(nodes, edges)
    g.buildGraph()
    rp, col, val = csr_of(g)
    c = R.widen_float(c_float)
    out = dict(input=flat_of(nodes, edges), csr=dict(row_ptr=rp, col=col, val=hexes(val)),
               outdeg=[rp[i + 1] - rp[i] for i in range(len(nodes))],
               damping_float=c_float, damping_double=c.hex(), seeds=[])
    for s in seeds:
        entry = dict(seed=s, ranks={}, thresholds={}, recommendation={}, top={})
        for n in iters:
            m = R.Model(g, c, s)
            m.run(int(n))
            entry["ranks"][str(n)] = hexes(m.rank)
        for thr in thresholds:
            m = R.Model(g, c, s)
            t = (1 / R.DOUBLE_MAX) * g.size() if thr == "default" else float(thr)   # Model.cs:53
            n_it = run_threshold_capped(m, t, 2000)
            if n_it is None:       # the reference would spin forever (no bitwise fixed point): not a fixture
                entry["thresholds"][str(thr)] = dict(iters=None)
            else:
                entry["thresholds"][str(thr)] = dict(iters=n_it, rank=hexes(m.rank))
        rec = R.Recommender(g)
        if s in edges:
            n = iters[-1]
            full = rec.Recommendation(s, c_float, n)
            entry["recommendation"] = dict(n_iter=n, ids=[p[0] for p in full], scores=hexes([p[1] for p in full]))
            for k in top_ns:
                t = rec.Recommendation(s, c_float, n, k)
                entry["top"][str(k)] = dict(ids=[p[0] for p in t], scores=hexes([p[1] for p in t]))
        else:
            entry["recommendation"] = "KeyNotFoundException"
        out["seeds"].append(entry)
    out["uniform"] = {}
    for n in uniform_iters:
        m = R.Model(g, c)
        m.run(int(n))
        out["uniform"][str(n)] = hexes(m.rank)
    return out


def random_case(rng: random.Random, n_user, n_item, n_etc, n_rel, frac_weights=True):
    """Small DataLoader-shaped graph: users, items, third parties; bidirectional LIKE/FRIENDSHIP/FOLLOW/
    AUTHORSHIP, directed fractional MENTION, some FRIENDSHIP retyped UNDEFINED (Experiment.cs:84-101),
    (target,type) dedup per source (DataLoader.cs:60-77), isolated and UNDEFINED-only nodes."""
    n = n_user + n_item + n_etc
    nodes = {}
    for i in range(n):
        if i < n_user:
            nodes[i] = R.Node(1000 + i, R.USER)
        elif i < n_user + n_item:
            nodes[i] = R.Node(5000 + rng.randrange(10 ** 6) * 1000 + i, R.ITEM)   # ids not monotone in index
        else:
            nodes[i] = R.Node(2000 + i, R.ETC_NODE)
    edges = {}

    def add(s, d, t, w):
        lst = edges.setdefault(s, [])
        for l in lst:
            if l.targetNode == d and l.type == t:
                return
        lst.append(R.ForwardLink(d, t, w))

    for _ in range(n_rel):
        kind = rng.random()
        u = min(int(rng.paretovariate(1.2)) - 1, n_user - 1)
        if kind < 0.55:
            it = n_user + min(int(rng.paretovariate(1.1)) - 1, n_item - 3)    # last 2 items stay isolated
            add(u, it, R.LIKE, 1.0); add(it, u, R.LIKE, 1.0)
        elif kind < 0.75:
            v = rng.randrange(n_user)
            if v != u:
                add(u, v, R.FRIENDSHIP, 1.0); add(v, u, R.FRIENDSHIP, 1.0)
        elif kind < 0.85 and n_etc > 1:
            x = n_user + n_item + rng.randrange(n_etc - 1)                    # last ETC stays isolated
            add(u, x, R.FOLLOW, 1.0); add(x, u, R.FOLLOW, 1.0)
        elif kind < 0.95:
            it = n_user + rng.randrange(n_item - 2)
            add(u, it, R.AUTHORSHIP, 1.0); add(it, u, R.AUTHORSHIP, 1.0)
        else:
            v = rng.randrange(n_user)
            if v != u:
                add(u, v, R.MENTION, (rng.randrange(1, 200) / 64.0) if frac_weights else 1.0)
    # retype a share of FRIENDSHIP links to UNDEFINED after dedup
    for lst in edges.values():
        for l in lst:
            if l.type == R.FRIENDSHIP and rng.random() < 0.3:
                l.type = R.E_UNDEFINED
    # one user whose only links are UNDEFINED (dangling although it has a key)
    lonely = n_user - 1
    edges[lonely] = [R.ForwardLink(0, R.E_UNDEFINED, 1.0), R.ForwardLink(1, R.E_UNDEFINED, 2.5)]
    return nodes, edges


# ------------------------------------------------------------------ pure-Python synthetic generator (repo spec)
def mix64(z):
    z &= M64
    z ^= z >> 30; z = (z * 0xbf58476d1ce4e5b9) & M64
    z ^= z >> 27; z = (z * 0x94d049bb133111eb) & M64
    z ^= z >> 31
    return z


def H(seed, j, k):
    return mix64((mix64((seed + 0x9E3779B97F4A7C15 * (j + 1)) & M64) + 0xD1B54A32D192ED03 * (k + 1)) & M64)


def ceil_log2(n):
    L = 0
    while (1 << L) < n:
        L += 1
    return L


def synth_py(spec):
    seed, U, T, X = spec["seed"], spec["n_users"], spec["n_items"], spec["n_third"]
    p1 = spec["p1_byte"]

    def draw(j, which, rng_):
        L = ceil_log2(rng_)
        v = 0
        for l in range(L):
            h = H(seed, j, which * 4 + (l >> 3))
            byte = (h >> (8 * (l & 7))) & 255
            v |= (1 if byte < p1 else 0) << l
        return v % rng_

    def perm(x, rng_, salt):
        if not spec["scramble"]:
            return x
        return (x * 2654435761 + mix64(seed ^ salt) % rng_) % rng_

    SU, ST = 0x1111111111111111, 0x2222222222222222
    keys = set()
    j = 0
    key = lambda s, c, d: (s << 31) | (c << 28) | d
    for t in range(T):
        if H(seed, j, 15) % 1000 < spec["authorship_per_mille"]:
            a = perm(draw(j, 0, U), U, SU)
            keys.add(key(a, 3, U + t)); keys.add(key(U + t, 3, a))
        j += 1
    for _ in range(spec["n_like"]):
        u = perm(draw(j, 0, U), U, SU); it = U + perm(draw(j, 1, T), T, ST)
        keys.add(key(u, 0, it)); keys.add(key(it, 0, u)); j += 1
    for _ in range(spec["n_friend"]):
        u = perm(draw(j, 0, U), U, SU); v = perm(draw(j, 1, U), U, SU)
        if u != v:
            keys.add(key(u, 1, v)); keys.add(key(v, 1, u))
        j += 1
    for _ in range(spec["n_follow"]):
        u = perm(draw(j, 0, U), U, SU); x = U + T + draw(j, 1, X)
        keys.add(key(u, 2, x)); keys.add(key(x, 2, u)); j += 1
    for _ in range(spec["n_mention"]):
        u = perm(draw(j, 0, U), U, SU); v = perm(draw(j, 1, U), U, SU)
        if u != v:
            keys.add(key(u, 4, v))
        j += 1
    src, dst, et, w = [], [], [], []
    for k in sorted(keys):
        s, c, d = k >> 31, (k >> 28) & 7, k & ((1 << 28) - 1)
        ty, wt = 0, 1.0
        if c == 0:
            ty = R.LIKE
        elif c == 1:
            lo, hi = min(s, d), max(s, d)
            und = mix64(seed ^ 0xF1E2D3C4B5A69788 ^ ((lo << 32) | hi)) % 1000 < spec["undefined_per_mille"]
            ty = R.E_UNDEFINED if und else R.FRIENDSHIP
        elif c == 2:
            ty = R.FOLLOW
        elif c == 3:
            ty = R.AUTHORSHIP
        else:
            ty = R.MENTION
            wt = (1 + (mix64(seed ^ 0xA5A5A5A55A5A5A5A ^ ((s << 32) | d)) & 127)) / 32.0
        src.append(s); dst.append(d); et.append(ty); w.append(wt)
    N = U + T + X
    node_id = [1000000000 + i if i < U else (5000000000000 + (i - U) if i < U + T else 2000000000 + (i - U - T))
               for i in range(N)]
    node_type = [R.USER if i < U else (R.ITEM if i < U + T else R.ETC_NODE) for i in range(N)]
    return dict(node_id=node_id, node_type=node_type, src=src, dst=dst, etype=et, w=hexes(w))


SYNTH_TINY = dict(seed=20260101, n_users=37, n_items=150, n_third=9, authorship_per_mille=700, n_like=600,
                  n_friend=160, n_follow=40, n_mention=25, undefined_per_mille=200, scramble=1, p1_byte=61, reserved=0)


def main():
    os.makedirs(OUT, exist_ok=True)
    nodes, edges = R.kat_graph_8c()
    kat = case_from(nodes, edges, seeds=[0, 1, 5], iters=[1, 2, 10],
                    thresholds=[1e-3, 1e-6, 1e-9, 1e-12, "default"], top_ns=(1, 2, 10))
    with open(os.path.join(OUT, "kat_8c.json"), "w") as f:
        json.dump(kat, f, indent=0)

    rng = random.Random(20260101)
    for name, (nu, ni, ne, nr) in dict(small_a=(12, 30, 4, 160), small_b=(25, 70, 6, 420)).items():
        nodes, edges = random_case(rng, nu, ni, ne, nr)
        case = case_from(nodes, edges, seeds=[0, 3, nu - 1, nu + 1], iters=[1, 2, 5, 15],
                         thresholds=[1e-3, 1e-9, 1e-11])
        with open(os.path.join(OUT, f"{name}.json"), "w") as f:
            json.dump(case, f, indent=0)

    syn = synth_py(SYNTH_TINY)
    with open(os.path.join(OUT, "synth_tiny.json"), "w") as f:
        json.dump(dict(spec=SYNTH_TINY, graph=syn), f, indent=0)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
