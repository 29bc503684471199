"""Random flattened inputs of the C ABI with the corner cases of SURVEY 8(a), shared by the CPU and GPU tests."""


def random_flat(rng, n_users, n_items, n_etc, n_links, zero_rows=0, nan_rows=0):
    """A random flattened input with the corner cases of SURVEY 8(a).  Links in (source asc, insertion) order."""
    n = n_users + n_items + n_etc
    node_type = [1] * n_users + [2] * n_items + [3] * n_etc
    node_id = [1000 + i for i in range(n_users)] + [5000 + rng.randrange(10 ** 6) * 0 + i for i in range(n_items)] + \
              [9000 + i for i in range(n_etc)]
    per_src = [[] for _ in range(n)]
    for _ in range(n_links):
        s, d = rng.randrange(n), rng.randrange(n)
        t = rng.choice([0, 1, 1, 1, 2, 2, 3, 4, 5, 7])                     # UNDEFINED links, every defined type
        w = 1.0 if t != 4 else rng.choice([0.5, 0.25, 1.75, rng.random() * 3])
        per_src[s].append((d, t, w))
        if rng.random() < 0.15:
            per_src[s].append((d, rng.choice([1, 5]), 1.0))                # the same target under another type: a multi-edge
    for s in rng.sample(range(n), zero_rows):                              # weights that sum to 0 -> 0/0 = NaN (Graph.cs:81)
        per_src[s] = [(rng.randrange(n), 2, 0.0), (rng.randrange(n), 3, 0.0)]
    for s in rng.sample(range(n), nan_rows):                               # +w and -w -> x / 0 = +-Inf
        d = rng.randrange(n)
        per_src[s] = [(d, 2, 2.0), (d, 3, -2.0), (rng.randrange(n), 1, 0.0)]
    for s in rng.sample(range(n), max(1, n // 10)):                        # UNDEFINED-only rows and rows without an entry
        per_src[s] = [(rng.randrange(n), 0, 1.0)] if rng.random() < 0.5 else []
    src, dst, et, w = [], [], [], []
    for s in range(n):
        for d, t, x in per_src[s]:
            src.append(s); dst.append(d); et.append(t); w.append(x)
    return dict(node_id=node_id, node_type=node_type, src=src, dst=dst, etype=et, w=w)
