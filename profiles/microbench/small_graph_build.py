"""Per-ego-network cost on the reference's scale (C1): Graph(nodes, edges) + buildGraph() + one Recommendation, as
Experiment.cs:104-109 does for every ego network."""
import sys, time; sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np, recommendersystems_b200 as rs, oracle as O
spec = dict(seed=20260101, n_users=1000, n_items=9000, n_third=200, authorship_per_mille=800, n_like=36000, n_friend=8000,
            n_follow=600, n_mention=400, undefined_per_mille=100, scramble=1, p1_byte=61, reserved=0)
L = O.synth_generate(spec)
seed = int(np.flatnonzero(np.bincount(L["src"], minlength=len(L["node_id"]))[:1000] > 0)[0])
def one():
    g = rs.Graph.from_arrays(L["node_id"], L["node_type"], L["src"], L["dst"], L["etype"], L["w"])
    t1 = time.perf_counter()
    g.buildGraph()
    t2 = time.perf_counter()
    top = rs.Recommender(g).Recommendation(seed, 0.15, 20, 10)
    t3 = time.perf_counter()
    b = g.info().build_ms
    g.close()
    return t1, t2, t3, b
for _ in range(5): one()
acc = np.zeros(4)
for _ in range(50):
    t0 = time.perf_counter(); t1, t2, t3, b = one(); t4 = time.perf_counter()
    acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t0]
acc /= 50
print(f"C1 ego network ({len(L['node_id'])} nodes, {len(L['src'])} links): create {acc[0]*1e3:.2f} ms, buildGraph {acc[1]*1e3:.2f} ms (device {b:.2f} ms), "
      f"first Recommendation {acc[2]*1e3:.2f} ms, total incl. destroy {acc[3]*1e3:.2f} ms")
