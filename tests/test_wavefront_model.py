"""The CPU model behind DESIGN.md section 3 (profiles/microbench/wavefront_model.py) on a tiny graph: its rebuilt label
order is a permutation with the hot nodes in descending-degree order, its edge stream holds every link once plus one
padding link per row without in-links, and its wavefront counts respect their obvious bounds."""
import importlib.util
import os

import numpy as np

import oracle as O
from conftest import C1_SPEC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("wavefront_model", os.path.join(ROOT, "profiles", "microbench", "wavefront_model.py"))
wm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(wm)


def test_model_rebuilds_labels_and_stream():
    g = O.synth_generate(dict(C1_SPEC, n_mention=0, undefined_per_mille=0))
    n, src, dst = len(g["node_id"]), g["src"], g["dst"]
    new_of_old, n_hot, deg = wm.relabel(n, src, dst)
    assert sorted(new_of_old.tolist()) == list(range(n))
    old_of_new = np.argsort(new_of_old)
    hot_deg = deg[old_of_new[:n_hot]]
    assert (hot_deg >= wm.HOT_MIN).all() and (np.diff(hot_deg) <= 0).all()
    for by_label in (False, True):
        stream = wm.build_stream(n, src, dst, new_of_old, by_label)
        indeg = np.bincount(dst, minlength=n)
        assert len(stream) == len(src) + int((indeg == 0).sum())
        assert int((stream == n).sum()) == int((indeg == 0).sum())
        assert np.array_equal(np.sort(stream[stream < n]), np.sort(new_of_old[src]))
        for layout in ("lane8", "lane2", "link"):
            m = wm.instr_matrix(stream, layout, n)
            assert m.shape[1] == 32 and m.size >= len(stream)
            g_wf, s_wf, g_l, s_l = wm.count(m, hub=64, elt=8)
            assert g_l + s_l == m.size and 0 < g_wf <= g_l and 0 < s_wf <= s_l
    plain, blocked, pads = wm.blocked_stats(n, src, dst, new_of_old, 4)
    assert blocked - pads == len(src) and blocked >= plain
