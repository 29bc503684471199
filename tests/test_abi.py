"""CPU tests of the drop-in boundary: librwr_b200.so loads, exports every symbol include/rwr_b200.h declares,
struct layouts match, and compute entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT

from recommendersystems_b200 import _native as N


def header_functions():
    src = open(os.path.join(ROOT, "include", "rwr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rwr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = N.lib()
    names = header_functions()
    assert len(names) >= 24
    for name in names:
        assert hasattr(L, name), f"librwr_b200.so does not export {name}"
    assert set(names) == set(N.SYMBOLS), "ctypes table and header disagree"
    assert L.rwr_abi_version() == 2


def test_struct_layouts(tmp_path):
    """The header is valid C, and ctypes mirrors agree with the C compiler on every struct size."""
    import subprocess
    csrc = tmp_path / "sz.c"
    csrc.write_text('#include <stdio.h>\n#include "rwr_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n",'
                    'sizeof(rwr_opts),sizeof(rwr_synth_spec),sizeof(rwr_graph_info),sizeof(rwr_run_info));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(csrc), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(N.rwr_opts), C.sizeof(N.rwr_synth_spec), C.sizeof(N.rwr_graph_info), C.sizeof(N.rwr_run_info)]
    import oracle as O
    assert C.sizeof(O.SynthSpec) == C.sizeof(N.rwr_synth_spec) == O.lib().orc_sizeof_synth_spec()


def test_no_torch_types_and_no_oracle_in_product():
    """The product package never imports the oracle, and the shared library does not link it."""
    pkg = os.path.join(ROOT, "recommendersystems_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cs", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "import oracle" not in text and "rwr_literal" not in text, f
                assert "libref" not in text and "ref_shim" not in text and "reference_rwr" not in text and "import ref" not in text, f
    import subprocess
    out = subprocess.run(["ldd", N.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out


def test_host_side_evaluate_matches_reference_arithmetic():
    import rwr_literal as R
    import recommendersystems_b200 as rs
    rec = [(50, .9), (40, .8), (30, .7), (20, .6), (10, .5)]
    for test in ({40, 10}, set(), {50}, {99}, {10, 20, 30, 40, 50}):
        assert rs.evaluate(rec, test) == R.evaluate(rec, test)


def test_fails_loudly_without_gpu():
    if N.lib().rwr_device_count() > 0:
        pytest.skip("a CUDA device is present")
    import recommendersystems_b200 as rs
    nodes = {0: rs.Node(1, rs.NodeType.USER), 1: rs.Node(2, rs.NodeType.ITEM)}
    edges = {0: [rs.ForwardLink(1, rs.EdgeType.LIKE, 1.0)]}
    with pytest.raises(rs.RwrError) as ei:
        rs.Graph(nodes, edges)
    assert ei.value.code == N.RWR_E_CUDA and "no CPU fallback" in str(ei.value)
    with pytest.raises(rs.RwrError):
        rs.Graph.synthetic(dict(seed=1, n_users=4, n_items=4, n_like=4, n_friend=2))


def test_argument_validation_needs_no_gpu():
    L = N.lib()
    out = C.c_void_p()
    assert L.rwr_graph_create(-1, None, None, 0, None, None, None, None, None, C.byref(out)) == N.RWR_E_INVALID
    assert L.rwr_graph_build(None) == N.RWR_E_INVALID
    assert b"NULL" in L.rwr_last_error()
    hits, ap = C.c_int32(), C.c_double()
    ids = np.array([3, 2, 1], np.int64)
    assert L.rwr_evaluate(ids.ctypes.data_as(C.c_void_p), 3, None, 0, C.byref(hits), C.byref(ap)) == 0
    assert hits.value == 0 and ap.value == 0.0
