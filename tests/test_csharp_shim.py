"""The C# P/Invoke shim (shipped as source, no C# toolchain here) must bind every symbol the header declares, with the
struct fields of the header in the same order -- checked textually."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*p):
    with open(os.path.join(ROOT, *p)) as f:
        return f.read()


def test_every_exported_symbol_has_a_dllimport():
    header = _read("include", "rwr_b200.h")
    cs = _read("recommendersystems_b200", "csharp", "RwrNative.cs")
    declared = set(re.findall(r"\b(rwr_[a-z_0-9]+)\s*\(", header))
    bound = set(re.findall(r"extern\s+[\w\[\]]+\s+(rwr_[a-z_0-9]+)\s*\(", cs))
    assert declared - bound == set(), f"no [DllImport] for {sorted(declared - bound)}"
    assert bound - declared == set(), f"[DllImport] of unknown symbols {sorted(bound - declared)}"


def _c_fields(header, name):
    body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        out += [n.strip() for n in names.split(",")]
    return out


def _cs_fields(cs, name):
    body = re.search(r"public struct " + name + r" \{(.*?)\n    \}", cs, re.S).group(1)
    out = []
    for decl in re.findall(r"public\s+\w+\s+([^;(]+);", body):
        out += [n.strip() for n in decl.split(",")]
    return out


def test_struct_fields_match_the_header():
    header = _read("include", "rwr_b200.h")
    cs = _read("recommendersystems_b200", "csharp", "RwrNative.cs")
    for c_name, cs_name in (("rwr_opts", "RwrOpts"), ("rwr_run_info", "RwrRunInfo"), ("rwr_graph_info", "RwrGraphInfo"),
                            ("rwr_synth_spec", "RwrSynthSpec")):
        assert _c_fields(header, c_name) == _cs_fields(cs, cs_name), c_name
