#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- compiles the reference's own C# sources of the RWR path for this image.

    python oracle/cs2cpp.py [--ref /root/reference] [--out oracle/_ref/reference_rwr.hpp]

No C# toolchain exists here (mono, mcs, csc, dotnet, msbuild: all absent), so the reference cannot be built the usual way.
Its hot path, however, is 265 lines of a C# subset whose statements are also C++ statements once the declarations are
respelt: this script reads Recommenders/RWRBased/{Graph,Model,Recommender}.cs WHERE THEY LIE under /root/reference and
writes one C++ header into oracle/_ref/ (git-ignored; reference text never enters the repository), which
oracle/ref_driver.cpp wraps in a C ABI -> oracle/_ref/libref.so.  Every rule below is syntactic and local; none knows
what the code computes, and no arithmetic expression, loop bound, comparison or statement order is touched:

  declarations   `public class C {..}` / `public struct S {..}` -> `struct C {..};` (+ `S() = default;`: C# structs have an
                 implicit parameterless constructor), `public enum E {..}` -> `enum class E : int {..};`, member access
                 modifiers dropped, `namespace A.B` -> `namespace A_B`, `using` lines dropped
  types          `T[]` -> `Array<T>`, `new T[n]` -> `Array<T>(n)`, `new Dictionary<..>(..)` / `new List<..>(..)` /
                 `new KeyValuePair<..>(..)` -> the same without `new` (oracle/ref_shim.hpp keeps their reference semantics),
                 `long` -> `long long`, `var` -> `auto`, `null` -> `nullptr`, `1d` -> `1.0`,
                 `double.MaxValue` -> `std::numeric_limits<double>::max()`
  references     a variable of class type `C x` becomes `C* x` and `x.` becomes `x->` inside the type that declares it;
                 `this.` -> `this->`; `E.MEMBER` -> `E::MEMBER` for the enums
  statements     `foreach (T x in e)` -> `for (T x : e)`; `(a, b) => {` -> `[&](auto a, auto b) {`;
                 properties `.Count` / `.Length` -> `.Count()` / `.Length()`; `x.CompareTo(y)` -> `CompareTo(x, y)`
  order          top-level types are emitted enums first, then structs, then classes in dependency order (C++ needs a
                 type complete before its first use; C# does not)

The header records the SHA-256 of each source file it was made from.  tests/test_reference_pin.py checks the translator on
its own rule table (no reference needed) and, where oracle/_ref/libref.so exists, holds the C++ oracle, the Python literal
restatement and the committed golden vectors to the transliterated reference bit for bit.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ("Recommenders/RWRBased/Recommender.cs", "Recommenders/RWRBased/Graph.cs", "Recommenders/RWRBased/Model.cs")


def split_top_level(text: str):
    """-> (namespace name, [(kind, name, block text)]) of one C# file: the enum / struct / class declarations directly inside
    its namespace, cut out by brace matching."""
    text = text.lstrip("﻿")
    m = re.search(r"\bnamespace\s+([\w.]+)\s*\{", text)
    if not m:
        raise ValueError("no namespace")
    ns = m.group(1)
    body_start = m.end()
    blocks = []
    pos = body_start
    decl = re.compile(r"\b(?:public\s+|internal\s+)?(enum|struct|class)\s+(\w+)[^{;]*\{")
    while True:
        d = decl.search(text, pos)
        if not d:
            break
        depth, i = 1, d.end()
        while depth:
            ch = text[i]
            if ch == "{":
                depth += 1
            elif ch == "}":
                depth -= 1
            elif text.startswith("//", i):           # a brace inside a line comment does not count
                i = text.index("\n", i)
                continue
            i += 1
        blocks.append((d.group(1), d.group(2), text[d.start():i]))
        pos = i
    return ns, blocks


def translate_block(kind: str, name: str, block: str, enums, classes) -> str:
    s = block
    if kind == "enum":
        s = re.sub(r"^(?:public\s+|internal\s+)?enum\s+(\w+)", r"enum class \1 : int", s)
        return s + ";"
    s = re.sub(r"^(?:public\s+|internal\s+)?(?:struct|class)\s+(\w+)", r"struct \1", s)
    if kind == "struct":                                  # the implicit parameterless constructor of a C# struct
        s = re.sub(r"^(struct\s+\w+\s*\{)", r"\1\n        " + name + "() = default;", s)
    # member access modifiers
    s = re.sub(r"(?m)^(\s*)(?:public|private|protected|internal)\s+", r"\1", s)
    # ---- types
    s = re.sub(r"\bnew\s+(Dictionary|List|KeyValuePair)\s*<", r"\1<", s)
    s = re.sub(r"\bnew\s+(\w+)\[([^\]]+)\]", r"Array<\1>(\2)", s)
    s = re.sub(r"\b(\w+)\[\]", r"Array<\1>", s)
    s = re.sub(r"\blong\b", "long long", s)
    s = re.sub(r"\bvar\b", "auto", s)
    s = re.sub(r"\bnull\b", "nullptr", s)
    s = re.sub(r"\b(\d+)d\b", r"\1.0", s)
    s = s.replace("double.MaxValue", "std::numeric_limits<double>::max()")
    # ---- variables of class type are references: `C x` -> `C* x`, `x.` -> `x->` (inside this type only)
    cls = "|".join(sorted(classes))
    if cls:
        refs = set(re.findall(r"\b(?:%s)\s+(\w+)\s*[;,)=]" % cls, s))
        s = re.sub(r"\b(%s)\s+(?=\w+\s*[;,)=])" % cls, r"\1* ", s)
        for v in sorted(refs):
            s = re.sub(r"(?<![\w.])%s\.|(?<=this\.)%s\." % (v, v), v + "->", s)
    s = s.replace("this.", "this->")
    for e in sorted(enums):
        s = re.sub(r"\b%s\.(?=[A-Z_])" % e, e + "::", s)
    # ---- statements
    s = re.sub(r"\bforeach\s*\(\s*([\w<>, ]+?)\s+(\w+)\s+in\s+", r"for (\1 \2 : ", s)
    s = re.sub(r"\(\s*(\w+)\s*,\s*(\w+)\s*\)\s*=>\s*\{", r"[&](auto \1, auto \2) {", s)
    s = re.sub(r"\.(Count|Length)\b(?!\s*\()", r".\1()", s)
    s = re.sub(r"->(Count|Length)\b(?!\s*\()", r"->\1()", s)
    s = re.sub(r"((?:\w+(?:\.|->))*\w+)\.CompareTo\(([^()]*)\)", r"CompareTo(\1, \2)", s)
    return s + ";"


def depends_on(block: str, other: str) -> bool:
    return re.search(r"\b%s\b" % other, block) is not None


def transliterate(files: dict) -> str:
    """files: {relative path: C# text} -> the C++ header text."""
    ns_name, items = None, []
    for rel, text in files.items():
        ns, blocks = split_top_level(text)
        ns_name = ns_name or ns
        if ns != ns_name:
            raise ValueError("sources of more than one namespace")
        items += [(k, n, b, rel) for k, n, b in blocks]
    enums = {n for k, n, _, _ in items if k == "enum"}
    classes = {n for k, n, _, _ in items if k == "class"}
    ordered = [it for it in items if it[0] == "enum"] + [it for it in items if it[0] == "struct"]
    pending = [it for it in items if it[0] == "class"]
    while pending:                                        # classes: a type after every class its text names
        for it in pending:
            if not any(o is not it and depends_on(it[2], o[1]) for o in pending):
                ordered.append(it)
                pending.remove(it)
                break
        else:
            raise ValueError("cyclic dependency between classes: " + ", ".join(p[1] for p in pending))
    out = ["// GENERATED by oracle/cs2cpp.py -- the reference's own statements, declarations respelt for C++.  Never commit.",
           "#pragma once", '#include "../ref_shim.hpp"', ""]
    for rel, text in files.items():
        out.append("// source: %s  sha256 %s" % (rel, hashlib.sha256(text.encode("utf-8")).hexdigest()))
    out += ["", "namespace %s {" % ns_name.replace(".", "_"), "using namespace bcl;", ""]
    for kind, name, block, rel in ordered:
        out.append("// ---- %s %s (%s)" % (kind, name, rel))
        out.append("    " + translate_block(kind, name, block, enums, classes))
        out.append("")
    out.append("}  // namespace")
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "_ref", "reference_rwr.hpp"))
    args = ap.parse_args()
    files = {}
    for rel in SOURCES:
        with open(os.path.join(args.ref, rel), encoding="utf-8-sig") as f:
            files[rel] = f.read()
    text = transliterate(files)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text)
    print("wrote %s (%d lines from %d reference files)" % (args.out, text.count("\n"), len(files)))


if __name__ == "__main__":
    main()
