#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- compiles the reference's own C# sources for this image.

    python oracle/cs2cpp.py [--ref /root/reference] [--out-dir oracle/_ref]

No C# toolchain exists here (mono, mcs, csc, dotnet, msbuild: all absent), so the reference cannot be built the usual way.
Its RWR path and the callers either side of it, however, are a few hundred lines of a C# subset whose statements are also
C++ statements once the declarations are respelt.  This script reads the sources WHERE THEY LIE under /root/reference and
writes two C++ headers into a TEMPORARY directory that `make -C oracle ref` deletes once g++ has consumed them: only the
binaries (oracle/_ref/libref.so, caller_reference) and the SHA-256 list of the sources stay, git-ignored; reference text never
enters the repository or its working tree:

  reference_rwr.hpp         Recommenders/RWRBased/{Graph,Model,Recommender}.cs, whole            (the hot path)
  reference_experiment.hpp  TweetRecommender/DataLoader.cs, whole; of Experiment.cs the enums, ThreadParams and the k-fold
                            loop of runKFoldCrossValidation (`for (int fold ...) {...}`, cut out by brace matching and
                            wrapped in a function whose parameters are the locals it reads)       (N1-N4)

oracle/ref_driver.cpp wraps both in a C ABI -> oracle/_ref/libref.so; the SQLite binding DataLoader calls
(TweetRecommender/SQLiteAdapter.cs over System.Data.SQLite, a NuGet package that is not vendored) is replaced by an
in-memory table store with the same six methods (oracle/ref_shim.hpp).  Every rule below is syntactic and local; none
knows what the code computes, and no arithmetic expression, loop bound, comparison or statement order is touched:

  declarations   `public class C {..}` / `public struct S {..}` -> `struct C {..};` (+ `S() = default;`: C# structs have an
                 implicit parameterless constructor), `public enum E {..}` -> `enum class E : int {..};`, member access
                 modifiers dropped, `namespace A.B` -> `namespace A_B`, `using A.B;` -> `using namespace A_B;` when A.B
                 is one of the translated namespaces, dropped otherwise
  types          `T[]` -> `Array<T>`, `new T[n]` -> `Array<T>(n)`, `new Dictionary<..>(..)` / `List` / `HashSet` /
                 `KeyValuePair` -> the same without `new` (oracle/ref_shim.hpp keeps their reference semantics),
                 `long` -> `long long`, `string` -> `std::string`, `var` -> `auto`, `null` -> `nullptr`, `1d` -> `1.0`,
                 `double.MaxValue` -> `std::numeric_limits<double>::max()`
  references     a variable of class type `C x` becomes `C* x` and `x.` becomes `x->` inside the type that declares it;
                 `this.` -> `this->`; `E.MEMBER` -> `E::MEMBER` for the enums
  statements     `foreach (T x in e)` -> `for (T x : e)`; `(a, b) => {` -> `[&](auto a, auto b) {`; properties `.Count` /
                 `.Length` / `.Keys` / `.Values` -> calls; `x.CompareTo(y)` -> `CompareTo(x, y)`;
                 static BCL calls `long.Parse(` / `Path.X(` / `Math.X(` -> `ParseLong(` / `Path::X(` / `Math::X(`;
                 a `lock (..) {..}` block that holds nothing but `Console.` statements is dropped (progress output)
  order          top-level types are emitted enums first, then structs, then classes in dependency order (C++ needs a
                 type complete before its first use; C# does not)

oracle/_ref/sources.sha256 records the SHA-256 of each source file the library was made from.  tests/test_reference_pin.py checks the translator on
its own rule table (no reference needed) and, where oracle/_ref/libref.so exists, holds the C++ oracle, the Python
restatements and the committed golden vectors to the transliterated reference.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ("Recommenders/RWRBased/Recommender.cs", "Recommenders/RWRBased/Graph.cs", "Recommenders/RWRBased/Model.cs")
CALLER_SOURCES = ("TweetRecommender/Experiment.cs", "TweetRecommender/DataLoader.cs")
GENERIC_BCL = "Dictionary|List|HashSet|KeyValuePair"


def strip_bom(text: str) -> str:
    return text.lstrip("﻿")


def match_brace(text: str, open_end: int) -> int:
    """Index just past the `}` that closes the `{` ending at `open_end`; braces in line comments and strings do not count."""
    depth, i = 1, open_end
    while depth:
        ch = text[i]
        if text.startswith("//", i):
            i = text.index("\n", i)
            continue
        if ch == '"':
            i += 1
            while text[i] != '"':
                i += 2 if text[i] == "\\" else 1
        elif ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
        i += 1
    return i


def split_top_level(text: str):
    """-> (namespace, [using names], [(kind, name, block text)]) of one C# file: the enum / struct / class declarations
    directly inside its namespace, cut out by brace matching."""
    text = strip_bom(text)
    usings = re.findall(r"(?m)^\s*using\s+([\w.]+)\s*;", text)
    m = re.search(r"\bnamespace\s+([\w.]+)\s*\{", text)
    if not m:
        raise ValueError("no namespace")
    blocks = []
    pos = m.end()
    decl = re.compile(r"\b(?:public\s+|internal\s+)?(enum|struct|class)\s+(\w+)[^{;]*\{")
    while True:
        d = decl.search(text, pos)
        if not d:
            break
        end = match_brace(text, d.end())
        blocks.append((d.group(1), d.group(2), text[d.start():end]))
        pos = end
    return m.group(1), usings, blocks


def drop_console_locks(s: str) -> str:
    """`lock (..) { Console.Write(..); .. }` -> a comment, when every statement inside is console output."""
    out, pos = [], 0
    for m in re.finditer(r"\block\s*\([^)]*\)\s*\{", s):
        if m.start() < pos:
            continue
        end = match_brace(s, m.end())
        inner = re.sub(r"//[^\n]*", "", s[m.end():end - 1])
        statements = [x.strip() for x in inner.split(";") if x.strip()]
        if statements and all("Console." in x for x in statements):
            out.append(s[pos:m.start()] + "/* console output (a lock block of Console statements) dropped by cs2cpp */")
            pos = end
    out.append(s[pos:])
    return "".join(out)


def respell(s: str, enums, classes, structs=()) -> str:
    """The rules that apply inside a type or a statement fragment."""
    s = drop_console_locks(s)
    # static BCL calls (before `long` is respelt)
    s = re.sub(r"\blong\.Parse\(", "ParseLong(", s)
    s = re.sub(r"\b(Path|Math)\.(?=\w+\()", r"\1::", s)
    # ---- types
    s = re.sub(r"\bnew\s+(%s)\s*<" % GENERIC_BCL, r"\1<", s)
    if structs:
        s = re.sub(r"\bnew\s+(%s)\s*\(" % "|".join(sorted(structs)), r"\1(", s)
    s = re.sub(r"\bnew\s+(\w+)\[([^\]]+)\]", r"Array<\1>(\2)", s)
    s = re.sub(r"\b(\w+)\[\]", r"Array<\1>", s)
    s = re.sub(r"\blong\b", "long long", s)
    s = re.sub(r"\bstring\b", "std::string", s)
    s = re.sub(r"\bvar\b", "auto", s)
    s = re.sub(r"\bnull\b", "nullptr", s)
    s = re.sub(r"\b(\d+)d\b", r"\1.0", s)
    s = s.replace("double.MaxValue", "std::numeric_limits<double>::max()")
    # ---- variables of class type are references: `C x` -> `C* x`, `x.` -> `x->` (inside this type / fragment only)
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
    s = re.sub(r"\bforeach\s*\(\s*([\w<>:, ]+?)\s+(\w+)\s+in\s+", r"for (\1 \2 : ", s)
    s = re.sub(r"\(\s*(\w+)\s*,\s*(\w+)\s*\)\s*=>\s*\{", r"[&](auto \1, auto \2) {", s)
    s = re.sub(r"(\.|->)(Count|Length|Keys|Values)\b(?!\s*\()", r"\1\2()", s)
    s = re.sub(r"((?:\w+(?:\.|->))*\w+)\.CompareTo\(([^()]*)\)", r"CompareTo(\1, \2)", s)
    return s


def translate_block(kind: str, name: str, block: str, enums, classes, structs=()) -> str:
    s = block
    if kind == "enum":
        return re.sub(r"^(?:public\s+|internal\s+)?enum\s+(\w+)", r"enum class \1 : int", s) + ";"
    s = re.sub(r"^(?:public\s+|internal\s+)?(?:struct|class)\s+(\w+)", r"struct \1", s)
    if kind == "struct":                                  # the implicit parameterless constructor of a C# struct
        s = re.sub(r"^(struct\s+\w+\s*\{)", r"\1\n        " + name + "() = default;", s)
    s = re.sub(r"(?m)^(\s*)(?:public|private|protected|internal)\s+", r"\1", s)      # member access modifiers
    return respell(s, enums, classes, structs) + ";"


def depends_on(block: str, other: str) -> bool:
    return re.search(r"\b%s\b" % other, block) is not None


def header(files: dict, includes) -> list:
    out = ["// GENERATED by oracle/cs2cpp.py -- the reference's own statements, declarations respelt for C++.  Never commit.",
           "#pragma once"] + ['#include "%s"' % i for i in includes] + [""]
    for rel, text in files.items():
        out.append("// source: %s  sha256 %s" % (rel, hashlib.sha256(text.encode("utf-8")).hexdigest()))
    return out + [""]


def transliterate(files: dict, includes=("ref_shim.hpp",), known_namespaces=(), known_enums=(), external_classes=(),
                  known_structs=(), only=None, fragments=()) -> str:
    """files: {relative path: C# text} -> the C++ header text.
    only:       {file: set of top-level type names to emit}; a file that is not a key is emitted whole
    fragments:  [(file, class name, regex of the statement's head, C++ function signature, {regex: replacement} extra rules)]:
                the braced statement of `class name` that starts at the regex, emitted as the body of that function"""
    ns_name, items, usings = None, [], []
    for rel, text in files.items():
        ns, us, blocks = split_top_level(text)
        ns_name = ns_name or ns
        if ns != ns_name:
            raise ValueError("sources of more than one namespace")
        usings += [u for u in us if u not in usings]
        items += [(k, n, b, rel) for k, n, b in blocks]
    enums = {n for k, n, _, _ in items if k == "enum"} | set(known_enums)
    classes = {n for k, n, _, _ in items if k == "class"} | set(external_classes)
    structs = {n for k, n, _, _ in items if k == "struct"} | set(known_structs)
    emit = [it for it in items if only is None or it[3] not in only or it[1] in only[it[3]]]
    ordered = [it for it in emit if it[0] == "enum"] + [it for it in emit if it[0] == "struct"]
    pending = [it for it in emit if it[0] == "class"]
    while pending:                                        # classes: a type after every class its text names
        for it in pending:
            if not any(o is not it and depends_on(it[2], o[1]) for o in pending):
                ordered.append(it)
                pending.remove(it)
                break
        else:
            raise ValueError("cyclic dependency between classes: " + ", ".join(p[1] for p in pending))
    out = header(files, includes)
    out += ["namespace %s {" % ns_name.replace(".", "_"), "using namespace bcl;"]
    out += ["using namespace %s;" % u.replace(".", "_") for u in usings if u in known_namespaces]
    out.append("")
    for kind, name, block, rel in ordered:
        out.append("// ---- %s %s (%s)" % (kind, name, rel))
        out.append("    " + translate_block(kind, name, block, enums, classes, structs))
        out.append("")
    for rel, cname, head, signature, extra in fragments:
        (block,) = [b for k, n, b, r in items if r == rel and n == cname]
        m = re.search(head, block)
        if not m or block[m.end() - 1] != "{":
            raise ValueError("fragment %r not found in %s" % (head, cname))
        body = respell(block[m.start():match_brace(block, m.end())], enums, classes, structs)
        for pat, rep in extra.items():
            body = re.sub(pat, rep, body)
        out.append("// ---- the statement `%s` of %s (%s), as the body of a function over the locals it reads" % (head, cname, rel))
        out += ["    " + signature + " {", "                    " + body, "        return true;", "    }", ""]
    out.append("}  // namespace")
    return "\n".join(out) + "\n"


def read_sources(ref_root: str, rels) -> dict:
    files = {}
    for rel in rels:
        with open(os.path.join(ref_root, rel), encoding="utf-8-sig") as f:
            files[rel] = f.read()
    return files


# the k-fold loop of Experiment.runKFoldCrossValidation reads these locals of the method (Experiment.cs:37-39, :46, :49, :61-66)
FOLD_LOOP_SIGNATURE = ("inline bool runFolds(std::string dbFile, int nFolds, int nIterations, Methodology methodology, "
                       "Dictionary<EvaluationMetric, double> finalResult, List<EvaluationMetric> metrics, int& cntLikes)")


def transliterate_rwr(ref_root: str) -> str:
    return transliterate(read_sources(ref_root, SOURCES))


def transliterate_callers(ref_root: str) -> str:
    return transliterate(
        read_sources(ref_root, CALLER_SOURCES), includes=("ref_shim.hpp", "reference_rwr.hpp"),
        known_namespaces=("Recommenders.RWRBased",), known_enums=("NodeType", "EdgeType"),
        external_classes=("SQLiteAdapter", "Graph", "Model", "Recommender"), known_structs=("Node", "ForwardLink"),
        only={"TweetRecommender/Experiment.cs": {"Methodology", "Feature", "EvaluationMetric", "ThreadParams"}},
        fragments=[("TweetRecommender/Experiment.cs", "Experiment", r"for \(int fold = 0; fold < nFolds; fold\+\+\) \{",
                    FOLD_LOOP_SIGNATURE, {r"\breturn;": "return false;"})])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out-dir", default=os.path.join(HERE, "_ref"), help="where the generated headers go (the Makefile passes a temporary directory)")
    ap.add_argument("--hashes", default=None, help="write `sha256  file` of every source read (kept next to libref.so)")
    args = ap.parse_args()
    os.makedirs(args.out_dir, exist_ok=True)
    for name, text in (("reference_rwr.hpp", transliterate_rwr(args.ref)), ("reference_experiment.hpp", transliterate_callers(args.ref))):
        with open(os.path.join(args.out_dir, name), "w") as f:
            f.write(text)
        print("wrote %s (%d lines)" % (os.path.join(args.out_dir, name), text.count("\n")))
    if args.hashes:
        with open(args.hashes, "w") as f:
            for rel, text in {**read_sources(args.ref, SOURCES), **read_sources(args.ref, CALLER_SOURCES)}.items():
                f.write("%s  %s\n" % (hashlib.sha256(text.encode("utf-8")).hexdigest(), rel))


if __name__ == "__main__":
    main()
