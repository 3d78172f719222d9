#!/usr/bin/env python
"""Extract the data / inits tuples of the reference's own example fixtures into JSON.

Run HERE (build container, where /root/reference is mounted):

    python tests/golden/make_fixtures.py

It reads the Julia named-tuple literals `data = (...)`, `inits = (...)`,
`inits_alternative = (...)` (and their `_realistic` / `_simplistic` variants) from
`/root/reference/JuliaBUGS/src/BUGSExamples/Volume_1/*.jl` and writes
`tests/golden/examples_vol1.json`.  Only numbers are extracted; the model text used by the
tests is written independently in `juliabugs.jl_b200/jbugs/examples.py`.
The golden log-joint constants pinned by the reference's tests
(`JuliaBUGS/test/model/evaluation.jl:355-379`) are recorded in the same file.

Matrices are stored row-major as nested lists (rows of the Julia literal); `missing` -> null.
"""
import json
import os
import re
import sys

REF = "/root/reference/JuliaBUGS/src/BUGSExamples/Volume_1"
FILES = {
    "rats": "01_Rats.jl",
    "pumps": "02_Pumps.jl",
    "dogs": "03_Dogs.jl",
    "seeds": "04_Seeds.jl",
    "surgical": "05_Surgical.jl",
    "epil": "11_Epil.jl",
    "blockers": "12_Blocker.jl",
    "bones": "15_Bones.jl",
    "salm": "07_Salm.jl",
    "equiv": "08_Equiv.jl",
    "dyes": "09_Dyes.jl",
    "stacks": "10_Stacks.jl",
    "oxford": "13_Oxford.jl",
    "lsat": "14_LSAT.jl",
    "magnesium": "06_Magnesium.jl",
    "inhalers": "16_Inhalers.jl",
    "mice": "17_Mice.jl",
    "kidney": "18_Kidney.jl",
    "leuk": "19_Leuk.jl",
    "leukfr": "20_LeukFr.jl",
}
# Volume 2 examples used by the expression-plate tests (written to examples_vol2.json)
REF2 = "/root/reference/JuliaBUGS/src/BUGSExamples/Volume_2"
FILES2 = {"dugongs": "01_Dugongs.jl", "air": "07_Air.jl", "orange": "02_Orange_trees.jl", "cervix": "08_Cervix.jl"}
TUPLES = re.compile(r"^(data|inits\w*|inits_alternative\w*)\s*=\s*\(", re.M)

# golden constants: JuliaBUGS/test/model/evaluation.jl:355-379 (rtol 1e-6 in the reference test)
GOLDEN_LOGJOINT = {
    "rats": {"transformed": -174029.38703951868, "untransformed": -174029.38703951868},
    "blockers": {"transformed": -8417.497576569103, "untransformed": -8417.497576569103},
    "bones": {"transformed": -139.03545202758394, "untransformed": -139.03545202758394},
    "dogs": {"transformed": -1243.3996613167667, "untransformed": -1243.188922285352},
}


def balanced(text, start):
    """text[start] == '(' -> index just past the matching ')'."""
    depth, i, in_str = 0, start, False
    while i < len(text):
        c = text[i]
        if in_str:
            if c == '"':
                in_str = False
        elif c == '"':
            in_str = True
        elif c == "#":
            while i < len(text) and text[i] != "\n":
                i += 1
            continue
        elif c in "([":
            depth += 1
        elif c in ")]":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced")


def strip_comments(s):
    return re.sub(r"#[^\n]*", "", s)


def split_top(s, sep=","):
    out, depth, cur, in_str = [], 0, "", False
    for c in s:
        if c == '"':
            in_str = not in_str
        if not in_str:
            if c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
            elif c == sep and depth == 0:
                out.append(cur)
                cur = ""
                continue
        cur += c
    if cur.strip():
        out.append(cur)
    return out


def number(tok):
    tok = tok.strip()
    if tok in ("missing", "NaN"):
        return None
    tok = tok.replace("E", "e")
    if re.fullmatch(r"[-+]?\d+", tok):
        return int(tok)
    return float(tok)


def value(src):
    src = src.strip()
    m = re.fullmatch(r"fill\(\s*([^,]+),\s*(\d+)\s*\)", src)
    if m:
        return [number(m.group(1))] * int(m.group(2))
    m = re.fullmatch(r"zeros\(\s*(?:\w+\s*,\s*)?(\d+)\s*\)", src)
    if m:
        return [0] * int(m.group(1))
    if src.startswith("["):
        body = src[1:-1]
        if "," in body and not re.search(r"\d\s+[-\d.m]", body.replace(",", " , ").replace(" , ", ",")):
            return [number(t) for t in body.replace("\n", " ").split(",") if t.strip()]
        if "," in body:
            return [number(t) for t in re.split(r"[,\s]+", body.strip()) if t.strip()]
        rows = [r for r in re.split(r"[;\n]", body) if r.strip()]
        mat = [[number(t) for t in r.split()] for r in rows]
        return mat[0] if len(mat) == 1 else mat
    if src.startswith("("):
        return named_tuple(src)
    return number(src)


def named_tuple(src):
    src = strip_comments(src.strip())
    assert src[0] == "(" and src[-1] == ")", src[:40]
    out = {}
    for item in split_top(src[1:-1]):
        if not item.strip():
            continue
        k, v = item.split("=", 1)
        k = k.strip()
        m = re.fullmatch(r'var"([^"]+)"', k)
        if m:
            k = m.group(1)
        out[k] = value(v)
    return out


def extract(ref, files, volume):
    out = {}
    for name, fn in files.items():
        path = os.path.join(ref, fn)
        if not os.path.exists(path):
            continue
        text = open(path).read()
        entry = {"source": f"JuliaBUGS/src/BUGSExamples/{volume}/{fn}"}
        for m in TUPLES.finditer(text):
            end = balanced(text, m.end() - 1)
            try:
                entry[m.group(1)] = named_tuple(text[m.end() - 1:end])
            except Exception as e:  # keep going: not every example is needed by the hot path
                entry[m.group(1)] = {"__unparsed__": repr(e)}
        if name in GOLDEN_LOGJOINT:
            entry["golden_logjoint"] = GOLDEN_LOGJOINT[name]
        out[name] = entry
    return out


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not mounted; fixtures can only be regenerated in the build container")
    here = os.path.dirname(os.path.abspath(__file__))
    for ref, files, volume, fn in ((REF, FILES, "Volume_1", "examples_vol1.json"), (REF2, FILES2, "Volume_2", "examples_vol2.json")):
        out = extract(ref, files, volume)
        dst = os.path.join(here, fn)
        with open(dst, "w") as f:
            json.dump(out, f, separators=(",", ":"))
        print("wrote", dst, os.path.getsize(dst), "bytes;", ", ".join(out))


if __name__ == "__main__":
    main()
