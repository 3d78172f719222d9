#!/usr/bin/env python
"""Instruction mix of the main (unrolled) group loop of a plate-kernel object file, per observation.

    python scripts/count_fp64_sass.py juliabugs.jl_b200/csrc/spec_seeds_128.o --obs-per-iter 2 [--write seeds]

The loop is found as the innermost backward branch that contains MUFU.RCP64H / the long DFMA runs (largest FP64
density).  Prints counts by class; --write NAME stores {"fp64_per_obs": ..} in profiles/fp64_counts.json, which bench.py
uses for the FP64-pipe roofline of the logit plate kernels."""
import argparse
import json
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("--obs-per-iter", type=int, default=2)
    ap.add_argument("--write", default="")
    a = ap.parse_args()
    ins = sass(a.obj)
    addr = {ad: k for k, (ad, _) in enumerate(ins)}
    best = None
    for k, (ad, text) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= ad or tgt not in addr:
            continue
        body = [t for _, t in ins[addr[tgt]:k + 1]]
        fp = sum(1 for t in body if re.match(r"(@!?U?P\d+\s+)?D(FMA|ADD|MUL|SETP|MNMX)", t))
        if fp >= 20 and (best is None or fp / len(body) > best[0]):
            best = (fp / len(body), body)
    if best is None:
        sys.exit("no FP64-dense loop found")
    body = best[1]
    cls = Counter()
    for t in body:
        op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
        cls[op] += 1
    n = a.obs_per_iter
    fp64 = sum(v for k, v in cls.items() if k in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
    print(f"loop body: {len(body)} instructions for {n} observations = {len(body) / n:.1f} per observation")
    print(f"FP64 pipe: {fp64} = {fp64 / n:.1f} per observation  ({ {k: cls[k] for k in ('DFMA', 'DADD', 'DMUL', 'DSETP') if cls[k]} })")
    print("other:", {k: v for k, v in cls.most_common() if k not in ("DFMA", "DADD", "DMUL", "DSETP")})
    if a.write:
        path = os.path.join(ROOT, "profiles", "fp64_counts.json")
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[a.write] = {"fp64_per_obs": fp64 / n, "issued_per_obs": len(body) / n,
                      "source": f"cuobjdump -sass {os.path.relpath(a.obj, ROOT)}: main group loop ({n} observations per iteration), "
                                "scripts/count_fp64_sass.py"}
        json.dump(d, open(path, "w"), indent=1)
        print("wrote", path)


if __name__ == "__main__":
    main()
