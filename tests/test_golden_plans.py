"""The lowering pass is pinned byte for byte: re-lowering every fixture must reproduce the committed plan blob
(tests/golden/plans/*.bin, written by tests/golden/make_plans.py) and the digests of the data slots.  The same files are
what the Julia implementation of the pass (juliabugs.jl_b200/julia/lower_to_plan.jl) is checked against on a machine that
has Julia (julia/check_golden_plans.jl) -- SURVEY section 7 step 1: "the pass written twice with byte-identical plans"."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_plans  # noqa: E402

MANIFEST = json.load(open(os.path.join(HERE, "golden", "plans", "manifest.json")))
CASES = {name: (text, data, kw) for name, text, data, kw in make_plans.cases()}


def test_manifest_covers_every_case():
    assert set(MANIFEST) == set(CASES) and len(CASES) >= 20


@pytest.mark.parametrize("name", sorted(CASES))
def test_relowering_reproduces_the_golden_plan(name):
    import jbugs
    text, data, kw = CASES[name]
    plan, lw = jbugs.lower(text, data, **kw)
    blob, rec = make_plans.record(plan, lw)
    want = MANIFEST[name]
    gold = open(os.path.join(HERE, "golden", "plans", name + ".bin"), "rb").read()
    assert hashlib.sha256(gold).hexdigest() == want["sha256"]
    assert blob == gold, f"{name}: the lowering no longer emits the committed plan (run tests/golden/make_plans.py if intended)"
    assert rec["theta_labels"] == want["theta_labels"]
    assert [(s["dtype"], s["len"], s["sha256"]) for s in rec["data_slots"]] == \
           [(s["dtype"], s["len"], s["sha256"]) for s in want["data_slots"]]


def test_plan_header_fields_of_a_golden_blob():
    """the blob layout a second implementation has to hit: 64-byte header, then the record arrays in header order"""
    from jbugs import plan as P
    gold = open(os.path.join(HERE, "golden", "plans", "bones.bin"), "rb").read()
    h = np.frombuffer(gold[:64], dtype=P.HEADER_DT)[0]
    assert h["magic"] == b"JBUGSPL1" and h["version"] == 1 and h["D"] == 13 and h["n_plates"] == 1 and h["n_xops"] == 32
    need = (64 + 80 * h["n_globals"] + 40 * h["n_gtape"] + 24 * h["n_data"] + 128 * h["n_plates"] + 104 * h["n_locals"]
            + 32 * h["n_terms"] + 56 * h["n_dense"] + 32 * h["n_xops"])
    assert need == len(gold)
