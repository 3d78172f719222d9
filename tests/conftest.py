import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "juliabugs.jl_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _arr(v):
    if isinstance(v, list):
        return np.array([[np.nan if t is None else t for t in r] if isinstance(r, list) else (np.nan if r is None else r)
                         for r in v], dtype=float)
    return v


@pytest.fixture(scope="session")
def fixtures():
    """data / inits of the reference's own examples (tests/golden/make_fixtures.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "examples_vol1.json")) as f:
        raw = json.load(f)
    out = {}
    for name, entry in raw.items():
        out[name] = {k: ({kk: _arr(vv) for kk, vv in v.items()} if isinstance(v, dict) and k != "golden_logjoint" else v)
                     for k, v in entry.items()}
    return out


def rel_err(a, b, floor=1e-300):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if not a.size:
        return 0.0
    same = (a == b) | (np.isnan(a) & np.isnan(b))  # identical values, including equal infinities (-Inf log densities)
    with np.errstate(invalid="ignore"):
        err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    return float(np.max(np.where(same, 0.0, err)))
