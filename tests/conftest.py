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
    with open(os.path.join(ROOT, "tests", "golden", "examples_vol2.json")) as f:
        raw.update(json.load(f))
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


def grad_err(g, g0):
    """max relative gradient error per chain.  A component that is a sum of many terms of mixed sign (the
    hyper-parameter slots: tau * sum_i (alpha_i - alpha.c)) can cancel to far below the size of its summands, and two
    correct evaluations that add in a different order then differ by eps * sum|terms|, not eps * |result|; so components
    are measured against max(|g_d|, 1e-4 * max_d |g_d|) of their chain.  The ONE scale used by every gradient parity
    test, CPU and GPU, against the plan-level C restatement and against the independent per-node oracle."""
    g = np.asarray(g, dtype=float)
    g0 = np.asarray(g0, dtype=float)
    if g.ndim == 1:
        g, g0 = g[:, None], g0[:, None]
    if not g0.size:
        return 0.0
    scale = np.maximum(np.abs(g0), 1e-4 * np.max(np.abs(g0), axis=0, keepdims=True))
    scale = np.maximum(scale, 1e-300)
    return float(np.max(np.abs(g - g0) / scale))
