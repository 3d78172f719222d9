#!/usr/bin/env python
"""Golden kernel plans: the byte-exact output of the lowering pass for every fixture model.

    python tests/golden/make_plans.py          # rewrites tests/golden/plans/*.bin and manifest.json

`<name>.bin` is the plan blob (`include/jbugs_plan.h`: header + records) exactly as `jbugs_plan_create` receives it;
`manifest.json` holds, per model, the sha256 of the blob, the theta labels in slot order, and for every data slot its
name, dtype, length and the sha256 of its bytes (the arrays travel separately, `jbugs_plan_upload_data`).

Purpose: a second implementation of the pass -- the Julia one in `juliabugs.jl_b200/julia/lower_to_plan.jl`, which a
machine with Julia can run -- is checked byte for byte against these files (`julia/check_golden_plans.jl`), and
`tests/test_golden_plans.py` keeps the Python pass from drifting.  Both orders of the reference are covered for Rats
(generated-mode order = the default here, and UseGraph's order passed as `theta_order`)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "juliabugs.jl_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def fixtures():
    from conftest import _arr
    raw = json.load(open(os.path.join(HERE, "examples_vol1.json")))
    raw.update(json.load(open(os.path.join(HERE, "examples_vol2.json"))))
    return {name: {k: ({kk: _arr(vv) for kk, vv in v.items()} if isinstance(v, dict) and k != "golden_logjoint" else v)
                   for k, v in entry.items()} for name, entry in raw.items()}


def cases():
    """(case name, model text, data, kwargs of jbugs.lower)"""
    import jbugs
    from helpers import ZOO, synth_rats, synth_seeds
    fx = fixtures()
    for name, text, key, _ in ZOO:
        yield name, text, fx[key]["data"], {}
    yield "rats_untransformed", jbugs.examples.RATS, fx["rats"]["data"], {"transformed": False}
    import graph_oracle as go
    gm = go.compile(jbugs.examples.RATS, fx["rats"]["data"], fx["rats"]["inits"])
    yield "rats_graph_order", jbugs.examples.RATS, fx["rats"]["data"], {"theta_order": gm.param_names()}
    yield "rats_synth_2000", jbugs.examples.RATS, synth_rats(2000)[0], {}
    yield "seeds_synth_5000", jbugs.examples.SEEDS, synth_seeds(5000)[0], {}
    rng = np.random.default_rng(20)
    X = np.asfortranarray(rng.standard_normal((300, 7)))
    y = (rng.random(300) < 0.5).astype(float)
    yield "dense_logistic_300x7", jbugs.examples.DENSE_LOGISTIC, {"N": 300, "P": 7, "X": X, "y": y}, {}
    # marginalised plates (has_obs == 3): the reference's mixture test models (J/test/model/auto_marginalization.jl)
    import test_marginalization as tm
    yield "gmm2_marginalised", tm.GMM2 % (0.3, 0.7), {"N": 4, "y": np.array([-1.5, 2.3, -2.1, 1.8])}, {"marginalize": True}
    yield "gmm3_marginalised", tm.GMM3, {"N": 3, "y": np.array([-2.5, 0.5, 3.2])}, {"marginalize": True}
    yield "gmm_partially_observed_latent", tm.GMM_SHARED, {"N": 4, "y": np.array([1.0, 2.0, -1.0, 3.0]),
                                                          "z": np.array([2.0, np.nan, 1.0, np.nan])}, {"marginalize": True}
    # Bernoulli latents: in arithmetic, behind a derived subscript and partially observed; Cervix (vol. 2) at full size
    # -- two observed children per latent, a parameter array picked by a latent-derived and a data subscript
    rng = np.random.default_rng(23)
    y5 = 2.0 * rng.standard_normal(5)
    yield "bern_latent_arith", tm.BERN_ARITH, {"N": 5, "y": y5}, {"marginalize": True}
    yield "bern_latent_subscript_partially_observed", tm.BERN_SUBSCRIPT, {
        "N": 5, "y": y5, "q": 0.1 + 0.8 * rng.random(5), "z": np.array([1.0, np.nan, 0.0, np.nan, np.nan])}, {"marginalize": True}
    yield "cervix_marginalised", jbugs.examples.CERVIX, tm._cervix_data(), {"marginalize": True}
    # truncated priors, T(lower, upper): every bijector the bounds can pick, and the families whose mass needs the
    # incomplete gamma / beta functions (include/jbugs_trunc.h)
    yield "truncated_priors", """model {
      s ~ dnorm(0, 0.25) T(0, )
      m ~ dnorm(1, 4) T(-1, 2.5)
      u ~ dlogis(0, 1) T(, 0.5)
      g ~ dgamma(2.5, 1.5) T(0.4, 3.0)
      b ~ dbeta(2.5, 4.0) T(0.2, 0.9)
      h ~ dt(0, 1.0, 3) T(0, )
      for (i in 1:N) { y[i] ~ dnorm(m + u + b * x[i], s + g + h) }
    }""", {"N": 6, "y": np.array([0.3, 1.1, -0.4, 2.0, 0.8, 1.5]), "x": np.linspace(-1.0, 1.0, 6)}, {}


def record(plan, lw):
    blob = plan.to_bytes()
    slots = []
    for s in plan.data:
        a = np.ascontiguousarray(s.values)
        slots.append({"name": s.name, "dtype": "int64" if a.dtype == np.int64 else "float64", "len": int(a.size),
                      "rank": int(s.rank), "dim0": int(s.dim0 or a.size), "sha256": hashlib.sha256(a.tobytes()).hexdigest()})
    names = lw.param_names() if plan.D <= 20000 else None
    return blob, {"sha256": hashlib.sha256(blob).hexdigest(), "nbytes": len(blob), "D": int(plan.D),
                  "n_plates": len(plan.plates), "n_xops": sum(len(p.xops) for p in plan.plates),
                  "theta_labels": names, "data_slots": slots}


def main(write=True):
    import jbugs
    out = {}
    dst = os.path.join(HERE, "plans")
    os.makedirs(dst, exist_ok=True)
    for name, text, data, kw in cases():
        plan, lw = jbugs.lower(text, data, **kw)
        blob, rec = record(plan, lw)
        out[name] = rec
        if write:
            with open(os.path.join(dst, name + ".bin"), "wb") as f:
                f.write(blob)
    if write:
        with open(os.path.join(dst, "manifest.json"), "w") as f:
            json.dump(out, f, indent=1)
        print("wrote", len(out), "plans,", sum(r["nbytes"] for r in out.values()), "bytes")
    return out


if __name__ == "__main__":
    main()
