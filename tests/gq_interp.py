"""Test infrastructure: a numpy reading of a lowered generated-quantities program (include/jbugs_gq.h) -- deterministic
statements only -- so that the lowering (jbugs/gq_device.py) can be checked on a box without a GPU.  Mirrors the
statement semantics of the kernel op for op; never imported by the product."""
import numpy as np

from jbugs import plan as P
from jbugs.gq_device import (LEAF_CONST, LEAF_DATA_CELL, LEAF_Q, LEAF_Q_CELL, LEAF_THETA, LEAF_THETA_CELL, LOGICAL)

_UN = {P.OP_IDENTITY: lambda x: x, P.OP_EXP: np.exp, P.OP_LOGISTIC: lambda x: 1.0 / (1.0 + np.exp(-x)),
       P.OP_CEXPEXP: lambda x: -np.expm1(-np.exp(x)), P.OP_LOG: np.log, P.OP_SQRT: np.sqrt, P.OP_NEG: lambda x: -x,
       40: np.abs, 41: lambda x: (x >= 0).astype(float), 42: lambda x: np.log(x / (1 - x)), 43: lambda x: np.log(-np.log1p(-x)),
       45: np.sin, 46: np.cos, 47: np.round, 48: np.trunc}
_BIN = {P.OP_ADD: np.add, P.OP_SUB: np.subtract, P.OP_MUL: np.multiply, P.OP_DIV: np.divide, P.OP_POW: np.power,
        60: lambda a, b: (a == b).astype(float), 61: np.maximum, 62: np.minimum}


def constrain(tr, lb, ub, y):
    if tr == P.TR_LOG:
        return lb + np.exp(y)
    if tr == P.TR_UPPER:
        return ub - np.exp(y)
    if tr == P.TR_LOGIT:
        return lb + (ub - lb) / (1.0 + np.exp(-y))
    return y


def run_program(program, theta):
    """theta: [D, S] unconstrained.  Returns Q [n_rows, S] with NaN rows for stochastic statements."""
    from scipy.special import ndtr
    st, ops, ip, dp = program.arrays()
    S = theta.shape[1]
    Q = np.full((program.n_rows, S), np.nan)
    x = np.stack([constrain(program.tr[d], program.lb[d], program.ub[d], theta[d]) for d in range(program.D)]) if program.D else theta
    for s in st:
        if s["family"] != LOGICAL:
            continue
        for cell in range(int(s["n_cells"])):
            v = []
            for o in ops[s["first_op"]:s["first_op"] + s["n_ops"]]:
                op = int(o["op"])
                if op == P.OP_LEAF:
                    k, off = int(o["leaf_kind"]), int(o["offset"])
                    if k == LEAF_CONST:
                        v.append(np.full(S, o["value"]))
                    elif k == LEAF_THETA:
                        v.append(x[off])
                    elif k == LEAF_THETA_CELL:
                        v.append(x[ip[off + cell]])
                    elif k == LEAF_Q:
                        v.append(Q[off])
                    elif k == LEAF_Q_CELL:
                        v.append(Q[ip[off + cell]])
                    elif k == LEAF_DATA_CELL:
                        v.append(np.full(S, dp[off + cell]))
                elif op == P.OP_PHI:
                    v.append(ndtr(v[o["a"]]))
                elif op in _UN:
                    with np.errstate(all="ignore"):
                        v.append(_UN[op](v[o["a"]]))
                else:
                    with np.errstate(all="ignore"):
                        v.append(_BIN[op](v[o["a"]], v[o["b"]]))
            Q[int(s["out_row0"]) + cell] = v[-1]
    return Q
