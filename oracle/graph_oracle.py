"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Per-node CPU restatement of the reference's hot path, written from the reference's
*behaviour* (no Julia is installed here, the reference cannot be executed; see
DESIGN.md "oracle").  It follows, function by function:

* `compile`                      JuliaBUGS/src/JuliaBUGS.jl:328-371
    - loop unrolling             JuliaBUGS/src/compiler_pass.jl:6-45      (`analyze_block`)
    - array sizes                JuliaBUGS/src/compiler_pass.jl:61-136    (`CollectVariables`)
    - transformed data           JuliaBUGS/src/compiler_pass.jl:497-547 + JuliaBUGS.jl:175-187
    - vertices / edges           JuliaBUGS/src/compiler_pass.jl:816-932   (`AddVertices`, `AddEdges`)
    - dependency extraction      JuliaBUGS/src/compiler_pass.jl:589-657
* node order                     JuliaBUGS/src/model/bugsmodel.jl:156-158 -> Graphs.jl
                                 `topological_sort_by_dfs` (third party, Graphs 1.x: iterate
                                 vertices ascending, iterative DFS taking the first unvisited
                                 out-neighbour, emit on finish, reverse)
* classification / GQ discovery  JuliaBUGS/src/model/bugsmodel.jl:48-69,154-246, graphs.jl:27-71
* theta layout, getparams        JuliaBUGS/src/model/bugsmodel.jl:400-475,619-665
* the node loop                  JuliaBUGS/src/model/evaluation.jl:217-275 (`evaluate_with_values!!`)
* generated-mode program order   JuliaBUGS/src/source_gen.jl:5-223,459-552,864-920 and
                                 bugsmodel.jl:770-799 (`CollectSortedNodes` over the reconstructed program)
* boundary semantics             JuliaBUGS/src/model/logdensityproblems.jl:19-27,219-232

Gradients are obtained the way the reference's `AutoForwardDiff` backend obtains them:
by pushing dual numbers through the very same node loop.

Pure-Python loops: intended for small cases only (Rats 30x5 = 367 vertices).
Parity pin: checked against the golden log-joint constants of
`JuliaBUGS/test/model/evaluation.jl:355-379` and the theta-order example at `:131-141`
(tests/test_oracle_goldens.py).  Gradient values are not pinned by any reference test.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
# the oracle reads BUGS text with its own front end (oracle/bugs_ast.py): nothing under oracle/ imports the product
from bugs_ast import (  # noqa: E402
    BinOp, Call, Colon, Det, For, Index, Neg, Num, Range, Stoch, Var, parse,
)

import dists  # noqa: E402
from numeric import FLOAT_OR_DUAL, DomainError, Dual, _val  # noqa: E402

MISSING = None


def _is_missing(x):
    return x is None or (isinstance(x, float) and math.isnan(x))


# ----------------------------------------------------------------------------- expression evaluation


def _to_int(x):
    v = _val(x)
    if float(v) != int(v):
        raise ValueError(f"non-integer index {v}")
    return int(v)


class Evaluator:
    """Evaluates AST expressions against an environment (the node function of
    `make_function_expr`, compiler_pass.jl:751-814, including the Int() cast on indices)."""

    def __init__(self, M=FLOAT_OR_DUAL):
        self.M = M

    def ev(self, e, env, lv):
        M = self.M
        if isinstance(e, Num):
            return e.value
        if isinstance(e, Var):
            if e.name in lv:
                return lv[e.name]
            v = env[e.name]
            if isinstance(v, np.ndarray):
                if v.ndim == 0:
                    return v[()]
                return v
            return v
        if isinstance(e, Index):
            arr = env[e.name]
            idx = []
            for d, ie in enumerate(e.idx):
                if isinstance(ie, Colon):
                    idx.append(slice(None))
                elif isinstance(ie, Range):
                    lo = _to_int(self.ev(ie.lo, env, lv))
                    hi = _to_int(self.ev(ie.hi, env, lv))
                    idx.append(slice(lo - 1, hi))
                else:
                    idx.append(_to_int(self.ev(ie, env, lv)) - 1)
            return arr[tuple(idx)]
        if isinstance(e, Neg):
            return -self.ev(e.a, env, lv)
        if isinstance(e, BinOp):
            a = self.ev(e.a, env, lv)
            b = self.ev(e.b, env, lv)
            if _is_missing(a) or _is_missing(b):
                return MISSING
            if e.op == "+":
                return a + b
            if e.op == "-":
                return a - b
            if e.op == "*":
                return a * b
            if e.op == "/":
                return a / b
            if e.op == "^":
                return a**b
            raise ValueError(e.op)
        if isinstance(e, Call):
            args = [self.ev(a, env, lv) for a in e.args]
            for a in args:
                if _is_missing(a):
                    return MISSING
                if isinstance(a, np.ndarray) and a.dtype == object and any(_is_missing(t) for t in a.ravel()):
                    return MISSING
                if isinstance(a, np.ndarray) and a.dtype != object and np.isnan(a).any():
                    return MISSING
            return self.call(e.fn, args)
        raise TypeError(f"cannot evaluate {e!r}")

    def call(self, fn, a):
        M = self.M
        # BUGSPrimitives/functions.jl
        if fn == "exp":
            return M.exp(a[0])
        if fn == "log":
            return M.log(a[0])
        if fn == "sqrt":
            return M.sqrt(a[0])
        if fn in ("logistic", "ilogit"):
            return M.logistic(a[0])
        if fn == "logit":
            return M.log(a[0] / (1.0 - a[0]))
        if fn in ("cexpexp", "icloglog"):
            return M.cexpexp(a[0])
        if fn == "cloglog":
            return M.log(-M.log(1.0 - a[0]))
        if fn == "phi":
            return M.phi(a[0])
        if fn == "pow":
            return a[0] ** a[1]
        if fn == "abs":
            return M.abs(a[0])
        if fn == "step":  # functions.jl: _step(x) = ifelse(x >= 0, 1, 0)
            return 1.0 if _val(a[0]) >= 0 else 0.0
        if fn == "equals":
            return 1.0 if _val(a[0]) == _val(a[1]) else 0.0
        if fn == "max":
            return a[0] if _val(a[0]) >= _val(a[1]) else a[1]
        if fn == "min":
            return a[0] if _val(a[0]) <= _val(a[1]) else a[1]
        if fn == "sum":
            return _sum(list(np.asarray(a[0], dtype=object).ravel()))
        if fn == "mean":
            xs = list(np.asarray(a[0], dtype=object).ravel())
            return _sum(xs) / len(xs)
        if fn == "sd":  # BUGSPrimitives/functions.jl:132-134: Statistics.std (corrected, n - 1)
            xs = list(np.asarray(a[0], dtype=object).ravel())
            m = _sum(xs) / len(xs)
            return M.sqrt(_sum([(x - m) * (x - m) for x in xs]) / (len(xs) - 1))
        if fn == "inprod":  # functions.jl:33-35: dot product
            xs = list(np.asarray(a[0], dtype=object).ravel())
            ys = list(np.asarray(a[1], dtype=object).ravel())
            return _sum([x * y for x, y in zip(xs, ys)])
        if fn == "loggam":
            return M.lgamma(a[0])
        if fn == "logfact":
            return M.lgamma(a[0] + 1.0)
        if fn == "sin":
            return M.sin(a[0])
        if fn == "cos":
            return M.cos(a[0])
        raise NotImplementedError(f"function {fn} is outside the restated subset")


def _sum(xs):
    # pairwise-free left fold (Julia's `sum` is pairwise above 1024 elements; below that a
    # simple loop -- every fixture here is far below the threshold)
    s = 0.0
    for x in xs:
        s = s + x
    return s


# ----------------------------------------------------------------------------- compile


class Node:
    __slots__ = ("name", "idx", "stmt", "is_stochastic", "is_observed", "lv", "vid")

    def __init__(self, name, idx, stmt, is_stochastic, is_observed, lv, vid):
        self.name, self.idx, self.stmt = name, idx, stmt
        self.is_stochastic, self.is_observed, self.lv, self.vid = is_stochastic, is_observed, lv, vid

    @property
    def label(self):
        if self.idx == ():
            return self.name
        return f"{self.name}[{','.join(str(i) for i in self.idx)}]"

    def __repr__(self):
        return self.label


def _unroll(stmts, env, lv, fn, ev):
    """`analyze_block` (compiler_pass.jl:6-45): program order, loops fully unrolled,
    a loop whose lower bound exceeds its upper bound is skipped."""
    for s in stmts:
        if isinstance(s, For):
            lo = _to_int(ev.ev(s.lo, env, lv))
            hi = _to_int(ev.ev(s.hi, env, lv))
            for k in range(lo, hi + 1):
                lv2 = dict(lv)
                lv2[s.var] = k
                _unroll(s.body, env, lv2, fn, ev)
        else:
            fn(s, lv)


def _flatten(stmts, loops=(), out=None):
    """statements in program order with their loop nests (stmt ids of source_gen.jl:5-20)."""
    if out is None:
        out = []
    for s in stmts:
        if isinstance(s, For):
            _flatten(s.body, loops + ((s.var, s.lo, s.hi),), out)
        else:
            out.append((s, loops))
    return out


def _flatten_bounds(stmts):
    """censored(d(args...), lower, upper) -> a plain call `censored:d`(args..., lower, upper) with +-Inf for an empty
    bound, so that the bounds take part in evaluation and dependency tracking like any other distribution argument"""
    out = []
    for s in stmts:
        if isinstance(s, For):
            out.append(For(s.var, s.lo, s.hi, _flatten_bounds(s.body)))
        elif isinstance(s, Stoch) and s.dist.fn in ("censored", "truncated"):
            inner, lo, hi = s.dist.args
            lo = Num(-math.inf, False) if lo == Var("nothing") else lo
            hi = Num(math.inf, False) if hi == Var("nothing") else hi
            out.append(Stoch(s.lhs, Call(s.dist.fn + ":" + inner.fn, tuple(inner.args) + (lo, hi))))
        else:
            out.append(s)
    return tuple(out)


class GraphModel:
    """What `compile` hands to the evaluator: environment, graph, node order, theta layout."""

    def __init__(self, text_or_stmts, data, inits=None, transformed=True):
        self.stmts = _flatten_bounds(parse(text_or_stmts) if isinstance(text_or_stmts, str) else tuple(text_or_stmts))
        self.transformed = transformed
        self.ev = Evaluator(FLOAT_OR_DUAL)
        self._flat = _flatten(self.stmts)
        self._stmt_id = {id(s): k + 1 for k, (s, _) in enumerate(self._flat)}
        self._build_env(data)
        self._data_transformation()
        self._add_vertices()
        self._add_edges()
        self._fill_inits(inits or {})
        self.sorted_nodes = [self.nodes[v] for v in self._toposort(self.out, len(self.nodes))]
        self._classify()
        self.mode = "graph"

    # ---- environment ------------------------------------------------------------------------
    def _lhs_index(self, lhs, env, lv):
        if isinstance(lhs, Var):
            return lhs.name, ()
        idx = []
        for ie in lhs.idx:
            if isinstance(ie, Range):
                idx.append((_to_int(self.ev.ev(ie.lo, env, lv)), _to_int(self.ev.ev(ie.hi, env, lv))))
            elif isinstance(ie, Colon):
                raise NotImplementedError("colon on the lhs")
            else:
                idx.append(_to_int(self.ev.ev(ie, env, lv)))
        return lhs.name, tuple(idx)

    def _build_env(self, data):
        env = {}
        for k, v in data.items():
            if isinstance(v, (list, tuple)):
                v = np.asarray(v, dtype=float)
            if isinstance(v, np.ndarray):
                a = np.empty(v.shape, dtype=object)
                flat = v.astype(float).ravel()
                af = a.ravel()
                for t in range(flat.size):
                    af[t] = MISSING if math.isnan(flat[t]) else float(flat[t])
                env[k] = af.reshape(v.shape)
            else:
                env[k] = float(v)
        # CollectVariables: sizes of arrays that are not (fully) given as data
        sizes, scalars = {}, set()

        def collect(s, lv):
            name, idx = self._lhs_index(s.lhs, env, lv)
            if idx == ():
                scalars.add(name)
                return
            top = tuple(i[1] if isinstance(i, tuple) else i for i in idx)
            cur = sizes.get(name, (0,) * len(top))
            sizes[name] = tuple(max(a, b) for a, b in zip(cur, top))

        _unroll(self.stmts, env, {}, collect, self.ev)
        for name in scalars:
            env.setdefault(name, MISSING)
        for name, shp in sizes.items():
            if name in env:
                if tuple(env[name].shape) != shp and any(a < b for a, b in zip(env[name].shape, shp)):
                    raise ValueError(f"data array {name} is smaller than the model requires")
            else:
                env[name] = np.full(shp, MISSING, dtype=object)
        self.env = env
        self.data_names = set(data.keys())

    def _get(self, name, idx):
        if idx == ():
            return self.env[name]
        return self.env[name][tuple(i - 1 for i in idx)]

    def _set(self, env, name, idx, val):
        if idx == ():
            env[name] = val
        else:
            env[name][tuple(i - 1 for i in idx)] = val

    def _data_transformation(self):
        """Evaluate every logical statement whose rhs is computable from data, to a fixpoint
        (compiler_pass.jl:497-547 driven by JuliaBUGS.jl:175-187)."""
        self.transformed_data = set()
        changed = True
        while changed:
            changed = False

            def visit(s, lv):
                nonlocal changed
                if not isinstance(s, Det):
                    return
                name, idx = self._lhs_index(s.lhs, self.env, lv)
                if not _is_missing(self._get(name, idx)):
                    return
                try:
                    val = self.ev.ev(s.rhs, self.env, lv)
                except (TypeError, DomainError):
                    return
                if _is_missing(val) or isinstance(val, np.ndarray):
                    return
                self._set(self.env, name, idx, float(val))
                self.transformed_data.add((name, idx))
                changed = True

            _unroll(self.stmts, self.env, {}, visit, self.ev)

    # ---- graph ------------------------------------------------------------------------------
    def _add_vertices(self):
        self.nodes = []
        self.vid = {}

        def visit(s, lv):
            name, idx = self._lhs_index(s.lhs, self.env, lv)
            if any(isinstance(i, tuple) for i in idx):
                raise NotImplementedError("multivariate nodes are outside the hot-path scope")
            resolved = not _is_missing(self._get(name, idx))
            if isinstance(s, Det):
                if resolved:
                    return  # transformed data: no vertex
                st, ob = False, False
            else:
                st, ob = True, resolved
            n = Node(name, idx, s, st, ob, dict(lv), len(self.nodes))
            self.nodes.append(n)
            self.vid[(name, idx)] = n.vid

        _unroll(self.stmts, self.env, {}, visit, self.ev)

    def _deps(self, e, lv, acc):
        """`evaluate_and_track_dependencies` (compiler_pass.jl:589-657): returns the value if it
        is resolvable from the environment, else MISSING, and appends (name, index-ranges)."""
        env = self.env
        if isinstance(e, Num):
            return e.value
        if isinstance(e, Var):
            if e.name in lv:
                return lv[e.name]
            v = env[e.name]
            if isinstance(v, np.ndarray):
                raise ValueError(f"{e.name} is an array but referenced as a scalar")
            if _is_missing(v):
                acc.append((e.name, ()))
                return MISSING
            return v
        if isinstance(e, Range):
            lo, hi = self._deps(e.lo, lv, acc), self._deps(e.hi, lv, acc)
            if _is_missing(lo) or _is_missing(hi):
                return MISSING
            return (_to_int(lo), _to_int(hi))
        if isinstance(e, Index):
            arr = env[e.name]
            idx = []
            for d, ie in enumerate(e.idx):
                if isinstance(ie, Colon):
                    idx.append((1, arr.shape[d]))
                else:
                    idx.append(self._deps(ie, lv, acc))
            if all(not _is_missing(i) for i in idx):
                idx = [i if isinstance(i, tuple) else _to_int(i) for i in idx]
                sl = tuple(slice(i[0] - 1, i[1]) if isinstance(i, tuple) else i - 1 for i in idx)
                val = arr[sl]
                if isinstance(val, np.ndarray):
                    if all(not _is_missing(t) for t in val.ravel()):
                        return val
                elif not _is_missing(val):
                    return val
                acc.append((e.name, tuple(idx)))
            else:
                acc.append(
                    (e.name, tuple((1, arr.shape[d]) if _is_missing(i) else (i if isinstance(i, tuple) else _to_int(i))
                                   for d, i in enumerate(idx)))
                )
            return MISSING
        if isinstance(e, Neg):
            a = self._deps(e.a, lv, acc)
            return MISSING if _is_missing(a) else -a
        if isinstance(e, BinOp):
            a, b = self._deps(e.a, lv, acc), self._deps(e.b, lv, acc)
            if _is_missing(a) or _is_missing(b):
                return MISSING
            return _binop(e.op, a, b)
        if isinstance(e, Call):
            args = [self._deps(a, lv, acc) for a in e.args]
            if any(_is_missing(a) for a in args):
                return MISSING
            try:
                return self.ev.call(e.fn, args)
            except NotImplementedError:
                return MISSING
        raise TypeError(e)

    def _add_edges(self):
        n = len(self.nodes)
        out = [set() for _ in range(n)]

        def visit(s, lv):
            name, idx = self._lhs_index(s.lhs, self.env, lv)
            if (name, idx) not in self.vid:
                return
            me = self.vid[(name, idx)]
            if self.nodes[me].stmt is not s:
                return
            acc = []
            rhs = s.rhs if isinstance(s, Det) else s.dist
            self._deps(rhs, lv, acc)
            for dname, didx in acc:
                if didx == ():
                    cands = [(dname, ())]
                else:
                    ranges = [range(i[0], i[1] + 1) if isinstance(i, tuple) else (i,) for i in didx]
                    cands = [(dname, tuple(t)) for t in _product(ranges)]
                for c in cands:
                    v = self.vid.get(c)
                    if v is None:
                        if _is_missing(self._get(*c)):
                            raise ValueError(f"{c} is referenced but not defined")
                        continue
                    if v != me:
                        out[v].add(me)

        _unroll(self.stmts, self.env, {}, visit, self.ev)
        self.out = [sorted(s) for s in out]

    @staticmethod
    def _toposort(out, n):
        """Graphs.jl `topological_sort_by_dfs` (third party, restated from its published source)."""
        color = [0] * n
        verts = []
        for v in range(n):
            if color[v] != 0:
                continue
            stack = [v]
            color[v] = 1
            while stack:
                u = stack[-1]
                w = -1
                for x in out[u]:
                    if color[x] == 1:
                        raise ValueError("the graph contains a cycle")
                    if color[x] == 0:
                        w = x
                        break
                if w >= 0:
                    color[w] = 1
                    stack.append(w)
                else:
                    color[u] = 2
                    verts.append(u)
                    stack.pop()
        return verts[::-1]

    def _fill_inits(self, inits):
        """inits for unobserved stochastic nodes; anything still missing becomes 0.0
        (JuliaBUGS.jl:346-362; the reference then draws `rand(dist)` for parameters without an
        init, bugsmodel.jl:441-447 -- not reproducible without its RNG, tests always pass theta)."""
        for k, v in inits.items():
            if isinstance(v, (list, tuple, np.ndarray)):
                a = np.asarray(v, dtype=float)
                tgt = self.env[k]
                for pos in np.ndindex(a.shape):
                    if not math.isnan(a[pos]) and _is_missing(tgt[pos]):
                        tgt[pos] = float(a[pos])
            elif _is_missing(self.env.get(k)):
                self.env[k] = float(v)
        for k, v in self.env.items():
            if isinstance(v, np.ndarray):
                f = v.ravel()
                for t in range(f.size):
                    if _is_missing(f[t]):
                        f[t] = 0.0
                self.env[k] = f.reshape(v.shape)
            elif _is_missing(v):
                self.env[k] = 0.0

    # ---- classification -----------------------------------------------------------------------
    def _classify(self):
        """graphs.jl:27-71 + bugsmodel.jl:48-69,181-199."""
        n = len(self.nodes)
        memo = {}

        def reach(v):
            # dfs_can_reach_observations (graphs.jl:48-71); plate graphs are only a few levels deep
            if v in memo:
                return memo[v]
            nd = self.nodes[v]
            if nd.is_stochastic and nd.is_observed:
                memo[v] = True
                return True
            r = False
            for c in self.out[v]:
                if reach(c):
                    r = True
                    break
            memo[v] = r
            return r

        gq = set()
        for nd in self.nodes:
            if not (nd.is_stochastic and nd.is_observed):
                if not reach(nd.vid):
                    gq.add(nd.vid)
        has_obs = any(nd.is_stochastic and nd.is_observed for nd in self.nodes)
        if not has_obs:
            memo2 = {}

            def reach_stoch(v):
                if v in memo2:
                    return memo2[v]
                r = False
                for c in self.out[v]:
                    if self.nodes[c].is_stochastic or reach_stoch(c):
                        r = True
                        break
                memo2[v] = r
                return r

            gq = {v for v in gq if not self.nodes[v].is_stochastic and not reach_stoch(v)}
        self.gq = gq
        self._refresh_layout()

    def _refresh_layout(self):
        self.var_type = {}
        self.parameters = []
        for nd in self.sorted_nodes:
            if nd.is_stochastic and nd.is_observed:
                t = "Observation"
            elif nd.vid in self.gq:
                t = "GeneratedQuantity"
            elif nd.is_stochastic:
                t = "ModelParameter"
                self.parameters.append(nd)
            else:
                t = "TransformedParameter"
            self.var_type[nd.vid] = t
        self.D = len(self.parameters)  # every parameter in scope is scalar

    # ---- generated mode order -------------------------------------------------------------------
    def use_generated_order(self):
        """Rebuild `sorted_nodes` the way `set_evaluation_mode(model, UseGeneratedLogDensityFunction())`
        does (bugsmodel.jl:737-801): statement-level ordering graph, DFS topological sort of the
        statements, full fission, regrouping of consecutive statements with identical loop nests
        (source_gen.jl:180-223), then the unrolled program order of that reconstructed program."""
        flat = self._flat
        ns = len(flat)
        var_stmt = {}
        for nd in self.nodes:
            var_stmt[nd.vid] = self._stmt_id[id(nd.stmt)]
        coarse = [set() for _ in range(ns + 1)]
        deg = [0] * (ns + 1)
        for u in range(len(self.nodes)):
            for v in self.out[u]:
                a, b = var_stmt[u], var_stmt[v]
                if b not in coarse[a]:
                    coarse[a].add(b)
                    deg[a] += 1
                    deg[b] += 1
        # ordering graph: self edges are loop-carried recurrences (dropped when sequentially valid,
        # source_gen.jl:864-920); the plate models in scope have none.
        order_out = [sorted(x for x in coarse[a] if x != a) for a in range(ns + 1)]
        for a in range(1, ns + 1):
            if a in coarse[a]:
                raise NotImplementedError("loop-carried self dependence: outside the hot-path scope")
        topo = self._toposort([[x - 1 for x in order_out[a + 1]] for a in range(ns)], ns)
        fissioned = []
        for sid0 in topo:
            s, loops = flat[sid0]
            if isinstance(s, Det) and deg[sid0 + 1] == 0:
                continue  # transformed data statement
            fissioned.append((loops, s))
        # regroup consecutive statements with identical loop nests
        groups = []
        for loops, s in fissioned:
            if groups and groups[-1][0] == loops:
                groups[-1][1].append(s)
            else:
                groups.append((loops, [s]))
        order = []

        def emit(loops, body, lv):
            if not loops:
                for s in body:
                    name, idx = self._lhs_index(s.lhs, self.env, lv)
                    v = self.vid.get((name, idx))
                    if v is not None and self.nodes[v].stmt is s:
                        order.append(v)
                return
            (var, lo, hi) = loops[0]
            lo_v = _to_int(self.ev.ev(lo, self.env, lv))
            hi_v = _to_int(self.ev.ev(hi, self.env, lv))
            for k in range(lo_v, hi_v + 1):
                lv2 = dict(lv)
                lv2[var] = k
                emit(loops[1:], body, lv2)

        for loops, body in groups:
            emit(loops, body, {})
        assert sorted(order) == list(range(len(self.nodes))), "generated order must cover every vertex"
        self.sorted_nodes = [self.nodes[v] for v in order]
        self.mode = "generated"
        self._refresh_layout()
        return self

    # ---- theta ----------------------------------------------------------------------------------
    def _dist_of(self, nd, env, ev):
        args = []
        for a in nd.stmt.dist.args:
            args.append(ev.ev(a, env, nd.lv))
        if nd.stmt.dist.fn == "dcat":
            args = [list(np.asarray(args[0], dtype=object).ravel())]
        return dists.make_dist(ev.M, nd.stmt.dist.fn, args)

    def getparams(self):
        """bugsmodel.jl:619-665: parameters in `sorted_nodes` order, linked when `transformed`."""
        theta = []
        M = FLOAT_OR_DUAL
        for nd in self.parameters:
            x = self._get(nd.name, nd.idx)
            if self.transformed:
                d = self._dist_of(nd, self.env, self.ev)
                theta.append(float(dists.link(M, d, x)))
            else:
                theta.append(float(x))
        return np.asarray(theta)

    def param_names(self):
        return [nd.label for nd in self.parameters]

    # ---- the node loop ----------------------------------------------------------------------------
    def evaluate_with_values(self, theta, temperature=1.0, M=None, transformed=None):
        """`evaluate_with_values!!` (evaluation.jl:217-275)."""
        ev = self.ev if M is None else Evaluator(M)
        M = ev.M
        transformed = self.transformed if transformed is None else transformed
        env = {}
        for k, v in self.env.items():  # smart_copy_evaluation_env (utils.jl:161-169)
            if isinstance(v, np.ndarray):
                env[k] = v.copy()
                if M.name == "mpmath":
                    f = env[k].ravel()
                    for t in range(f.size):
                        f[t] = M.const(f[t])
            else:
                env[k] = M.const(v) if M.name == "mpmath" else v
        cur = 0
        logprior, loglik = 0.0, 0.0
        for nd in self.sorted_nodes:
            t = self.var_type[nd.vid]
            if t == "GeneratedQuantity":
                continue
            if not nd.is_stochastic:
                val = ev.ev(nd.stmt.rhs, env, nd.lv)
                self._set(env, nd.name, nd.idx, val)
            elif not nd.is_observed:
                d = self._dist_of(nd, env, ev)
                y = theta[cur]
                cur += 1
                if transformed:
                    x, logjac = dists.invlink_with_logjac(M, d, y)
                else:
                    x, logjac = y, 0.0
                logprior = logprior + d.logpdf(M, x) + logjac
                self._set(env, nd.name, nd.idx, x)
            else:
                d = self._dist_of(nd, env, ev)
                loglik = loglik + d.logpdf(M, self._get_env(env, nd))
        assert cur == len(theta), "all of theta must be consumed"
        return env, {"logprior": logprior, "loglikelihood": loglik,
                     "tempered_logjoint": logprior + temperature * loglik}

    # ---- UseAutoMarginalization (evaluation.jl:739-961) ------------------------------------------------------
    def _is_latent(self, nd):
        return (nd.is_stochastic and not nd.is_observed and self.var_type[nd.vid] != "GeneratedQuantity"
                and nd.stmt.dist.fn in ("dcat", "dbern"))

    def marginal_param_names(self):
        """the continuous parameters, in `sorted_nodes` order: theta of the marginalised model"""
        return [nd.label for nd in self.parameters if not self._is_latent(nd)]

    def evaluate_with_marginalization_values(self, theta, temperature=1.0, M=None):
        """`evaluate_with_marginalization_values!!` / `_marginalize_recursive` (evaluation.jl:739-961) without the
        memo table: the node loop, with every unobserved finite discrete node summed over its support --
        log_prior_marg = logsumexp_z(log p(z) + rest_prior), log_joint_marg = logsumexp_z(that + rest_lik),
        log_lik = log_joint_marg - log_prior_marg.  Exponential in the number of latents: small test cases only."""
        ev = self.ev if M is None else Evaluator(M)
        M = ev.M
        nodes = [nd for nd in self.sorted_nodes if self.var_type[nd.vid] != "GeneratedQuantity"]
        slot, cur = {}, 0
        for nd in nodes:
            if nd.is_stochastic and not nd.is_observed and not self._is_latent(nd):
                slot[nd.vid] = cur
                cur += 1
        assert cur == len(theta), "theta holds the continuous parameters only"

        def lse(xs):
            m = max(float(_val(x)) for x in xs)
            if m == -math.inf:
                return -math.inf
            acc = 0.0
            for x in xs:
                acc = acc + M.exp(x - m)
            return m + M.log(acc)

        def rec(k, env):
            if k == len(nodes):
                return 0.0, 0.0
            nd = nodes[k]
            if not nd.is_stochastic:
                env = dict(env)
                env[nd.name] = env[nd.name].copy() if isinstance(env[nd.name], np.ndarray) else env[nd.name]
                self._set(env, nd.name, nd.idx, ev.ev(nd.stmt.rhs, env, nd.lv))
                return rec(k + 1, env)
            d = self._dist_of(nd, env, ev)
            if nd.is_observed:
                pr, lk = rec(k + 1, env)
                return pr, lk + d.logpdf(M, self._get_env(env, nd))
            if self._is_latent(nd):
                priors, joints = [], []
                # the support of the finite discrete node (evaluation.jl:860-890 enumerates `support(dist)`)
                for val in (range(1, len(d.p) + 1) if isinstance(d, dists.Categorical) else (0, 1)):
                    env2 = dict(env)
                    env2[nd.name] = env2[nd.name].copy() if isinstance(env2[nd.name], np.ndarray) else env2[nd.name]
                    self._set(env2, nd.name, nd.idx, float(val))
                    pr, lk = rec(k + 1, env2)
                    tot = d.logpdf(M, val) + pr
                    priors.append(tot)
                    joints.append(tot + lk)
                lp, lj = lse(priors), lse(joints)
                return lp, (lj - lp if math.isfinite(float(_val(lp))) else lj)
            y = theta[slot[nd.vid]]
            x, logjac = dists.invlink_with_logjac(M, d, y) if self.transformed else (y, 0.0)
            env = dict(env)
            env[nd.name] = env[nd.name].copy() if isinstance(env[nd.name], np.ndarray) else env[nd.name]
            self._set(env, nd.name, nd.idx, x)
            pr, lk = rec(k + 1, env)
            return d.logpdf(M, x) + logjac + pr, lk

        env0 = {k: (v.astype(object) if isinstance(v, np.ndarray) else v) for k, v in self.env.items()}
        pr, lk = rec(0, env0)
        return {"logprior": pr, "loglikelihood": lk, "tempered_logjoint": pr + temperature * lk}

    def marginal_logdensity_and_gradient(self, theta):
        n = len(theta)
        duals = [Dual.seed(float(theta[i]), i, n) for i in range(n)]
        lj = self.evaluate_with_marginalization_values(duals)["tempered_logjoint"]
        return lj.v, np.array(lj.d, dtype=float)

    @staticmethod
    def _get_env(env, nd):
        if nd.idx == ():
            return env[nd.name]
        return env[nd.name][tuple(i - 1 for i in nd.idx)]

    # ---- LogDensityProblems boundary -----------------------------------------------------------------
    def dimension(self):
        return self.D

    def logdensity(self, theta):
        """logdensityproblems.jl:19-27: DomainError -> -Inf."""
        try:
            _, r = self.evaluate_with_values(list(map(float, theta)))
            return float(_val(r["tempered_logjoint"]))
        except DomainError:
            return -math.inf

    def logdensity_and_gradient(self, theta):
        """logdensityproblems.jl:219-232 with a forward-mode backend: DomainError -> (-Inf, NaN)."""
        n = len(theta)
        try:
            duals = [Dual.seed(float(theta[i]), i, n) for i in range(n)]
            _, r = self.evaluate_with_values(duals)
            lj = r["tempered_logjoint"]
            if not isinstance(lj, Dual):
                return float(lj), np.zeros(n)
            g = lj.d if isinstance(lj.d, np.ndarray) else np.zeros(n)
            return lj.v, np.array(g, dtype=float)
        except DomainError:
            return -math.inf, np.full(n, math.nan)

    def logdensity_mp(self, theta, dps=50):
        from numeric import mp_math

        M = mp_math(dps)
        _, r = self.evaluate_with_values([M.const(float(t)) for t in theta], M=M)
        return r["tempered_logjoint"]


def _binop(op, a, b):
    if op == "+":
        return a + b
    if op == "-":
        return a - b
    if op == "*":
        return a * b
    if op == "/":
        return a / b
    if op == "^":
        return a**b
    raise ValueError(op)


def _product(ranges):
    if not ranges:
        yield ()
        return
    for x in ranges[0]:
        for rest in _product(ranges[1:]):
            yield (x,) + rest


def compile(model_text, data, inits=None, transformed=True):  # noqa: A001 - mirrors the reference name
    return GraphModel(model_text, data, inits, transformed)
