"""Generated quantities on the device (include/jbugs_gq.h).

The reference fills in the nodes outside the target by re-running the node loop once per draw on the host
(`J/src/model/abstractmcmc.jl:92-110`, `J/src/model/evaluation.jl:354-378`).  Here the statements outside the target
(and the logical statements they read) are lowered ONCE into a small program -- per statement an SSA tape over the
statement's loop domain, with the row numbers / data values that vary over the domain precomputed per cell -- and
`libjbugs_cuda` evaluates it for all draws of a run in one pass: thread = draw, the draws matrix stays in HBM.
`jbugs/gq.py` is the host-side (numpy) twin of the same statement semantics; the tests check both against the per-node
oracle's evaluation environment.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import plan as P
from .cuda import _check, load
from .lowering import Lowering, LoweringError, _grids
from .parser import BinOp, Call, Colon, Index, Neg, Num, Range, Var, free_vars

MAX_OPS = 96
LOGICAL = 0xFFFFFFFF
LEAF_CONST, LEAF_THETA, LEAF_THETA_CELL, LEAF_Q, LEAF_Q_CELL, LEAF_DATA_CELL = range(6)
UNARY = dict(P.UNARY_OF, abs=40, step=41, logit=42, cloglog=43, probit=44, sin=45, cos=46, round=47, trunc=48,
             loggam=49, logfact=50)
BINARY = {"equals": 60, "max": 61, "min": 62, "pow": P.OP_POW}
BINOPS = {"+": P.OP_ADD, "-": P.OP_SUB, "*": P.OP_MUL, "/": P.OP_DIV, "^": P.OP_POW}

OP_DT = np.dtype([("op", "<u4"), ("a", "<u2"), ("b", "<u2"), ("leaf_kind", "<u4"), ("pad", "<u4"), ("offset", "<i8"),
                  ("value", "<f8")])
STMT_DT = np.dtype([("n_cells", "<i8"), ("out_row0", "<i8"), ("first_op", "<u4"), ("n_ops", "<u4"), ("family", "<u4"),
                    ("n_args", "<u4"), ("arg_slot", "<u4", (3,)), ("pad", "<u4")])
assert OP_DT.itemsize == 32 and STMT_DT.itemsize == 48


class _Tape:
    def __init__(self):
        self.ops: List[tuple] = []
        self.memo: Dict[str, int] = {}

    def push(self, key, op, a=0, b=0, leaf_kind=0, offset=0, value=0.0):
        if key is not None and key in self.memo:
            return self.memo[key]
        if len(self.ops) >= MAX_OPS:
            raise LoweringError(f"a generated quantity needs more than {MAX_OPS} tape ops")
        self.ops.append((op, a, b, leaf_kind, 0, offset, value))
        if key is not None:
            self.memo[key] = len(self.ops) - 1
        return len(self.ops) - 1


class GQProgram:
    """statements outside the target, lowered for the device; `outputs[name] = (rows, is_gq)` with `rows` an int array
    over the variable's dims (-1: not computed)."""

    def __init__(self, lw: Lowering, only_gq: bool = False):
        """`only_gq`: keep the statements outside the target and the logical statements they read, nothing else (what a
        sampler's monitored rows can feed); by default every logical statement is part of the program."""
        self._keep_sids = None
        if only_gq:
            full = GQProgram(lw)
            by_name: Dict[str, List[int]] = {}
            for sid, nm in full._sid_name.items():
                by_name.setdefault(nm, []).append(sid)
            keep = {sid for sid in full._sid_name if full._sid_gq[sid]}
            todo = list(keep)
            while todo:
                for nm in full._reads.get(todo.pop(), ()):
                    for sid in by_name.get(nm, ()):
                        if sid not in keep:
                            keep.add(sid)
                            todo.append(sid)
            self._keep_sids = keep
        self._reads: Dict[int, set] = {}
        self._sid_name: Dict[int, str] = {}
        self._sid_gq: Dict[int, bool] = {}
        self.lw = lw
        self.D = lw.D
        self.stmts: List[tuple] = []
        self.ops: List[tuple] = []
        self.ipool: List[np.ndarray] = []
        self.dpool: List[np.ndarray] = []
        self.n_ipool = self.n_dpool = 0
        self.n_rows = 0
        self.outputs: Dict[str, tuple] = {}
        self.skipped: Dict[str, str] = {}
        self._theta_rows: Dict[str, np.ndarray] = {}
        self._q_rows: Dict[str, np.ndarray] = {}
        self._bijectors()
        self._param_tables()
        self._lower_statements()

    # ---- theta: bijector per row, row number per parameter element ------------------------------------------------
    def _bijectors(self):
        plan = self.lw.plan
        self.tr = np.zeros(self.D, dtype=np.int32)
        self.lb = np.zeros(self.D)
        self.ub = np.zeros(self.D)
        self._fams = {g.name: (g.transform, g.lb, g.ub) for g in plan.globals}
        for pl in plan.plates:
            for l in pl.locals:
                self._fams[l.name] = (l.transform, l.lb, l.ub)

    @staticmethod
    def _grow(table: Optional[np.ndarray], top) -> np.ndarray:
        if table is None:
            return np.full(top, -1, dtype=np.int64)
        if all(c >= t for c, t in zip(table.shape, top)):
            return table
        new = np.full(tuple(max(c, t) for c, t in zip(table.shape, top)), -1, dtype=np.int64)
        new[tuple(slice(0, c) for c in table.shape)] = table
        return new

    def _lhs_index(self, s):
        glv = _grids(s.extents, s.loop_vars)
        return [np.broadcast_to(np.asarray(self.lw.de.ev(ie, glv)), s.shape if s.shape else ()).astype(np.int64)
                for ie in s.node.lhs.idx]

    def _param_tables(self):
        lw = self.lw
        for s in lw.stmts:
            if s.kind != "param" or s.is_gq or (s.sid not in lw.theta_map and lw.theta_order is None):
                continue
            try:
                rows = lw.theta_index(s)
            except LoweringError:
                continue
            base = s.name
            # globals may carry an element label ("beta[2]") as their record name
            if isinstance(s.node.lhs, Var):
                self._theta_rows[base] = np.asarray(int(rows), dtype=np.int64)
                names = [base]
                flat = [int(rows)]
            else:
                idx = self._lhs_index(s)
                top = tuple(int(i.max()) for i in idx)
                t = self._grow(self._theta_rows.get(base), top)
                t[tuple(i - 1 for i in idx)] = rows
                self._theta_rows[base] = t
                names = lw._labels_for(s) if s.size <= 4096 else []
                flat = np.asarray(rows).ravel().tolist()
            bij = self._fams.get(base)
            for k, r in enumerate(flat):
                b = bij or (self._fams.get(names[k]) if k < len(names) else None)
                if b is None:
                    continue
                self.tr[r], self.lb[r], self.ub[r] = b

    # ---- statements ------------------------------------------------------------------------------------------------
    def _known(self, e, loop_vars) -> bool:
        return all((v in self.lw.data and v not in self.lw.de.unknown) or v in loop_vars or v in self._theta_rows or v in self._q_rows
                   for v in free_vars(e))

    def _lower_statements(self):
        lw = self.lw
        pending = [s for s in lw.stmts if s.kind == "logical" or (s.kind == "param" and s.is_gq)]
        if self._keep_sids is not None:
            pending = [s for s in pending if s.sid in self._keep_sids]
        progress = True
        while pending and progress:
            progress = False
            for s in list(pending):
                exprs = [s.node.rhs] if s.kind == "logical" else list(s.node.dist.args)
                if not all(self._known(e, s.loop_vars) for e in exprs):
                    continue
                needs = {v for e in exprs for v in free_vars(e)}
                if any(t is not s and t.name in needs for t in pending):
                    continue   # another pending statement still has to fill an array this one reads
                pending.remove(s)
                progress = True
                mark = (len(self.ops), len(self.ipool), len(self.dpool), self.n_ipool, self.n_dpool)
                try:
                    self._lower_one(s, exprs)
                except (LoweringError, KeyError, IndexError, ValueError) as err:
                    # same policy as the host twin: a statement outside the vocabulary is left out, loudly recorded
                    del self.ops[mark[0]:], self.ipool[mark[1]:], self.dpool[mark[2]:]
                    self.n_ipool, self.n_dpool = mark[3], mark[4]
                    self.skipped[s.name] = str(err)
        for s in pending:
            self.skipped.setdefault(s.name, "depends on a quantity that could not be lowered")

    def _lower_one(self, s, exprs):
        lw = self.lw
        if s.ragged:
            raise LoweringError("ragged loop bounds")
        shape = tuple(s.shape)
        n_cells = int(np.prod(shape)) if shape else 1
        lv = _grids(s.extents, s.loop_vars)
        tape = _Tape()
        slots = [self._expr(e, s, lv, shape, n_cells, tape) for e in exprs]
        if s.kind == "logical":
            fam, n_args = LOGICAL, 0
            if slots[0] != len(tape.ops) - 1:   # the statement's value must be the last op
                slots[0] = tape.push(None, P.OP_IDENTITY, slots[0])
        else:
            fam = lw._family(s.node.dist)
            if fam in (P.FAM_FLAT, P.FAM_CATEGORICAL):
                raise LoweringError(f"forward sampling of {s.node.dist.fn} is not provided")
            n_args = len(slots)
        self._reads[s.sid] = {v for e in exprs for v in free_vars(e) if v in self._q_rows}
        self._sid_name[s.sid] = s.name
        self._sid_gq[s.sid] = bool(s.is_gq)
        first = len(self.ops)
        self.ops.extend(tape.ops)
        arg_slot = (slots + [0, 0, 0])[:3] if fam != LOGICAL else [0, 0, 0]
        self.stmts.append((n_cells, self.n_rows, first, len(tape.ops), fam, n_args, arg_slot, 0))
        rows = np.arange(self.n_rows, self.n_rows + n_cells, dtype=np.int64).reshape(shape if shape else ())
        self.n_rows += n_cells
        if isinstance(s.node.lhs, Var):
            self._q_rows[s.name] = rows
        else:
            idx = self._lhs_index(s)
            top = tuple(int(i.max()) for i in idx)
            t = self._grow(self._q_rows.get(s.name), top)
            t[tuple(i - 1 for i in idx)] = rows
            self._q_rows[s.name] = t
        prev = self.outputs.get(s.name)
        self.outputs[s.name] = (self._q_rows[s.name], bool(s.is_gq) or bool(prev and prev[1]))

    def _rows_leaf(self, e, table, kinds, s, lv, shape, n_cells, tape):
        """leaf reading theta / Q rows `table[subscripts]` for every cell"""
        if isinstance(e, Var):
            if table.ndim != 0:
                raise LoweringError(f"'{e.name}' is an array used without subscripts")
            rows = np.full(n_cells, int(table), dtype=np.int64)
        else:
            if table.ndim != len(e.idx):
                raise LoweringError(f"'{e.name}' is referenced with {len(e.idx)} subscripts")
            idx = []
            for ie in e.idx:
                if isinstance(ie, (Range, Colon)):
                    raise LoweringError("an index range outside sum / mean / sd / inprod")
                if not self.lw.de.known(ie, s.loop_vars):
                    raise LoweringError("array index depends on a parameter")
                idx.append(np.broadcast_to(np.asarray(self.lw.de.ev(ie, lv)), shape if shape else ()).astype(np.int64).ravel() - 1)
            if any((i < 0).any() or (i >= d).any() for i, d in zip(idx, table.shape)):
                raise LoweringError(f"'{e.name}': subscript out of range")
            rows = table[tuple(idx)].reshape(-1)
        if (rows < 0).any():
            raise LoweringError(f"an element of '{e.name}' is neither a parameter nor a computed quantity")
        if (rows == rows[0]).all():
            return tape.push(repr(e), P.OP_LEAF, leaf_kind=kinds[0], offset=int(rows[0]))
        off = self.n_ipool
        self.ipool.append(np.ascontiguousarray(rows))
        self.n_ipool += rows.size
        return tape.push(repr(e), P.OP_LEAF, leaf_kind=kinds[1], offset=off)

    def _elements(self, e, s):
        """scalar element expressions of a vector argument `x[lo:hi]` / `x[i, 1:n]` (static bounds)"""
        if not isinstance(e, Index):
            raise LoweringError("sum / mean / sd / inprod over something that is not an array section")
        dims = None
        if e.name in self._theta_rows:
            dims = self._theta_rows[e.name].shape
        elif e.name in self._q_rows:
            dims = self._q_rows[e.name].shape
        elif e.name in self.lw.data:
            dims = np.shape(self.lw.data[e.name])
        if dims is None or len(dims) != len(e.idx):
            raise LoweringError(f"'{e.name}': unknown shape")
        choices = []
        for d, ie in enumerate(e.idx):
            if isinstance(ie, Colon):
                choices.append([Num(float(k), True) for k in range(1, dims[d] + 1)])
            elif isinstance(ie, Range):
                if not (self.lw.de.known(ie.lo, ()) and self.lw.de.known(ie.hi, ())):
                    raise LoweringError("loop-dependent index range over a parameter array")
                lo, hi = int(self.lw.de.ev(ie.lo, {})), int(self.lw.de.ev(ie.hi, {}))
                choices.append([Num(float(k), True) for k in range(lo, hi + 1)])
            else:
                choices.append([ie])
        out = [()]
        for c in choices:
            out = [o + (x,) for o in out for x in c]
        if not 1 <= len(out) <= 40:
            raise LoweringError("a reduction over more than 40 parameter elements")
        return [Index(e.name, o) for o in out]

    def _expr(self, e, s, lv, shape, n_cells, tape: _Tape) -> int:
        key = repr(e)
        if key in tape.memo:
            return tape.memo[key]
        lw = self.lw
        rec = lambda x: self._expr(x, s, lv, shape, n_cells, tape)  # noqa: E731
        if lw.de.known(e, s.loop_vars):
            vals = np.broadcast_to(np.asarray(lw.de.ev(e, lv), dtype=np.float64), shape if shape else ()).reshape(-1)
            if vals.size == 0 or (vals == vals[0]).all() or np.isnan(vals).all():
                return tape.push(key, P.OP_LEAF, leaf_kind=LEAF_CONST, value=float(vals[0]) if vals.size else 0.0)
            off = self.n_dpool
            self.dpool.append(np.ascontiguousarray(vals))
            self.n_dpool += vals.size
            return tape.push(key, P.OP_LEAF, leaf_kind=LEAF_DATA_CELL, offset=off)
        if isinstance(e, (Var, Index)):
            if e.name in self._q_rows:   # (a generated quantity shadows nothing: names are unique per model)
                return self._rows_leaf(e, self._q_rows[e.name], (LEAF_Q, LEAF_Q_CELL), s, lv, shape, n_cells, tape)
            if e.name in self._theta_rows:
                return self._rows_leaf(e, self._theta_rows[e.name], (LEAF_THETA, LEAF_THETA_CELL), s, lv, shape, n_cells, tape)
            raise LoweringError(f"'{e.name}' is not available")
        if isinstance(e, Neg):
            return tape.push(key, P.OP_NEG, rec(e.a))
        if isinstance(e, BinOp):
            return tape.push(key, BINOPS[e.op], rec(e.a), rec(e.b))
        if isinstance(e, Call):
            fn = {"ilogit": "logistic", "icloglog": "cexpexp"}.get(e.fn, e.fn)
            if fn in UNARY and len(e.args) == 1:
                return tape.push(key, UNARY[fn], rec(e.args[0]))
            if fn in BINARY and len(e.args) == 2:
                return tape.push(key, BINARY[fn], rec(e.args[0]), rec(e.args[1]))
            if fn in ("sum", "mean", "sd") and len(e.args) == 1:
                els = [rec(x) for x in self._elements(e.args[0], s)]
                n = len(els)
                acc = els[0]
                for x in els[1:]:
                    acc = tape.push(None, P.OP_ADD, acc, x)
                if fn == "sum":
                    tape.memo[key] = acc
                    return acc
                cn = tape.push(None, P.OP_LEAF, leaf_kind=LEAF_CONST, value=float(n))
                mean = tape.push(None, P.OP_DIV, acc, cn)
                if fn == "mean":
                    tape.memo[key] = mean
                    return mean
                ss = None
                for x in els:
                    d = tape.push(None, P.OP_SUB, x, mean)
                    d2 = tape.push(None, P.OP_MUL, d, d)
                    ss = d2 if ss is None else tape.push(None, P.OP_ADD, ss, d2)
                c1 = tape.push(None, P.OP_LEAF, leaf_kind=LEAF_CONST, value=float(n - 1))
                out = tape.push(None, P.OP_SQRT, tape.push(None, P.OP_DIV, ss, c1))
                tape.memo[key] = out
                return out
            if fn == "inprod" and len(e.args) == 2:
                xs, ys = self._elements(e.args[0], s), self._elements(e.args[1], s)
                if len(xs) != len(ys):
                    raise LoweringError("inprod of sections of different lengths")
                acc = None
                for x, y in zip(xs, ys):
                    p = tape.push(None, P.OP_MUL, rec(x), rec(y))
                    acc = p if acc is None else tape.push(None, P.OP_ADD, acc, p)
                tape.memo[key] = acc
                return acc
        raise LoweringError(f"unsupported expression in a generated quantity: {e!r}")

    # ---- packing ---------------------------------------------------------------------------------------------------
    def arrays(self):
        ops = np.array(self.ops, dtype=OP_DT) if self.ops else np.zeros(0, dtype=OP_DT)
        st = np.zeros(len(self.stmts), dtype=STMT_DT)
        for k, (n_cells, row0, first, n_ops, fam, n_args, arg_slot, _) in enumerate(self.stmts):
            st[k] = (n_cells, row0, first, n_ops, fam, n_args, arg_slot, 0)
        ip = np.concatenate(self.ipool).astype(np.int64) if self.ipool else np.zeros(0, dtype=np.int64)
        dp = np.concatenate(self.dpool).astype(np.float64) if self.dpool else np.zeros(0, dtype=np.float64)
        return st, ops, ip, dp

    def theta_rows_needed(self) -> np.ndarray:
        need = set()
        _, ops, ip, _ = self.arrays()
        for (n_cells, _, first, n_ops, *_rest) in self.stmts:
            for o in ops[first:first + n_ops]:
                if o["op"] == P.OP_LEAF and o["leaf_kind"] == LEAF_THETA:
                    need.add(int(o["offset"]))
                elif o["op"] == P.OP_LEAF and o["leaf_kind"] == LEAF_THETA_CELL:
                    need.update(int(r) for r in ip[o["offset"]:o["offset"] + n_cells])
        return np.array(sorted(need), dtype=np.int64)


def _bind(L):
    c = ctypes
    L.jbugs_gq_create.argtypes = [c.c_int, c.c_int64, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p,
                                  c.c_int64, c.c_void_p, c.c_int64, c.c_void_p, c.c_int64, c.c_int64, c.POINTER(c.c_void_p)]
    L.jbugs_gq_destroy.argtypes = [c.c_void_p]
    L.jbugs_gq_destroy.restype = None
    L.jbugs_gq_run.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_int64, c.c_void_p, c.c_int64, c.c_uint64, c.c_void_p,
                               c.c_int64, c.c_void_p]
    L.jbugs_gq_launch_count.argtypes = [c.c_void_p]
    L.jbugs_gq_launch_count.restype = c.c_int64


class DeviceGQ:
    """the program of a model on one GPU.  `run(theta, rows)` -> name -> array of shape dims + (S,)."""

    def __init__(self, lowering: Lowering, device: int = 0, only_gq: bool = False):
        self.program = GQProgram(lowering, only_gq=only_gq)
        self.L = load()
        _bind(self.L)
        st, ops, ip, dp = self.program.arrays()
        self._keep = (st, ops, ip, dp)
        h = ctypes.c_void_p()
        pr = self.program
        _check(self.L.jbugs_gq_create(device, pr.D, pr.tr.ctypes.data, pr.lb.ctypes.data, pr.ub.ctypes.data,
                                      st.ctypes.data, len(st), ops.ctypes.data, len(ops), ip.ctypes.data, len(ip),
                                      dp.ctypes.data, len(dp), pr.n_rows, ctypes.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.L.jbugs_gq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self) -> int:
        return int(self.L.jbugs_gq_launch_count(self.h))

    def run_raw(self, theta_ptr: int, ld: int, R: int, rows: Optional[np.ndarray], S: int, seed: int, out_ptr: int, ld_out: int,
                stream=None):
        rp = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        _check(self.L.jbugs_gq_run(self.h, theta_ptr, ld, R, None if rp is None else rp.ctypes.data, S, seed, out_ptr, ld_out, stream))

    def run(self, theta: np.ndarray, rows: Optional[Sequence[int]] = None, seed: int = 0, only_gq: bool = True) -> Dict[str, np.ndarray]:
        """theta: [R, S] UNCONSTRAINED draws on the host (row r is theta index rows[r]; all D rows when `rows` is None)"""
        th = np.ascontiguousarray(theta, dtype=np.float64)
        R, S = th.shape
        out = np.full((max(self.program.n_rows, 1), S), np.nan)
        self.run_raw(th.ctypes.data, S, R, None if rows is None else np.asarray(rows), S, seed, out.ctypes.data, S)
        return self.collect(out, only_gq)

    def collect(self, out: np.ndarray, only_gq: bool = True) -> Dict[str, np.ndarray]:
        S = out.shape[1]
        res = {}
        for name, (rows, is_gq) in self.program.outputs.items():
            if only_gq and not is_gq:
                continue
            r = np.asarray(rows)
            vals = np.full(r.shape + (S,), np.nan)
            ok = r >= 0
            vals[ok] = out[r[ok]]
            res[name] = vals
        return res
