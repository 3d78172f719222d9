"""BUGS-language front end (subset) -> small immutable AST.

The reference turns BUGS text into Julia `Expr` trees in
`JuliaBUGS/src/parser/bugs_parser.jl` and rewrites link functions on the left
hand side in `JuliaBUGS/src/parser/bugs_macro.jl:159-193`
(`logit(p) <- e` becomes `p = logistic(e)`, `log` -> `exp`,
`cloglog` -> `cexpexp`, `probit` -> `phi`).  The front end is host-side text
handling and is OUT of the hot path; this module only
exists so that models can be written down for the lowering pass and for the
oracle.  It is a from-scratch recursive-descent parser, not a translation.

Grammar (subset):
    program  := ['model'] '{' stmts '}' | stmts
    stmt     := 'for' '(' NAME 'in' expr ':' expr ')' '{' stmts '}'
              | lhs '~' call [';']
              | lhs '<-' expr | lhs '=' expr | LINK '(' lhs ')' '<-' expr
    lhs      := NAME | NAME '[' index {',' index} ']'
    index    := [expr [':' expr]]          (empty index == whole axis)
    expr     := usual arithmetic with + - * / ^, unary minus, calls, indexing
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Tuple, Union

# --------------------------------------------------------------------------- AST


@dataclass(frozen=True)
class Num:
    value: float
    is_int: bool = False


@dataclass(frozen=True)
class Var:
    name: str


@dataclass(frozen=True)
class Range:
    lo: "Expr"
    hi: "Expr"


@dataclass(frozen=True)
class Colon:
    """An empty index (`x[]` in BUGS, `x[:]` in Julia): the whole axis."""


@dataclass(frozen=True)
class Index:
    name: str
    idx: Tuple["Expr", ...]


@dataclass(frozen=True)
class Call:
    fn: str
    args: Tuple["Expr", ...]


@dataclass(frozen=True)
class BinOp:
    op: str
    a: "Expr"
    b: "Expr"


@dataclass(frozen=True)
class Neg:
    a: "Expr"


Expr = Union[Num, Var, Range, Colon, Index, Call, BinOp, Neg]


@dataclass(frozen=True)
class Stoch:
    lhs: Union[Var, Index]
    dist: Call


@dataclass(frozen=True)
class Det:
    lhs: Union[Var, Index]
    rhs: Expr


@dataclass(frozen=True)
class For:
    var: str
    lo: Expr
    hi: Expr
    body: Tuple["Stmt", ...]


Stmt = Union[Stoch, Det, For]

# link function on the lhs -> inverse applied to the rhs (bugs_macro.jl:159-182)
LINK_INVERSE = {"logit": "logistic", "log": "exp", "cloglog": "cexpexp", "probit": "phi"}

# --------------------------------------------------------------------------- lexer

_TOKEN = re.compile(
    r"""
    (?P<ws>\s+|\#[^\n]*)
  | (?P<num>(\d+\.\d*|\.\d+|\d+)([eE][+-]?\d+)?)
  | (?P<name>[A-Za-z_][A-Za-z_0-9]*(\.[A-Za-z_0-9]+)*)
  | (?P<op><-|[~=+\-*/^(){}\[\],:;])
    """,
    re.VERBOSE,
)


class BUGSSyntaxError(ValueError):
    pass


def _lex(text: str):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if m is None:
            raise BUGSSyntaxError(f"unexpected character {text[pos]!r} at offset {pos}")
        pos = m.end()
        if m.lastgroup == "ws":
            continue
        out.append((m.lastgroup, m.group(m.lastgroup)))
    out.append(("eof", ""))
    return out


class _Parser:
    def __init__(self, text: str):
        self.toks = _lex(text)
        self.i = 0

    # -- helpers
    def peek(self, k=0):
        return self.toks[self.i + k]

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "name"):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise BUGSSyntaxError(f"expected {val!r}, found {self.peek()[1]!r} (token {self.i})")

    # -- statements
    def program(self):
        if self.peek() == ("name", "model") and self.peek(1)[1] == "{":
            self.next()
        if self.accept("{"):
            body = self.stmts("}")
            self.expect("}")
        else:
            body = self.stmts(None)
        if self.peek()[0] != "eof":
            raise BUGSSyntaxError(f"trailing input at token {self.i}: {self.peek()[1]!r}")
        return tuple(body)

    def stmts(self, closer):
        out = []
        while True:
            while self.accept(";"):
                pass
            kind, val = self.peek()
            if kind == "eof" or (closer is not None and val == closer):
                return out
            out.append(self.stmt())

    def stmt(self):
        if self.peek() == ("name", "for") and self.peek(1)[1] == "(":
            self.next()
            self.expect("(")
            kind, var = self.next()
            if kind != "name":
                raise BUGSSyntaxError("loop variable expected")
            self.expect("in")
            lo = self.additive()
            self.expect(":")
            hi = self.additive()
            self.expect(")")
            self.expect("{")
            body = self.stmts("}")
            self.expect("}")
            return For(var, lo, hi, tuple(body))
        # link(lhs) <- expr
        kind, val = self.peek()
        if kind == "name" and val in LINK_INVERSE and self.peek(1)[1] == "(":
            save = self.i
            self.next()
            self.expect("(")
            try:
                lhs = self.lhs()
                if self.accept(")") and (self.accept("<-") or self.accept("=")):
                    rhs = self.expr()
                    return Det(lhs, Call(LINK_INVERSE[val], (rhs,)))
            except BUGSSyntaxError:
                pass
            self.i = save
        lhs = self.lhs()
        if self.accept("~"):
            d = self.primary()
            if not isinstance(d, Call):
                raise BUGSSyntaxError("distribution call expected after '~'")
            # censoring / truncation suffix: dweib(r, mu) C(lower, upper), T(lower, upper) (I(,) is WinBUGS' spelling of
            # C); either bound may be empty.  Same tree as the reference builds (bugs_parser.jl:493):
            # censored(dist, lower, upper) with `nothing` for an empty bound
            if self.peek()[0] == "name" and self.peek()[1] in ("C", "T", "I") and self.peek(1)[1] == "(":
                which = self.next()[1]
                self.next()
                bounds = []
                for closer in (",", ")"):
                    bounds.append(Var("nothing") if self.peek()[1] == closer else self.expr())
                    self.expect(closer)
                d = Call("truncated" if which == "T" else "censored", (d, bounds[0], bounds[1]))
            return Stoch(lhs, d)
        if self.accept("<-") or self.accept("="):
            return Det(lhs, self.expr())
        raise BUGSSyntaxError(f"'~' or '<-' expected, found {self.peek()[1]!r}")

    def lhs(self):
        kind, name = self.next()
        if kind != "name":
            raise BUGSSyntaxError(f"variable expected, found {name!r}")
        if self.accept("["):
            return Index(name, self.indices())
        return Var(name)

    def indices(self):
        idx = []
        while True:
            if self.peek()[1] in (",", "]"):
                idx.append(Colon())
            else:
                idx.append(self.expr())
            if self.accept("]"):
                return tuple(idx)
            self.expect(",")

    # -- expressions (':' binds loosest, as in Julia)
    def expr(self):
        if self.peek()[1] == ":" and self.peek(1)[1] in (",", "]"):
            self.next()
            return Colon()
        a = self.additive()
        if self.accept(":"):
            return Range(a, self.additive())
        return a

    def additive(self):
        a = self.term()
        while self.peek()[1] in ("+", "-") and self.peek()[0] == "op":
            op = self.next()[1]
            a = BinOp(op, a, self.term())
        return a

    def term(self):
        a = self.unary()
        while self.peek()[1] in ("*", "/") and self.peek()[0] == "op":
            op = self.next()[1]
            a = BinOp(op, a, self.unary())
        return a

    def unary(self):
        if self.peek() == ("op", "-"):
            self.next()
            return Neg(self.unary())
        if self.peek() == ("op", "+"):
            self.next()
            return self.unary()
        return self.power()

    def power(self):
        a = self.primary()
        if self.accept("^"):
            return BinOp("^", a, self.unary())  # right associative, binds tighter than unary minus on the left
        return a

    def primary(self):
        kind, val = self.next()
        if kind == "num":
            is_int = re.fullmatch(r"\d+", val) is not None
            return Num(float(val), is_int)
        if kind == "name":
            if self.accept("("):
                args = []
                if not self.accept(")"):
                    while True:
                        args.append(self.expr())
                        if self.accept(")"):
                            break
                        self.expect(",")
                return Call(val, tuple(args))
            if self.accept("["):
                return Index(val, self.indices())
            return Var(val)
        if (kind, val) == ("op", "("):
            e = self.expr()
            self.expect(")")
            return e
        raise BUGSSyntaxError(f"unexpected token {val!r}")


def parse(text: str) -> Tuple[Stmt, ...]:
    """Parse BUGS model text into a tuple of statements."""
    return _Parser(text).program()


def free_vars(e: Expr, acc=None):
    """All variable names referenced by an expression (arrays and scalars)."""
    if acc is None:
        acc = []
    if isinstance(e, Var):
        if e.name not in acc:
            acc.append(e.name)
    elif isinstance(e, Index):
        if e.name not in acc:
            acc.append(e.name)
        for i in e.idx:
            free_vars(i, acc)
    elif isinstance(e, Call):
        for a in e.args:
            free_vars(a, acc)
    elif isinstance(e, BinOp):
        free_vars(e.a, acc)
        free_vars(e.b, acc)
    elif isinstance(e, Neg):
        free_vars(e.a, acc)
    elif isinstance(e, Range):
        free_vars(e.lo, acc)
        free_vars(e.hi, acc)
    return acc
