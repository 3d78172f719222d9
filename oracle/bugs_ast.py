"""TEST INFRASTRUCTURE ONLY -- the oracle's OWN reader of BUGS model text.

The per-node oracle (graph_oracle.py) used to import the product's parser (`jbugs/parser.py`), so a
front-end bug was common-mode: invisible to every parity test.  This module is an independent
reader -- a different technique (one regular-expression scanner feeding a precedence-climbing
expression parser driven by a binding-power table) written without reference to the product's
recursive-descent code -- so that `tests/test_oracle_goldens.py::test_oracle_and_product_front_ends_agree`
can compare the two trees statement by statement.  Nothing under `oracle/` imports `jbugs`.

What it follows in the reference (paths relative to /root/reference/JuliaBUGS):
  * token classes and the dotted-name rule (`alpha.c`, `tau.b1`)   src/parser/bugs_parser.jl (`process_variable!`)
  * statements: `for (i in a:b) { ... }`, `lhs ~ dist(...)`, `lhs <- expr` / `lhs = expr`
                                                                    src/parser/bugs_parser.jl (`process_for!`, `process_assignment!`)
  * link functions on the left-hand side: `logit(p[i]) <- e` means `p[i] = logistic(e)`; `log` -> `exp`,
    `cloglog` -> `cexpexp`, `probit` -> `phi`                       src/parser/bugs_macro.jl:159-182
  * an empty index `x[]` / `x[,1]` is the whole axis (Julia `:`)    src/parser/bugs_parser.jl (`process_indexing!`)
  * operator precedence is R's / Julia's: `^` (right-assoc) binds tighter than unary minus, which binds
    tighter than `* /`, then `+ -`; `a:b` binds loosest.

Node types are plain tuples-with-names (collections.namedtuple), with the same field names the
oracle's evaluator reads: Num(value, is_int) Var(name) Range(lo, hi) Colon() Index(name, idx)
Call(fn, args) BinOp(op, a, b) Neg(a) | Stoch(lhs, dist) Det(lhs, rhs) For(var, lo, hi, body).
"""
from __future__ import annotations

import re
from collections import namedtuple

Num = namedtuple("Num", "value is_int")
Var = namedtuple("Var", "name")
Range = namedtuple("Range", "lo hi")
Colon = namedtuple("Colon", "")
Index = namedtuple("Index", "name idx")
Call = namedtuple("Call", "fn args")
BinOp = namedtuple("BinOp", "op a b")
Neg = namedtuple("Neg", "a")
Stoch = namedtuple("Stoch", "lhs dist")
Det = namedtuple("Det", "lhs rhs")
For = namedtuple("For", "var lo hi body")

INVERSE_OF_LINK = {"logit": "logistic", "log": "exp", "cloglog": "cexpexp", "probit": "phi"}


class ModelTextError(ValueError):
    pass


_SCAN = re.compile(r"""
      (?P<skip>   [ \t\r\n]+ | \#[^\n]* )
    | (?P<number> (?: \d+ (?: \.\d* )? | \.\d+ ) (?: [eE][+-]?\d+ )? )
    | (?P<ident>  [A-Za-z_]\w* (?: \.\w+ )* )
    | (?P<arrow>  <- )
    | (?P<punct>  [~=+\-*/^(){}\[\],:;] )
""", re.X)


def scan(text):
    """list of (kind, text) with kind in number / ident / sym, terminated by ('end', '')"""
    toks, at = [], 0
    n = len(text)
    while at < n:
        m = _SCAN.match(text, at)
        if not m:
            raise ModelTextError(f"cannot read model text at offset {at}: {text[at:at + 12]!r}")
        at = m.end()
        k = m.lastgroup
        if k == "skip":
            continue
        toks.append(("sym" if k in ("arrow", "punct") else k, m.group()))
    toks.append(("end", ""))
    return toks


# binding powers of the infix operators: (left, right); right < left makes `^` right-associative
_INFIX = {"+": (10, 11), "-": (10, 11), "*": (20, 21), "/": (20, 21), "^": (40, 39)}
_PREFIX_MINUS = 30   # tighter than * and /, looser than ^   (-x^2 == -(x^2),  -a*b == (-a)*b)


class Reader:
    def __init__(self, text):
        self.t = scan(text)
        self.k = 0

    # ---- token access
    def look(self, ahead=0):
        return self.t[min(self.k + ahead, len(self.t) - 1)]

    def take(self):
        tok = self.t[self.k]
        self.k += 1
        return tok

    def at_sym(self, s, ahead=0):
        return self.look(ahead) == ("sym", s)

    def eat(self, s):
        if self.at_sym(s):
            self.k += 1
            return True
        return False

    def need(self, s):
        if not self.eat(s):
            raise ModelTextError(f"expected '{s}' but found '{self.look()[1]}' (token {self.k})")

    def at_word(self, w, ahead=0):
        return self.look(ahead) == ("ident", w)

    # ---- program and statements
    def model(self):
        if self.at_word("model") and self.at_sym("{", 1):
            self.take()
        braced = self.eat("{")
        body = self.block()
        if braced:
            self.need("}")
        if self.look()[0] != "end":
            raise ModelTextError(f"unexpected '{self.look()[1]}' after the model")
        return tuple(body)

    def block(self):
        out = []
        while True:
            while self.eat(";"):
                pass
            if self.look()[0] == "end" or self.at_sym("}"):
                return out
            out.append(self.statement())

    def statement(self):
        if self.at_word("for") and self.at_sym("(", 1):
            self.take()
            self.need("(")
            kind, name = self.take()
            if kind != "ident":
                raise ModelTextError("a loop needs a variable name")
            if not self.at_word("in"):
                raise ModelTextError("'in' expected in the loop header")
            self.take()
            lo = self.arith(0)
            self.need(":")
            hi = self.arith(0)
            self.need(")")
            self.need("{")
            body = self.block()
            self.need("}")
            return For(name, lo, hi, tuple(body))
        # link(lhs) <- e      (only when the parenthesised thing is a plain variable reference followed by an assignment)
        if self.look()[0] == "ident" and self.look()[1] in INVERSE_OF_LINK and self.at_sym("(", 1):
            mark = self.k
            link = self.take()[1]
            self.take()
            try:
                target = self.reference()
                if self.eat(")") and (self.eat("<-") or self.eat("=")):
                    return Det(target, Call(INVERSE_OF_LINK[link], (self.expression(),)))
            except ModelTextError:
                pass
            self.k = mark
        target = self.reference()
        if self.eat("~"):
            d = self.atom()
            if not isinstance(d, Call):
                raise ModelTextError("a distribution call must follow '~'")
            # dist(...) C(lo, hi) / T(lo, hi) / I(lo, hi): censored(dist, lo, hi) / truncated(dist, lo, hi), an empty
            # bound is `nothing` (J/src/parser/bugs_parser.jl:470-500)
            if any(self.at_word(w) for w in ("C", "T", "I")) and self.at_sym("(", 1):
                which = self.take()[1]
                self.take()
                bounds = []
                for closer in (",", ")"):
                    if self.at_sym(closer):
                        bounds.append(Var("nothing"))
                    else:
                        bounds.append(self.expression())
                    self.need(closer)
                d = Call("truncated" if which == "T" else "censored", (d, bounds[0], bounds[1]))
            return Stoch(target, d)
        if self.eat("<-") or self.eat("="):
            return Det(target, self.expression())
        raise ModelTextError(f"'~' or '<-' expected after {target!r}")

    def reference(self):
        kind, name = self.take()
        if kind != "ident":
            raise ModelTextError(f"a variable was expected, found '{name}'")
        if self.eat("["):
            return Index(name, self.subscripts())
        return Var(name)

    def subscripts(self):
        subs = []
        while True:
            if self.at_sym(",") or self.at_sym("]"):
                subs.append(Colon())        # empty slot: the whole axis
            else:
                subs.append(self.expression())
            if self.eat("]"):
                return tuple(subs)
            self.need(",")

    # ---- expressions
    def expression(self):
        """arith [ ':' arith ]   -- a bare ':' inside a subscript list is the whole axis"""
        if self.at_sym(":") and (self.at_sym(",", 1) or self.at_sym("]", 1)):
            self.take()
            return Colon()
        left = self.arith(0)
        if self.eat(":"):
            return Range(left, self.arith(0))
        return left

    def arith(self, min_bp):
        tok = self.look()
        if tok == ("sym", "-"):
            self.take()
            left = Neg(self.arith(_PREFIX_MINUS))
        elif tok == ("sym", "+"):
            self.take()
            left = self.arith(_PREFIX_MINUS)
        else:
            left = self.atom()
        while True:
            kind, op = self.look()
            if kind != "sym" or op not in _INFIX:
                return left
            lbp, rbp = _INFIX[op]
            if lbp < min_bp:
                return left
            self.take()
            left = BinOp(op, left, self.arith(rbp))

    def atom(self):
        kind, s = self.take()
        if kind == "number":
            return Num(float(s), s.isdigit())
        if kind == "ident":
            if self.eat("("):
                args = []
                if not self.eat(")"):
                    while True:
                        args.append(self.expression())
                        if self.eat(")"):
                            break
                        self.need(",")
                return Call(s, tuple(args))
            if self.eat("["):
                return Index(s, self.subscripts())
            return Var(s)
        if (kind, s) == ("sym", "("):
            inner = self.expression()
            self.need(")")
            return inner
        raise ModelTextError(f"unexpected '{s}' in an expression")


def parse(text):
    """BUGS model text -> tuple of statements"""
    return Reader(text).model()


def as_plain(node):
    """a tree of plain tuples ('TypeName', fields...) -- for comparing trees built from different node classes"""
    if isinstance(node, tuple) and hasattr(node, "_fields"):
        return (type(node).__name__,) + tuple(as_plain(getattr(node, f)) for f in node._fields)
    if isinstance(node, tuple):
        return tuple(as_plain(x) for x in node)
    if hasattr(node, "__dataclass_fields__"):
        return (type(node).__name__,) + tuple(as_plain(getattr(node, f)) for f in node.__dataclass_fields__)
    return node
