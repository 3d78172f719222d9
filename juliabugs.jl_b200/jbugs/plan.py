"""Flat kernel plan: Python mirror of `include/jbugs_plan.h` (records + (de)serialisation).

The plan is what the lowering pass (`lowering.py`) emits and what `libjbugs_cuda`
(`jbugs_plan_create`) consumes.  The numpy structured dtypes below are laid out byte for byte
like the C structs; `tests/test_plan_abi.py` checks the sizes against the compiled header.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

MAGIC = b"JBUGSPL1"
VERSION = 1
FLAG_TRANSFORMED = 1

# enum jbugs_arg_kind
ARG_CONST, ARG_GVAL, ARG_LOCAL, ARG_DATA_I, ARG_DATA_J, ARG_DATA_IJ, ARG_ONE, ARG_XVAL, ARG_GVEC_J = range(9)
# enum jbugs_family
(FAM_NORMAL, FAM_GAMMA, FAM_EXPONENTIAL, FAM_UNIFORM, FAM_BETA, FAM_LOGNORMAL, FAM_BERNOULLI,
 FAM_BINOMIAL, FAM_POISSON, FAM_FLAT, FAM_LOGISTIC, FAM_LAPLACE, FAM_WEIBULL, FAM_NEGBIN, FAM_CHISQ, FAM_T,
 FAM_GEOMETRIC, FAM_CATEGORICAL, FAM_BETABIN) = range(19)
FAMILY_OF = {
    "dnorm": FAM_NORMAL, "dgamma": FAM_GAMMA, "dexp": FAM_EXPONENTIAL, "dunif": FAM_UNIFORM,
    "dbeta": FAM_BETA, "dlnorm": FAM_LOGNORMAL, "dbern": FAM_BERNOULLI, "dbin": FAM_BINOMIAL,
    "dpois": FAM_POISSON, "dflat": FAM_FLAT, "dlogis": FAM_LOGISTIC, "ddexp": FAM_LAPLACE, "dweib": FAM_WEIBULL,
    "dnegbin": FAM_NEGBIN, "dchisqr": FAM_CHISQ, "dt": FAM_T, "dgeom": FAM_GEOMETRIC, "dcat": FAM_CATEGORICAL,
    "dbetabin": FAM_BETABIN,
}
FAMILY_NAME = {v: k for k, v in FAMILY_OF.items()}
FAMILY_NARGS = {FAM_NORMAL: 2, FAM_GAMMA: 2, FAM_EXPONENTIAL: 1, FAM_UNIFORM: 2, FAM_BETA: 2,
                FAM_LOGNORMAL: 2, FAM_BERNOULLI: 1, FAM_BINOMIAL: 2, FAM_POISSON: 1, FAM_FLAT: 0,
                FAM_LOGISTIC: 2, FAM_LAPLACE: 2, FAM_WEIBULL: 2, FAM_NEGBIN: 2, FAM_CHISQ: 1, FAM_T: 3,
                FAM_GEOMETRIC: 1, FAM_CATEGORICAL: 1, FAM_BETABIN: 3}
DISCRETE = {FAM_BERNOULLI, FAM_BINOMIAL, FAM_POISSON, FAM_NEGBIN, FAM_GEOMETRIC, FAM_CATEGORICAL, FAM_BETABIN}
# enum jbugs_transform
TR_IDENTITY, TR_LOG, TR_LOGIT, TR_UPPER = range(4)
# enum jbugs_op
OP_IDENTITY, OP_EXP, OP_LOGISTIC, OP_CEXPEXP, OP_PHI, OP_LOG, OP_SQRT, OP_NEG = range(8)
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW = 16, 17, 18, 19, 20
OP_LEAF, OP_SELECT, OP_PICK_K, OP_GVEC_AT = 32, 33, 34, 35   # expression tapes only
MAX_XOPS = 64
MAX_LOC = 4    # local parameters per plate (csrc/jbugs_kernels.cuh)
LINK_OF = {"exp": OP_EXP, "logistic": OP_LOGISTIC, "ilogit": OP_LOGISTIC, "cexpexp": OP_CEXPEXP,
           "icloglog": OP_CEXPEXP, "phi": OP_PHI}
UNARY_OF = dict(LINK_OF, log=OP_LOG, sqrt=OP_SQRT)
LEVEL_GROUP, LEVEL_OBS = 0, 1
MAX_LINK = 4

ARG_DT = np.dtype([("kind", "<u4"), ("id", "<i4"), ("value", "<f8")])
GLOBAL_DT = np.dtype([("theta_idx", "<i8"), ("transform", "<u4"), ("family", "<u4"),
                      ("lb", "<f8"), ("ub", "<f8"), ("args", ARG_DT, (3,))])
GOP_DT = np.dtype([("op", "<u4"), ("pad", "<u4"), ("a", ARG_DT), ("b", ARG_DT)])
DATA_DT = np.dtype([("len", "<i8"), ("dim0", "<i8"), ("rank", "<u4"), ("dtype", "<u4")])
LOCAL_DT = np.dtype([("level", "<u4"), ("transform", "<u4"), ("theta_base", "<i8"),
                     ("stride_i", "<i8"), ("stride_j", "<i8"), ("idx_slot", "<i4"),
                     ("family", "<u4"), ("lb", "<f8"), ("ub", "<f8"), ("args", ARG_DT, (3,))])
TERM_DT = np.dtype([("coef", ARG_DT), ("cov", ARG_DT)])
PLATE_DT = np.dtype([("N", "<i8"), ("T", "<i8"), ("first_local", "<u4"), ("n_locals", "<u4"),
                     ("first_term", "<u4"), ("n_terms", "<u4"), ("n_link", "<u4"),
                     ("link", "<u4", (MAX_LINK,)), ("family", "<u4"), ("eta_arg", "<u4"),
                     ("has_obs", "<u4"), ("y", ARG_DT), ("aux", ARG_DT, (3,))])
HEADER_DT = np.dtype([("magic", "S8"), ("version", "<u4"), ("flags", "<u4"), ("D", "<i8"),
                      ("n_globals", "<u4"), ("n_gtape", "<u4"), ("n_data", "<u4"),
                      ("n_plates", "<u4"), ("n_locals", "<u4"), ("n_terms", "<u4"),
                      ("n_dense", "<u4"), ("n_xops", "<u4"), ("reserved", "<u8")])
DENSE_DT = np.dtype([("N", "<i8"), ("P", "<i8"), ("theta_base", "<i8"), ("theta_stride", "<i8"),
                     ("x_slot", "<i4"), ("y_slot", "<i4"), ("n_slot", "<i4"),
                     ("family", "<u4"), ("link", "<u4"), ("pad", "<u4")])
XOP_DT = np.dtype([("op", "<u4"), ("a", "<u2"), ("b", "<u2"), ("c", "<u2"), ("pad", "<u2"), ("pad2", "<u4"), ("leaf", ARG_DT)])
assert XOP_DT.itemsize == 32
assert ARG_DT.itemsize == 16 and GLOBAL_DT.itemsize == 80 and GOP_DT.itemsize == 40
assert DATA_DT.itemsize == 24 and LOCAL_DT.itemsize == 104 and TERM_DT.itemsize == 32
assert PLATE_DT.itemsize == 128 and HEADER_DT.itemsize == 64 and DENSE_DT.itemsize == 56


@dataclass(frozen=True)
class Arg:
    kind: int
    id: int = 0
    value: float = 0.0

    @staticmethod
    def const(v):
        return Arg(ARG_CONST, 0, float(v))

    @staticmethod
    def gval(i):
        return Arg(ARG_GVAL, int(i))

    @staticmethod
    def local(i):
        return Arg(ARG_LOCAL, int(i))

    @staticmethod
    def one():
        return Arg(ARG_ONE, 0, 1.0)

    @staticmethod
    def xval(i):
        return Arg(ARG_XVAL, int(i))

    def rec(self):
        return (self.kind, self.id, self.value)


ZERO_ARG = Arg.const(0.0)


@dataclass
class Global:
    name: str
    theta_idx: int
    transform: int
    family: int
    args: List[Arg]
    lb: float = -np.inf
    ub: float = np.inf


@dataclass
class GOp:
    op: int
    a: Arg
    b: Arg = ZERO_ARG
    name: str = ""


@dataclass
class DataSlot:
    name: str
    values: np.ndarray  # float64 (or int64 for index arrays), 1-D (rank-2 arrays flattened column-major)
    dim0: int = 0
    rank: int = 1

    @property
    def dtype_code(self):
        return 1 if self.values.dtype == np.int64 else 0


@dataclass
class Local:
    name: str
    level: int
    transform: int
    theta_base: int
    stride_i: int
    stride_j: int
    family: int
    args: List[Arg]
    lb: float = -np.inf
    ub: float = np.inf
    idx_slot: int = -1


@dataclass
class Term:
    coef: Arg
    cov: Arg


@dataclass
class XOp:
    """one op of a plate's expression tape (static single assignment: produces value slot = its position)"""
    op: int
    a: int = 0
    b: int = 0
    c: int = 0
    leaf: Arg = ZERO_ARG


@dataclass
class Plate:
    N: int
    T: int
    locals: List[Local] = field(default_factory=list)
    terms: List[Term] = field(default_factory=list)
    link: List[int] = field(default_factory=list)
    family: int = FAM_NORMAL
    eta_arg: int = 0
    has_obs: bool = True
    y: Arg = ZERO_ARG
    aux: List[Arg] = field(default_factory=list)
    name: str = ""
    n_real: int = -1               # observed cells when the plate is padded (weight given), else N * T
    weight: Optional[Arg] = None   # 0/1 per cell (DATA_I | DATA_IJ): padded cells of a plate gathered by a data index
    xops: List[XOp] = field(default_factory=list)   # expression plate (has_obs == 2 in the record): replaces terms/link
    n_states: int = 0              # > 0: one finite discrete latent per observation is summed out (has_obs == 3)
    logprior_slot: int = -1        # tape slot of log p(z = k)
    zobs_slot: int = -1            # tape slot of a data leaf holding the latent's observed state per cell (0 = missing)


@dataclass
class Dense:
    """y[i] ~ dbern|dbin(logistic(inprod(X[i, 1:P], beta[1:P])) [, n[i]]) -- evaluated as FP64 GEMMs over all chains."""
    N: int
    P: int
    theta_base: int
    theta_stride: int
    x_slot: int
    y_slot: int
    n_slot: int = -1
    family: int = FAM_BERNOULLI
    link: int = OP_LOGISTIC
    name: str = ""


@dataclass
class Plan:
    D: int
    transformed: bool
    globals: List[Global] = field(default_factory=list)
    gtape: List[GOp] = field(default_factory=list)
    data: List[DataSlot] = field(default_factory=list)
    plates: List[Plate] = field(default_factory=list)
    dense: List[Dense] = field(default_factory=list)
    param_names: Optional[List[str]] = None  # theta slot -> variable label (chain naming)

    @property
    def n_gvals(self):
        return len(self.globals) + len(self.gtape)

    def n_obs(self):
        return (sum((p.n_real if p.n_real >= 0 else p.N * p.T) for p in self.plates if p.has_obs)
                + sum(d.N for d in self.dense))

    # ------------------------------------------------------------------ serialisation
    def to_bytes(self) -> bytes:
        def args3(a):
            a = list(a) + [ZERO_ARG] * (3 - len(a))
            return [x.rec() for x in a[:3]]

        g = np.zeros(len(self.globals), GLOBAL_DT)
        for k, x in enumerate(self.globals):
            g[k] = (x.theta_idx, x.transform, x.family, x.lb, x.ub, args3(x.args))
        t = np.zeros(len(self.gtape), GOP_DT)
        for k, x in enumerate(self.gtape):
            t[k] = (x.op, 0, x.a.rec(), x.b.rec())
        d = np.zeros(len(self.data), DATA_DT)
        for k, x in enumerate(self.data):
            d[k] = (x.values.size, x.dim0 or x.values.size, x.rank, x.dtype_code)
        nl = sum(len(p.locals) for p in self.plates)
        nt = sum(len(p.terms) for p in self.plates)
        nx = sum(len(p.xops) for p in self.plates)
        loc = np.zeros(nl, LOCAL_DT)
        trm = np.zeros(nt, TERM_DT)
        xop = np.zeros(nx, XOP_DT)
        pl = np.zeros(len(self.plates), PLATE_DT)
        il = it = ix = 0
        for k, p in enumerate(self.plates):
            if len(p.link) > MAX_LINK:
                raise ValueError("link tape longer than JBUGS_MAX_LINK")
            if len(p.xops) > MAX_XOPS:
                raise ValueError("expression tape longer than JBUGS_MAX_XOPS")
            link = list(p.link) + [OP_IDENTITY] * (MAX_LINK - len(p.link))
            is_x = bool(p.xops)
            if p.n_states > 0:
                link[0] = p.logprior_slot
                link[1] = p.zobs_slot + 1
            pl[k] = (p.N, p.T, il, len(p.locals), ix if is_x else it, len(p.xops) if is_x else len(p.terms),
                     len(p.link), link, p.family,
                     p.n_states if p.n_states > 0 else p.eta_arg,
                     ((3 if p.n_states > 0 else 2) if is_x else 1) if p.has_obs else 0, p.y.rec(),
                     args3(p.aux if p.weight is None else (list(p.aux) + [ZERO_ARG, ZERO_ARG])[:2] + [p.weight]))
            for x in p.xops:
                xop[ix] = (x.op, x.a, x.b, x.c, 0, 0, x.leaf.rec())
                ix += 1
            for x in p.locals:
                loc[il] = (x.level, x.transform, x.theta_base, x.stride_i, x.stride_j, x.idx_slot,
                           x.family, x.lb, x.ub, args3(x.args))
                il += 1
            for x in p.terms:
                trm[it] = (x.coef.rec(), x.cov.rec())
                it += 1
        h = np.zeros(1, HEADER_DT)
        h[0] = (MAGIC, VERSION, FLAG_TRANSFORMED if self.transformed else 0, self.D,
                len(self.globals), len(self.gtape), len(self.data), len(self.plates), nl, nt,
                len(self.dense), nx, 0)
        dn = np.zeros(len(self.dense), DENSE_DT)
        for k, x in enumerate(self.dense):
            dn[k] = (x.N, x.P, x.theta_base, x.theta_stride, x.x_slot, x.y_slot, x.n_slot, x.family, x.link, 0)
        return b"".join(x.tobytes() for x in (h, g, t, d, pl, loc, trm, dn, xop))

    def describe(self) -> str:
        out = [f"plan D={self.D} transformed={self.transformed} globals={len(self.globals)} "
               f"gtape={len(self.gtape)} data={len(self.data)} plates={len(self.plates)}"]
        for g in self.globals:
            out.append(f"  global {g.name}: theta[{g.theta_idx}] tr={g.transform} {FAMILY_NAME[g.family]}")
        for p in self.plates:
            out.append(f"  plate {p.name}: N={p.N} T={p.T} fam={FAMILY_NAME[p.family]} link={p.link} "
                       f"terms={len(p.terms)} locals={[l.name for l in p.locals]}")
        return "\n".join(out)
