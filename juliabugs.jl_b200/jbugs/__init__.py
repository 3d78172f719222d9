"""jbugs -- host-side mirror of the JuliaBUGS hot path over libjbugs_cuda (B200, sm_100a).

`from jbugs import compile, getparams, LogDensityProblems, UseAutoMarginalization, UseCUDABatch, ...`
"""
from . import examples, plan  # noqa: F401
from .lowering import Lowering, LoweringError, lower  # noqa: F401
from .model import (  # noqa: F401
    BUGSModel, EvaluationMode, LogDensityOrder, LogDensityProblems, UseAutoMarginalization, UseCUDABatch,
    UseGeneratedLogDensityFunction, UseGraph, compile, condition, decondition, evaluate_with_values, getparams,
    set_evaluation_mode, settrans,
)
from .cuda import CudaPlan, JBugsCudaError  # noqa: F401
from . import dist, gq, gq_device, hmc  # noqa: F401,E402
