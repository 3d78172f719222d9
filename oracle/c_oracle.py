"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/libjbugs_oracle.so (jbugs_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product path (juliabugs.jl_b200/jbugs) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libjbugs_oracle.so")
    src = os.path.join(_HERE, "jbugs_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "jbugs_plan.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.jbo_plan_create.restype = ctypes.c_void_p
        L.jbo_plan_create.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        L.jbo_plan_destroy.argtypes = [ctypes.c_void_p]
        L.jbo_dim.restype = ctypes.c_int64
        L.jbo_dim.argtypes = [ctypes.c_void_p]
        L.jbo_plan_set_data.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        L.jbo_plan_set_shard.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64]
        L.jbo_eval_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.jbo_max_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


class COracle:
    """Evaluates a `jbugs.plan.Plan` on the CPU, one theta per thread."""

    def __init__(self, plan):
        self.L = lib()
        blob = plan.to_bytes()
        self._blob = ctypes.create_string_buffer(blob, len(blob))
        self.h = self.L.jbo_plan_create(self._blob, len(blob))
        if not self.h:
            raise ValueError("oracle rejected the plan blob")
        self.D = int(self.L.jbo_dim(self.h))
        self._keep = []
        for k, slot in enumerate(plan.data):
            a = np.ascontiguousarray(slot.values)
            self._keep.append(a)
            rc = self.L.jbo_plan_set_data(self.h, k, a.ctypes.data, a.nbytes)
            if rc:
                raise ValueError(f"data slot {k}: size mismatch")

    def __del__(self):
        try:
            self.L.jbo_plan_destroy(self.h)
        except Exception:
            pass

    def max_threads(self):
        return int(self.L.jbo_max_threads())

    def eval_batch(self, theta, layout="chain_fast", want_grad=True, nthreads=0):
        """theta: (D, C) array. layout 'chain_fast' expects C-contiguous (D, C); 'param_fast' expects
        an F-contiguous (D, C) (== Julia Matrix{Float64}(D, C)).  Returns logp (C,), grad like theta, status."""
        theta = np.asarray(theta, dtype=np.float64)
        if theta.ndim == 1:
            theta = theta.reshape(-1, 1)
        D, C = theta.shape
        assert D == self.D, (D, self.D)
        if layout == "chain_fast":
            th = np.ascontiguousarray(theta)
            ld, lay = C, 0
            grad = np.zeros((D, C), order="C")
        else:
            th = np.asfortranarray(theta)
            ld, lay = D, 1
            grad = np.zeros((D, C), order="F")
        logp = np.zeros(C)
        status = np.zeros(C, dtype=np.int32)
        rc = self.L.jbo_eval_batch(self.h, th.ctypes.data, ld, C, lay, logp.ctypes.data,
                                   grad.ctypes.data if want_grad else None, status.ctypes.data,
                                   1 if want_grad else 0, nthreads)
        if rc:
            raise RuntimeError(f"jbo_eval_batch failed: {rc}")
        return logp, (grad if want_grad else None), status
