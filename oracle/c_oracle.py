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
    hdrs = [os.path.join(_HERE, "..", "include", h) for h in ("jbugs_plan.h", "jbugs_trunc.h")]

    def stale():
        return not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs)

    if force or stale():
        # several pytest-xdist workers may get here at once: one rebuilds, the others wait for the lock and find it done
        import fcntl
        with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            if force or stale():
                subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _bind(ctypes.CDLL(build()))
    return _LIB


def _bind(L):
    if True:
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
    return L


_NATIVE = None
NATIVE_FLAGS = ["-O3", "-march=native", "-fPIC", "-fopenmp", "-std=gnu11", "-fno-fast-math"]


def native_lib():
    """the same source compiled ON THIS MACHINE with -O3 -march=native (FMA contraction allowed): the strongest build of
    the CPU port, used only as bench.py's timed baseline (the checker stays the strict -O2 -ffp-contract=off build, whose
    results do not depend on the host).  Falls back to the checker build when there is no compiler."""
    global _NATIVE
    if _NATIVE is None:
        import tempfile
        try:
            out = os.path.join(tempfile.mkdtemp(prefix="jbugs_oracle_native_"), "libjbugs_oracle_native.so")
            subprocess.check_call(["gcc"] + NATIVE_FLAGS + ["-shared", "-o", out, os.path.join(_HERE, "jbugs_oracle.c"), "-lm"],
                                  stderr=subprocess.DEVNULL)
            _NATIVE = (_bind(ctypes.CDLL(out)), " ".join(["gcc"] + NATIVE_FLAGS))
        except Exception:
            _NATIVE = (lib(), "gcc -O2 -ffp-contract=off (no compiler on this box for the native build)")
    return _NATIVE


class COracle:
    """Evaluates a `jbugs.plan.Plan` on the CPU, one theta per thread."""

    def __init__(self, plan, native=False):
        self.L = native_lib()[0] if native else lib()
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
