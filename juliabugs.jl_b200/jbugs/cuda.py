"""ctypes binding of `csrc/libjbugs_cuda.so` (C ABI in `include/jbugs_cuda.h`).

This is the Python twin of the Julia `ccall` stubs in `julia/JuliaBUGSCUDABatch.jl`.  The product
path has no CPU fallback: if the shared library is missing or no CUDA device is present, every
compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JBUGS_CUDA_LIB") or os.path.join(os.path.dirname(_HERE), "csrc", "libjbugs_cuda.so")

LAYOUT_CHAIN_FAST, LAYOUT_PARAM_FAST = 0, 1
EVAL_ASYNC = 1

SYMBOLS = [
    "jbugs_last_error", "jbugs_version", "jbugs_device_count", "jbugs_plan_create", "jbugs_plan_destroy",
    "jbugs_dim", "jbugs_plan_upload_data", "jbugs_plan_set_shard", "jbugs_plan_set_dense_shard", "jbugs_eval_batch",
    "jbugs_batch_alloc",
    "jbugs_batch_free", "jbugs_host_alloc", "jbugs_host_free", "jbugs_comm_unique_id", "jbugs_comm_init",
    "jbugs_comm_destroy", "jbugs_launch_count", "jbugs_plan_describe", "jbugs_plan_set_option",
    "jbugs_last_timing", "jbugs_hmc_create", "jbugs_hmc_destroy", "jbugs_hmc_init", "jbugs_hmc_run",
    "jbugs_hmc_get_state", "jbugs_hmc_run_nuts", "jbugs_hmc_nuts_stats", "jbugs_plan_specialize", "jbugs_hmc_draw_momentum", "jbugs_timing_history", "jbugs_host_alloc_ex",
    "jbugs_bind_host_to_device", "jbugs_fp64_peak", "jbugs_eval_batch_tempered", "jbugs_jit_last_compile_seconds", "jbugs_plan_specialize_async",
    "jbugs_gq_create", "jbugs_gq_destroy", "jbugs_gq_run", "jbugs_gq_launch_count", "jbugs_plan_synchronize",
]


class JBugsCudaError(RuntimeError):
    pass


_lib = None


def load():
    """Load libjbugs_cuda.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JBugsCudaError(
            f"{LIB_PATH} is missing: build it with `make -C juliabugs.jl_b200/csrc` "
            "(or __graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    L.jbugs_last_error.restype = c.c_char_p
    L.jbugs_version.restype = c.c_int
    L.jbugs_device_count.restype = c.c_int
    L.jbugs_plan_create.argtypes = [c.c_void_p, c.c_size_t, c.c_int, c.POINTER(c.c_void_p)]
    L.jbugs_plan_destroy.argtypes = [c.c_void_p]
    L.jbugs_dim.restype = c.c_int64
    L.jbugs_dim.argtypes = [c.c_void_p]
    L.jbugs_plan_upload_data.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_size_t, c.c_int]
    L.jbugs_plan_set_shard.argtypes = [c.c_void_p, c.c_int, c.c_int64, c.c_int64]
    L.jbugs_plan_set_dense_shard.argtypes = [c.c_void_p, c.c_int64, c.c_int64]
    L.jbugs_plan_specialize.argtypes = [c.c_void_p, c.c_char_p]
    L.jbugs_jit_last_compile_seconds.restype = c.c_double
    L.jbugs_plan_specialize_async.argtypes = [c.c_void_p, c.c_char_p]
    L.jbugs_plan_synchronize.argtypes = [c.c_void_p, c.c_void_p]
    L.jbugs_eval_batch.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_int64, c.c_int, c.c_void_p, c.c_void_p,
                                   c.c_void_p, c.c_int, c.c_int, c.c_void_p]
    L.jbugs_eval_batch_tempered.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_int64, c.c_int, c.c_double, c.c_void_p,
                                            c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_void_p]
    L.jbugs_batch_alloc.argtypes = [c.c_void_p, c.c_int64, c.POINTER(c.c_void_p), c.POINTER(c.c_void_p),
                                    c.POINTER(c.c_void_p), c.POINTER(c.c_void_p), c.POINTER(c.c_int64)]
    L.jbugs_batch_free.argtypes = [c.c_void_p]
    L.jbugs_host_alloc.argtypes = [c.POINTER(c.c_void_p), c.c_size_t]
    L.jbugs_host_free.argtypes = [c.c_void_p]
    L.jbugs_comm_unique_id.argtypes = [c.c_void_p]
    L.jbugs_comm_init.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_void_p]
    L.jbugs_comm_destroy.argtypes = [c.c_void_p]
    L.jbugs_launch_count.restype = c.c_int64
    L.jbugs_launch_count.argtypes = [c.c_void_p]
    L.jbugs_plan_describe.argtypes = [c.c_void_p, c.c_char_p, c.c_size_t]
    L.jbugs_plan_set_option.argtypes = [c.c_void_p, c.c_char_p, c.c_int64]
    L.jbugs_last_timing.argtypes = [c.c_void_p, c.POINTER(c.c_float), c.POINTER(c.c_float)]
    L.jbugs_timing_history.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.POINTER(c.c_int)]
    L.jbugs_host_alloc_ex.argtypes = [c.POINTER(c.c_void_p), c.c_size_t, c.c_int]
    L.jbugs_fp64_peak.argtypes = [c.c_int, c.POINTER(c.c_double)]
    L.jbugs_bind_host_to_device.argtypes = [c.c_int, c.POINTER(c.c_int), c.POINTER(c.c_int)]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise JBugsCudaError(f"libjbugs_cuda error {rc}: {load().jbugs_last_error().decode()}")


def device_count() -> int:
    return int(load().jbugs_device_count())


def _ptr(x):
    """address of a numpy array (host) or of anything exposing data_ptr() (torch CUDA tensor), or int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


class CudaPlan:
    """Owns one `jbugs_plan*`: device copies of the model data plus the evaluation workspaces."""

    def __init__(self, plan, device: int = 0, upload: bool = True):
        self.L = load()
        self.plan = plan
        blob = plan.to_bytes()
        h = ctypes.c_void_p()
        _check(self.L.jbugs_plan_create(blob, len(blob), device, ctypes.byref(h)))
        self.h = h
        self.device = device
        self.D = int(self.L.jbugs_dim(self.h))
        if upload:
            for k, slot in enumerate(plan.data):
                self.upload(k, slot.values)

    def close(self):
        if getattr(self, "h", None):
            self.L.jbugs_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data ----------------------------------------------------------------------------------
    def upload(self, slot: int, values):
        a = np.ascontiguousarray(values)
        dtype = 1 if a.dtype == np.int64 else 0
        if dtype == 0:
            a = np.ascontiguousarray(a, dtype=np.float64)
        _check(self.L.jbugs_plan_upload_data(self.h, slot, a.ctypes.data, a.nbytes, dtype))

    def set_shard(self, plate: int, i0: int, i1: int):
        _check(self.L.jbugs_plan_set_shard(self.h, plate, i0, i1))

    def specialize(self, cache_dir: Optional[str] = None):
        """compile the plate kernel for the exact shape of every plate that has no in-tree specialisation, expression
        tapes included (NVRTC in-process, cubins cached in a per-user directory; "none" disables the cache); without it
        those plates run the interpreter instantiations."""
        _check(self.L.jbugs_plan_specialize(self.h, cache_dir.encode() if cache_dir else None))

    def synchronize(self, stream=None):
        """wait for earlier asynchronous evaluations (flags = JBUGS_EVAL_ASYNC) of this plan"""
        _check(self.L.jbugs_plan_synchronize(self.h, stream))

    def specialize_async(self, cache_dir: Optional[str] = None):
        """`specialize` without blocking: the compiler runs on a worker thread, evaluations switch over when it is done"""
        _check(self.L.jbugs_plan_specialize_async(self.h, cache_dir.encode() if cache_dir else None))

    def last_compile_seconds(self) -> float:
        """seconds NVRTC spent in the last `specialize` call of this process (0.0: everything came from the cache)"""
        return float(self.L.jbugs_jit_last_compile_seconds())

    def set_dense_shard(self, i0: int, i1: int):
        _check(self.L.jbugs_plan_set_dense_shard(self.h, i0, i1))

    def set_option(self, key: str, value: int):
        _check(self.L.jbugs_plan_set_option(self.h, key.encode(), int(value)))

    def describe(self) -> str:
        buf = ctypes.create_string_buffer(4096)
        _check(self.L.jbugs_plan_describe(self.h, buf, len(buf)))
        return buf.value.decode()

    def launch_count(self) -> int:
        return int(self.L.jbugs_launch_count(self.h))

    def last_timing(self):
        a, b = ctypes.c_float(), ctypes.c_float()
        _check(self.L.jbugs_last_timing(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def timing_history(self, cap: int = 1024):
        """(plate_ms, total_ms) arrays of the evaluations since set_option("timing", 2), oldest first."""
        a = np.zeros(cap, dtype=np.float32)
        b = np.zeros(cap, dtype=np.float32)
        n = ctypes.c_int()
        _check(self.L.jbugs_timing_history(self.h, a.ctypes.data, b.ctypes.data, cap, ctypes.byref(n)))
        return a[:n.value].astype(float), b[:n.value].astype(float)

    # -- evaluation -------------------------------------------------------------------------------
    def eval_raw(self, theta, ld, C, layout, logp, grad, status, want_grad=True, flags=0, stream=None):
        """thin call: every buffer is an address / numpy array / torch tensor, nothing is allocated."""
        _check(self.L.jbugs_eval_batch(self.h, _ptr(theta), ld, C, layout, _ptr(logp), _ptr(grad), _ptr(status),
                                       1 if want_grad else 0, flags, stream))

    def eval_host(self, theta: np.ndarray, layout="chain_fast", want_grad=True):
        """theta: (D, C) float64 numpy array (C-contiguous for chain_fast, F-contiguous for param_fast).
        Returns (logp (C,), grad (D, C) in the same memory order, status (C,))."""
        theta = np.asarray(theta, dtype=np.float64)
        if theta.ndim == 1:
            theta = theta.reshape(-1, 1)
        D, C = theta.shape
        if D != self.D:
            raise ValueError(f"theta has {D} rows, model dimension is {self.D}")
        if layout == "chain_fast":
            th = np.ascontiguousarray(theta)
            grad = np.empty((D, C), order="C") if want_grad else None
            ld, lay = C, LAYOUT_CHAIN_FAST
        else:
            th = np.asfortranarray(theta)
            grad = np.empty((D, C), order="F") if want_grad else None
            ld, lay = D, LAYOUT_PARAM_FAST
        logp = np.empty(C)
        status = np.empty(C, dtype=np.int32)
        self.eval_raw(th, ld, C, lay, logp, grad, status, want_grad)
        return logp, grad, status

    def eval_tempered(self, theta: np.ndarray, temperature: float = 1.0, layout="param_fast", want_grad=True):
        """(logprior, loglikelihood, tempered_logjoint, grad of the tempered log joint, status) for a (D, C) batch"""
        theta = np.asarray(theta, dtype=np.float64)
        if theta.ndim == 1:
            theta = theta.reshape(-1, 1)
        D, C = theta.shape
        if layout == "chain_fast":
            th, grad, ld, lay = np.ascontiguousarray(theta), np.empty((D, C), order="C"), C, LAYOUT_CHAIN_FAST
        else:
            th, grad, ld, lay = np.asfortranarray(theta), np.empty((D, C), order="F"), D, LAYOUT_PARAM_FAST
        lpr, ll, lp = np.empty(C), np.empty(C), np.empty(C)
        status = np.empty(C, dtype=np.int32)
        _check(self.L.jbugs_eval_batch_tempered(self.h, th.ctypes.data, ld, C, lay, float(temperature), lpr.ctypes.data,
                                                ll.ctypes.data, lp.ctypes.data, grad.ctypes.data if want_grad else None,
                                                status.ctypes.data, 1 if want_grad else 0, None))
        return lpr, ll, lp, (grad if want_grad else None), status

    def batch_alloc(self, C: int):
        th, g, lp, st = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        ld = ctypes.c_int64()
        _check(self.L.jbugs_batch_alloc(self.h, C, ctypes.byref(th), ctypes.byref(g), ctypes.byref(lp),
                                        ctypes.byref(st), ctypes.byref(ld)))
        return th.value, g.value, lp.value, st.value, int(ld.value)

    def batch_free(self):
        _check(self.L.jbugs_batch_free(self.h))

    # -- multi GPU -----------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        _check(load().jbugs_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        _check(self.L.jbugs_comm_init(self.h, nranks, rank, ctypes.c_char_p(unique_id)))


def host_alloc(nbytes: int, write_combined: bool = False) -> Optional[int]:
    p = ctypes.c_void_p()
    _check(load().jbugs_host_alloc_ex(ctypes.byref(p), nbytes, 1 if write_combined else 0))
    return p.value


def fp64_peak(device: int = 0) -> float:
    """measured FP64 fused multiply-adds per second of the device (independent DFMA chains)."""
    v = ctypes.c_double()
    _check(load().jbugs_fp64_peak(device, ctypes.byref(v)))
    return v.value


def bind_host_to_device(device: int):
    """pin this process to the CPUs / memory of the GPU's NUMA node; returns (node, n_cpus), node -1 if unknown."""
    node, ncpu = ctypes.c_int(-1), ctypes.c_int(0)
    _check(load().jbugs_bind_host_to_device(device, ctypes.byref(node), ctypes.byref(ncpu)))
    return node.value, ncpu.value


def host_free(ptr: int):
    _check(load().jbugs_host_free(ptr))
