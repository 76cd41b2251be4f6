"""Loader for the native C++ runtime (``csrc/runtime`` -> ``_lib/libdtf_runtime.so``).

TF implements accumulators, FIFO queues, the tensor-bundle checkpoint format
and the step tracer in C++ (SURVEY §2.2 "Native in TF?").  This framework does
the same: ``libdtf_runtime.so`` (plain C ABI, loaded with ctypes) provides
them, with Python fallbacks of identical semantics when the library has not
been built (``python __graft_entry__.py build`` builds it).
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional

_LIB = None
_LIB_TRIED = False
_LOCK = threading.Lock()


def lib_path() -> str:
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(here, "_lib", "libdtf_runtime.so")


def load() -> Optional[ctypes.CDLL]:
    global _LIB, _LIB_TRIED
    with _LOCK:
        if _LIB_TRIED:
            return _LIB
        _LIB_TRIED = True
        if os.environ.get("DTF_DISABLE_NATIVE_RUNTIME") == "1":
            return None
        p = lib_path()
        if os.path.exists(p):
            try:
                _LIB = ctypes.CDLL(p)
                from . import _native_bindings
                _native_bindings.declare(_LIB)
            except OSError:
                _LIB = None
        return _LIB


def available() -> bool:
    return load() is not None


def make_accumulator(name: str):
    lib = load()
    if lib is not None:
        from ._native_bindings import NativeAccumulator
        return NativeAccumulator(lib, name)
    from ..parallel.ps_state import ConditionalAccumulator
    return ConditionalAccumulator(name=name)


def make_queue(name: str):
    lib = load()
    if lib is not None:
        from ._native_bindings import NativeQueue
        return NativeQueue(lib, name)
    from ..parallel.ps_state import FIFOQueue
    return FIFOQueue(name=name)


def cpu_optimizer_apply(kind: int, var, m, v, g, lr: float, momentum: float = 0.0, nesterov: bool = False,
                        beta1: float = 0.0, beta2: float = 0.0, eps: float = 0.0) -> bool:
    """Fused in-place apply on contiguous fp32 CPU tensors through ``csrc/runtime/cpu_kernels.cpp`` (kind 0 sgd,
    1 momentum, 2 TF-Adam with ``lr`` = the bias-corrected ``lr_t``).  Returns False when the native library or the
    layout does not allow it -- the caller then runs the equivalent torch ops."""
    import torch
    lib = load()
    if lib is None or not hasattr(lib, "dtf_cpu_optimizer_apply"):
        return False
    for t in (var, m, v, g):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.device.type != "cpu"):
            return False
    if g.numel() != var.numel():
        return False
    rc = lib.dtf_cpu_optimizer_apply(kind, var.data_ptr(), None if m is None else m.data_ptr(),
                                     None if v is None else v.data_ptr(), g.data_ptr(), var.numel(), lr, momentum,
                                     int(bool(nesterov)), beta1, beta2, eps)
    return rc == 0
