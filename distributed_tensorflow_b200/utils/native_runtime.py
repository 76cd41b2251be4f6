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
