"""ctypes bindings for ``libdtf_runtime.so`` (native accumulator / token queue / bundle I/O / tracer)."""
from __future__ import annotations

import ctypes
import threading
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint32, c_void_p
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ..framework import errors

_OK, _CANCELLED, _DEADLINE, _CLOSED, _BAD = 0, 1, 2, 3, 4


def declare(lib: ctypes.CDLL) -> None:
    if hasattr(lib, "dtf_cpu_optimizer_apply"):
        lib.dtf_cpu_optimizer_apply.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong,
                                                ctypes.c_float, ctypes.c_float, c_int, ctypes.c_float, ctypes.c_float,
                                                ctypes.c_float]
        lib.dtf_cpu_optimizer_apply.restype = c_int
    if hasattr(lib, "dtf_cpu_philox_fill"):
        lib.dtf_cpu_philox_fill.argtypes = [c_void_p, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_ulonglong,
                                            c_int, ctypes.c_float, ctypes.c_float]
        lib.dtf_cpu_philox_fill.restype = c_int
        lib.dtf_cpu_philox_words.argtypes = [c_void_p, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_ulonglong]
        lib.dtf_cpu_philox_words.restype = c_int
    if hasattr(lib, "dtf_gather_rows"):
        lib.dtf_gather_rows.argtypes = [c_void_p, ctypes.c_longlong, c_void_p, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_int]
        lib.dtf_gather_rows.restype = c_int
    lib.dtf_acc_create.restype = c_void_p
    lib.dtf_acc_destroy.argtypes = [c_void_p]
    lib.dtf_acc_apply_grad.argtypes = [c_void_p, c_void_p, c_int64, c_int64]
    lib.dtf_acc_apply_grad.restype = c_int
    lib.dtf_acc_take_grad.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_double]
    lib.dtf_acc_take_grad.restype = c_int
    lib.dtf_acc_wait_count.argtypes = [c_void_p, c_int64, c_void_p, c_double]
    lib.dtf_acc_wait_count.restype = c_int
    for n in ("dtf_acc_size", "dtf_acc_num_accumulated", "dtf_acc_global_step", "dtf_acc_dropped"):
        getattr(lib, n).argtypes = [c_void_p]
        getattr(lib, n).restype = c_int64
    lib.dtf_acc_set_global_step.argtypes = [c_void_p, c_int64]
    lib.dtf_acc_close.argtypes = [c_void_p]
    lib.dtf_queue_create.restype = c_void_p
    lib.dtf_queue_destroy.argtypes = [c_void_p]
    lib.dtf_queue_enqueue_many.argtypes = [c_void_p, c_int64, c_int64]
    lib.dtf_queue_enqueue_many.restype = c_int
    lib.dtf_queue_enqueue_values.argtypes = [c_void_p, c_void_p, c_int64]
    lib.dtf_queue_enqueue_values.restype = c_int
    lib.dtf_queue_dequeue.argtypes = [c_void_p, POINTER(c_int64), c_void_p, c_double]
    lib.dtf_queue_dequeue.restype = c_int
    lib.dtf_queue_size.argtypes = [c_void_p]
    lib.dtf_queue_size.restype = c_int64
    lib.dtf_queue_close.argtypes = [c_void_p]
    lib.dtf_crc32.argtypes = [c_void_p, c_int64]
    lib.dtf_crc32.restype = c_uint32
    lib.dtf_bundle_write.argtypes = [c_char_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int]
    lib.dtf_bundle_write.restype = c_int
    lib.dtf_bundle_read.argtypes = [c_char_p, c_int, c_void_p, c_void_p, c_void_p, c_int]
    lib.dtf_bundle_read.restype = c_int
    lib.dtf_tracer_create.argtypes = [c_int64]
    lib.dtf_tracer_create.restype = c_void_p
    lib.dtf_tracer_destroy.argtypes = [c_void_p]
    lib.dtf_tracer_now_ns.restype = c_int64
    lib.dtf_tracer_record.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32]
    lib.dtf_tracer_drain.argtypes = [c_void_p, c_void_p, c_int64]
    lib.dtf_tracer_drain.restype = c_int64


_SLICE = 0.02          # seconds a native wait blocks before Python re-checks cancellation (GIL released meanwhile)


def _blocking(call, cancel, timeout: Optional[float], what: str) -> None:
    """Run ``call(slice_seconds) -> rc`` until it succeeds.  The native wait loops sleep on a condition variable (woken
    by the producer immediately); every ``_SLICE`` they come back so that cancellation -- a ``threading.Event`` or a
    :class:`parallel.rpc.PeerAwareCancel` -- and the caller's deadline are honoured WITHOUT a watcher thread per call
    (a thread per token dequeue / take_grad was 2/3 of all samples in a profile of sync training)."""
    import time as _time
    deadline = None if timeout is None else _time.time() + timeout
    while True:
        if cancel is not None and cancel.is_set():
            raise errors.CancelledError("%s cancelled" % what)
        step = _SLICE if deadline is None else max(0.0, min(_SLICE, deadline - _time.time()))
        rc = call(step)
        if rc == _OK:
            return
        if rc != _DEADLINE:
            _raise(rc, what)
        if deadline is not None and _time.time() >= deadline:
            raise errors.DeadlineExceededError("%s timed out" % what)


def _raise(rc: int, what: str):
    if rc == _CANCELLED:
        raise errors.CancelledError("%s cancelled" % what)
    if rc == _DEADLINE:
        raise errors.DeadlineExceededError("%s timed out" % what)
    if rc == _CLOSED:
        raise errors.CancelledError("%s: resource closed" % what)
    raise errors.OpError("%s failed (code %d)" % (what, rc))


class NativeAccumulator:
    """Conditional accumulator in the native runtime (fp32 sums on the host: the common case of the control-plane tier).
    Gradients of another dtype (float64, bf16) or living on a GPU keep their dtype and device instead: the first such
    ``apply_grad`` hands this accumulator over to the tensor-based implementation (no D2H round trip per push, no silent
    cast), carrying the accumulator's time step along."""

    def __init__(self, lib: ctypes.CDLL, name: str = "accumulator"):
        self._lib, self.name = lib, name
        self._h = lib.dtf_acc_create()
        self._shape, self._dtype, self._device = None, torch.float32, torch.device("cpu")
        self._py = None
        self._used = False

    def _delegate(self):
        from ..parallel.ps_state import ConditionalAccumulator
        py = ConditionalAccumulator(name=self.name)
        py.set_global_step(int(self._lib.dtf_acc_global_step(self._h)))
        self._py = py
        return py

    def apply_grad(self, grad: torch.Tensor, local_step: int) -> bool:
        if self._py is None and not self._used and (grad.dtype != torch.float32 or grad.device.type != "cpu"):
            self._delegate()
        if self._py is not None:
            return self._py.apply_grad(grad, local_step)
        self._used = True
        self._shape, self._device = tuple(grad.shape), grad.device
        g = grad.detach().to(device="cpu", dtype=torch.float32).contiguous()
        rc = self._lib.dtf_acc_apply_grad(self._h, g.data_ptr(), g.numel(), int(local_step))
        if rc < 0:
            raise errors.InvalidArgumentError("accumulator %s: gradient shape changed" % self.name)
        return rc == 1

    def take_grad(self, num_required: int, cancel: Optional[threading.Event] = None,
                  timeout: Optional[float] = None) -> torch.Tensor:
        if self._py is not None:
            return self._py.take_grad(num_required, cancel, timeout)
        what = "take_grad on %s" % self.name
        _blocking(lambda t: _OK if self._py is not None else self._lib.dtf_acc_wait_count(self._h, int(num_required), None, t),
                  cancel, timeout, what)
        if self._py is not None:             # handed over while this consumer was waiting
            return self._py.take_grad(num_required, cancel, timeout)
        n = self._lib.dtf_acc_size(self._h)
        out = torch.empty(n, dtype=torch.float32)
        rc = self._lib.dtf_acc_take_grad(self._h, int(num_required), out.data_ptr(), n, None, 0.0)
        if rc != _OK:
            _raise(rc, what)
        if self._shape is not None:
            out = out.reshape(self._shape)
        return out.to(self._device) if self._device.type != "cpu" else out

    def set_global_step(self, s: int) -> None:
        if self._py is not None:
            return self._py.set_global_step(s)
        self._lib.dtf_acc_set_global_step(self._h, int(s))

    def num_accumulated(self) -> int:
        if self._py is not None:
            return self._py.num_accumulated()
        return int(self._lib.dtf_acc_num_accumulated(self._h))

    @property
    def global_step(self) -> int:
        if self._py is not None:
            return self._py.global_step
        return int(self._lib.dtf_acc_global_step(self._h))

    @property
    def num_dropped(self) -> int:
        if self._py is not None:
            return self._py.num_dropped
        return int(self._lib.dtf_acc_dropped(self._h))

    def close(self) -> None:
        if self._py is not None:
            self._py.close()
        self._lib.dtf_acc_close(self._h)

    def __del__(self):
        try:
            self._lib.dtf_acc_destroy(self._h)
        except Exception:
            pass


class NativeQueue:
    def __init__(self, lib: ctypes.CDLL, name: str = "fifo_queue"):
        self._lib, self.name = lib, name
        self._h = lib.dtf_queue_create()

    def enqueue(self, value: int) -> None:
        self.enqueue_many([value])

    def enqueue_many(self, values: Sequence[int]) -> None:
        arr = np.asarray([int(v) for v in values], dtype=np.int64)
        rc = self._lib.dtf_queue_enqueue_values(self._h, arr.ctypes.data, arr.size)
        if rc != _OK:
            raise errors.CancelledError("queue %s is closed" % self.name)

    def dequeue(self, cancel: Optional[threading.Event] = None, timeout: Optional[float] = None) -> int:
        out = c_int64(0)

        def once(t):
            rc = self._lib.dtf_queue_dequeue(self._h, ctypes.byref(out), None, t)
            if rc == _CLOSED:
                raise errors.OutOfRangeError("queue %s is closed and empty" % self.name)
            return rc
        _blocking(once, cancel, timeout, "dequeue on %s" % self.name)
        return int(out.value)

    def size(self) -> int:
        return int(self._lib.dtf_queue_size(self._h))

    def close(self, cancel_pending_enqueues: bool = False) -> None:
        self._closed = True
        self._lib.dtf_queue_close(self._h)

    def is_closed(self) -> bool:
        return getattr(self, "_closed", False)

    def __del__(self):
        try:
            self._lib.dtf_queue_destroy(self._h)
        except Exception:
            pass


def bundle_write(lib: ctypes.CDLL, path: str, blobs: List[Tuple[int, bytes]], total_bytes: int, threads: int = 4) -> None:
    n = len(blobs)
    bufs = [ctypes.create_string_buffer(b, len(b)) if len(b) else ctypes.create_string_buffer(1) for _, b in blobs]
    ptrs = (c_void_p * max(n, 1))(*[ctypes.addressof(b) for b in bufs])
    sizes = (c_int64 * max(n, 1))(*[len(b) for _, b in blobs])
    offs = (c_int64 * max(n, 1))(*[o for o, _ in blobs])
    rc = lib.dtf_bundle_write(path.encode(), n, ptrs, sizes, offs, total_bytes, threads)
    if rc:
        raise OSError(-rc, "bundle write failed for %s" % path)


def crc32(lib: ctypes.CDLL, data: bytes) -> int:
    return int(lib.dtf_crc32(data, len(data)))
