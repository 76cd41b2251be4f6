"""PS-resident synchronisation state: conditional gradient accumulators and FIFO token queues.

These are the two stateful resources ``SyncReplicasOptimizer`` keeps on the
parameter server:

* :class:`ConditionalAccumulator` -- ``apply_grad(g, local_step)`` drops
  gradients stamped older than the accumulator's ``global_step`` (stale);
  ``take_grad(n)`` blocks until at least ``n`` fresh gradients arrived, returns
  their **mean**, resets the sum and advances nothing by itself (the chief sets
  the new global step explicitly).
* :class:`FIFOQueue` -- the ``sync_token_q``: chief enqueues
  ``total_num_replicas`` tokens carrying the new global step, every worker
  dequeues one per step.

The native C++ implementations live in ``csrc/runtime/ps_state.cpp`` (loaded
through ``utils/native_runtime.py`` when the shared object was built); the
Python classes below define the semantics, are the CPU oracle in tests, and
are the fallback when the library is not built.  The on-device equivalents
(gradient slots + stamps + token flags in NVLink-visible memory) are in
``parallel/ps_engine.py`` / ``csrc/ps_apply.cu``.
"""
from __future__ import annotations

import threading
import time
from collections import deque
from typing import Any, Deque, List, Optional

import torch

from ..framework import errors

__all__ = ["ConditionalAccumulator", "FIFOQueue"]

_POLL = 0.05


class ConditionalAccumulator:
    def __init__(self, dtype=torch.float32, shape=None, name: str = "accumulator"):
        self.name, self.dtype, self.shape = name, dtype, shape
        self._sum: Optional[torch.Tensor] = None
        self._count = 0
        self._global_step = 0
        self._cv = threading.Condition()
        self._closed = False
        self.num_dropped = 0
        self.num_applied = 0

    # -- producer side (every worker) ------------------------------------------------------------
    def apply_grad(self, grad: torch.Tensor, local_step: int) -> bool:
        """Accumulate ``grad`` unless it is stale.  Returns True when accepted."""
        with self._cv:
            if int(local_step) < self._global_step:
                self.num_dropped += 1
                return False
            g = grad.detach()
            if self._sum is None:
                self._sum = g.clone()
            else:
                self._sum.add_(g.to(self._sum.device))
            self._count += 1
            self.num_applied += 1
            self._cv.notify_all()
            return True

    # -- consumer side (chief) ---------------------------------------------------------------------
    def take_grad(self, num_required: int, cancel: Optional[threading.Event] = None,
                  timeout: Optional[float] = None) -> torch.Tensor:
        deadline = None if timeout is None else time.time() + timeout
        with self._cv:
            while self._count < int(num_required):
                if self._closed:
                    raise errors.CancelledError("accumulator %s closed" % self.name)
                if cancel is not None and cancel.is_set():
                    raise errors.CancelledError("take_grad on %s cancelled" % self.name)
                if deadline is not None and time.time() > deadline:
                    raise errors.DeadlineExceededError("take_grad on %s timed out" % self.name)
                self._cv.wait(_POLL)
            mean = self._sum / float(self._count)
            self._sum = None
            self._count = 0
            self._global_step += 1       # TF: TakeGrad bumps the accumulator's own time step
            return mean

    def set_global_step(self, new_global_step: int) -> None:
        with self._cv:
            # never moves backwards (TF semantic)
            self._global_step = max(self._global_step, int(new_global_step))

    def num_accumulated(self) -> int:
        with self._cv:
            return self._count

    @property
    def global_step(self) -> int:
        return self._global_step

    def close(self) -> None:
        with self._cv:
            self._closed = True
            self._cv.notify_all()


class FIFOQueue:
    def __init__(self, capacity: int = -1, name: str = "fifo_queue"):
        self.name = name
        self.capacity = capacity
        self._q: Deque[Any] = deque()
        self._cv = threading.Condition()
        self._closed = False

    def enqueue(self, value: Any) -> None:
        self.enqueue_many([value])

    def enqueue_many(self, values: List[Any]) -> None:
        with self._cv:
            if self._closed:
                raise errors.CancelledError("queue %s is closed" % self.name)
            self._q.extend(values)
            self._cv.notify_all()

    def dequeue(self, cancel: Optional[threading.Event] = None, timeout: Optional[float] = None) -> Any:
        deadline = None if timeout is None else time.time() + timeout
        with self._cv:
            while not self._q:
                if self._closed:
                    raise errors.OutOfRangeError("queue %s is closed and empty" % self.name)
                if cancel is not None and cancel.is_set():
                    raise errors.CancelledError("dequeue on %s cancelled" % self.name)
                if deadline is not None and time.time() > deadline:
                    raise errors.DeadlineExceededError("dequeue on %s timed out" % self.name)
                self._cv.wait(_POLL)
            return self._q.popleft()

    def size(self) -> int:
        with self._cv:
            return len(self._q)

    def is_closed(self) -> bool:
        with self._cv:
            return self._closed

    def close(self, cancel_pending_enqueues: bool = False) -> None:
        with self._cv:
            self._closed = True
            self._cv.notify_all()
