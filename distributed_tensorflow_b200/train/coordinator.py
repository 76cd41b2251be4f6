"""Thread coordination: ``Coordinator`` and ``QueueRunner``.

The chief's synchronous-aggregation loop is a queue runner: a
daemon thread that keeps running ``sync_op`` until the coordinator asks it to
stop.  ``QueueRunner.create_threads(sess, coord, daemon, start)`` keeps TF's
signature so hooks written against the reference API work unchanged.
"""
from __future__ import annotations

import sys
import threading
import time
from typing import Any, List, Optional, Sequence

from ..framework import errors

__all__ = ["Coordinator", "QueueRunner"]


class Coordinator:
    def __init__(self, clean_stop_exception_types=None):
        self._stop_event = threading.Event()
        self._lock = threading.Lock()
        self._threads: List[threading.Thread] = []
        self._exc_info = None
        self._clean = tuple(clean_stop_exception_types or (errors.OutOfRangeError, errors.CancelledError))

    def should_stop(self) -> bool:
        return self._stop_event.is_set()

    def request_stop(self, ex: Optional[BaseException] = None) -> None:
        with self._lock:
            if ex is not None and not isinstance(ex, self._clean) and self._exc_info is None:
                self._exc_info = (type(ex), ex, ex.__traceback__)
            self._stop_event.set()

    def clear_stop(self) -> None:
        with self._lock:
            self._stop_event.clear()
            self._exc_info = None

    def wait_for_stop(self, timeout: Optional[float] = None) -> bool:
        return self._stop_event.wait(timeout)

    def register_thread(self, thread: threading.Thread) -> None:
        with self._lock:
            self._threads.append(thread)

    def join(self, threads: Optional[Sequence[threading.Thread]] = None, stop_grace_period_secs: float = 10.0,
             ignore_live_threads: bool = True) -> None:
        with self._lock:
            ts = list(self._threads) + list(threads or [])
        deadline = time.time() + stop_grace_period_secs
        for t in ts:
            t.join(max(0.0, deadline - time.time()))
        with self._lock:
            exc, self._exc_info = self._exc_info, None
        if exc is not None:
            raise exc[1].with_traceback(exc[2])

    def stop_on_exception(self):
        coord = self

        class _Ctx:
            def __enter__(self_inner):
                return coord

            def __exit__(self_inner, et, ev, tb):
                if ev is not None:
                    coord.request_stop(ev)
                    return True
                return False
        return _Ctx()


class QueueRunner:
    def __init__(self, queue=None, enqueue_ops: Optional[Sequence[Any]] = None, close_op=None):
        self.queue = queue
        self.enqueue_ops = list(enqueue_ops or [])
        self.close_op = close_op          # run once when the coordinator stops (closes the queue for its consumers)
        self.exceptions_raised: List[BaseException] = []

    def _close_on_stop(self, sess, coord: Coordinator) -> None:
        coord.wait_for_stop()
        try:
            sess.run(self.close_op)
        except Exception:                 # noqa: BLE001 - the session / ps may already be gone at shutdown
            pass

    def _run(self, sess, op, coord: Optional[Coordinator]) -> None:
        try:
            while coord is None or not coord.should_stop():
                try:
                    sess.run(op)
                except (errors.OutOfRangeError, errors.CancelledError):
                    return
                except (errors.DeadlineExceededError,):
                    continue
        except (errors.AbortedError, errors.UnavailableError) as e:
            # the ps went away: the recoverable session rebuilds the loop after recovery
            self.exceptions_raised.append(e)
        except RuntimeError as e:
            if "closed Session" in str(e):
                return
            self.exceptions_raised.append(e)
            if coord is not None:
                coord.request_stop(e)
        except BaseException as e:  # noqa: BLE001
            self.exceptions_raised.append(e)
            if coord is not None:
                coord.request_stop(e)

    def create_threads(self, sess, coord: Optional[Coordinator] = None, daemon: bool = False, start: bool = False
                       ) -> List[threading.Thread]:
        threads = []
        for op in self.enqueue_ops:
            t = threading.Thread(target=self._run, args=(sess, op, coord), name="dtf-queue-runner", daemon=daemon)
            if coord is not None:
                coord.register_thread(t)
            threads.append(t)
        if self.close_op is not None and coord is not None:
            t = threading.Thread(target=self._close_on_stop, args=(sess, coord), name="dtf-queue-closer", daemon=True)
            threads.append(t)             # not registered with the coordinator: it only starts working once that stops
        if start:
            for t in threads:
                t.start()
        return threads
